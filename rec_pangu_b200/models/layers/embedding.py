"""EmbeddingLayer — same constructor, parameters and state_dict keys as the reference
(rec_pangu/models/layers/embedding.py:11-71), forward = ONE multi-table gather kernel (rpb_gather_fwd)."""
from typing import Dict, Optional, Sequence

import torch
from torch import nn

from ... import ops
from ..utils import sparse_feature_names, dense_feature_names


class EmbeddingLayer(nn.Module):
    def __init__(self, enc_dict: Dict[str, dict], embedding_dim: int) -> None:
        super().__init__()
        self.enc_dict = enc_dict
        self.embedding_dim = embedding_dim
        # nn.Embedding modules are parameter containers only (state_dict key
        # `embedding_layer.<col>.weight`, [vocab_size+1, D], embedding.py:31-34); their forward is never called.
        self.embedding_layer = nn.ModuleDict()
        self.emb_feature = []
        from ... import dist as _dist
        kw = {'device': 'meta'} if _dist.deferred_tables.active else {}      # shape only: shards are attached later
        for col in sparse_feature_names(enc_dict):
            self.emb_feature.append(col)
            self.embedding_layer.update({col: nn.Embedding(num_embeddings=enc_dict[col]['vocab_size'] + 1,
                                                           embedding_dim=embedding_dim, **kw)})
        self.dense_feature = dense_feature_names(enc_dict)
        # 'dense': backward returns fresh zero-filled [rows, D] grads exactly like nn.Embedding (reference behaviour);
        # 'persistent': grads live in persistent buffers that are re-zeroed sparsely (ops.GradStore) — same .grad
        # contents, O(batch) instead of O(vocabulary) traffic per step.
        self.grad_mode = 'dense'
        self._grad_store = ops.GradStore()
        self._shards = None                # dist.ShardedTables when the tables are row-sharded over several GPUs

    def attach_shards(self, st):
        """Switch to row-sharded tables in NVLink peer memory (rec_pangu_b200.dist.shard_model_tables): each
        nn.Embedding container now holds only this rank's shard [ceil(rows/G), D] as its Parameter."""
        self._shards = st
        for f, c in enumerate(self.emb_feature):
            trainable = self.embedding_layer[c].weight.requires_grad
            self.embedding_layer[c].weight = nn.Parameter(st.weights[f], requires_grad=trainable)
            self.embedding_layer[c].weight._rpb_shards = st          # lets another layer's gather find these shards (LR tables)

    def set_weights(self, col_name: str, embedding_matrix: torch.Tensor, trainable: Optional[bool] = True) -> None:
        """embedding.py:36-47."""
        self.embedding_layer[col_name].weight = nn.Parameter(embedding_matrix)
        if not trainable:
            self.embedding_layer[col_name].weight.requires_grad = False

    def tables(self):
        return [self.embedding_layer[c].weight for c in self.emb_feature]

    def feature_row(self, X: Dict[str, torch.Tensor], with_dense: bool = True, want_fm: bool = False,
                    lr_tables: Optional[Sequence[torch.Tensor]] = None, want_x: bool = True):
        """Fused entry used by the model forwards: returns (x [B, ldx], fm [B] | None, lr_in | None) where
        x = [emb_0 | ... | emb_{F-1} | dense_0..dense_{Nd-1} | 0-pad] (see include/rec_pangu_b200.h)."""
        idx = [X[c] for c in self.emb_feature]
        dense = [X[c] for c in self.dense_feature] if with_dense else []
        if self._shards is not None:
            x, fm, _ = ops.gather_sharded(self._shards, self.tables(), idx, dense, want_fm=want_fm)
            lr_in = None
            if lr_tables is not None:
                # the D = 1 tables of the LR_Layer are row-sharded the same way (owner = id mod G): a second gather launch
                # over their shards yields [lr_f[idx_f] (F) | dense (Nd) | 0-pad] = the LR input row (shallow.py:22-27)
                st_lr = getattr(lr_tables[0], '_rpb_shards', None)
                if st_lr is None:
                    raise RuntimeError('row-sharded embedding tables need row-sharded LR tables (dist.shard_model_tables shards both)')
                lr_in, _, _ = ops.gather_sharded(st_lr, list(lr_tables), idx, dense)
            return x, fm, lr_in
        return ops.gather(self.tables(), idx, dense, lr_tables=lr_tables, want_fm=want_fm,
                          grad_store=self._grad_store if self.grad_mode == 'persistent' else None, want_x=want_x)

    def clean_grads(self):
        """Sparse re-zero of the persistent grad buffers (no-op in 'dense' mode)."""
        self._grad_store.clean()
        if self._shards is not None:
            ops.sharded_clean(self._shards)

    def forward(self, X: Dict[str, torch.Tensor], name: Optional[str] = None) -> torch.Tensor:
        """[B, F, D] (name=None) — embedding.py:58-63; a strided view of the feature row, no stack copy."""
        if name is None:
            x, _, _ = self.feature_row(X, with_dense=False)
            F, D = len(self.emb_feature), self.embedding_dim
            return x[:, :F * D].view(x.shape[0], F, D)
        if 'seq' in name:
            raise NotImplementedError('sequence-feature lookup belongs to the sequence-recall workload (out of scope)')
        x, _, _ = ops.gather([self.embedding_layer[name].weight], [X[name]])
        return x[:, :self.embedding_dim].view(x.shape[0], 1, self.embedding_dim)
