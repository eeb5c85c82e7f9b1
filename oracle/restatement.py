"""CPU restatement of the reference hot path (TEST INFRASTRUCTURE — see oracle/__init__.py).

Every function restates one reference symbol as plain functional torch-CPU code over a
``state_dict`` that uses the reference's own key names, so that
(i) golden vectors produced by the real reference classes pin it, and (ii) torch
autograd over it yields the reference's gradients (the reference's backward *is*
autograd over these same ATen ops).  Citations are ``/root/reference/rec_pangu/...``.

Nothing here is imported by the product package.
"""
from itertools import combinations
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

__all__ = [
    "sparse_cols", "dense_cols", "embedding_layer", "get_linear_input", "fm_layer",
    "bi_interaction", "mlp", "lr_layer", "crossnet", "cin", "senet", "bilinear_field_interaction",
    "mhsa", "bce_mean", "deepfm", "xdeepfm", "autoint", "dcn", "fibinet", "afm", "fm", "wdl", "nfm", "mmoe",
    "sharebottom", "omoe", "mlmmoe", "essm", "masknet", "lr", "aitm", "MODEL_FORWARDS",
]


# ---------------------------------------------------------------- feature bookkeeping
def sparse_cols(enc_dict) -> List[str]:
    """Field order = enc_dict insertion order of keys having 'vocab_size' (models/layers/embedding.py:28-30)."""
    return [c for c in enc_dict if 'vocab_size' in enc_dict[c]]


def dense_cols(enc_dict) -> List[str]:
    """Dense order = enc_dict keys having 'min' (models/utils.py:133-135)."""
    return [c for c in enc_dict if 'min' in enc_dict[c]]


# ---------------------------------------------------------------- layers
def embedding_layer(sd, prefix: str, enc_dict, data) -> torch.Tensor:
    """EmbeddingLayer.forward (models/layers/embedding.py:49-63): per-field row lookup, stack -> [B,F,D].

    ``.long()`` cast as embedding.py:61; table has V+1 rows (embedding.py:32).  Out-of-range
    indices raise IndexError exactly like aten::embedding on CPU.
    """
    outs = []
    for col in sparse_cols(enc_dict):
        w = sd[f"{prefix}.embedding_layer.{col}.weight"]
        idx = data[col].long().view(-1, 1)
        outs.append(F.embedding(idx, w))
    return torch.stack(outs, dim=1).squeeze(2)


def get_linear_input(enc_dict, data) -> torch.Tensor:
    """models/utils.py:122-137: stack dense columns -> [B,Nd]."""
    return torch.stack([data[c] for c in dense_cols(enc_dict)], dim=1)


def bi_interaction(e: torch.Tensor) -> torch.Tensor:
    """InnerProductLayer 'Bi_interaction_pooling' (models/layers/interaction.py:36-42): 0.5((sum e)^2 - sum e^2) -> [B,D]."""
    return (torch.sum(e, dim=1) ** 2 - torch.sum(e ** 2, dim=1)) * 0.5


def fm_layer(e: torch.Tensor) -> torch.Tensor:
    """FM_Layer / 'product_sum_pooling' (interaction.py:36-44, 225-235) -> [B,1]."""
    return bi_interaction(e).sum(dim=-1, keepdim=True)


def mlp(sd, prefix: str, x: torch.Tensor, n_hidden: int, stride: int, has_out: bool = True) -> torch.Tensor:
    """MLP (models/layers/deep.py:62-84): Linear->ReLU[->Dropout] * n_hidden [+ Linear(out)].

    ``stride`` is the nn.Sequential index stride between Linear layers: 2 when dropout_rates=0
    (DeepFM/WDL/NFM/FiBiNet), 3 when the default Dropout(0.1) module is present (xDeepFM/AutoInt),
    SURVEY.md App. A-6.  Dropout is identity here (oracle runs eval()/p=0 semantics).
    """
    for i in range(n_hidden):
        x = F.relu(F.linear(x, sd[f"{prefix}.net.{i * stride}.weight"], sd[f"{prefix}.net.{i * stride}.bias"]))
    if has_out:
        k = n_hidden * stride
        x = F.linear(x, sd[f"{prefix}.net.{k}.weight"], sd[f"{prefix}.net.{k}.bias"])
    return x


def lr_layer(sd, prefix: str, enc_dict, data) -> torch.Tensor:
    """LR_Layer (models/layers/shallow.py:14-27): D=1 embedding per field ++ dense -> Linear(F+Nd,1)."""
    sparse = embedding_layer(sd, f"{prefix}.emb_layer", enc_dict, data).squeeze(-1)
    x = torch.cat((sparse, get_linear_input(enc_dict, data)), dim=1)
    return F.linear(x, sd[f"{prefix}.fc.weight"], sd[f"{prefix}.fc.bias"])


def crossnet(sd, prefix: str, x0: torch.Tensor, num_layers: int) -> torch.Tensor:
    """CrossNet (interaction.py:119-141): x_{l+1} = x_l + (w_l . x_l) x_0 + b_l."""
    xi = x0
    for l in range(num_layers):
        w = sd[f"{prefix}.cross_net.{l}.weight.weight"]      # [1, dim]
        b = sd[f"{prefix}.cross_net.{l}.bias"]               # [dim]
        xi = xi + (F.linear(xi, w) * x0 + b)
    return xi


def cin(sd, prefix: str, e: torch.Tensor, units: Sequence[int]) -> torch.Tensor:
    """CompressedInteractionNet (interaction.py:144-171): outer product over fields, 1x1 conv, sum-pool over D, fc."""
    B, _, D = e.shape
    x0, xi, pooled = e, e, []
    for i in range(len(units)):
        had = torch.einsum("bhd,bmd->bhmd", x0, xi).reshape(B, -1, D)
        w = sd[f"{prefix}.cin_layer.layer_{i + 1}.weight"]   # [U, Cin, 1]
        b = sd[f"{prefix}.cin_layer.layer_{i + 1}.bias"]
        xi = F.conv1d(had, w, b).view(B, -1, D)
        pooled.append(xi.sum(dim=-1))
    return F.linear(torch.cat(pooled, dim=-1), sd[f"{prefix}.fc.weight"], sd[f"{prefix}.fc.bias"])


def senet(sd, prefix: str, e: torch.Tensor) -> torch.Tensor:
    """SENET_Layer (interaction.py:238-251): Z=mean_d; A=relu(W2 relu(W1 Z)); V=E*A."""
    z = torch.mean(e, dim=-1)
    a = F.relu(F.linear(F.relu(F.linear(z, sd[f"{prefix}.excitation.0.weight"])), sd[f"{prefix}.excitation.2.weight"]))
    return e * a.unsqueeze(-1)


def bilinear_field_interaction(sd, prefix: str, e: torch.Tensor) -> torch.Tensor:
    """BilinearInteractionLayer 'field_interaction' (interaction.py:55-81): (v_i W_p^T) * v_j, pairs in combinations order."""
    F_ = e.shape[1]
    outs = []
    for p, (i, j) in enumerate(combinations(range(F_), 2)):
        w = sd[f"{prefix}.bilinear_layer.{p}.weight"]
        outs.append(F.linear(e[:, i:i + 1, :], w) * e[:, j:j + 1, :])
    return torch.cat(outs, dim=1)


def mhsa(sd, prefix: str, x: torch.Tensor, num_heads: int, attention_dim: int) -> torch.Tensor:
    """MultiHeadSelfAttention(align_to='output') (models/layers/attention.py:35-101) with the raw
    ``.view(B*H, -1, d)`` head regroup (attention.py:73-75): no scale, softmax(dim=2),
    residual through W_res when D != H*d, ReLU at the end."""
    B = x.shape[0]
    q = F.linear(x, sd[f"{prefix}.W_q.weight"]).view(B * num_heads, -1, attention_dim)
    k = F.linear(x, sd[f"{prefix}.W_k.weight"]).view(B * num_heads, -1, attention_dim)
    v = F.linear(x, sd[f"{prefix}.W_v.weight"]).view(B * num_heads, -1, attention_dim)
    att = torch.softmax(torch.bmm(q, k.transpose(1, 2)), dim=2)
    out = torch.bmm(att, v).view(B, -1, num_heads * attention_dim)
    res = x
    if f"{prefix}.W_res.weight" in sd:
        res = F.linear(x, sd[f"{prefix}.W_res.weight"])
    return F.relu(out + res)


def bce_mean(pred: torch.Tensor, label: torch.Tensor) -> torch.Tensor:
    """torch.nn.BCELoss() mean reduction (e.g. ranking/deepfm.py:31,63); log clamped at -100 like ATen."""
    return F.binary_cross_entropy(pred, label)


# ---------------------------------------------------------------- model forwards
def _finish(logit, data, is_training):
    pred = torch.sigmoid(logit)
    out = {'pred': pred, 'logit': logit}
    if is_training:
        out['loss'] = bce_mean(pred.squeeze(-1), data['label'])
    return out


def _emb_dense(sd, enc_dict, data):
    e = embedding_layer(sd, "embedding_layer", enc_dict, data)
    return e, torch.cat((e.flatten(start_dim=1), get_linear_input(enc_dict, data)), dim=1)


def deepfm(sd, enc_dict, data, is_training=True, hidden_units=(64, 64, 64)):
    """DeepFM.forward (models/ranking/deepfm.py:41-67)."""
    e, x = _emb_dense(sd, enc_dict, data)
    return _finish(fm_layer(e) + mlp(sd, "dnn", x, len(hidden_units), 2), data, is_training)


def xdeepfm(sd, enc_dict, data, is_training=True, dnn_hidden_units=(64, 64, 64), cin_layer_units=(16, 16, 16)):
    """xDeepFM.forward (models/ranking/xdeepfm.py:48-79); MLP has the default Dropout modules (stride 3)."""
    e, x = _emb_dense(sd, enc_dict, data)
    logit = lr_layer(sd, "lr_layer", enc_dict, data) + cin(sd, "cin", e, cin_layer_units) \
        + mlp(sd, "dnn", x, len(dnn_hidden_units), 3)
    return _finish(logit, data, is_training)


def autoint(sd, enc_dict, data, is_training=True, dnn_hidden_units=(64, 64, 64), attention_layers=1,
            num_heads=1, attention_dim=8):
    """AutoInt.forward (models/ranking/autoint.py:59-88)."""
    e, x = _emb_dense(sd, enc_dict, data)
    a = e
    for l in range(attention_layers):
        a = mhsa(sd, f"self_attention.{l}", a, num_heads, attention_dim)
    logit = F.linear(a.flatten(start_dim=1), sd["fc.weight"], sd["fc.bias"])
    logit = logit + mlp(sd, "dnn", x, len(dnn_hidden_units), 3)
    logit = logit + lr_layer(sd, "lr_layer", enc_dict, data)
    return _finish(logit, data, is_training)


def dcn(sd, enc_dict, data, is_training=True, crossing_layers=3):
    """DCN.forward (models/ranking/dcn.py:46-68): cross net over [emb, dense] then fc; no deep tower (App. A-7)."""
    _, x = _emb_dense(sd, enc_dict, data)
    c = crossnet(sd, "crossnet", x, crossing_layers)
    return _finish(F.linear(c, sd["fc.weight"], sd["fc.bias"]), data, is_training)


def fibinet(sd, enc_dict, data, is_training=True, hidden_units=(64, 64, 64)):
    """FiBiNet.forward (models/ranking/fibinet.py:46-77): one bilinear layer shared by raw and SENET embeddings."""
    e = embedding_layer(sd, "embedding_layer", enc_dict, data)
    p = bilinear_field_interaction(sd, "bilinear_interaction", e)
    q = bilinear_field_interaction(sd, "bilinear_interaction", senet(sd, "senet_layer", e))
    comb = torch.flatten(torch.cat([p, q], dim=1), start_dim=1)
    comb = torch.cat([comb, get_linear_input(enc_dict, data)], dim=1)
    logit = lr_layer(sd, "lr", enc_dict, data) + mlp(sd, "dnn", comb, len(hidden_units), 2)
    return _finish(logit, data, is_training)


def fm(sd, enc_dict, data, is_training=True):
    """FM.forward (models/ranking/fm.py)."""
    return _finish(fm_layer(embedding_layer(sd, "embedding_layer", enc_dict, data)), data, is_training)


def wdl(sd, enc_dict, data, is_training=True, hidden_units=(64, 64, 64)):
    """WDL.forward (models/ranking/wdl.py)."""
    _, x = _emb_dense(sd, enc_dict, data)
    return _finish(lr_layer(sd, "lr", enc_dict, data) + mlp(sd, "dnn", x, len(hidden_units), 2), data, is_training)


def nfm(sd, enc_dict, data, is_training=True, hidden_units=(64, 64, 64)):
    """NFM.forward (models/ranking/nfm.py): LR + MLP(bi-interaction pooled [B,D])."""
    e = embedding_layer(sd, "embedding_layer", enc_dict, data)
    logit = lr_layer(sd, "lr", enc_dict, data) + mlp(sd, "dnn", bi_interaction(e), len(hidden_units), 2)
    return _finish(logit, data, is_training)


def mmoe(sd, enc_dict, data, gates: List[torch.Tensor], gates_bias: List[torch.Tensor], is_training=True,
         num_task=2, hidden_dim=(128, 64), bn_training: bool = False, bn_eps: float = 1e-5):
    """MMOE.forward/.loss (models/multi_task/mmoe.py:70-130).  Gates are unregistered python lists in the
    reference (mmoe.py:43-47) so they are passed explicitly.  Towers = Linear->BatchNorm1d->Dropout (no
    activation, App. A-10); Dropout is identity here; BatchNorm uses batch statistics when ``bn_training``."""
    _, hidden = _emb_dense(sd, enc_dict, data)
    experts_out = torch.einsum('ij,jkl->ikl', hidden, sd["experts"]) + sd["experts_bias"]
    outs = []
    for g, gb in zip(gates, gates_bias):
        gate = torch.softmax(hidden @ g + gb, dim=-1)
        outs.append(torch.sum(experts_out * gate.unsqueeze(1), dim=2))
    out = {}
    loss = 0
    for t in range(num_task):
        x = outs[t]
        p = f"task_{t + 1}_dnn"
        for j in range(len(hidden_dim)):
            x = F.linear(x, sd[f"{p}.ctr_hidden_{j}.weight"], sd[f"{p}.ctr_hidden_{j}.bias"])
            x = F.batch_norm(x, sd[f"{p}.ctr_batchnorm_{j}.running_mean"].clone(),
                             sd[f"{p}.ctr_batchnorm_{j}.running_var"].clone(),
                             sd[f"{p}.ctr_batchnorm_{j}.weight"], sd[f"{p}.ctr_batchnorm_{j}.bias"],
                             training=bn_training, momentum=0.1, eps=bn_eps)
        x = torch.sigmoid(F.linear(x, sd[f"{p}.task_last_layer.weight"], sd[f"{p}.task_last_layer.bias"]))
        out[f'task{t + 1}_pred'] = x
        if is_training:
            # mmoe.py:127-128: weight 1/T, BCE on pred + 1e-6
            loss = loss + (1.0 / num_task) * F.binary_cross_entropy(x.squeeze(-1) + 1e-6, data[f'task{t + 1}_label'])
    if is_training:
        out['loss'] = loss
    return out


def _task_towers(sd, tower_inputs, data, is_training, hidden_dim, bn_training, eps=0.0, bn_eps=1e-5):
    """The per-task tower + weighted-BCE block shared by ShareBottom / OMOE / MLMMOE (sharebottom.py:69-92, omoe.py:82-107,
    mlmmoe.py:118-143): Linear -> BatchNorm1d -> Dropout (identity here) per hidden layer, Linear(->1), Sigmoid;
    loss = sum_t (1/T) BCE(pred_t, task{t}_label) — no +1e-6 (that is MMOE only)."""
    T = len(tower_inputs)
    out, loss = {}, 0
    for t in range(T):
        x = tower_inputs[t]
        p = f"task_{t + 1}_dnn"
        for j in range(len(hidden_dim)):
            x = F.linear(x, sd[f"{p}.ctr_hidden_{j}.weight"], sd[f"{p}.ctr_hidden_{j}.bias"])
            x = F.batch_norm(x, sd[f"{p}.ctr_batchnorm_{j}.running_mean"].clone(),
                             sd[f"{p}.ctr_batchnorm_{j}.running_var"].clone(),
                             sd[f"{p}.ctr_batchnorm_{j}.weight"], sd[f"{p}.ctr_batchnorm_{j}.bias"],
                             training=bn_training, momentum=0.1, eps=bn_eps)
        x = torch.sigmoid(F.linear(x, sd[f"{p}.task_last_layer.weight"], sd[f"{p}.task_last_layer.bias"]))
        out[f'task{t + 1}_pred'] = x
        if is_training:
            loss = loss + (1.0 / T) * F.binary_cross_entropy(x.squeeze(-1) + eps if eps else x.squeeze(-1),
                                                             data[f'task{t + 1}_label'])
    if is_training:
        out['loss'] = loss
    return out


def sharebottom(sd, enc_dict, data, is_training=True, num_task=2, hidden_units=(128, 64), bn_training: bool = False):
    """ShareBottom.forward/.loss (models/multi_task/sharebottom.py:54-92): every task tower reads the same
    cat(embeddings.flatten(1), dense) row."""
    _, hidden = _emb_dense(sd, enc_dict, data)
    return _task_towers(sd, [hidden] * num_task, data, is_training, hidden_units, bn_training)


def omoe(sd, enc_dict, data, is_training=True, num_task=2, hidden_dim=(128, 64), bn_training: bool = False):
    """OMOE.forward/.loss (models/multi_task/omoe.py:58-107): experts einsum + bias (omoe.py:73-74), ONE parameter-only
    gate softmax(dim=0) shared by all tasks (omoe.py:79-80)."""
    _, hidden = _emb_dense(sd, enc_dict, data)
    experts_out = torch.einsum('ij,jkl->ikl', hidden, sd["experts"]) + sd["experts_bias"]
    gate = torch.softmax(sd["gate"], dim=0)
    gate_out = torch.einsum('abc,cd->abd', experts_out, gate).squeeze(-1)
    return _task_towers(sd, [gate_out] * num_task, data, is_training, hidden_dim, bn_training)


def mlmmoe(sd, enc_dict, data, level_gates: List[torch.Tensor], gates: List[torch.Tensor], gates_bias: List[torch.Tensor],
           is_training=True, num_task=2, hidden_dim=(128, 64), bn_training: bool = False):
    """MLMMOE.forward/.loss (models/multi_task/mlmmoe.py:74-143): experts (mlmmoe.py:90-91), parameter-only second-level
    gates (mlmmoe.py:96-102), per-task input gates (mlmmoe.py:104-110) and the gated sum (mlmmoe.py:112-117).  The three
    gate lists are unregistered python lists in the reference and are passed explicitly."""
    _, hidden = _emb_dense(sd, enc_dict, data)
    experts_out = torch.einsum('ij,jkl->ikl', hidden, sd["experts"]) + sd["experts_bias"]
    level_out = torch.cat([torch.einsum('abc,cd->abd', experts_out, torch.softmax(g, dim=0)) for g in level_gates], dim=-1)
    outs = []
    for g, gb in zip(gates, gates_bias):
        gate = torch.softmax(hidden @ g + gb, dim=-1)
        outs.append(torch.sum(level_out * gate.unsqueeze(1), dim=2))
    return _task_towers(sd, outs[:num_task], data, is_training, hidden_dim, bn_training)


def essm(sd, enc_dict, data, is_training=True, hidden_dim=(128, 64), dropouts=(0.2, 0.2), w_ctr=0.5):
    """ESSM.forward/.loss (models/multi_task/essm.py:38-75): two MLPs over the flattened embeddings (no dense part),
    click / conversion = sigmoid; loss = BCE(click*conversion, task2_label) + 0.5 * BCE(click, task1_label) — the product
    is what the reference hands to its loss as `conversion` (essm.py:52-56).  Dropout is identity (eval)."""
    hidden = embedding_layer(sd, "embedding_layer", enc_dict, data).flatten(start_dim=1)
    stride = 3 if any(p > 0 for p in dropouts) else 2
    click = torch.sigmoid(mlp(sd, "ctr_layer", hidden, len(hidden_dim), stride))
    conversion = torch.sigmoid(mlp(sd, "cvr_layer", hidden, len(hidden_dim), stride))
    out = {'task1_pred': click, 'task2_pred': conversion}
    if is_training:
        pctrcvr = click * conversion
        out['loss'] = F.binary_cross_entropy(pctrcvr.squeeze(-1), data['task2_label']) + \
            w_ctr * F.binary_cross_entropy(click.squeeze(-1), data['task1_label'])
    return out


def afm(sd, enc_dict, data, is_training=True, hidden_units=(64, 64, 64)):
    """AFM.forward (models/ranking/afm.py:38-67) — the same statements as FiBiNet.forward (fibinet.py:46-77)."""
    return fibinet(sd, enc_dict, data, is_training=is_training, hidden_units=hidden_units)


def _mask_block(sd, prefix, net, mask_input):
    """MaskBlock.forward (layers/interaction.py:279-283): LN_out(W_h (LN_in(net) * W_2 relu(W_1 mask_input)))."""
    n = F.layer_norm(net, (net.shape[1],), sd[f"{prefix}._input_layer_norm.weight"], sd[f"{prefix}._input_layer_norm.bias"])
    m = F.linear(F.relu(F.linear(mask_input, sd[f"{prefix}._mask_layer.0.weight"], sd[f"{prefix}._mask_layer.0.bias"])),
                 sd[f"{prefix}._mask_layer.2.weight"], sd[f"{prefix}._mask_layer.2.bias"])
    h = F.linear(n * m, sd[f"{prefix}._hidden_layer.weight"], sd[f"{prefix}._hidden_layer.bias"])
    return F.layer_norm(h, (h.shape[1],), sd[f"{prefix}._layer_norm.weight"], sd[f"{prefix}._layer_norm.bias"])


def masknet(sd, enc_dict, data, hidden_units=(64, 64, 64), block_num=3, use_parallel=True, **_):
    """MaskNet.forward (ranking/masknet.py:57-86); the MLP keeps its default Dropout(0.1) modules (eval: identity), so the
    Linear layers sit at net.{0,3,6,...}."""
    emb = embedding_layer(sd, "embedding_layer", enc_dict, data)
    x = torch.cat([emb.flatten(start_dim=1), get_linear_input(enc_dict, data)], dim=1)
    if use_parallel:
        out = torch.stack([_mask_block(sd, f"mask_block_list.{i}", x, x) for i in range(block_num)], dim=1).mean(dim=1)
    else:
        out = x
        for i in range(block_num):
            out = _mask_block(sd, f"mask_block_list.{i}", out, x)
    logit = mlp(sd, "mlp", out, len(hidden_units), 3)
    pred = torch.sigmoid(logit)
    res = {"logit": logit, "pred": pred}
    if "label" in data:
        res["loss"] = bce_mean(pred.squeeze(-1), data["label"])
    return res


def lr(sd, enc_dict, data, **_):
    """LR.forward (ranking/lr.py:42-55): sigmoid(LR_Layer(data))."""
    logit = lr_layer(sd, "lr_layer", enc_dict, data)
    pred = torch.sigmoid(logit)
    res = {"logit": logit, "pred": pred}
    if "label" in data:
        res["loss"] = bce_mean(pred.squeeze(-1), data["label"])
    return res


def aitm(sd, enc_dict, data, tower_dims=(400, 400, 400), **_):
    """AITM.forward / .loss (multi_task/aitm.py:61-120), eval mode (dropout off)."""
    x = embedding_layer(sd, "embedding_layer", enc_dict, data).flatten(start_dim=1)
    n = len(tower_dims)
    tc = mlp(sd, "click_tower", x, n, 3, has_out=False)
    tv = mlp(sd, "conversion_tower", x, n, 3, has_out=False)
    info = F.relu(F.linear(tc, sd["info_layer.0.weight"], sd["info_layer.0.bias"]))
    tok = torch.stack([tv, info], dim=1)                                    # [B, 2, d]
    ait = mhsa(sd, "attention_layer", tok, num_heads=1, attention_dim=tok.shape[2]).sum(dim=1)
    click = torch.sigmoid(F.linear(tc, sd["click_layer.0.weight"], sd["click_layer.0.bias"])).squeeze(1)
    conv = torch.sigmoid(F.linear(ait, sd["conversion_layer.0.weight"], sd["conversion_layer.0.bias"])).squeeze(1)
    res = {"task1_pred": click, "task2_pred": conv}
    if "task1_label" in data:
        res["loss"] = (F.binary_cross_entropy(click, data["task1_label"]) + F.binary_cross_entropy(conv, data["task2_label"])
                       + 0.6 * torch.clamp(conv - click, min=0).sum())
    return res


MODEL_FORWARDS = {
    'AFM': afm,
    'DeepFM': deepfm, 'xDeepFM': xdeepfm, 'AutoInt': autoint, 'DCN': dcn, 'FiBiNet': fibinet,
    'FM': fm, 'WDL': wdl, 'NFM': nfm, 'MMOE': mmoe, 'ShareBottom': sharebottom, 'OMOE': omoe, 'MLMMOE': mlmmoe, 'ESSM': essm,
    'MaskNet': masknet, 'LR': lr, 'AITM': aitm,
}
