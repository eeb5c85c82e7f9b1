"""rec_pangu_b200 — B200-native (sm_100a) hot path behind the rec_pangu ranking / multi-task API.

Host side mirrors the reference's import paths (rec_pangu.models.ranking.DeepFM, rec_pangu.trainer.RankTrainer,
rec_pangu.model_pipeline.train_model, rec_pangu.dataset.get_dataloader ...); compute is the C-ABI kernel library
``librec_pangu_b200.so`` (include/rec_pangu_b200.h).  See DESIGN.md.
"""
__version__ = '0.1.0'
