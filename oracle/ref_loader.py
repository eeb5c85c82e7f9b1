"""Import the UNMODIFIED reference package from oracle/_ref (TEST INFRASTRUCTURE; see oracle/build_ref.py).

Only `bench.py`'s reference arm / `cpu_baseline` leg (and tests) use this, in a process that never imports this repo's own
`rec_pangu` alias package: both are called `rec_pangu`, so `load()` refuses to run when the alias is already imported.
`faiss` and `dgl` (used only by the reference's recall / graph models, outside the hot path) are stubbed exactly as
SURVEY.md App. B-2 describes.
"""
import importlib.machinery
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, '_ref')


def available() -> bool:
    return os.path.isdir(os.path.join(REF_DIR, 'rec_pangu', 'models'))


def load():
    """Returns the reference's `rec_pangu.models.ranking` and `rec_pangu.models.multi_task` modules."""
    if not available():
        raise ImportError('oracle/_ref is not built (python oracle/build_ref.py in the build container)')
    mod = sys.modules.get('rec_pangu')
    if mod is not None and not os.path.abspath(getattr(mod, '__file__', '') or '').startswith(REF_DIR):
        raise ImportError('this process already imported the rec_pangu alias package of this repo; the reference must be '
                          'loaded in a process of its own')
    for name in ['faiss', 'dgl', 'dgl.function']:
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__spec__ = importlib.machinery.ModuleSpec(name, None)
            sys.modules[name] = m
    sys.modules['dgl'].function = sys.modules['dgl.function']
    if not hasattr(sys.modules['dgl'], 'DGLGraph'):
        sys.modules['dgl'].DGLGraph = object
    os.environ.setdefault('WANDB_MODE', 'disabled')
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    elif sys.path[0] != REF_DIR:
        sys.path.remove(REF_DIR)
        sys.path.insert(0, REF_DIR)
    import rec_pangu.models.ranking as ranking
    import rec_pangu.models.multi_task as multi_task
    assert os.path.abspath(sys.modules['rec_pangu'].__file__).startswith(REF_DIR), sys.modules['rec_pangu'].__file__
    return ranking, multi_task
