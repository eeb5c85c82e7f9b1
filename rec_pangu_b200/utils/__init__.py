"""Small host-side helpers kept from the reference's utils package (rec_pangu/utils/__init__.py): only what the
trainer uses.  The pypi version-check thread and the faiss recall evaluators are out of scope."""
from .json_utils import beautify_json
from .gpu_utils import get_gpu_usage, set_device
