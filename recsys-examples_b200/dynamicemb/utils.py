"""`dynamicemb.utils` — import path of the reference (utils.py:41-57)."""
from .types import torch_to_dyn_emb  # noqa: F401
