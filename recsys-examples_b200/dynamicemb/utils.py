"""`dynamicemb.utils` — import path of the reference (utils.py): dtype helper and the TorchRec collection types the examples test with
`isinstance` (the empty set when torchrec is not installed)."""
from .types import DTYPE_NUM_BYTES, torch_to_dyn_emb  # noqa: F401

try:
    from torchrec.modules.embedding_modules import EmbeddingBagCollection, EmbeddingCollection
    TORCHREC_TYPES = {EmbeddingBagCollection, EmbeddingCollection}
except ImportError:          # no torchrec: no module can be one of its collections
    TORCHREC_TYPES = set()
