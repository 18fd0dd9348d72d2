"""hstu — B200-native HSTU jagged attention behind `hstu_attn_varlen_func`
(reference: third_party/FBGEMM/fbgemm_gpu/experimental/hstu/hstu/__init__.py, cuda_hstu_attention.py:677)."""
from .cuda_hstu_attention import hstu_attn_varlen_func  # noqa: F401
from . import hstu_ops_gpu  # noqa: F401
