"""`hstu.hstu_blackwell` — the import path the reference's fused layer tries for the sm_100 attention ops
(examples/hstu/ops/fused_hstu_op.py:50-58: `from hstu.hstu_blackwell import hstu_ops_gpu`).  The ops are this package's
`hstu.hstu_ops_gpu` (hand-written tcgen05 kernels behind the C ABI), same names and parameters as the reference's
`hstu_blackwell/hstu_ops_gpu.py` (`tests/test_api_signatures_cpu.py`)."""
from .. import hstu_ops_gpu  # noqa: F401
