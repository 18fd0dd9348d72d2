"""Same ops under the reference's module path (`hstu.hstu_blackwell.hstu_ops_gpu`)."""
from ..hstu_ops_gpu import hstu_varlen_bwd_100, hstu_varlen_fwd_100  # noqa: F401
