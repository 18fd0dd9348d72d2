"""Top-level `dynamicemb_extensions` — the import name of the reference's pybind module (corelib/dynamicemb/src/module_bind.cu:22-43,
`from dynamicemb_extensions import get_table_range, segmented_unique_cuda`).  Everything lives in `dynamicemb.dynamicemb_extensions`
(ctypes over librecsys_b200.so); this module only re-exports it under the name the reference's Python imports."""
from dynamicemb.dynamicemb_extensions import *  # noqa: F401,F403
from dynamicemb.dynamicemb_extensions import EvictStrategy, InitializerMode, InsertResult, OptimizerType, ScorePolicy  # noqa: F401
