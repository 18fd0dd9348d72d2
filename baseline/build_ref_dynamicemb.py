"""Build the UNMODIFIED reference DynamicEmb CUDA extension (dynamicemb_extensions) from
/root/reference/corelib/dynamicemb/src into baseline/_ref/ (git-ignored, travels with gpurun).

Test/bench infrastructure only: used on the GPU box as the *reference kernels* to generate
golden fixtures (tests/golden/) and as a parity cross-check.  Never imported by the product.
Mirrors the flags of corelib/dynamicemb/setup.py:108-142 restricted to sm_100.
"""
import os, sys, glob
os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
os.environ.setdefault("MAX_JOBS", "6")
from torch.utils.cpp_extension import load

REF = "/root/reference/corelib/dynamicemb/src"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref", "dynamicemb_ext")
os.makedirs(OUT, exist_ok=True)
EXCL = {"lookup_torch_binding.cu", "get_table_range_torch_binding.cu", "expand_table_ids_torch_binding.cu"}
srcs = []
for root, _, files in os.walk(REF):
    for f in files:
        if f in EXCL:
            continue
        if f.endswith((".cu", ".cpp")):
            srcs.append(os.path.join(root, f))
srcs.sort()
print(len(srcs), "sources")
load(
    name="dynamicemb_extensions",
    sources=srcs,
    extra_include_paths=[REF],
    extra_cflags=["-O3", "-w", "-DDEMB_USE_PYBIND11"],
    extra_cuda_cflags=["-O3", "-lineinfo", "--expt-relaxed-constexpr", "--expt-extended-lambda", "--use_fast_math",
                       "-gencode", "arch=compute_100,code=sm_100", "-w",
                       "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__",
                       "-U__CUDA_NO_HALF2_OPERATORS__", "-U__CUDA_NO_BFLOAT16_CONVERSIONS__", "-DDEMB_USE_PYBIND11"],
    extra_ldflags=["-Wl,--no-as-needed", "-lcuda"],
    build_directory=OUT,
    verbose=True,
    is_python_module=False,
)
print("built into", OUT)
