"""ctypes binding of librecsys_b200.so (the C ABI declared in include/dynamicemb_b200.h).

The product path has no CPU fallback: if the CUDA library is missing, importing this module
raises.  Device pointers are passed as integers (`tensor.data_ptr()`), the stream as the raw
`cudaStream_t` of torch's current stream.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "librecsys_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
        "There is no CPU fallback for the DynamicEmb / HSTU hot paths."
    )
lib = ctypes.CDLL(LIB_PATH)

P = ctypes.c_void_p
I64 = ctypes.c_int64
I32 = ctypes.c_int
U64 = ctypes.c_uint64
F32 = ctypes.c_float

_SIGS = {
    "demb_table_init": (I32, [P, I64, I64, I32, P]),
    "demb_table_lookup": (I32, [P, P, I64, I32, I64, P, P, I32, P, U64, P, P, P, P]),
    "demb_table_insert_workspace_bytes": (I64, [I64]),
    "demb_table_insert": (I32, [P, P, I64, I32, I64, P, I64, P, P, I32, P, U64, P, I32, P, P, P, P, P, P, P, P, P, I64, P]),
    "demb_table_erase": (I32, [P, P, I64, I32, P, I64, P, P, P, P]),
    "demb_counter_update": (I32, [P, P, P, P, I64, I64, I32, P]),
    "demb_table_export": (I32, [P, P, I64, I32, I64, I64, I64, U64, I32, I32, P, P, P, P, P]),
    "demb_get_table_range": (I32, [P, P, I32, I64, P, P]),
    "demb_segmented_unique_workspace_bytes": (I64, [I64, I32]),
    "demb_unique_scratch_bytes": (I64, [I64, I32]),
    "demb_unique_scratch_init": (I32, [P, I64, P]),
    "demb_segmented_unique": (I32, [I64, P, P, P, I32, P, P, P, P, P, P, P, P, P, I64, P]),
    "demb_flagged_compact_workspace_bytes": (I64, [I64]),
    "demb_flagged_compact": (I32, [I64, P, P, P, P, P, I32, P, I64, P]),
    "demb_expand_table_ids": (I32, [P, I32, I64, P, P]),
    "demb_lookup_forward": (I32, [P, P, I64, I32, P, I64, I32, P, I64, P, P, I32, P, I64, I32, I32, P, I32, F32, P, P, P]),
    "demb_gather_forward": (I32, [P, I64, I32, I64, P, P, P, I64, I32, I32, P, I32, P, P]),
    "demb_rows_from_slots": (I32, [I64, P, P, P, P, P]),
    "demb_init_rows": (I32, [P, I64, I32, I64, P, P, I32, F32, F32, F32, F32, U64, P, P, F32, P, P, P]),
    "demb_copy_rows": (I32, [P, I64, I32, I64, P, P, I64, I32, P]),
    "demb_backward_workspace_bytes": (I64, [I64, I32]),
    "demb_backward": (I32, [P, I64, I32, I64, P, I64, P, P, I64, P, I64, I32, I32, I32, F32, F32, F32, F32, F32, F32, F32, P, P, P, P, P, I64, P]),
    "demb_backward_sort": (I32, [I32, I64, P, I64, P, I64, I32, I32, P, P, P, I64, P]),
    "demb_backward_apply": (I32, [P, I64, I32, I64, P, I64, P, P, I64, P, I64, I32, I32, I32, F32, F32, F32, F32, F32, F32, F32, P, P, P, P, I64, P]),
    "demb_update_rows": (I32, [P, I64, I32, I64, P, P, I64, I32, F32, F32, F32, F32, F32, F32, F32, P]),
    "demb_flat_table_copy": (I32, [P, P, I64, P, I64, P, P, I64, P, I64, I64, I32, I32, P]),
    "demb_flat_table_update": (I32, [P, P, P, I64, P, P, I64, P, I64, I32, F32, F32, F32, F32, F32, F32, F32, P]),
    "demb_bucket_of": (I32, [P, I64, I64, P, P, P, P]),
    "demb_fill_i32": (I32, [P, I64, I32, P]),
    "demb_train_prefetch_workspace_bytes": (I64, [I64, I32]),
    "demb_train_prefetch": (I32, [P, P, I64, I32, P, P, P, P, I64, I32, P, I64, P, P, P, I32, P, I32, P, U64, I32, I32, F32, F32, F32, F32, U64, P, F32,
                                  P, P, P, P, P, P, P, P, P, I64, P]),
    "demb_counter_update_n": (I32, [P, P, P, P, I64, P, I64, I32, P]),
    "demb_set_option": (I32, [I32, I32]),
    "demb_get_option": (I32, [I32]),
    "demb_profile_enable": (I32, [I32]),
    "demb_profile_read": (I32, [P]),
    "demb_shard_layout": (I32, [I32, I64, I64, I32, P]),
    "demb_shard_route_workspace_bytes": (I64, [I64, I32, I32]),
    "demb_shard_route": (I32, [I32, I32, I32, I32, I64, I64, P, P, I64, P, P, P, P, P, P, P, P, I64, P]),
    "demb_shard_recv_workspace_bytes": (I64, [I32, I32]),
    "demb_shard_recv": (I32, [I32, I32, I32, I32, I64, I64, I64, P, P, P, P, P, P, P, P, I64, P]),
    "demb_shard_gather_to_peers": (I32, [P, I64, I32, I64, P, P, P, P, P]),
    "demb_shard_gather_to_peers_part": (I32, [P, I64, I32, I64, P, P, P, P, P, I32, P]),
    "demb_train_prefetch_hook": (I32, [P, P]),
    "demb_peer_barrier": (I32, [I32, I32, I64, I64, I32, P, P, I32, P, P]),
    "demb_zero_i64": (I32, [P, I64, P]),
    "demb_ipc_alloc": (I32, [I64, P, P]),
    "demb_ipc_open": (I32, [P, P]),
    "demb_ipc_close": (I32, [P]),
    "demb_ipc_free": (I32, [P]),
    "demb_bucketize_workspace_bytes": (I64, [I64, I32]),
    "demb_block_bucketize_sparse_features": (I32, [I64, I64, I32, P, P, P, P, P, P, P, P, P, P, I64, P]),
    "demb_block_bucketize_sparse_features_n": (I32, [I64, I64, I32, I64, P, P, P, P, P, P, P, P, P, P, I64, P]),
}
for _name, (_res, _args) in _SIGS.items():
    _f = getattr(lib, _name)
    _f.restype = _res
    _f.argtypes = _args

DEMB_ERR_ARG = -1000
DEMB_ERR_WORKSPACE = -1001


def check(rc: int, what: str = "") -> None:
    if rc == 0:
        return
    if rc == DEMB_ERR_ARG:
        raise ValueError(f"{what}: invalid argument")
    if rc == DEMB_ERR_WORKSPACE:
        raise RuntimeError(f"{what}: workspace too small")
    raise RuntimeError(f"{what}: CUDA error {-rc} ({torch.cuda.get_device_name() if torch.cuda.is_available() else 'no GPU'})")


def ptr(t):
    """Device pointer of a tensor, or NULL for None."""
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


# ---- measurement hook (bench.py): CUDA events around selected C-ABI calls + a count of this library's kernel launches
PROFILE = None            # dict name -> list[(start_event, end_event)] when enabled
LAUNCHES = [0]            # number of librecsys_b200 kernels launched (ours, excluding cub / memset nodes)


def launch(name: str, kernels: int, fn, *args):
    LAUNCHES[0] += kernels
    if PROFILE is None:
        return fn(*args)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    rc = fn(*args)
    e.record()
    PROFILE.setdefault(name, []).append((s, e))
    return rc


_ws_cache = {}


def workspace(nbytes: int, device) -> torch.Tensor:
    """Per-(device, stream) grow-only scratch buffer: the C ABI never allocates."""
    key = (str(device), torch.cuda.current_stream(device).cuda_stream)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf
