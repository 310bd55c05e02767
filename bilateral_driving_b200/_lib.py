"""ctypes binding of libbds_b200.so (the C ABI declared in include/bds.h).

The library is the product: there is NO Python / torch / CPU fallback.  Importing this module
without the built library raises, and every call checks the return code and raises
``BdsError`` with ``bds_last_error()``.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# BDS_LIB selects another build of the same library (profiling aid: libbds_b200_stats.so); never a fallback
LIB_PATH = os.path.join(_HERE, os.environ.get("BDS_LIB", "libbds_b200.so"))
MAX_LEVELS = 4
TILE = 16
COUNTERS_LEN = 4096  # BDS_COUNTERS_LEN (include/bds.h)
ABI_VERSION = 5      # BDS_ABI_VERSION: 5 = + compact SH gradient exchange, tv levels; 4 = + bds_densify_stats; 3 = + bds_project_bwd_extras; 2 = moment-form gradient records
SPLAT_FLOATS = 12


class BdsError(RuntimeError):
    pass


class BilateralDesc(C.Structure):
    _fields_ = [
        ("n_levels", C.c_int),
        ("L", C.c_int * MAX_LEVELS),
        ("GY", C.c_int * MAX_LEVELS),
        ("GX", C.c_int * MAX_LEVELS),
        ("factor", C.c_int * MAX_LEVELS),
    ]

    @classmethod
    def make(cls, sizes_xyl, factors):
        """sizes_xyl: sequence of (grid_X, grid_Y, grid_W) as in the reference ctor;
        factors: per-level guidance factor or None (full-resolution guidance)."""
        d = cls()
        d.n_levels = len(sizes_xyl)
        if not 1 <= d.n_levels <= MAX_LEVELS:
            raise ValueError(f"1..{MAX_LEVELS} bilateral levels supported, got {d.n_levels}")
        for i, (gx, gy, gl) in enumerate(sizes_xyl):
            d.GX[i], d.GY[i], d.L[i] = int(gx), int(gy), int(gl)
            d.factor[i] = 0 if factors is None else int(factors[i])
        return d


class RenderDesc(C.Structure):
    _fields_ = [
        ("n_gauss", C.c_int), ("n_cams", C.c_int), ("width", C.c_int), ("height", C.c_int),
        ("near_plane", C.c_float), ("far_plane", C.c_float), ("radius_clip", C.c_float), ("eps2d", C.c_float),
        ("antialiased", C.c_int), ("row_begin", C.c_int), ("row_end", C.c_int),
        ("raw_params", C.c_int), ("sh_degree", C.c_int), ("sh_K", C.c_int),
    ]


class EpilogueDesc(C.Structure):
    _fields_ = [
        ("mode", C.c_int), ("channels", C.c_int), ("expected_depth", C.c_int), ("bil", BilateralDesc),
    ]


def _load():
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA library is the product and there is no fallback. "
            "Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C bilateral_driving_b200/csrc`).")
    lib = C.CDLL(LIB_PATH)
    lib.bds_last_error.restype = C.c_char_p
    lib.bds_abi_version.restype = C.c_int
    if lib.bds_abi_version() != ABI_VERSION:
        raise ImportError(f"{LIB_PATH} has ABI version {lib.bds_abi_version()}, this package needs {ABI_VERSION}: rebuild it")
    lib.bds_device_arch.restype = C.c_int
    if hasattr(lib, "bds_launch_count"):
        lib.bds_launch_count.restype = C.c_ulonglong
    for name in ("bds_bilateral_workspace_bytes", "bds_bin_count_workspace_bytes",
                 "bds_bin_sort_workspace_bytes", "bds_composite_workspace_bytes"):
        if hasattr(lib, name):
            getattr(lib, name).restype = C.c_size_t
    return lib


lib = _load()

# every symbol include/bds.h declares; tests/test_abi.py checks they are all exported
ABI_SYMBOLS = (
    "bds_last_error", "bds_abi_version", "bds_device_arch", "bds_launch_count",
    "bds_bilateral_workspace_bytes", "bds_bilateral_fwd", "bds_bilateral_bwd",
    "bds_bilagrid_slice_fwd", "bds_bilagrid_slice_bwd", "bds_tv_fwd_bwd", "bds_tv_levels_fwd_bwd",
    "bds_sh_fwd", "bds_sh_bwd",
    "bds_project_fwd", "bds_project_bwd", "bds_project_bwd_extras", "bds_project_bwd_compact_sh", "bds_sh_expand_bwd",
    "bds_bin_count_workspace_bytes", "bds_bin_count", "bds_bin_sort_workspace_bytes", "bds_bin_sort",
    "bds_composite_workspace_bytes", "bds_composite_fwd", "bds_composite_bwd",
    "bds_slot_keep", "bds_composite_fwd_masked",
    "bds_loss_fwd_bwd", "bds_densify_stats",
)


def check(rc: int, what: str = ""):
    if rc != 0:
        raise BdsError(f"{what or 'bds call'} failed (rc={rc}): {lib.bds_last_error().decode()}")


def ptr(t):
    """Device pointer of a torch tensor (or NULL for None)."""
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())


def ptr_array(tensors):
    """Host array of device pointers (entries may be None -> NULL)."""
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = 0 if t is None else t.data_ptr()
    return arr


def stream_ptr():
    """Current torch stream of the current device.  The C ABI launches on the CURRENT device, so every op runs under
    ``device_scoped`` (current device = the device of its tensors)."""
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _first_cuda_device(objs):
    import torch

    for o in objs:
        if torch.is_tensor(o):
            if o.is_cuda:
                return o.device
        elif isinstance(o, (list, tuple)):
            d = _first_cuda_device(o)
            if d is not None:
                return d
    return None


def device_scoped(fn):
    """Decorator for the forward / backward of an autograd Function: run it with the device of its (first CUDA)
    tensor argument as the current device, so that kernels, memsets, function attributes, the stream taken by
    ``stream_ptr()`` and torch's own allocations all belong to the device the tensors live on (a model on cuda:1
    without torch.cuda.set_device would otherwise launch on cuda:0)."""
    import functools

    @functools.wraps(fn)
    def scoped(ctx, *args):
        import torch

        dev = _first_cuda_device(args)
        if dev is None:
            try:
                dev = _first_cuda_device(ctx.saved_tensors)
            except Exception:
                dev = None
        if dev is None:   # CPU tensors: the op itself raises BdsError (no CPU fallback)
            return fn(ctx, *args)
        with torch.cuda.device(dev):
            return fn(ctx, *args)

    return scoped


def require_cuda(*tensors):
    import torch

    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise BdsError("bds operators run on CUDA tensors only (no CPU fallback exists)")
        if t.dtype not in (torch.float32, torch.int32, torch.int64):
            raise BdsError(f"unsupported dtype {t.dtype}: the path computes in fp32")
