"""ctypes binding of liby3b200.so (C ABI declared in include/y3b200.h).

The shared library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  Nothing here computes anything:
if the library is missing, or no B200 is present, every compute entry point raises -- there is no CPU fallback.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# liby3b200.so is the release library (no environment knobs, no debug branches).  Measurement tools set Y3_PROF_LIB=1 to
# load liby3b200_prof.so instead: the same sources compiled with -DY3_PROFILING (Y3_* experiment knobs, per-CTA
# timestamps, kernel ablations).  Same ABI, same results with the knobs at their defaults.
LIB_PATH = os.path.join(_HERE, "lib", "liby3b200_prof.so" if os.environ.get("Y3_PROF_LIB") == "1" else "liby3b200.so")

Y3_OK, Y3_ERR_INVALID, Y3_ERR_UNSUPPORTED, Y3_ERR_CUDA, Y3_ERR_STATE = 0, 1, 2, 3, 4
OP_CONV, OP_SHORTCUT, OP_UPSAMPLE, OP_CONCAT, OP_YOLO, OP_MAXPOOL = range(6)


class LayerDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("op", "src0", "src1", "ksize", "stride", "filters", "pad", "batch_normalize", "activation")]


class LayerPlan(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("H", "W", "C", "kernel", "fused_add", "fused_upsample", "buffer", "chan_offset", "pix_stride",
                 "block_n", "swizzle", "stages", "flat", "padded")] + [("arena_offset", C.c_int64)]


class ChainStep(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("layer", "posts", "chained", "dep_step", "res_step", "tiles", "ctas", "tiles_n", "rows_per_group",
                 "rev", "rot", "run_first", "run_len", "vshift")]


class Y3Error(RuntimeError):
    """CUDA / state failure reported by the library."""


class Y3Unsupported(NotImplementedError):
    """The graph or shape is valid in the reference but not implemented on the B200 path (no fallback exists)."""


_p = C.c_void_p
_i = C.c_int
_f = C.c_float
_i64 = C.c_int64

# name -> (restype, argtypes); also the list of symbols the CPU test-suite checks for
SIGNATURES = {
    "y3_last_error": (C.c_char_p, []),
    "y3_version": (_i, []),
    "y3_crc32c": (C.c_uint32, [C.c_uint32, _p, _i64]),
    "y3_ctx_create": (_i, [_i, C.POINTER(_p)]),
    "y3_ctx_destroy": (None, [_p]),
    "y3_ctx_sm_count": (_i, [_p]),
    "y3_net_create": (_i, [_p, C.POINTER(LayerDesc), _i, _i, _i, _i, _i, C.POINTER(_p)]),
    "y3_net_destroy": (None, [_p]),
    "y3_net_num_convs": (_i, [_p]),
    "y3_net_num_outputs": (_i, [_p]),
    "y3_net_get_plan": (_i, [_p, C.POINTER(LayerPlan), _i]),
    "y3_net_arena_bytes": (_i64, [_p]),
    "y3_net_read_layer": (_i, [_p, _i, _i, _p]),
    "y3_net_output_shape": (_i, [_p, _i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "y3_net_load_conv": (_i, [_p, _i, _p, _p, _p, _p, _p, _p, _f]),
    "y3_net_forward": (_i, [_p, _p, _i, C.POINTER(_p), _i, _p]),
    "y3_net_forward_pitched": (_i, [_p, _p, _i, C.POINTER(_p), C.POINTER(_i), _i, _p]),
    "y3_net_forward_u8": (_i, [_p, _p, _i, C.POINTER(_p), C.POINTER(_i), _i, _p]),
    "y3_net_num_steps": (_i, [_p]),
    "y3_net_chain_plan": (_i, [_p, _i, C.POINTER(ChainStep), _i]),
    "y3_net_forward_timed": (_i, [_p, _p, _i, C.POINTER(_p), _i, _p, _p, _p, _i]),
    "y3_decode": (_i, [_p, C.POINTER(_p), C.POINTER(_i), C.POINTER(_i), _i, _p, _i, _i, _p, _p, _p, _p, _p, _p]),
    "y3_decode_pitched": (_i, [_p, C.POINTER(_p), C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), _i, _p, _i, _i, _p, _p, _p, _p, _p, _p]),
    "y3_class_reduce": (_i, [_p, _p, _p, _i, _i, _i, _p, _p, _p]),
    "y3_nms": (_i, [_p, _p, _p, _i, _i, _i, _f, _f, _p, _p, _p, _p]),
    "y3_gather_detections": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _p, _p, _p, _p]),
    "y3_gather_detections_packed": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _p, _p, _p, _p, _p]),
    "y3_preprocess": (_i, [_p, _p, _i, _i, _i, _i, _p, _p]),
    "y3_evaluate": (_i, [_p, _p, _p, _p, _i, _p, _p, _p, _i, _i, _i, _f, _p, _p]),
    "y3_conv_block_n": (_i, [_i, _i]),
    "y3_conv2d_bf16": (_i, [_p, _p, _i, _i, _i, _i, _i64, _p, _p, _i, _i, _i, _i, _p, _i64, _p, _i64, _i, _i, _p]),
    "y3_conv2d_flat_bf16": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _i, _i, _p, _i64, _p, _i64, _p]),
    "y3_conv2d_stem_f32": (_i, [_p, _p, _i, _i, _i, _p, _p, _i, _i, _p, _i64, _p]),
    "y3_dbg_umma_shift": (_i, [_p, _p, _i, _p, _i, _i, _i, _p, _p]),
    "y3_dbg_tma_tile": (_i, [_p, _p, _i, _i, _i, _i, _i64, _i, _i, _i, _i, _i, _i, _i, _p, _p]),
    "y3_dbg_timestamps": (_i, [_p]),
    "y3_dbg_timestamps_net": (_i, [_p]),
    "y3_dbg_set_chain_runs": (_i, [_i]),
    "y3_watchdog_code": (_i, [_p]),
}

_lib = None


def lib():
    """Load (once) and return the ctypes handle; raises if the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Y3Error(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          f"(there is no CPU fallback)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc):
    """Map a status code to the exception type the reference raises for the same mistake."""
    if rc == Y3_OK:
        return
    msg = lib().y3_last_error().decode()
    if rc == Y3_ERR_INVALID:
        raise ValueError(msg)
    if rc == Y3_ERR_UNSUPPORTED:
        raise Y3Unsupported(msg)
    raise Y3Error(msg)


class Context:
    """One per GPU per process.  device < 0 gives a planning-only context (host logic tests, no CUDA)."""

    def __init__(self, device=0):
        h = _p()
        check(lib().y3_ctx_create(int(device), C.byref(h)))
        self.handle = h
        self.device = int(device)

    def close(self):
        if getattr(self, "handle", None):
            lib().y3_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def sm_count(self):
        return lib().y3_ctx_sm_count(self.handle)

    def watchdog_code(self):
        return lib().y3_watchdog_code(self.handle)


_contexts = {}


def context(device=None):
    """Process-wide context for a CUDA device (default: torch's current device)."""
    if device is None:
        import torch
        if not torch.cuda.is_available():
            raise Y3Error("no CUDA device: the YOLOv3 path runs only on a B200 (there is no CPU fallback)")
        device = torch.cuda.current_device()
    device = int(device)
    if device not in _contexts:
        _contexts[device] = Context(device)
    return _contexts[device]


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else _p(t.data_ptr())


def stream_ptr():
    import torch
    return _p(torch.cuda.current_stream().cuda_stream)
