"""ctypes binding of libsdemc_b200.so (include/sdemc_b200.h).

This is the only place where Python meets the CUDA engine.  There is no CPU fallback: if the shared library is
missing or no sm_100 GPU is present, the solvers raise.  The structures mirror the C header field by field.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SDEMC_B200_LIB", os.path.join(_HERE, "libsdemc_b200.so"))

MAX_DIM = 4

# enums (sdemc_b200.h)
FAMILY_GEOMETRIC, FAMILY_ARITHMETIC, FAMILY_HESTON, FAMILY_USER = 0, 1, 2, 3
SCHEME_EULER, SCHEME_HESTON, SCHEME_MILSTEIN = 0, 1, 2
MARKS_NONE, MARKS_LOGNORMAL, MARKS_ICDF = 0, 1, 2
JUMPS_AUTO, JUMPS_QUEUE, JUMPS_INLINE = 0, 1, 2
INDEX_TERMINAL, INDEX_ADAPTED = 0, 1
(PAYOFF_EURO_CALL, PAYOFF_EURO_PUT, PAYOFF_BINARY_AON, PAYOFF_BASKET_ARITH, PAYOFF_BASKET_GEOM, PAYOFF_RAINBOW,
 PAYOFF_DIGITAL, PAYOFF_ASIAN_CALL, PAYOFF_HESTON_RAINBOW, PAYOFF_BEST_OF) = range(10)


DRAWS_BROWNIAN, DRAWS_QUEUE, DRAWS_INLINE, DRAWS_PACKED = range(4)
SHORT_AUTO, SHORT_OFF, SHORT_ALIGNED, SHORT_PACKED, SHORT_PACKED_GENERIC = range(5)
OUT_NO_TMA = 1
RANGE_COUNT_ON_HOST = 1


class _Sized(C.Structure):
    """Every struct handed to the library starts with struct_size = sizeof(struct) (the ABI's layout handshake);
    the constructor fills it, positional arguments start at the second field."""

    def __init__(self, *args, **kw):
        super().__init__(C.sizeof(type(self)), *args, **kw)


class SdemcSde(_Sized):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("family", C.c_int32), ("scheme", C.c_int32), ("dim", C.c_int32), ("m", C.c_int32), ("marks", C.c_int32),
        ("num_steps", C.c_int32), ("max_jumps", C.c_int32), ("exact_jumps", C.c_int32), ("asian", C.c_int32),
        ("jump_strategy", C.c_int32), ("queue_depth", C.c_int32), ("short_path", C.c_int32),
        ("T", C.c_float), ("x0", C.c_float * MAX_DIM), ("chol", C.c_float * (MAX_DIM * MAX_DIM)),
        ("a", C.c_float * MAX_DIM), ("b1", C.c_float * MAX_DIM), ("b2", C.c_float * MAX_DIM),
        ("c", C.c_float * MAX_DIM), ("rate", C.c_float), ("mark_p", C.c_float * 12), ("heston", C.c_float * 4),
        ("user_p", C.c_float * 16),
    ]


class SdemcPayoff(_Sized):
    _fields_ = [("struct_size", C.c_uint32), ("kind", C.c_int32), ("log", C.c_int32), ("index_mode", C.c_int32),
                ("strike", C.c_float), ("transform_discount", C.c_float), ("aux", C.c_float), ("df", C.c_float)]


class SdemcRange(_Sized):
    """SdemcRange(seed, path_lo, n_paths[, d_range]); d_range: device pointer to a (path_lo, n_paths) pair of uint64
    written by sdemc_plan_mc / sdemc_plan_mlmc -- the kernels then read their range when they run"""
    _fields_ = [("struct_size", C.c_uint32), ("flags", C.c_uint32), ("seed", C.c_uint64), ("path_lo", C.c_uint64),
                ("n_paths", C.c_uint64), ("d_range", C.c_void_p)]

    def __init__(self, seed=0, path_lo=0, n_paths=0, d_range=None, flags=0):
        super().__init__(flags, seed, path_lo, n_paths, d_range)


class SdemcInject(_Sized):
    """SdemcInject(d_z, d_zc, d_jump_times, d_marks, K, total_steps=0)"""
    _fields_ = [("struct_size", C.c_uint32), ("K", C.c_int32), ("d_z", C.c_void_p), ("d_zc", C.c_void_p),
                ("d_jump_times", C.c_void_p), ("d_marks", C.c_void_p), ("total_steps", C.c_int32),
                ("reserved", C.c_int32)]

    def __init__(self, d_z=None, d_zc=None, d_jump_times=None, d_marks=None, K=0, total_steps=0):
        super().__init__(K, d_z, d_zc, d_jump_times, d_marks, total_steps, 0)


class SdemcPathsOut(_Sized):
    _fields_ = [("struct_size", C.c_uint32), ("flags", C.c_uint32),
                ("d_paths", C.c_void_p), ("d_left", C.c_void_p), ("d_times", C.c_void_p), ("d_jumps", C.c_void_p),
                ("d_normals", C.c_void_p), ("d_payoffs", C.c_void_p), ("d_iters", C.c_void_p),
                ("d_terminal", C.c_void_p), ("d_total_steps", C.c_void_p), ("pitch_state", C.c_int64),
                ("pitch_times", C.c_int64), ("pitch_normals", C.c_int64)]

    def __init__(self, d_paths=None, d_left=None, d_times=None, d_jumps=None, d_normals=None, d_payoffs=None,
                 d_iters=None, d_total_steps=None, pitch_state=0, pitch_times=0, pitch_normals=0, d_terminal=None,
                 flags=0):
        super().__init__(flags, d_paths, d_left, d_times, d_jumps, d_normals, d_payoffs, d_iters, d_terminal,
                         d_total_steps, pitch_state, pitch_times, pitch_normals)


class SdemcMlp(_Sized):
    _fields_ = [("struct_size", C.c_uint32), ("cv_steps", C.c_uint32), ("d_w", C.c_void_p * 4),
                ("d_b", C.c_void_p * 4), ("in_dim", C.c_int32), ("hidden", C.c_int32), ("out_dim", C.c_int32),
                ("n_hidden_layers", C.c_int32)]


class SdemcCoeffsF64(_Sized):
    _fields_ = [("struct_size", C.c_uint32), ("reserved", C.c_uint32), ("T", C.c_double), ("x0", C.c_double * MAX_DIM),
                ("chol", C.c_double * (MAX_DIM * MAX_DIM)), ("a", C.c_double * MAX_DIM), ("b1", C.c_double * MAX_DIM),
                ("b2", C.c_double * MAX_DIM), ("c", C.c_double * MAX_DIM), ("rate", C.c_double),
                ("mark_p", C.c_double * 12), ("strike", C.c_double), ("transform_discount", C.c_double),
                ("aux", C.c_double), ("df", C.c_double)]


class SdemcInjectF64(_Sized):
    _fields_ = [("struct_size", C.c_uint32), ("K", C.c_int32), ("d_z", C.c_void_p), ("d_zc", C.c_void_p),
                ("d_jump_times", C.c_void_p), ("d_marks", C.c_void_p)]


MOMENT_FIELDS = ("sum", "sumsq", "sum_c", "sumsq_c", "sum_pc", "n", "iters", "reserved")
NUM_MOMENTS = len(MOMENT_FIELDS)

_SIGNATURES = {
    "sdemc_version": (C.c_int, []),
    "sdemc_strerror": (C.c_char_p, [C.c_int]),
    "sdemc_last_cuda_error": (C.c_char_p, []),
    "sdemc_device_info": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_uint64)]),
    "sdemc_workspace_bytes": (C.c_uint64, []),
    "sdemc_abi_layout": (C.c_int, [C.POINTER(C.c_uint32), C.c_int]),
    "sdemc_mc_moments": (C.c_int, [C.POINTER(SdemcSde), C.POINTER(SdemcPayoff), C.POINTER(SdemcRange),
                                   C.POINTER(SdemcPathsOut), C.c_void_p, C.c_void_p, C.c_void_p]),
    "sdemc_plan_mc": (C.c_int, [C.c_void_p, C.c_uint64, C.c_double, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int32, C.c_int32,
                                C.c_void_p, C.c_void_p, C.c_void_p]),
    "sdemc_plan_mlmc": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_uint64, C.c_double, C.c_double, C.c_uint64,
                                  C.c_uint64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sdemc_debug_draws": (C.c_int, [C.POINTER(SdemcSde), C.POINTER(SdemcRange), C.c_int32, C.c_int32, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    "sdemc_eval_payoff": (C.c_int, [C.POINTER(SdemcPayoff), C.c_int32, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]),
    "sdemc_solve_paths": (C.c_int, [C.POINTER(SdemcSde), C.POINTER(SdemcPayoff), C.POINTER(SdemcRange),
                                    C.POINTER(SdemcInject), C.POINTER(SdemcPathsOut), C.c_void_p, C.c_void_p]),
    "sdemc_mlmc_pair": (C.c_int, [C.POINTER(SdemcSde), C.POINTER(SdemcPayoff), C.c_int32, C.c_int32,
                                  C.POINTER(SdemcRange), C.POINTER(SdemcInject), C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p]),
    "sdemc_mlmc_pair_f64": (C.c_int, [C.POINTER(SdemcSde), C.POINTER(SdemcCoeffsF64), C.POINTER(SdemcPayoff), C.c_int32,
                                      C.c_int32, C.POINTER(SdemcRange), C.POINTER(SdemcInjectF64), C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "sdemc_mc_cv": (C.c_int, [C.POINTER(SdemcSde), C.POINTER(SdemcPayoff), C.c_float, C.c_float, C.POINTER(SdemcMlp),
                              C.POINTER(SdemcMlp), C.POINTER(SdemcRange), C.POINTER(SdemcInject), C.c_void_p,
                              C.c_void_p, C.c_void_p, C.c_void_p]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


class SdemcError(RuntimeError):
    pass


def load():
    """Load the engine (once).  Raises if the library has not been built -- there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SdemcError("libsdemc_b200.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; "
                             "g.build()'` or `make -C sde_mc_b200/csrc`; sde_mc_b200 has no CPU fallback" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.sdemc_version() != ABI_VERSION:
            raise SdemcError("%s speaks ABI version %d, this binding %d -- rebuild it (make -C sde_mc_b200/csrc)"
                             % (LIB_PATH, lib.sdemc_version(), ABI_VERSION))
        theirs = (C.c_uint32 * 9)()
        lib.sdemc_abi_layout(theirs, 9)
        if list(theirs) != struct_sizes():
            raise SdemcError("struct layouts of %s %s differ from this binding's %s" % (LIB_PATH, list(theirs),
                                                                                      struct_sizes()))
        _lib = lib
    return _lib


ABI_VERSION = 4


def struct_sizes():
    """sizes in the order of sdemc_abi_layout()"""
    return [C.sizeof(SdemcSde), C.sizeof(SdemcPayoff), C.sizeof(SdemcRange), C.sizeof(SdemcInject), 8 * NUM_MOMENTS,
            C.sizeof(SdemcPathsOut), C.sizeof(SdemcMlp), C.sizeof(SdemcCoeffsF64), C.sizeof(SdemcInjectF64)]


def load_abi_version():
    return ABI_VERSION


def check(rc):
    if rc != 0:
        lib = load()
        msg = lib.sdemc_strerror(rc).decode()
        if rc == -3:
            msg += " [" + lib.sdemc_last_cuda_error().decode() + "]"
        rearm_workspaces()
        raise SdemcError("sdemc error %d: %s" % (rc, msg))


def rearm_workspaces():
    """A launch that failed mid-grid leaves the reduction ticket of its workspace part-counted; zero every workspace
    so that later reductions on those streams start clean."""
    for ws in _workspaces.values():
        try:
            ws.zero_()
        except RuntimeError:      # the CUDA context itself is gone: nothing left to protect
            pass


def require_cuda(device):
    dev = torch.device(device)
    if dev.type != "cuda":
        raise SdemcError("sde_mc_b200 runs its solvers on B200 GPUs only (got device=%r); there is no CPU path. "
                         "Pass device='cuda'." % (device,))
    if not torch.cuda.is_available():
        raise SdemcError("no CUDA device available; sde_mc_b200 has no CPU fallback")
    return dev


_workspaces = {}


def workspace(device):
    """Zero-initialised scratch buffer per (device, stream) for the moment reductions."""
    dev = torch.device(device)
    key = (dev.index if dev.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(dev).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None:
        ws = torch.zeros(int(load().sdemc_workspace_bytes()), dtype=torch.uint8, device=dev)
        _workspaces[key] = ws
    return ws


def stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(torch.device(device)).cuda_stream)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
