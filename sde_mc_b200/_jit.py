"""JIT build of the engine for user-defined SDEs (SDEMC_FAMILY_USER).

The reference lets users subclass `Sde` and write drift / diffusion / jumps as Python tensor functions
(/root/reference/sde_mc/sde.py:63-152); its solvers call them every step.  To run such a model on the fused kernels
the subclass also states the coefficients as CUDA expressions (`kernel_code()`); this module substitutes them into
csrc/user_model.cu.in, compiles that translation unit with nvcc for sm_100a against the engine's own headers, caches
the shared library by content hash and binds it with ctypes.  No nvcc -> SdemcError (there is no fallback path).
"""
import ctypes as C
import hashlib
import os
import shutil
import subprocess

from . import _lib as L

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")


def _default_cache():
    """in the package when it is writable (a source checkout), else the user's cache directory (an installed package)"""
    local = os.path.join(_HERE, "_jit")
    if os.access(local if os.path.isdir(local) else _HERE, os.W_OK):
        return local
    return os.path.join(os.environ.get("XDG_CACHE_HOME", os.path.join(os.path.expanduser("~"), ".cache")),
                        "sde_mc_b200", "jit")


_CACHE = os.environ.get("SDEMC_B200_JIT_DIR") or _default_cache()
_loaded = {}


def _nvcc():
    cand = os.environ.get("SDEMC_NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise L.SdemcError("user-defined SDEs are JIT-compiled with nvcc, which was not found (set SDEMC_NVCC); "
                           "sde_mc_b200 has no CPU fallback")
    return cand


def _cases(exprs, dim, what):
    if exprs is None:
        return ""
    if isinstance(exprs, str):
        exprs = [exprs] * dim
    if len(exprs) != dim:
        raise ValueError("%s needs one expression per component (%d), got %d" % (what, dim, len(exprs)))
    return "\n".join("    case %d: return (float)(%s);" % (i, e) for i, e in enumerate(exprs))


def _headers_digest():
    h = hashlib.sha1()
    for name in sorted(os.listdir(_CSRC)):
        if name.endswith((".cuh", ".in")):
            with open(os.path.join(_CSRC, name), "rb") as fh:
                h.update(fh.read())
    with open(os.path.join(_HERE, "..", "include", "sdemc_b200.h"), "rb") as fh:
        h.update(fh.read())
    return h.hexdigest()


def source_for(dim, marks, code):
    with open(os.path.join(_CSRC, "user_model.cu.in")) as fh:
        src = fh.read()
    return (src.replace("@DRIFT_CASES@", _cases(code["drift"], dim, "drift"))
               .replace("@DIFFUSION_CASES@", _cases(code["diffusion"], dim, "diffusion"))
               .replace("@JUMP_CASES@", _cases(code.get("jump"), dim, "jump"))
               .replace("@DIM@", str(int(dim))).replace("@MARKS@", str(int(marks))))


def build(dim, marks, code, verbose=False):
    """Compile (or find in the cache) the library for this model shape and code; returns its path."""
    src = source_for(dim, marks, code)
    key = hashlib.sha1((src + _headers_digest()).encode()).hexdigest()[:20]
    os.makedirs(_CACHE, exist_ok=True)
    so = os.path.join(_CACHE, "user_%s.so" % key)
    if os.path.exists(so):
        return so
    # under torchrun every rank builds at once: each writes and compiles its own copy of the source and only the
    # finished files are renamed into place
    cu = os.path.join(_CACHE, "user_%s.%d.cu" % (key, os.getpid()))
    with open(cu, "w") as fh:
        fh.write(src)
    tmp = so + ".tmp.%d" % os.getpid()
    cmd = [_nvcc(), "-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
           "--expt-relaxed-constexpr", "-shared", "-cudart", "static", "-I", _CSRC, cu, "-o", tmp]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        os.unlink(cu)
        raise L.SdemcError("nvcc failed on the user-defined SDE (check the CUDA expressions in kernel_code()):\n" +
                           res.stderr[-4000:])
    if verbose:
        print(res.stderr)
    os.replace(tmp, so)
    os.replace(cu, os.path.join(_CACHE, "user_%s.cu" % key))
    return so


class UserLibrary:
    """ctypes view of one JIT-built library: the two entry points user models support."""

    def __init__(self, path):
        lib = C.CDLL(path)
        if lib.sdemc_user_version() != L.load_abi_version():
            raise L.SdemcError("stale JIT library %s (ABI mismatch); delete the cache directory %s" % (path, _CACHE))
        lib.sdemc_user_last_cuda_error.restype = C.c_char_p
        lib.sdemc_user_mc_moments.restype = C.c_int
        lib.sdemc_user_mc_moments.argtypes = [C.POINTER(L.SdemcSde), C.POINTER(L.SdemcPayoff), C.POINTER(L.SdemcRange),
                                              C.POINTER(L.SdemcPathsOut), C.c_void_p, C.c_void_p, C.c_void_p]
        lib.sdemc_user_solve_paths.restype = C.c_int
        lib.sdemc_user_solve_paths.argtypes = [C.POINTER(L.SdemcSde), C.POINTER(L.SdemcPayoff), C.POINTER(L.SdemcRange),
                                               C.POINTER(L.SdemcPathsOut), C.c_void_p, C.c_void_p]
        self.lib, self.path = lib, path

    def check(self, rc):
        if rc != 0:
            msg = L.load().sdemc_strerror(rc).decode()
            if rc == -3:
                msg += " [" + self.lib.sdemc_user_last_cuda_error().decode() + "]"
            raise L.SdemcError("sdemc error %d in the user-model library: %s" % (rc, msg))

    def sdemc_mc_moments(self, sde, po, rng, per_path, mom, ws, stream):
        return self.lib.sdemc_user_mc_moments(sde, po, rng, per_path, mom, ws, stream)

    def sdemc_solve_paths(self, sde, po, rng, inj, out, ws, stream):
        if inj is not None:
            raise L.SdemcError("injected noise is not available for user-defined SDEs")
        return self.lib.sdemc_user_solve_paths(sde, po, rng, out, ws, stream)


def library_for(spec):
    """UserLibrary for a KernelSpec of the USER family (built on first use, cached in-process and on disk)."""
    code = spec.user_code
    key = (spec.dim, spec.marks, repr(code))
    lib = _loaded.get(key)
    if lib is None:
        lib = UserLibrary(build(spec.dim, spec.marks, code))
        _loaded[key] = lib
    return lib
