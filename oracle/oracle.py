"""TEST INFRASTRUCTURE -- ctypes wrapper of oracle/libsde_oracle.so, the CPU restatement of the reference hot path.

Only tests/, __graft_entry__.smoke() and the cpu_baseline / `--impl reference` legs of bench.py import this module.
The product package (sde_mc_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libsde_oracle.so")


class OracleSde(C.Structure):
    _fields_ = [("family", C.c_int32), ("scheme", C.c_int32), ("dim", C.c_int32), ("m", C.c_int32),
                ("marks", C.c_int32), ("num_steps", C.c_int32), ("max_jumps", C.c_int32), ("exact_jumps", C.c_int32),
                ("asian", C.c_int32), ("pad_", C.c_int32), ("T", C.c_double), ("x0", C.c_double * 4),
                ("chol", C.c_double * 16), ("a", C.c_double * 4), ("b1", C.c_double * 4), ("b2", C.c_double * 4),
                ("c", C.c_double * 4), ("rate", C.c_double), ("mark_p", C.c_double * 12), ("heston", C.c_double * 4)]


class OraclePayoff(C.Structure):
    _fields_ = [("kind", C.c_int32), ("log", C.c_int32), ("index_mode", C.c_int32), ("pad_", C.c_int32),
                ("strike", C.c_double), ("transform_discount", C.c_double), ("aux", C.c_double), ("df", C.c_double)]


class OracleMlp(C.Structure):
    _fields_ = [("w", C.c_void_p * 4), ("b", C.c_void_p * 4), ("in_dim", C.c_int32), ("hidden", C.c_int32),
                ("out_dim", C.c_int32), ("n_hidden_layers", C.c_int32)]


def build():
    """Compile the oracle (gcc, a second or two).  Called by __graft_entry__.build() and lazily by load()."""
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)


_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        _lib = C.CDLL(_LIB)
    return _lib


def sde_struct(spec, T, num_steps, max_jumps=0, exact_jumps=False):
    """spec: any object with the KernelSpec fields (family, dim, m, marks, asian, x0, chol, a, b1, b2, c, rate,
    mark_p, heston) holding python floats."""
    s = OracleSde()
    s.family, s.scheme, s.dim, s.m, s.marks = spec.family, getattr(spec, "scheme", 0), spec.dim, spec.m, spec.marks
    s.num_steps, s.max_jumps, s.exact_jumps, s.asian = int(num_steps), int(max_jumps), int(bool(exact_jumps)), spec.asian
    s.T = float(T)
    for i in range(4):
        s.x0[i], s.a[i], s.b1[i], s.b2[i], s.c[i] = spec.x0[i], spec.a[i], spec.b1[i], spec.b2[i], spec.c[i]
        s.heston[i] = spec.heston[i]
    for i in range(16):
        s.chol[i] = spec.chol[i]
    s.rate = spec.rate
    for i in range(12):
        s.mark_p[i] = spec.mark_p[i]
    return s


def payoff_struct(kind, strike, log=False, transform_discount=1.0, aux=1.0, df=1.0, index_mode=1):
    p = OraclePayoff()
    p.kind, p.log, p.index_mode = int(kind), int(bool(log)), int(index_mode)
    p.strike, p.transform_discount, p.aux, p.df = float(strike), float(transform_discount), float(aux), float(df)
    return p


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else C.c_void_p(0)


def _c(a, dt):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


def diffusion(sde, z, dtype=np.float32):
    """returns (paths (n,S+1,d), normals (n,S,d[,m]))"""
    lib = load()
    z = _c(z, dtype)
    n, S, d, m = z.shape
    paths = np.empty((n, S + 1, d), dtype)
    normals = np.empty((n, S, d, m), dtype)
    fn = lib.oracle_diffusion_f32 if dtype == np.float32 else lib.oracle_diffusion_f64
    fn(C.byref(sde), C.c_int64(n), _p(z), _p(paths), _p(normals))
    return paths, (normals[..., 0] if m == 1 else normals)


def diffusion_pair(sde, fine, coarse, z):
    lib = load()
    z = _c(z, np.float32)
    n, d = z.shape[0], z.shape[2]
    pf = np.empty((n, fine + 1, d), np.float32)
    pc = np.empty((n, coarse + 1, d), np.float32)
    lib.oracle_diffusion_pair_f32(C.byref(sde), C.c_int64(n), C.c_int(fine), C.c_int(coarse), _p(z), _p(pf), _p(pc))
    return pf, pc


def jump(sde, z, zc, jump_times, marks, dtype=np.float32, full=True):
    """returns dict(paths, left, times, jumps, normals, iters, total_steps); arrays span all K slots"""
    lib = load()
    z, zc, jt, mk = _c(z, dtype), _c(zc, dtype), _c(jump_times, dtype), _c(marks, dtype)
    n, K, d = z.shape
    m = sde.m
    paths = np.empty((n, K + 1, d), dtype)
    left = np.empty((n, K + 1, d), dtype) if full else None
    times = np.empty((n, K + 1), dtype) if full else None
    jumps = np.empty((n, K + 1, d), dtype) if full else None
    normals = np.empty((n, K, d, m), dtype) if full else None
    iters = np.empty((n,), np.int32)
    fn = lib.oracle_jump_f32 if dtype == np.float32 else lib.oracle_jump_f64
    fn.restype = C.c_int
    total = fn(C.byref(sde), C.c_int64(n), C.c_int(K), _p(z), _p(zc), _p(jt), _p(mk), _p(paths), _p(left), _p(times),
               _p(jumps), _p(normals), _p(iters))
    if normals is not None and m == 1:
        normals = normals[..., 0]
    return dict(paths=paths, left=left, times=times, jumps=jumps, normals=normals, iters=iters, total_steps=total)


def jump_pair(sde, fine, coarse, z, zc, jump_times, marks, dtype=np.float64):
    lib = load()
    z, zc, jt, mk = _c(z, dtype), _c(zc, dtype), _c(jump_times, dtype), _c(marks, dtype)
    n, d = z.shape[0], z.shape[2]
    K = mk.shape[1]
    fl = np.empty((n, d), dtype)
    cl = np.empty((n, d), dtype)
    iters = np.empty((n,), np.int32)
    fn = lib.oracle_jump_pair_f32 if dtype == np.float32 else lib.oracle_jump_pair_f64
    fn.restype = C.c_int
    total = fn(C.byref(sde), C.c_int64(n), C.c_int(fine), C.c_int(coarse), C.c_int(K), _p(z), _p(zc), _p(jt), _p(mk),
               _p(fl), _p(cl), _p(iters))
    return fl, cl, iters, total


def payoff(po, x, dtype=np.float32):
    lib = load()
    x = _c(x, dtype)
    n, d = x.shape
    out = np.empty((n,), dtype)
    fn = lib.oracle_payoff_f32 if dtype == np.float32 else lib.oracle_payoff_f64
    fn(C.byref(po), C.c_int64(n), C.c_int(d), _p(x), _p(out))
    return out


def icdf(mark_p, u):
    lib = load()
    u = _c(u, np.float32)
    out = np.empty_like(u)
    mp = (C.c_double * 12)(*[float(v) for v in mark_p])
    lib.oracle_icdf_f32(mp, C.c_int64(u.size), _p(u), _p(out))
    return out


def mlp_struct(weights, biases):
    """weights/biases: 4 numpy arrays each, torch Linear layout (out, in)."""
    m = OracleMlp()
    keep = []
    for i in range(4):
        w, b = _c(weights[i], np.float32), _c(biases[i], np.float32)
        keep += [w, b]
        m.w[i] = w.ctypes.data
        m.b[i] = b.ctypes.data
    m.in_dim, m.hidden, m.out_dim, m.n_hidden_layers = weights[0].shape[1], weights[0].shape[0], weights[3].shape[0], 3
    m._keep = keep
    return m


def remove_steps(tol, steps, time_interval):
    """helpers.py:71-74: index of the last step kept when `tol` is cut off the end of the interval"""
    return int(np.floor(steps - tol / (time_interval / steps)))


def cv_gamma_jump(sde, res, payoffs, disc_rate, jump_mean, f, g, brownian_steps=0):
    """per-path gamma from the arrays returned by jump(..., full=True); brownian_steps: integrate_cv's tol cut"""
    lib = load()
    n, K1, d = res["paths"].shape
    normals = res["normals"].reshape(n, K1 - 1, -1)
    gamma = np.empty((n,), np.float32)
    lib.oracle_cv_gamma_jump_f32(C.byref(sde), C.c_int64(n), C.c_int(K1 - 1), C.c_int(int(res["total_steps"])),
                                 C.c_int(int(brownian_steps)), C.c_double(disc_rate), C.c_double(jump_mean), C.byref(f),
                                 C.byref(g) if g is not None else None, _p(res["paths"]), _p(res["left"]),
                                 _p(res["times"]), _p(res["jumps"]), _p(_c(normals, np.float32)),
                                 _p(_c(payoffs, np.float32)), _p(gamma))
    return gamma


def cv_gamma_diffusion(sde, paths, normals, payoffs, disc_rate, f, brownian_steps=0):
    lib = load()
    n = paths.shape[0]
    gamma = np.empty((n,), np.float32)
    lib.oracle_cv_gamma_diffusion_f32(C.byref(sde), C.c_int64(n), C.c_int(int(brownian_steps)), C.c_double(disc_rate), C.byref(f),
                                      _p(_c(paths, np.float32)), _p(_c(normals.reshape(n, paths.shape[1] - 1, -1), np.float32)),
                                      _p(_c(payoffs, np.float32)), _p(gamma))
    return gamma


def philox(c0, c1, c2, c3, k0, k1):
    lib = load()
    out = (C.c_uint32 * 4)()
    lib.oracle_philox4x32_10(C.c_uint32(c0), C.c_uint32(c1), C.c_uint32(c2), C.c_uint32(c3), C.c_uint32(k0),
                             C.c_uint32(k1), out)
    return list(out)
