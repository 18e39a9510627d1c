"""TEST INFRASTRUCTURE -- numpy restatement of how the kernels turn Philox4x32-10 words into noise.

The product draws its noise in registers (sde_mc_b200/csrc/philox.cuh); nothing here is imported by it.  This module
restates, on the CPU and in float64 libm arithmetic, the counter layout and the bits -> (normal, gap, mark) maps of
that header, so that the Philox-driven MOMENTS kernels (diffusion.cuh FAST1D, jump1d.cuh, jump_flat.cuh) can be
compared per path with the oracle: derive the noise arrays here, feed them to oracle.diffusion / oracle.jump (the
restatement of the reference loops solvers.py:68-88,164-226), compare with the kernels' per-path outputs.

Pinned by tests/test_oracle_golden.py: `philox4x32_10` against the Random123 known-answer vectors (through the C
oracle's scalar implementation), the stream maps against stored outputs of the path-storing kernel on the GPU.

Counter (128 bit): (block index, stream id, path id lo, path id hi); key (64 bit): the seed (philox.cuh:8-11).
"""
import numpy as np

STREAM_DIFFUSION, STREAM_JUMP_QUEUE, STREAM_JUMP_INLINE, STREAM_PACKED = 0, 1, 2, 3   # philox.cuh:22-27
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, seed):
    """vectorised Philox4x32-10 (philox.cuh:42-56): uint32 arrays in, four uint32 arrays out"""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) & _MASK for c in np.broadcast_arrays(c0, c1, c2, c3)]
    k0, k1 = int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = c0 * _M0, c2 * _M1
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)
        c1, c3, c0, c2 = p1 & _MASK, p0 & _MASK, n0, n2
        k0, k1 = (k0 + _W0) & 0xFFFFFFFF, (k1 + _W1) & 0xFFFFFFFF
    return [c.astype(np.uint32) for c in (c0, c1, c2, c3)]


def blocks(seed, stream, path_ids, nblocks):
    """Philox output words of blocks 0..nblocks-1 of `stream` for every path: list of four (n, nblocks) uint32"""
    pid = np.asarray(path_ids, dtype=np.uint64)[:, None]
    b = np.arange(nblocks, dtype=np.uint64)[None, :]
    return philox4x32_10(b, np.uint64(stream), pid & _MASK, pid >> np.uint64(32), seed)


def u01_open0(w):
    """bits_to_u01_open0 (philox.cuh:66): 23 mantissa bits -> (0, 1] with 2^-23 spacing"""
    return 1.0 - (np.asarray(w, np.uint32) & np.uint32(0x7FFFFF)).astype(np.float64) * 2.0 ** -23


def u01(w):
    """bits_to_u01 (philox.cuh:64): [0, 1)"""
    return (np.asarray(w, np.uint32) & np.uint32(0x7FFFFF)).astype(np.float64) * 2.0 ** -23


def exp1(w):
    """exp1_from_bits (philox.cuh:150): -ln(u)"""
    return -np.log(u01_open0(w))


def box_muller_full(wa, wb):
    """box_muller (philox.cuh:82-97): 23-bit radius uniform from wa, 23-bit angle from wb"""
    r = np.sqrt(-2.0 * np.log(u01_open0(wa)))
    ang = 2.0 * np.pi * u01(wb)
    return r * np.cos(ang), r * np.sin(ang)


def polar3(o):
    """philox_polar3 (philox.cuh:109-122): three (radius, cos, sin) triples from one block -- 23-bit radius uniforms
    from the low bits of o0, o1, o2; 16-bit angles from o3 low half, o3 high half, top bytes of (o0, o1)"""
    o0, o1, o2, o3 = [np.asarray(w, np.uint32) for w in o]
    k = [o3 & np.uint32(0xFFFF), o3 >> np.uint32(16), (o0 >> np.uint32(24)) | ((o1 >> np.uint32(24)) << np.uint32(8))]
    r = [np.sqrt(-2.0 * np.log(u01_open0(w))) for w in (o0, o1, o2)]
    ang = [2.0 * np.pi * kk.astype(np.float64) / 65536.0 for kk in k]
    return r, [np.cos(a) for a in ang], [np.sin(a) for a in ang]


def normals6(o):
    """philox_normals6 (philox.cuh:139-147): six unit normals per block in slot order r0c0 r0s0 r1c1 r1s1 r2c2 r2s2;
    returns (..., 6)"""
    r, c, s = polar3(o)
    return np.stack([r[0] * c[0], r[0] * s[0], r[1] * c[1], r[1] * s[1], r[2] * c[2], r[2] * s[2]], axis=-1)


def brownian_normals(seed, path_ids, count):
    """the first `count` unit normals of STREAM_DIFFUSION per path, in consumption order: (n, count) float32"""
    nb = (count + 5) // 6
    z = normals6(blocks(seed, STREAM_DIFFUSION, path_ids, nb)).reshape(len(path_ids), nb * 6)
    return z[:, :count].astype(np.float32)


def queue_jumps(seed, path_ids, njumps, rate, marks):
    """the first `njumps` (cumulative jump time, raw mark draw) pairs of the sparse-jump queue (queue_refill,
    jump.cuh:66-93): global group g serves jumps 4g..4g+3 from blocks 2g (four gap words) and 2g+1 (marks: two full
    Box-Muller pairs for lognormal marks, four uniforms for icdf marks).  The sequence does not depend on the queue
    depth.  Times are accumulated in float32 like the kernel (fmaf(gap, 1/rate, tau))."""
    n = len(path_ids)
    ng = (njumps + 3) // 4
    o = blocks(seed, STREAM_JUMP_QUEUE, path_ids, 2 * ng)
    gaps = np.stack([w[:, 0::2] for w in o], axis=-1).reshape(n, ng * 4)          # block 2g, words 0..3
    mw = [w[:, 1::2] for w in o]                                                  # block 2g+1
    if marks == "lognormal":
        a0, a1 = box_muller_full(mw[0], mw[1])
        b0, b1 = box_muller_full(mw[2], mw[3])
        raw = np.stack([a0, a1, b0, b1], axis=-1).reshape(n, ng * 4)
    else:
        raw = np.stack([u01(w) for w in mw], axis=-1).reshape(n, ng * 4)
    inv_rate = np.float32(1.0) / np.float32(rate)
    tau = np.zeros((n,), np.float32)
    times = np.empty((n, ng * 4), np.float32)
    e = exp1(gaps).astype(np.float32)
    for j in range(ng * 4):
        tau = (e[:, j].astype(np.float64) * np.float64(inv_rate) + tau.astype(np.float64)).astype(np.float32)
        times[:, j] = tau
    return times[:, :njumps], raw[:, :njumps].astype(np.float32)


def packed_draws(seed, path_ids, niter):
    """STREAM_PACKED (jump_flat.cuh:95-106): block k//2 serves iterations k = 2b, 2b+1 of a 1-D lognormal-mark path.
    Returns per-iteration (z, gap, raw mark), each (n, niter): z = Brownian unit normal, gap ~ Exp(1) candidate,
    raw = N(0,1) mark candidate."""
    nb = (niter + 1) // 2
    o0, o1, o2, o3 = blocks(seed, STREAM_PACKED, path_ids, nb)
    ang_z = 2.0 * np.pi * (o3 & np.uint32(0xFFFF)).astype(np.float64) / 65536.0
    ang_m = 2.0 * np.pi * (o3 >> np.uint32(16)).astype(np.float64) / 65536.0
    r_z = np.sqrt(-2.0 * np.log(u01_open0(o0)))
    r_m = np.sqrt(-2.0 * np.log(u01_open0(o1)))
    odd = (o0 >> np.uint32(23)) | ((o1 >> np.uint32(23)) << np.uint32(9)) | ((o2 >> np.uint32(23)) << np.uint32(18))
    z = np.stack([r_z * np.cos(ang_z), r_z * np.sin(ang_z)], axis=-1).reshape(len(path_ids), nb * 2)
    raw = np.stack([r_m * np.cos(ang_m), r_m * np.sin(ang_m)], axis=-1).reshape(len(path_ids), nb * 2)
    gap = np.stack([exp1(o2), exp1(odd)], axis=-1).reshape(len(path_ids), nb * 2)
    return z[:, :niter].astype(np.float32), gap[:, :niter].astype(np.float32), raw[:, :niter].astype(np.float32)


def inline_draws(seed, path_ids, niter, marks):
    """STREAM_JUMP_INLINE (InlineJumps, jump.cuh:131-167): block k//2 serves iterations 2b, 2b+1.  Lognormal marks:
    gaps from words 0, 1 and one full Box-Muller pair from words (2, 3); icdf marks: (gap, uniform) from words
    (0, 1) and (2, 3).  Returns per-iteration (gap ~ Exp(1), raw mark draw), each (n, niter)."""
    nb = (niter + 1) // 2
    o0, o1, o2, o3 = blocks(seed, STREAM_JUMP_INLINE, path_ids, nb)
    if marks == "lognormal":
        g = np.stack([exp1(o0), exp1(o1)], axis=-1)
        a, b = box_muller_full(o2, o3)
        raw = np.stack([a, b], axis=-1)
    else:
        g = np.stack([exp1(o0), exp1(o2)], axis=-1)
        raw = np.stack([u01(o1), u01(o3)], axis=-1)
    n = len(path_ids)
    return g.reshape(n, nb * 2)[:, :niter].astype(np.float32), raw.reshape(n, nb * 2)[:, :niter].astype(np.float32)


def candidate_jumps(h0, T, rate, gap, raw, max_jumps):
    """Map the candidate-per-iteration jump strategies (INLINE / PACKED: a fresh (gap, mark) candidate every iteration,
    taken when the previous jump has been consumed) onto the reference's inputs (jump_times (n, max_jumps) cumulative,
    marks (n, K) raw draw read at the HIT iteration; solvers.py:143-148,212-217).  Only the clock of the jump-adapted
    loop is restated, in float32 like the kernels (solvers.py:190-193,212: dt = max(min(h0, min(tau, T) - t), 0),
    hit <=> |tau - t| <= 1e-12 + 1e-5 |t|); the states are left to oracle.jump.
    Returns (jump_times, marks, iters)."""
    gap, raw = np.asarray(gap, np.float32), np.asarray(raw, np.float32)
    n, K = gap.shape
    f32 = np.float32
    h0, T = f32(h0), f32(T)
    inv_rate = np.float64(f32(1.0) / f32(rate))
    t = np.zeros(n, f32)
    tau = np.zeros(n, f32)
    cur_raw = np.zeros(n, f32)
    need_pop = np.ones(n, bool)
    njump = np.zeros(n, np.int64)
    iters = np.zeros(n, np.int32)
    jump_times = np.full((n, max_jumps), np.inf, f32)
    marks = np.zeros((n, K), f32)
    rows = np.arange(n)
    for k in range(K):
        active = t < T
        pop = need_pop & active
        new_tau = (gap[:, k].astype(np.float64) * inv_rate + tau.astype(np.float64)).astype(f32)
        tau = np.where(pop, new_tau, tau)
        cur_raw = np.where(pop, raw[:, k], cur_raw)
        ok = pop & (njump < max_jumps)
        jump_times[rows[ok], njump[ok]] = tau[ok]
        njump = njump + pop
        dt = np.maximum(np.minimum(h0, np.minimum(tau, T) - t), f32(0)).astype(f32)
        t = np.where(active, (t + dt).astype(f32), t)
        hit = active & (np.abs(tau - t) <= (np.abs(t) * f32(1e-5) + f32(1e-12)).astype(f32))
        marks[:, k] = np.where(hit, cur_raw, f32(0))
        need_pop = np.where(active, hit, need_pop)
        iters += active
    return jump_times, marks, iters
