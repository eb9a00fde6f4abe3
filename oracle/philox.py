"""Philox4x32-10 counter-based Gaussian stream — numpy mirror of the generator in
csrc/dmip_rng.cuh (TEST INFRASTRUCTURE ONLY).  The reference draws its noise from
torch's global RNG (models/diffusion.py:32,42), which cannot be reproduced on a
GPU; production sampling therefore uses this keyed stream, and parity tests feed
its output to the oracle sampler as injected noise.

counter = (gidx_lo, gidx_hi, step, (stream << 16) | quad)   key = (seed_lo, seed_hi)
  gidx   : global particle index (obs * n_per_obs + n) — results are independent of
           how particles are split over CTAs or GPUs
  step   : SDE step index; 0xFFFFFFFF for the initial draw x0
  stream : 0 = state noise, 1 = CDiffE observation re-diffusion noise,
           2 = Metropolis proposal noise, 3 = Metropolis acceptance uniform (first output word, quad 0)
  quad   : element index // 4; the four 32-bit outputs give elements 4q..4q+3
Box–Muller on (u1,u2) = ((r>>8)+0.5)/2^24:  n0 = R cos(th), n1 = R sin(th).
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
STEP_INIT = 0xFFFFFFFF


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(v, dtype=np.uint32) for v in np.broadcast_arrays(c0, c1, c2, c3))
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for r in range(10):
            p0 = c0.astype(np.uint64) * M0
            p1 = c2.astype(np.uint64) * M1
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = p0.astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = p1.astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32(k0 + W0)
            k1 = np.uint32(k1 + W1)
    return c0, c1, c2, c3


def _u01(r):
    return ((r >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)


def normals(gidx, step, stream, dim, seed):
    """(len(gidx), dim) float32 standard normals for the given particles/step/stream."""
    gidx = np.asarray(gidx, dtype=np.uint64)
    nq = (dim + 3) // 4
    q = np.arange(nq, dtype=np.uint32)[None, :]
    c0 = (gidx & np.uint64(0xFFFFFFFF)).astype(np.uint32)[:, None]
    c1 = (gidx >> np.uint64(32)).astype(np.uint32)[:, None]
    c2 = np.uint32(step)
    c3 = (np.uint32(stream) << np.uint32(16)) | q
    r0, r1, r2, r3 = philox4x32_10(c0, c1, c2, c3, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    out = np.empty((len(gidx), nq, 4), dtype=np.float32)
    for j, (ra, rb) in enumerate(((r0, r1), (r2, r3))):
        u1 = _u01(ra)
        u2 = _u01(rb)
        R = np.sqrt(np.maximum(np.float32(-2.0) * np.log(u1), np.float32(0.0))).astype(np.float32)   # clamp as dmip_rng.cuh
        th = np.float32(2.0 * np.pi) * u2
        out[:, :, 2 * j] = R * np.cos(th)
        out[:, :, 2 * j + 1] = R * np.sin(th)
    return out.reshape(len(gidx), nq * 4)[:, :dim]


def uniforms(gidx, step, stream, seed):
    """(len(gidx),) float32 uniforms in (0,1): the first output word of the (gidx, step, stream, quad 0) block."""
    gidx = np.asarray(gidx, dtype=np.uint64)
    c0 = (gidx & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    c1 = (gidx >> np.uint64(32)).astype(np.uint32)
    r0, _, _, _ = philox4x32_10(c0, c1, np.uint32(step), np.uint32(stream) << np.uint32(16),
                                seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return _u01(r0)
