"""CPU oracle of the multi-resolution hash encoding  --  TEST INFRASTRUCTURE ONLY.

numpy restatement of HashEncoding.pytorch_fwd / hash_fn (/root/reference/fields/encodings.py:306-366):
  scalings = floor(min_res * g^level) in fp32 (:270-272); corners ceil / floor of x * scaling as int32 (:329-331);
  hash = (x*1) ^ (y*2654435761) ^ (z*805459861) in int64 WITHOUT 32-bit wrap, floor-mod T, + level*T (:317-322);
  trilinear blend in the reference's operation order, weight `offset` toward the ceil corner (:354-364).
Only tests/ may import it.  PARITY PINNING: tests/golden/make_hash_golden.py runs the unmodified reference
(implementation="torch") on seeded inputs and commits inputs + outputs as tests/golden/hash_*.npz;
tests/test_hash_encoding.py checks this oracle bit-for-bit against that fixture.
"""
import numpy as np


def scalings(num_levels=16, min_res=16, max_res=1024):
    levels = np.arange(num_levels)
    growth = np.exp((np.log(max_res) - np.log(min_res)) / (num_levels - 1))
    return np.floor((min_res * growth ** levels).astype(np.float32)).astype(np.float32)


def hash_fn(ix, iy, iz, level, log2_T):
    T = np.int64(1) << np.int64(log2_T)
    h = (ix.astype(np.int64) * np.int64(1)) ^ (iy.astype(np.int64) * np.int64(2654435761)) ^ (iz.astype(np.int64) * np.int64(805459861))
    return np.mod(h, T) + level.astype(np.int64) * T


def hash_encode(pts, table, scal, log2_T):
    """pts [N,3] fp32, table [L*T, F] fp32, scal [L] fp32 -> [N, L*F] fp32 (bit-exact restatement)."""
    pts = np.asarray(pts, np.float32); table = np.asarray(table, np.float32); scal = np.asarray(scal, np.float32)
    N, L, F = pts.shape[0], scal.shape[0], table.shape[1]
    scaled = pts[:, None, :] * scal[None, :, None]                      # [N,L,3] fp32
    c = np.ceil(scaled).astype(np.int32); f = np.floor(scaled).astype(np.int32)
    off = (scaled - f.astype(np.float32)).astype(np.float32)
    lvl = np.broadcast_to(np.arange(L)[None, :], (N, L))
    cx, cy, cz, fx, fy, fz = c[..., 0], c[..., 1], c[..., 2], f[..., 0], f[..., 1], f[..., 2]
    corners = [(cx, cy, cz), (cx, fy, cz), (fx, fy, cz), (fx, cy, cz), (cx, cy, fz), (cx, fy, fz), (fx, fy, fz), (fx, cy, fz)]
    fe = [table[hash_fn(x, y, z, lvl, log2_T)] for (x, y, z) in corners]     # each [N,L,F]
    ox, oy, oz = off[..., 0:1], off[..., 1:2], off[..., 2:3]
    one = np.float32(1.0)
    f03 = fe[0] * ox + fe[3] * (one - ox)
    f12 = fe[1] * ox + fe[2] * (one - ox)
    f56 = fe[5] * ox + fe[6] * (one - ox)
    f47 = fe[4] * ox + fe[7] * (one - ox)
    f0312 = f03 * oy + f12 * (one - oy)
    f4756 = f47 * oy + f56 * (one - oy)
    enc = f0312 * oz + f4756 * (one - oz)
    return enc.reshape(N, L * F).astype(np.float32)


def hash_encode_table_grad(pts, d_out, scal, log2_T, n_rows, F):
    """d loss / d table for d_out [N, L*F] (float64 accumulation; the CUDA kernel uses fp32 atomics)."""
    pts = np.asarray(pts, np.float32); scal = np.asarray(scal, np.float32)
    N, L = pts.shape[0], scal.shape[0]
    scaled = pts[:, None, :] * scal[None, :, None]
    c = np.ceil(scaled).astype(np.int32); f = np.floor(scaled).astype(np.int32)
    off = (scaled - f.astype(np.float32)).astype(np.float64)
    lvl = np.broadcast_to(np.arange(L)[None, :], (N, L))
    cx, cy, cz, fx, fy, fz = c[..., 0], c[..., 1], c[..., 2], f[..., 0], f[..., 1], f[..., 2]
    ox, oy, oz = off[..., 0], off[..., 1], off[..., 2]
    mx, my, mz = 1 - ox, 1 - oy, 1 - oz
    corners = [(cx, cy, cz, ox * oy * oz), (cx, fy, cz, ox * my * oz), (fx, fy, cz, mx * my * oz), (fx, cy, cz, mx * oy * oz),
               (cx, cy, fz, ox * oy * mz), (cx, fy, fz, ox * my * mz), (fx, fy, fz, mx * my * mz), (fx, cy, fz, mx * oy * mz)]
    g = np.zeros((n_rows, F), np.float64)
    d = np.asarray(d_out, np.float64).reshape(N, L, F)
    for x, y, z, w in corners:
        np.add.at(g, hash_fn(x, y, z, lvl, log2_T).reshape(-1), (w[..., None] * d).reshape(-1, F))
    return g
