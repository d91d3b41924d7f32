"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the noise stream the CUDA path draws in-kernel.

The reference (`/root/reference/src/pno_physics_bench/models.py:67-76, 264-275`) draws its noise with
`torch.randn_like` from torch's global generator.  Matching torch's CUDA Philox offsets bit for bit is brittle
(hard part 3), so the B200 path owns a counter-based stream and the parity harness injects
that stream into the reference.  This file restates the stream on the CPU (numpy, vectorised) so that
golden vectors and `-m "not gpu"` tests do not need a GPU:

    key     = (seed & 0xffffffff, seed >> 32)
    counter = (q & 0xffffffff, q >> 32, sample_idx, site)          q = "quad" index, 4 normals per quad
    x0..x3  = Philox4x32-10(counter, key)                          (Salmon et al., SC'11; same rounds/constants as cuRAND)
    (z0,z1) = BoxMuller(x0, x1),  (z2,z3) = BoxMuller(x2, x3)
    BoxMuller(a, b): u1 = ((a >> 9) + 0.5) * 2^-23,  u2 = ((b >> 9) + 0.5) * 2^-23
                     r = sqrt(-2 ln u1), t = 2*pi*(u2 - 0.5);  returns (r cos t, r sin t)

Element -> (quad, lane) maps (also implemented by `pno_philox_materialize` in include/pno_b200.h):
  * activation-shaped sites (lift / linear / project noise, README weight-space local noise):
    logical contiguous element index e (with the GLOBAL batch index) -> q = e >> 2, lane = e & 3.
  * spectral-weight sites: for block blk in {0,1} (weights1, weights2), logical element (i, o, kx, ky):
    q = (((blk*m1 + kx)*m2 + ky)*Ci + i)*ceil(Co/2) + (o >> 1);  lanes (0,1) = (eps_re, eps_im) of even o,
    lanes (2,3) = (eps_re, eps_im) of odd o.  This follows the kernel-native [kx][ky][i][o] parameter layout so
    that one Philox call feeds two adjacent complex weights.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10. Inputs: uint32 arrays (broadcastable) / ints. Returns 4 uint32 arrays."""
    c0 = np.asarray(c0, dtype=np.uint64)
    c1 = np.asarray(c1, dtype=np.uint64)
    c2 = np.asarray(c2, dtype=np.uint64)
    c3 = np.asarray(c3, dtype=np.uint64)
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for r in range(10):
        p0 = _M0 * c0            # < 2^64, exact in uint64
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32))


def _box_muller(a, b):
    u1 = ((a >> np.uint32(9)).astype(np.float64) + 0.5) * 2.0 ** -23
    u2 = ((b >> np.uint32(9)).astype(np.float64) + 0.5) * 2.0 ** -23
    r = np.sqrt(-2.0 * np.log(u1))
    t = 2.0 * np.pi * (u2 - 0.5)
    return r * np.cos(t), r * np.sin(t)


def quad_bits(seed, sample_idx, site, q):
    """Raw Philox output for quad indices q (uint64 array) -> uint32 array [..., 4]."""
    q = np.asarray(q, dtype=np.uint64)
    x = philox4x32_10(q & _MASK, q >> np.uint64(32), np.uint64(sample_idx & 0xFFFFFFFF),
                      np.uint64(site & 0xFFFFFFFF), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return np.stack(x, axis=-1)


def quad_normals(seed, sample_idx, site, q):
    """Four standard normals per quad index -> float64 array [..., 4]."""
    x = quad_bits(seed, sample_idx, site, q)
    z0, z1 = _box_muller(x[..., 0], x[..., 1])
    z2, z3 = _box_muller(x[..., 2], x[..., 3])
    return np.stack([z0, z1, z2, z3], axis=-1)


def normal_activation(seed, sample_idx, site, shape, batch_offset=0, dtype=np.float32):
    """eps for an activation-shaped site. shape = (B, ...); batch_offset = first GLOBAL batch index held locally."""
    shape = tuple(int(s) for s in shape)
    n = int(np.prod(shape))
    per_item = n // shape[0] if shape[0] else 0
    e = np.arange(n, dtype=np.uint64) + np.uint64(batch_offset * per_item)
    qn = quad_normals(seed, sample_idx, site, e >> np.uint64(2))
    z = np.take_along_axis(qn, (e & np.uint64(3)).astype(np.int64)[:, None], axis=1)[:, 0]
    return z.reshape(shape).astype(dtype)


def normal_spectral_weights(seed, sample_idx, site, ci, co, m1, m2, dtype=np.float32):
    """(eps_re, eps_im), each [2, Ci, Co, m1, m2] in the reference's logical layout (block 0 = weights1)."""
    cop = (co + 1) // 2
    blk, kx, ky, i, op = np.meshgrid(np.arange(2), np.arange(m1), np.arange(m2), np.arange(ci), np.arange(cop),
                                     indexing="ij")
    q = ((((blk * m1 + kx) * m2 + ky) * ci + i) * cop + op).astype(np.uint64)
    z = quad_normals(seed, sample_idx, site, q)            # [2, m1, m2, Ci, cop, 4]
    re = np.stack([z[..., 0], z[..., 2]], axis=-1).reshape(2, m1, m2, ci, 2 * cop)[..., :co]
    im = np.stack([z[..., 1], z[..., 3]], axis=-1).reshape(2, m1, m2, ci, 2 * cop)[..., :co]
    # kernel-native [blk][kx][ky][i][o]  ->  logical [blk][i][o][kx][ky]
    re = np.ascontiguousarray(re.transpose(0, 3, 4, 1, 2)).astype(dtype)
    im = np.ascontiguousarray(im.transpose(0, 3, 4, 1, 2)).astype(dtype)
    return re, im
