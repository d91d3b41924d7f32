"""Host-side model of the fp32-parity products the tensor-core kernels build from `kind::tf32` MMAs (csrc/pno_sm100.cuh: tf32_trunc,
tf32_lo_rn, x3b_split4).  The hardware reads a raw fp32 operand as its tf32 TRUNCATION (pinned on the GPU by
test_gpu_tc.py::test_tf32_mma_truncates_raw_fp32_operands); these tests pin, in numpy, why the kernels split the way they do:

  * hi = trunc(v), lo = v - hi fed raw to the MMA is truncated again, always towards zero and always with the sign of v: every
    product shrinks by ~2.7e-7 -- a bias that rel-L2 of fields does not show but a cancellation-heavy sum does;
  * rounding the lo part to tf32 first (tf32_lo_rn) leaves a few 1e-8 (the dropped lo x lo term when both operands are split by
    truncation, the round-half-up of values with 0-2 bits to drop);
  * the N-concatenated form of the forward transform keeps all four terms of (a_hi + a_lo)(b_hi + b_lo);
  * the tf32 + bf16 product of the local branch holds every term once only if ONE of its bf16 "full" copies is a copy of the hi part.
"""
import numpy as np

MASK = np.uint32(0xFFFFE000)


def trunc_tf32(v):
    return (v.astype(np.float32).view(np.uint32) & MASK).view(np.float32)


def rn_tf32(v):
    return ((v.astype(np.float32).view(np.uint32) + np.uint32(0x1000)) & MASK).view(np.float32)


def mma(a, b):
    """What one kind::tf32 MMA contributes for fp32 operands a, b: the product of their tf32 truncations (exact in fp32-ish; the
    accumulator is modelled in float64 so that only the operand handling is measured)."""
    return trunc_tf32(a).astype(np.float64) * trunc_tf32(b).astype(np.float64)


def operands(n=1 << 20, seed=7):
    rng = np.random.default_rng(seed)
    return rng.standard_normal(n).astype(np.float32), rng.standard_normal(n).astype(np.float32)


def signed_rel_error(got, a, b):
    want = a.astype(np.float64) * b.astype(np.float64)
    return float(np.mean((got - want) * np.sign(want)) / np.mean(np.abs(want)))


def rel_l2(got, a, b):
    want = a.astype(np.float64) * b.astype(np.float64)
    return float(np.linalg.norm(got - want) / np.linalg.norm(want))


def test_truncation_split_with_raw_lo_shrinks_every_product():
    a, b = operands()
    a_lo, b_lo = a - trunc_tf32(a), b - trunc_tf32(b)              # exact in fp32, up to 13 significant bits
    got = mma(a, b) + mma(a_lo, b) + mma(a, b_lo)
    assert rel_l2(got, a, b) < 1e-6                                 # looks fine as a field error ...
    assert signed_rel_error(got, a, b) < -1.5e-7                    # ... but every product is too small: a bias


def test_truncation_split_with_rounded_lo():
    """tf32_lo_rn: the lo operand is exact in tf32.  Both operands truncation-split and the lo x lo term dropped (the weight-gradient
    kernels): what is left is that term's mean (lo parts carry the sign of their value) minus the round-half-up of short values."""
    a, b = operands()
    a_lo, b_lo = rn_tf32(a - trunc_tf32(a)), rn_tf32(b - trunc_tf32(b))
    assert np.array_equal(trunc_tf32(a_lo), a_lo) and np.array_equal(trunc_tf32(b_lo), b_lo)      # exact tf32 operands
    got = mma(a, b) + mma(a_lo, b) + mma(a, b_lo)
    assert rel_l2(got, a, b) < 2e-7
    assert abs(signed_rel_error(got, a, b)) < 8e-8                  # 4.6x smaller than with the raw lo parts
    # one operand truncation-split, the other split by rounding (forward transform x tables, mix xhat x weights)
    b_hi = rn_tf32(b)
    got = mma(a, b_hi) + mma(a_lo, b_hi) + mma(a, b - b_hi)
    assert rel_l2(got, a, b) < 2e-7
    assert abs(signed_rel_error(got, a, b)) < 4e-8


def test_rounded_split_is_unbiased():
    """The classic 3xTF32 split (hi = RN(v), lo = v - hi, lo truncated by the MMA): lo has a random sign, no bias."""
    a, b = operands()
    a_hi, b_hi = rn_tf32(a), rn_tf32(b)
    got = mma(a_hi, b_hi) + mma(a - a_hi, b_hi) + mma(a_hi, b - b_hi)
    assert rel_l2(got, a, b) < 1e-7
    assert abs(signed_rel_error(got, a, b)) < 1e-9


def test_n_concatenated_product_keeps_the_fourth_term():
    """Forward transform, fp32 parity: A x [B_hi | B_lo] and A_lo x [B_hi | B_lo] with the halves added afterwards."""
    a, b = operands()
    a_lo = rn_tf32(a - trunc_tf32(a))
    b_hi = rn_tf32(b)                                               # tables are split on the host by rounding
    b_lo = rn_tf32((b.astype(np.float64) - b_hi.astype(np.float64)).astype(np.float32))
    main = mma(a, b_hi) + mma(a_lo, b_hi)
    corr = mma(a, b_lo) + mma(a_lo, b_lo)
    three = mma(a, b_hi) + mma(a_lo, b_hi) + mma(a, b_lo)
    assert rel_l2(main + corr, a, b) <= rel_l2(three, a, b) + 1e-9
    assert rel_l2(main + corr, a, b) < 1.5e-7
    assert abs(signed_rel_error(main + corr, a, b)) < 4e-8


def test_x3b_product_with_bf16_cross_terms():
    """Local branch: one tf32 MMA on the raw tiles + two bf16 MMAs.  bf16(h(a)) x bf16(l(b)) + bf16(l(a)) x bf16(b) holds every term
    of the product once; with bf16(a) in the first one (as first built) l(a) l(b) is counted twice: +1.2e-7 on every output."""
    import torch
    a, b = operands(1 << 18)

    def bf16(v):
        return torch.from_numpy(v.copy()).to(torch.bfloat16).to(torch.float64).numpy()       # round to nearest even

    a_lo, b_lo = a - trunc_tf32(a), b - trunc_tf32(b)
    twice = mma(a, b) + bf16(a) * bf16(b_lo) + bf16(a_lo) * bf16(b)
    once = mma(a, b) + bf16(trunc_tf32(a)) * bf16(b_lo) + bf16(a_lo) * bf16(b)
    assert rel_l2(once, a, b) < 3e-6 and rel_l2(twice, a, b) < 3e-6
    assert signed_rel_error(twice, a, b) > 8e-8
    assert abs(signed_rel_error(once, a, b)) < 2e-8
