"""The prepass turns uint16 millimetres into float64 metres without the division routine (csrc/bnv_frame.cuh mm_to_m):
q0 = RN(d * r), q = RN(q0 + RN_exact(d - 1000 q0) * r) with r = RN(1/1000).  Exhaustive proof that this is the correctly
rounded d / 1000.0 of load_depth (reference src/utils/common.py:93) for every uint16 d, in exact rational arithmetic
(float(Fraction) rounds to nearest even, i.e. it is the fused multiply-add's rounding)."""
from fractions import Fraction


def test_mm_to_m_is_the_correctly_rounded_quotient():
    r = Fraction(0.001)                       # the double nearest to 1/1000
    for d in range(65536):
        a = Fraction(d)
        q0 = Fraction(float(a * r))           # __dmul_rn
        rem = Fraction(float(a - q0 * 1000))  # __fma_rn(-q0, 1000, a)
        q = float(q0 + rem * r)               # __fma_rn(rem, r, q0)
        assert q == d / 1000.0, d
