"""The offsets -> sampling-location epilogue (csrc/gemm_tc.cu: div_rn_by) replaces div.rn.f32 by q0 = RN(x r), r = RN(1/n), and two
FMA residual corrections.  This restates the sequence with exact rational arithmetic (every RN to binary32 done exactly) and
checks that it returns the correctly rounded quotient RN(x / n) -- what the reference's `sampling_offsets / offset_normalizer`
(ops/modules/ms_deform_attn.py:190-192) computes -- for grid sizes and offsets of the magnitudes the path sees."""
from fractions import Fraction
import numpy as np


def rn32(fr):
    """Round a Fraction to the nearest binary32 (ties to even), normal range; returns a Fraction that is exactly a binary32."""
    if fr == 0:
        return Fraction(0)
    sgn = -1 if fr < 0 else 1
    a = abs(fr)
    e = a.numerator.bit_length() - a.denominator.bit_length()
    if Fraction(2) ** e > a:
        e -= 1                                   # 2^e <= a < 2^(e+1)
    assert -126 <= e <= 127
    scaled = a * Fraction(2) ** (23 - e)         # in [2^23, 2^24)
    m, rem = divmod(scaled.numerator, scaled.denominator)
    twice = 2 * rem
    if twice > scaled.denominator or (twice == scaled.denominator and (m & 1)):
        m += 1
    return sgn * Fraction(m) * Fraction(2) ** (e - 23)


def fma32(a, b, c):
    return rn32(a * b + c)


def div_rn_by(x, n):
    r = rn32(Fraction(1) / n)
    q = rn32(x * r)
    e = fma32(-q, n, x)
    q = fma32(e, r, q)
    e = fma32(-q, n, x)
    return fma32(e, r, q)


def test_two_corrections_give_the_ieee_quotient():
    rs = np.random.RandomState(0)
    divisors = [1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 14, 16, 20, 25, 28, 32, 40, 56, 64, 100, 112, 127, 224, 255, 511]
    xs = np.concatenate([rs.standard_normal(300), 10.0 ** rs.uniform(-6, 2, 300) * rs.choice([-1, 1], 300),
                         np.float32([0.0, 1.0, -1.0, 7.0, 1e-3, 0.1, 3.0000002, 16777215.0, 0.99999994])]).astype(np.float32)
    bad = 0
    for n in divisors:
        fn = Fraction(n)
        for x in xs:
            fx = Fraction(float(x))
            want = rn32(fx / fn)
            got = div_rn_by(fx, fn)
            bad += got != want
            assert float(want) == float(np.float32(x) / np.float32(n))      # rn32 itself agrees with IEEE division
    assert bad == 0


def test_one_correction_is_not_always_enough_but_two_are_on_hard_cases():
    # significands next to a power of two are where a single correction can stop one ulp short
    hard = [np.float32(1.0) + np.float32(k) * np.float32(2 ** -23) for k in range(1, 40)]
    hard += [np.float32(2.0) - np.float32(k) * np.float32(2 ** -23) for k in range(1, 40)]
    for n in (3, 7, 14, 28, 56, 11, 13, 49):
        for x in hard:
            fx, fn = Fraction(float(x)), Fraction(n)
            assert div_rn_by(fx, fn) == rn32(fx / fn)
