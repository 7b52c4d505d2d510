"""Limb-level model of the device Montgomery product (provekit_b200/csrc/fr.cuh: mont_row / fr_mul_t) and of the lazily reduced
butterfly (ntt.cu), run on the CPU with exact integers.

The device code keeps two 8-limb windows (E: limbs k..k+7, O: limbs k+1..k+8) and DROPS three things on purpose: the carry out of
the m*p_odd chain, any overflow of the top limb O[7] when a chain's carry is added to it, and the carry out of the final merge.
With fully reduced inputs (a, b < p) that is the classical CIOS bound.  The TMA NTT kernel feeds the product with a in [0, 4p]
(an unreduced butterfly difference) and skips the final conditional subtraction; this model asserts, on the extreme values of
those ranges and on random ones, that every dropped quantity is zero and that the result is the expected residue below 2p."""
import random

import pytest

P = 21888242871839275222246405745257275088548364400416034343698204186575808495617
R = 1 << 256
MASK = 0xFFFFFFFF
NP0 = 0xEFFFFFFF  # -p^-1 mod 2^32
PL = [(P >> (32 * k)) & MASK for k in range(8)]


def limbs(x):
    assert 0 <= x < R
    return [(x >> (32 * k)) & MASK for k in range(8)]


def chain4(acc, xs, b, first_carry=0):
    """mad.lo.cc / madc.hi.cc chain over four aligned pairs of `acc` (8 limbs): returns the carry out of acc[7]"""
    carry = first_carry
    for k, x in enumerate(xs):
        v = acc[2 * k] + (acc[2 * k + 1] << 32) + x * b + carry
        acc[2 * k], acc[2 * k + 1] = v & MASK, (v >> 32) & MASK
        carry = v >> 64
        assert carry <= 1
    return carry


def mont_row(first, e, o, a, bi, dropped):
    """one CIOS row on the rolling windows, statement by statement as in fr.cuh::mont_row"""
    if first:
        for k in range(4):  # mul4: plain wide products
            v = a[2 * k + 1] * bi
            o[2 * k], o[2 * k + 1] = v & MASK, v >> 32
            v = a[2 * k] * bi
            e[2 * k], e[2 * k + 1] = v & MASK, v >> 32
    else:
        # mad4_rshift: e0 += o[1] (carry c); o = (o >> 64) + a_odd * bi + c, top limbs start from zero
        v = e[0] + o[1]
        e[0], c = v & MASK, v >> 32
        shifted = o[2:8] + [0, 0]
        c = chain4(shifted, [a[1], a[3], a[5], a[7]], bi, c)
        dropped.append(("mad4_rshift carry out", c))  # madc.hi.u32 on the last limb: no carry out is kept
        o[:] = shifted
        # mad4_top(e, o[7], a_even, bi)
        c = chain4(e, [a[0], a[2], a[4], a[6]], bi)
        dropped.append(("o[7] overflow (a*b)", (o[7] + c) >> 32))
        o[7] = (o[7] + c) & MASK
    m = (e[0] * NP0) & MASK
    c = chain4(o, [PL[1], PL[3], PL[5], PL[7]], m)
    dropped.append(("m*p_odd carry out", c))  # (void)mad4(o, ...)
    c = chain4(e, [PL[0], PL[2], PL[4], PL[6]], m)
    dropped.append(("o[7] overflow (m*p)", (o[7] + c) >> 32))
    o[7] = (o[7] + c) & MASK
    assert e[0] == 0  # the limb the reduction clears


def device_mul(a_int, b_int, lazy):
    a, b = limbs(a_int), limbs(b_int)
    e, o, dropped = [0] * 8, [0] * 8, []
    wins = (e, o)
    for i in range(8):
        E, O = wins[i & 1], wins[(i & 1) ^ 1]
        mont_row(i == 0, E, O, a, b[i], dropped)
    # after an even number of rows: result limb k = e[k] + o[k+1] (+ carry), limb 7 = e[7] + carry
    r, carry = [], 0
    for k in range(8):
        v = e[k] + (o[k + 1] if k < 7 else 0) + carry
        r.append(v & MASK)
        carry = v >> 32
    dropped.append(("final merge carry out", carry))
    val = sum(x << (32 * k) for k, x in enumerate(r))
    for what, d in dropped:
        assert d == 0, f"{what} dropped {d} for a={a_int:#x} b={b_int:#x}"
    if not lazy:
        val = val - P if val >= P else val
    return val


R_INV = pow(R, -1, P)


def check_mul(a, b, lazy):
    got = device_mul(a, b, lazy)
    assert got % P == a * b * R_INV % P
    assert got < (2 * P if lazy else P)
    # the analytic bound the kernel comments quote: (a b + m p) / 2^256 < a * p / 2^256 + p
    assert got <= a * b // R + P


def test_strict_product_extremes_and_random():
    rng = random.Random(1)
    ext = [0, 1, P - 1, P - 2, (1 << 253), P >> 1, MASK, (MASK << 224) % P]
    for a in ext:
        for b in ext:
            check_mul(a, b, lazy=False)
    for _ in range(300):
        check_mul(rng.randrange(P), rng.randrange(P), lazy=False)


def test_lazy_product_accepts_unreduced_differences():
    """a in [0, 4p] (fr_sub_lazy output), b < p (a twiddle): nothing is dropped and the result stays below 2p"""
    rng = random.Random(2)
    a_ext = [4 * P, 4 * P - 1, 3 * P + 1, 2 * P, 2 * P - 1, P, 0, 1, (4 * P) & ~MASK, 4 * P - (1 << 224)]
    b_ext = [P - 1, P - 2, 1, 0, (1 << 253), (P - 1) & ~MASK, MASK]
    for a in a_ext:
        for b in b_ext:
            check_mul(a, b, lazy=True)
    for _ in range(300):
        check_mul(rng.randrange(4 * P + 1), rng.randrange(P), lazy=True)


def test_lazy_product_would_overflow_beyond_its_contract():
    """the contract is tight where it matters: 5p < 2^256 is what keeps the ninth limb; an operand near 2^256 with a large
    second operand must trip the model (so the assertions above are not vacuous)"""
    with pytest.raises(AssertionError):
        for b in (R - 1, R - (1 << 200)):
            device_mul(R - 1, b, lazy=True)


def test_lazy_butterfly_ranges():
    """fr_add_lazy / fr_sub_lazy / fr_neg_lazy / fr_reduce_2p_once (fr.cuh) keep every tile value in [0, 2p) and the 256-bit
    adds never wrap; the final fr_normalize_2p yields the canonical residue"""
    rng = random.Random(3)

    def reduce_2p_once(s):
        assert s < R
        return s - 2 * P if s >= 2 * P else s

    def add_lazy(a, b):
        assert a + b < R
        return reduce_2p_once(a + b)

    def sub_lazy(a, b):
        assert a + 2 * P < R and a + 2 * P - b >= 0
        return a + 2 * P - b

    def normalize(x):
        for _ in range(2):
            x = x - P if x >= P else x
        return x

    ext = [0, 1, P - 1, P, P + 1, 2 * P - 1]
    vals = ext + [rng.randrange(2 * P) for _ in range(200)]
    for a in vals:
        n = reduce_2p_once(sub_lazy(0, a))  # fr_neg_lazy
        assert 0 <= n < 2 * P and (n + a) % P == 0
        for b in ext + [rng.randrange(2 * P) for _ in range(5)]:
            s, d = add_lazy(a, b), sub_lazy(a, b)
            assert 0 <= s < 2 * P and s % P == (a + b) % P
            assert 0 <= d <= 4 * P and d % P == (a - b) % P
            assert 0 <= reduce_2p_once(d) < 2 * P  # the twiddle-1 branch
            w = rng.randrange(P)
            t = device_mul(d, w, lazy=True)  # the twiddle product of the butterfly
            assert t < 2 * P and t % P == (a - b) * w * R_INV % P
            assert normalize(s) == (a + b) % P and normalize(t) == (a - b) * w * R_INV % P


# ---- the table-driven Feistel round of the hash (skyscraper.cuh: sky_round_sum<1|2>) ---------------------------------------
def _parse_tables():
    import os
    import re
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "provekit_b200", "csrc", "skyscraper.cuh")).read()

    def rows(block):
        out = []
        for r in re.findall(r"\{([^{}]*)\}", block):
            w = [int(x.strip().rstrip("u"), 16) for x in r.split(",") if x.strip()]
            if len(w) == 8:
                out.append(sum(v << (32 * k) for k, v in enumerate(w)))
        return out

    rc = rows(re.search(r"SKY_RC\[18\]\[8\] = \{(.*?)\n\};", src, re.S).group(1))
    rcq = rows(re.search(r"#define PK_RCQ_ROWS \\\n(.*?)\n\n", src, re.S).group(1))
    qp = rows(re.search(r"SKY_QP\[6\]\[8\] = \{(.*?)\n\};", src, re.S).group(1))
    return rc, rcq, qp


def test_round_tables_hold_what_their_comments_say():
    rc, rcq, qp = _parse_tables()
    assert len(rc) == 18 and len(rcq) == 90 and len(qp) == 6
    assert all(x < P for x in rc) and rc[0] == 0 and rc[17] == 0
    for i in range(18):
        for q in range(5):
            assert rcq[5 * i + q] == (rc[i] - q * P) % R
    assert qp == [q * P for q in range(6)]


def test_table_round_keeps_the_lazy_range():
    """state < B = 2p + 5 * 2^224 is a fixed point of: s0 = r + F; q = floor(s0[7] / (p[7] + 1)); s = s0 + (rc - q p) mod 2^256,
    for F = a lazy square of a state (< x^2 / 2^256 + p) or a bar output (< p); q never exceeds the table (4) and nothing wraps."""
    rc, rcq, _ = _parse_tables()
    B = 2 * P + 5 * (1 << 224)
    f_sqr_max = (B - 1) * (B - 1) // R + P  # (x^2 + m p) / 2^256 with m < 2^256
    assert B - 1 + f_sqr_max < R  # r + F never wraps
    top_p1 = (P >> 224) + 1
    rng = random.Random(4)
    f_bar_max = P + 6 * (1 << 224) - 1  # a lazily reduced bar output: y - q p with q = floor(y[7] / (p[7] + 1)) <= 5
    assert f_bar_max < f_sqr_max  # the squaring bound covers it
    for y in (R - 1, 5 * P + 12345, 6 * (P >> 224 << 224), P, 0):  # sky_reduce<LAZY>: never negative, below the bound
        qy = (y >> 224) // ((P >> 224) + 1)
        assert qy <= 5 and 0 <= y - qy * P <= f_bar_max
    samples = [(B - 1, f_sqr_max), (B - 1, 0), (0, f_sqr_max), (0, 0), (B - 1, P - 1), (P, P), (B - 1, f_bar_max)]
    samples += [(rng.randrange(B), rng.randrange(f_sqr_max + 1)) for _ in range(2000)]
    for r_, f in samples:
        s0 = r_ + f
        q = (s0 >> 224) // top_p1
        assert q <= 4
        assert 0 <= s0 - q * P < P + 5 * (1 << 224)
        for i in (0, 1, 5, 6, 11, 16, 17):
            s = (s0 + rcq[5 * i + q]) % R
            assert s == s0 - q * P + rc[i]  # the modular wrap of the table entry cancels exactly
            assert s < B and s % P == (r_ + f + rc[i]) % P
