"""C oracle (oracle/kernels.c) == pure-Python restatement (oracle/pyref.py) on small seeded inputs."""
import ctypes

import numpy as np
import pytest

from helpers import from_mont, ptr, rand_fr, to_mont
from oracle import pyref as o

P = o.P
sz = ctypes.c_size_t


def _rand_ints(seed, n):
    rng = np.random.default_rng(seed)
    return from_mont(rand_fr(rng, n))  # any list of ints < p


@pytest.mark.parametrize("log_n", [1, 3, 6])
def test_wavelet(orc, log_n):
    xs = _rand_ints(log_n, 1 << log_n)
    a = to_mont(xs)
    orc.orc_evals_to_coeffs(ptr(a), log_n)
    assert from_mont(a) == o.evals_to_coeffs(xs)
    orc.orc_coeffs_to_evals(ptr(a), log_n)
    assert from_mont(a) == xs
    # definition check: evals of coefficient form == multilinear evaluation at boolean points
    c = o.evals_to_coeffs(xs)
    n = log_n
    for idx in (0, (1 << n) - 1, 1):
        pt = [(idx >> (n - 1 - j)) & 1 for j in range(n)]  # x_0 <-> MSB (whir convention)
        assert o.eval_coeffs_at_point(c, pt) == xs[idx]


@pytest.mark.parametrize("log_n,rate", [(4, 1), (5, 1), (6, 2), (7, 4), (8, 1)])
def test_rs_encode(orc, log_n, rate):
    xs = _rand_ints(100 + log_n, 1 << log_n)
    leaves = o.rs_encode_leaves(xs, rate, 4)
    L = len(leaves)
    out = np.zeros((L * 16, 4), np.uint64)
    orc.orc_rs_encode(ptr(to_mont(xs)), log_n, rate, 4, ptr(out), sz(16), sz(0))
    got = from_mont(out)
    assert [got[i * 16:(i + 1) * 16] for i in range(L)] == leaves
    # stacked layout (batch of 2): second polynomial at column offset 16
    out2 = np.zeros((L * 32, 4), np.uint64)
    orc.orc_rs_encode(ptr(to_mont(xs)), log_n, rate, 4, ptr(out2), sz(32), sz(16))
    got2 = from_mont(out2)
    assert [got2[i * 32 + 16:(i + 1) * 32] for i in range(L)] == leaves
    # the verifier's view (whir.go:139-142 + whir_utilities.go:180-186): folding the leaf with r equals
    # the folded polynomial evaluated at (g^16)^i
    r = _rand_ints(5, 4)
    folded = o.fold_coeffs(xs, r)
    g16 = pow(o.root_of_unity(log_n + rate), 16, P)
    for i in (0, 1, L - 1):
        assert o.eval_multilinear_coeffs(leaves[i], r) == o.eval_univariate(folded, pow(g16, i, P))


def test_univariate_fold_eq_mle_dot(orc):
    xs = _rand_ints(1, 1 << 13)
    z = _rand_ints(2, 1)[0]
    out = np.zeros(4, np.uint64)
    orc.orc_eval_univariate(ptr(to_mont(xs)), sz(len(xs)), ptr(to_mont([z])), ptr(out))
    assert from_mont(out)[0] == o.eval_univariate(xs, z)
    # univariate value == multilinear value at expand_from_univariate (utilities.go:182-190)
    assert o.eval_univariate(xs[:64], z) == o.eval_coeffs_at_point(xs[:64], o.expand_from_univariate(z, 6))
    r = _rand_ints(3, 4)
    outf = np.zeros((len(xs) // 16, 4), np.uint64)
    orc.orc_fold_coeffs(ptr(to_mont(xs)), 13, ptr(to_mont(r)), 4, ptr(outf))
    assert from_mont(outf) == o.fold_coeffs(xs, r)
    pt = _rand_ints(4, 7)
    acc0 = _rand_ints(5, 128)
    acc = to_mont(acc0)
    s = _rand_ints(6, 1)[0]
    orc.orc_eval_eq_accumulate(ptr(to_mont(pt)), 7, ptr(to_mont([s])), ptr(acc))
    eq = o.eval_eq(pt, s)
    assert from_mont(acc) == [(x + y) % P for x, y in zip(acc0, eq)]
    assert eq[5] == s * o.eq_poly_outside(pt, [(5 >> (6 - j)) & 1 for j in range(7)]) % P
    ev = _rand_ints(7, 128)
    orc.orc_mle_eval(ptr(to_mont(ev)), 7, ptr(to_mont(pt)), ptr(out))
    assert from_mont(out)[0] == sum(e * q for e, q in zip(ev, o.eval_eq(pt))) % P
    orc.orc_dot(ptr(to_mont(ev)), ptr(to_mont(acc0)), sz(128), ptr(out))
    assert from_mont(out)[0] == sum(e * q for e, q in zip(ev, acc0)) % P


@pytest.mark.parametrize("L,w", [(2, 16), (8, 32), (16, 3), (4, 1)])
def test_merkle(orc, L, w):
    xs = _rand_ints(L * w, L * w)
    leaves = [xs[i * w:(i + 1) * w] for i in range(L)]
    for version, comp in ((2, o.compress), (1, o.compress_v1)):
        nodes = np.zeros((2 * L, 4), np.uint64)
        orc.orc_merkle_build(ptr(to_mont(xs)), sz(L), sz(w), ptr(nodes), version)
        exp = o.merkle_tree(leaves, comp)
        assert from_mont(nodes)[1:] == exp[1:]
        idx = sorted(set(int(i) for i in np.random.default_rng(L).integers(0, L, size=3)))
        sib, pre, suf, _ = o.merkle_multipath(exp, idx)
        prev = []
        for i, s, k, sf in zip(idx, sib, pre, suf):
            prev = prev[:k] + sf
            assert o.merkle_verify_path(exp[1], i, leaves[i], s, prev, comp)


def test_zk_sumcheck_rounds(orc):
    log_n = 6
    arrs = [_rand_ints(10 + k, 1 << log_n) for k in range(4)]
    dev = [to_mont(x) for x in arrs]
    fold = None
    n = log_n
    for rnd in range(log_n):
        out3 = np.zeros((3, 4), np.uint64)
        f = None if fold is None else ptr(to_mont([fold]))
        orc.orc_zk_sumcheck_round(ptr(dev[0]), ptr(dev[1]), ptr(dev[2]), ptr(dev[3]), n, f, ptr(out3))
        exp3, arrs = o.zk_sumcheck_round(*arrs, fold)
        assert tuple(from_mont(out3)) == exp3
        if fold is not None:
            n -= 1
        for k in range(4):
            assert from_mont(dev[k][: 1 << n]) == arrs[k]
        fold = _rand_ints(50 + rnd, 1)[0]


def test_whir_sumcheck_rounds(orc):
    log_n = 6
    p, w = _rand_ints(20, 1 << log_n), _rand_ints(21, 1 << log_n)
    dp, dw = to_mont(p), to_mont(w)
    total = sum(a * b for a, b in zip(p, w)) % P
    fold = None
    n = log_n
    for rnd in range(log_n):
        out3 = np.zeros((3, 4), np.uint64)
        f = None if fold is None else ptr(to_mont([fold]))
        orc.orc_whir_sumcheck_round(ptr(dp), ptr(dw), n, f, ptr(out3))
        (h0, h1, h2), p, w = o.whir_sumcheck_round(p, w, fold)
        assert tuple(from_mont(out3)) == (h0, h1, h2)
        assert (h0 + h1) % P == total          # CheckSumOverBool, utilities.go:167-170
        if fold is not None:
            n -= 1
        assert from_mont(dp[: 1 << n]) == p and from_mont(dw[: 1 << n]) == w
        fold = _rand_ints(70 + rnd, 1)[0]
        # EvaluateQuadraticPolynomialFromEvaluationList, utilities.go:148-154
        inv2 = o.HALF
        b1 = (-h2 + 4 * h1 - 3 * h0) * inv2 % P
        b2 = (h2 - 2 * h1 + h0) * inv2 % P
        total = (fold * fold * b2 + fold * b1 + h0) % P
