"""Sponge-free pins of the sumcheck / fold conventions against the reference-produced proof (tests/golden/).

The Fiat-Shamir challenge VALUES cannot be re-derived here (spongefish is un-vendored), but a
sumcheck challenge is algebraically visible in the proof: round i's message must satisfy h_i(0) + h_i(1) = h_{i-1}(alpha),
so alpha is a root in Fr of a cubic (zk-sumcheck, provekit/prover/src/whir_r1cs.rs:280-345) or a quadratic (WHIR
sumcheck [whir]) whose coefficients are proof bytes.  A random cubic has a root with probability ~0.63 and a random
quadratic with 0.5, so 19 + 22 consecutive links all having roots (chance ~1e-4 * 2e-7) pins the message formats and
the chaining; and the recovered folding randomness then has to reproduce the final polynomial from the opened
leaves (9 equations over Fr, one combination of roots fits), which pins the fold order, the leaf/coefficient
convention, the query points and the final-polynomial encoding — the items SURVEY A.7 lists as unverifiable without
the sponge."""
import itertools

import pytest

import fixture_walk as fw
from oracle import pyref as o
from poly_roots import P, quad_from_evals, roots


@pytest.fixture(scope="module")
def proof():
    return fw.walk_proof()


def cubic_link(prev4, cur4):
    """zk-sumcheck: messages are the 4 coefficients of the cubic (lowest first); challenge candidates"""
    s = (2 * cur4[0] + cur4[1] + cur4[2] + cur4[3]) % P        # h_i(0) + h_i(1)
    f = list(prev4)
    f[0] = (f[0] - s) % P
    return roots(f)


def quad_link(prev3, cur3):
    """WHIR sumcheck: messages are the evaluations h(0), h(1), h(2); challenge candidates"""
    f = quad_from_evals(*prev3)
    f[0] = (f[0] - (cur3[0] + cur3[1])) % P
    return roots(f)


def test_zk_sumcheck_messages_chain(proof):
    zk = proof["zk_sumcheck"]
    assert len(zk) == 20
    cands = [cubic_link(zk[i - 1], zk[i]) for i in range(1, len(zk))]
    assert all(len(c) in (1, 3) for c in cands), [len(c) for c in cands]
    # every candidate really is a challenge consistent with the next message
    for i, c in enumerate(cands):
        for a in c:
            assert o.eval_univariate(zk[i], a) == (2 * zk[i + 1][0] + sum(zk[i + 1][1:])) % P
    # the wrong reading (messages as evaluations at 0, 1, 2, 3 of the cubic) does not chain
    bad = 0
    for i in range(1, len(zk)):
        s = (zk[i][0] + zk[i][1]) % P
        # Lagrange through (0..3) -> coefficients
        y = zk[i - 1]
        inv6, inv2 = pow(6, P - 2, P), pow(2, P - 2, P)
        c3 = (y[3] - 3 * y[2] + 3 * y[1] - y[0]) * inv6 % P
        c2 = ((y[2] - 2 * y[1] + y[0]) * inv2 - 3 * c3) % P
        c1 = (y[1] - y[0] - c2 - c3) % P
        bad += len(roots([(y[0] - s) % P, c1, c2, c3])) == 0
    assert bad >= 3


def whir_blocks(w):
    return [w["initial_sumcheck"]] + [r["sumcheck"] for r in w["rounds"]]


@pytest.mark.parametrize("which", ["whir_h", "whir_w"])
def test_whir_sumcheck_messages_chain(proof, which):
    w = proof[which]
    links = 0
    for blk in whir_blocks(w):
        assert len(blk) == 4
        for i in range(1, 4):
            assert len(quad_link(blk[i - 1], blk[i])) == 2
            links += 1
    if w["final_sumcheck"]:  # the final sumcheck continues the last block's claim (no new constraint in between)
        assert len(quad_link(whir_blocks(w)[-1][3], w["final_sumcheck"][0])) == 2
        links += 1
    assert links == (16 if which == "whir_w" else 6)


def test_final_fold_pins_fold_order_and_final_polynomial(proof):
    """fold(leaf_q, r) == finalPoly((g^16)^idx_q) for the 9 final queries of the witness WHIR, r = the last block's four
    sumcheck challenges in drawing order with r[j] binding bit j of the leaf index (computeFold = MultivarPoly(leaf, r),
    recursive-verifier/app/circuit/whir_utilities.go:180-186), g = generator of the last commitment's domain 2^18."""
    w = proof["whir_w"]
    blk, fin = w["rounds"][3]["sumcheck"], w["final_sumcheck"]
    cands = [quad_link(blk[i - 1], blk[i]) for i in range(1, 4)] + [quad_link(blk[3], fin[0])]
    leaves, idx, fc = w["final_answers"], w["final_multipath"][3], w["final_coeffs"]
    assert len(leaves) == 9 and len(fc) == 2
    gen = pow(o.root_of_unity(18), 16, P)
    pts = [pow(gen, i, P) for i in idx]
    fits = []
    for combo in itertools.product(*cands):
        r = list(combo)
        if all(o.eval_multilinear_coeffs(leaf, r) == o.eval_univariate(fc, z) for leaf, z in zip(leaves, pts)):
            fits.append(r)
    assert len(fits) == 1                                   # 16 candidate tuples, exactly one reproduces all 9 values
    r = fits[0]
    # the reversed binding order, or a neighbouring domain, fits for no tuple
    for combo in itertools.product(*cands):
        assert not all(o.eval_multilinear_coeffs(leaf, list(combo)[::-1]) == o.eval_univariate(fc, z) for leaf, z in zip(leaves, pts))
    for dl in (17, 19):
        g2 = pow(o.root_of_unity(dl), 16, P)
        assert not all(o.eval_multilinear_coeffs(leaf, r) == o.eval_univariate(fc, pow(g2, i, P)) for leaf, i in zip(leaves, idx))
    # CoefficientList::fold on the committed polynomial itself: its 32 coefficients follow from the opened leaves
    # (entry k of leaf i is c_k + c_{k+16} * Y_i); folding them with r gives the two final coefficients
    y0, y1 = pts[0], pts[1]
    inv = pow((y1 - y0) % P, P - 2, P)
    hi = [(leaves[1][k] - leaves[0][k]) * inv % P for k in range(16)]
    lo = [(leaves[0][k] - hi[k] * y0) % P for k in range(16)]
    coeffs = lo + hi
    for leaf, y in zip(leaves, pts):                        # all nine leaves lie on those 16 lines
        assert [(lo[k] + hi[k] * y) % P for k in range(16)] == leaf
    assert o.fold_coeffs(coeffs, r) == fc
    # and the RS-encode restatement reproduces the opened leaves from those coefficients (rate 2^-13)
    enc = o.rs_encode_leaves(coeffs, 13)
    assert [enc[i] for i in idx] == leaves


def test_intermediate_round_answers_fold_into_the_next_committed_polynomial(proof):
    """Round 3 of the witness WHIR opens 11 leaves of the round-2 commitment (domain 2^19, 16 values each).  Folded with
    that round's four sumcheck challenges they must be the values of the round-3 committed polynomial (its 32
    coefficients are known from the final openings) at (g_19^16)^idx: 11 equations; r1..r3 have two candidates each, r4
    (no following message to chain to) enters linearly and is solved from the first equation.  Exactly one candidate
    tuple satisfies all eleven; with it the claim update h'(0) + h'(1) = h(r4) + ood_ans + sum_q gamma^(q+1) fold_q
    (calculateShiftValue, recursive-verifier/app/circuit/whir_utilities.go) has a solution gamma in Fr."""
    w = proof["whir_w"]
    leaves, idx = w["final_answers"], w["final_multipath"][3]
    g18 = pow(o.root_of_unity(18), 16, P)
    y0, y1 = pow(g18, idx[0], P), pow(g18, idx[1], P)
    inv = pow((y1 - y0) % P, P - 2, P)
    hi = [(leaves[1][k] - leaves[0][k]) * inv % P for k in range(16)]
    f3 = [(leaves[0][k] - hi[k] * y0) % P for k in range(16)] + hi
    rd = w["rounds"][3]
    ans, qidx = rd["answers"], rd["multipath"][3]
    assert len(ans) == 11 and all(len(a) == 16 for a in ans)
    g19 = pow(o.root_of_unity(19), 16, P)
    target = [o.eval_univariate(f3, pow(g19, i, P)) for i in qidx]
    blk3, blk4 = w["rounds"][2]["sumcheck"], w["rounds"][3]["sumcheck"]
    cands = [quad_link(blk3[i - 1], blk3[i]) for i in range(1, 4)]
    fits = []
    for combo in itertools.product(*cands):
        r = list(combo)
        a = [o.eval_multilinear_coeffs(l[:8], r) for l in ans]      # fold = a + r4 * b: r4 binds the top bit of the leaf index
        b = [o.eval_multilinear_coeffs(l[8:], r) for l in ans]
        r4 = (target[0] - a[0]) * pow(b[0], P - 2, P) % P
        if all((x + r4 * y) % P == t for x, y, t in zip(a, b, target)):
            fits.append(r + [r4])
    assert len(fits) == 1
    r = fits[0]
    assert [o.eval_multilinear_coeffs(l, r) for l in ans] == target
    # claim update with the combination randomness gamma: constant + sum_q fold_q gamma^(q+1) = 0 has a root
    last = o.eval_univariate(quad_from_evals(*blk3[3]), r[3])
    s4 = (blk4[0][0] + blk4[0][1]) % P
    poly = [(last + rd["ood"][0] - s4) % P] + target
    assert len(roots(poly)) >= 1


def _solve3(rows, rhs):
    """3x3 linear system over Fr"""
    m = [list(r) + [b] for r, b in zip(rows, rhs)]
    for c in range(3):
        piv = next(i for i in range(c, 3) if m[i][c] % P)
        m[c], m[piv] = m[piv], m[c]
        inv = pow(m[c][c], P - 2, P)
        m[c] = [x * inv % P for x in m[c]]
        for i in range(3):
            if i != c and m[i][c]:
                f = m[i][c]
                m[i] = [(x - f * y) % P for x, y in zip(m[i], m[c])]
    return [m[i][3] for i in range(3)]


def test_blinding_commitment_is_fully_checkable(proof):
    """The blinding WHIR commits two 8-variable polynomials at rate 1/2: 2^9 evaluations = 32 leaves of 2 x 16 values, and
    its 122 STIR queries open ALL 32 leaves, so the reference's whole first codeword is in the proof.  Sponge-free pins:
      K1  every one of the 32 leaf columns is the evaluation of a polynomial of degree < 16 on the order-32 subgroup
          generated by (arkworks root of order 2^9)^16 (low-degree test), and the oracle's RS-encode reproduces all 32
          leaves from the recovered coefficients with the batch layout [poly 0: 16 | poly 1: 16];
      K3  the two OOD answers are the univariate values of the two polynomials at ONE common point
          (gcd(F - a_F, G - a_G) is linear);
      K4/K8  the round-0 committed polynomial (16 coefficients = every final leaf) is fold(F + b*G, r) with r = the four
          initial-sumcheck challenges, r[j] binding bit j: 16 equations for the two unknowns (b, r4); one candidate
          tuple fits all of them."""
    h = proof["whir_h"]
    r0 = h["rounds"][0]
    ans, idx = r0["answers"], r0["multipath"][3]
    assert idx == list(range(32)) and all(len(a) == 32 for a in ans)
    w32 = pow(o.root_of_unity(9), 16, P)
    assert pow(w32, 32, P) == 1 and pow(w32, 16, P) != 1
    inv32 = pow(32, P - 2, P)
    polys = []
    for b in range(2):
        coeffs = [0] * 256
        for k in range(16):
            v = [ans[i][16 * b + k] for i in range(32)]
            for t in range(32):
                c = sum(v[i] * pow(w32, (-i * t) % 32, P) for i in range(32)) * inv32 % P
                if t >= 16:
                    assert c == 0, "column is not a rate-1/2 codeword"
                else:
                    coeffs[k + 16 * t] = c
        polys.append(coeffs)
    F, G = polys
    enc_f, enc_g = o.rs_encode_leaves(F, 1), o.rs_encode_leaves(G, 1)
    assert [enc_f[i] + enc_g[i] for i in range(32)] == ans
    # K0 + create_masked_polynomial (provekit/common/src/utils/zk_utils.rs:3-11) + the blinding layout of
    # whir_r1cs.rs:211-225: in EVALUATION form (our wavelet convention) the first committed polynomial is
    # [4 coefficients of each of the m_0 = 20 blinding cubics | zero padding to 2^7 | 2^7 mask values]
    f_evals = o.coeffs_to_evals(F)
    assert all(f_evals[:80]) and not any(f_evals[80:128]) and all(f_evals[128:])
    assert o.evals_to_coeffs(f_evals) == F
    # "Sum of G over boolean hypercube" (whir_r1cs.rs:240-260) from those cubics: 2^(m_0 - 1) * sum_i (g_i(0) + g_i(1))
    m_0 = 20
    total = sum(f_evals[4 * i] + sum(f_evals[4 * i:4 * i + 4]) for i in range(m_0))
    assert proof["sum_g"] == pow(2, m_0 - 1, P) * total % P
    # OOD
    from poly_roots import pgcd
    a_f, a_g = proof["commit_h"]["ood"]
    g = pgcd([(F[0] - a_f) % P] + F[1:], [(G[0] - a_g) % P] + G[1:])
    assert len(g) == 2
    z = (-g[0]) * pow(g[1], P - 2, P) % P
    assert o.eval_univariate(F, z) == a_f and o.eval_univariate(G, z) == a_g
    assert o.eval_coeffs_at_point(F, o.expand_from_univariate(z, 8)) == a_f      # the multilinear reading of the same point
    # batching + first fold
    fl = h["final_answers"]
    assert len(fl) == 13 and all(x == fl[0] for x in fl) and len(fl[0]) == 16
    cp = fl[0]
    init = h["initial_sumcheck"]
    cands = [quad_link(init[i - 1], init[i]) for i in range(1, 4)]
    fits = []
    for combo in itertools.product(*cands):
        r = list(combo)
        f_lo = [o.eval_multilinear_coeffs(F[16 * t:16 * t + 8], r) for t in range(16)]
        f_hi = [o.eval_multilinear_coeffs(F[16 * t + 8:16 * t + 16], r) for t in range(16)]
        g_lo = [o.eval_multilinear_coeffs(G[16 * t:16 * t + 8], r) for t in range(16)]
        g_hi = [o.eval_multilinear_coeffs(G[16 * t + 8:16 * t + 16], r) for t in range(16)]
        # f_lo + b g_lo + r4 f_hi + (r4 b) g_hi = cp: linear in (b, r4, u = r4 b)
        b_, r4, u = _solve3([(g_lo[t], f_hi[t], g_hi[t]) for t in range(3)], [(cp[t] - f_lo[t]) % P for t in range(3)])
        if u == r4 * b_ % P and all((f_lo[t] + b_ * g_lo[t] + r4 * f_hi[t] + u * g_hi[t]) % P == cp[t] for t in range(16)):
            fits.append((b_, r + [r4]))
    assert len(fits) == 1
    b_, r = fits[0]
    batched = [(x + b_ * y) % P for x, y in zip(F, G)]
    assert o.fold_coeffs(batched, r) == cp
    # the round-0 commitment (rate 2^-4 on the domain 2^8) re-encodes to the opened final leaves
    enc = o.rs_encode_leaves(cp, 4)
    assert [enc[i] for i in h["final_multipath"][3]] == fl


def test_blinding_whir_final_equation_holds_without_the_sponge(proof):
    """Every verifier challenge of the blinding WHIR is recoverable from the proof by algebra (batching b, OOD point z,
    initial combination randomness, both blocks of folding randomness, the round's OOD point and combination
    randomness), so the verifier's final equation  last_claim == W(R) * final_polynomial  (computeWPoly,
    recursive-verifier/app/circuit/whir_utilities.go:133-166; whir.go:203 reverses the folding randomness) can be
    evaluated with the oracle's conventions and NO Fiat-Shamir layer.  Of the 72 candidate tuples exactly one satisfies
    it — a 254-bit equality that pins the weight conventions (OOD constraint first, powers of one challenge, eq over the
    reversed randomness, deferred linear-weight evaluation), the claim composition (ood_f + b ood_g) + gamma (s_f + b s_g)
    and the sumcheck/fold chain of a whole WHIR opening against the reference's own output."""
    from poly_roots import pgcd
    h = proof["whir_h"]
    r0 = h["rounds"][0]
    ans = r0["answers"]
    w32 = pow(o.root_of_unity(9), 16, P)
    inv32 = pow(32, P - 2, P)
    polys = []
    for b in range(2):
        coeffs = [0] * 256
        for k in range(16):
            v = [ans[i][16 * b + k] for i in range(32)]
            for t in range(16):
                coeffs[k + 16 * t] = sum(v[i] * pow(w32, (-i * t) % 32, P) for i in range(32)) * inv32 % P
        polys.append(coeffs)
    F, G = polys
    a_f, a_g = proof["commit_h"]["ood"]
    g = pgcd([(F[0] - a_f) % P] + F[1:], [(G[0] - a_g) % P] + G[1:])
    z = (-g[0]) * pow(g[1], P - 2, P) % P
    cp = h["final_answers"][0]
    init, blk1 = h["initial_sumcheck"], r0["sumcheck"]
    fit = []
    for combo in itertools.product(*[quad_link(init[i - 1], init[i]) for i in range(1, 4)]):
        r = list(combo)
        f_lo = [o.eval_multilinear_coeffs(F[16 * t:16 * t + 8], r) for t in range(16)]
        f_hi = [o.eval_multilinear_coeffs(F[16 * t + 8:16 * t + 16], r) for t in range(16)]
        g_lo = [o.eval_multilinear_coeffs(G[16 * t:16 * t + 8], r) for t in range(16)]
        g_hi = [o.eval_multilinear_coeffs(G[16 * t + 8:16 * t + 16], r) for t in range(16)]
        b_, r4, u = _solve3([(g_lo[t], f_hi[t], g_hi[t]) for t in range(3)], [(cp[t] - f_lo[t]) % P for t in range(3)])
        if u == r4 * b_ % P and all((f_lo[t] + b_ * g_lo[t] + r4 * f_hi[t] + u * g_hi[t]) % P == cp[t] for t in range(16)):
            fit.append((b_, r + [r4]))
    assert len(fit) == 1
    bq, r_a = fit[0]
    # initial claim: OOD constraint (weight 1) + linear statement (weight gamma0); "Polynomial sums" = (s_f, s_g)
    s_f, s_g = proof["blind_sums"]
    gamma0 = ((init[0][0] + init[0][1]) - (a_f + bq * a_g)) * pow((s_f + bq * s_g) % P, P - 2, P) % P
    # round 0: OOD point of the 16-coefficient polynomial, STIR points = all 32 leaf positions, combination randomness
    ood0, fconst = r0["ood"][0], h["final_coeffs"][0]
    z_cands = roots([(cp[0] - ood0) % P] + cp[1:])
    pts = [pow(w32, i, P) for i in r0["multipath"][3]]
    folds = [o.eval_univariate(cp, y) for y in pts]
    batched_leaves = [[(l[k] + bq * l[16 + k]) % P for k in range(16)] for l in ans]       # rlcBatchedLeaves
    assert [o.eval_multilinear_coeffs(l, r_a) for l in batched_leaves] == folds            # computeFold
    last_a = o.eval_univariate(quad_from_evals(*init[3]), r_a[3])
    g1_cands = roots([(last_a + ood0 - (blk1[0][0] + blk1[0][1])) % P] + folds)            # calculateShiftValue
    assert z_cands and g1_cands
    n, hits = 8, 0
    for combo in itertools.product(*[quad_link(blk1[i - 1], blk1[i]) for i in range(1, 4)]):
        r123 = list(combo)
        lo, hi = o.eval_multilinear_coeffs(cp[:8], r123), o.eval_multilinear_coeffs(cp[8:], r123)
        r4 = (fconst - lo) * pow(hi, P - 2, P) % P                                         # fold(c', r) = final constant
        big_r = (r_a + r123 + [r4])[::-1]
        last = o.eval_univariate(quad_from_evals(*blk1[3]), r4)
        for zp, g1 in itertools.product(z_cands, g1_cands):
            value = o.eq_poly_outside(o.expand_from_univariate(z, n), big_r[:n]) + gamma0 * h["deferred"][0]
            gp = 1
            for pt in [zp] + pts:
                value += gp * o.eq_poly_outside(o.expand_from_univariate(pt, 4), big_r[:4])
                gp = gp * g1 % P
            hits += last == value % P * fconst % P
    assert hits == 1
