"""Writes tests/golden/poseidon-1000.challenges.json: the Fiat-Shamir challenge VALUES of the reference-produced proof
(tests/golden/poseidon-1000.transcript.bin) that can be recovered from the proof bytes by algebra alone — no sponge
(see tests/test_fixture_algebra.py for how each one is pinned).  They are consecutive outputs of the reference's
Skyscraper duplex sponge with all absorbed inputs in between known, i.e. ready-made known-answer vectors for any
implementation of the transcript once its initial state (IV = hash of the domain-separator string) is known.

    python tests/golden/make_challenges.py        (needs only the committed transcript; ~1 min)
"""
import itertools
import json
import os
import struct
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import fixture_walk as fw  # noqa: E402
import test_fixture_algebra as T  # noqa: E402
from oracle import pyref as o  # noqa: E402
from poly_roots import P, pgcd, quad_from_evals, roots  # noqa: E402


def hx(v):
    return "0x%064x" % v


def main():
    pr = fw.walk_proof()
    h = pr["whir_h"]
    r0 = h["rounds"][0]
    F, G, _ = T.blinding_polys(pr)
    g, _, alphas = T._blinding_cubics_and_alphas(pr)
    zk = pr["zk_sumcheck"]
    rho = (2 * zk[0][0] + sum(zk[0][1:])) * pow(pr["sum_g"], P - 2, P) % P
    # rand (r_1..r_20): the unique tuple satisfying the Spartan relation
    m_0, inv2 = 20, pow(2, P - 2, P)

    def blind_round(i):
        prefix = sum(o.eval_univariate(g[j], alphas[j]) for j in range(i)) % P
        suffix = sum(g[j][0] + sum(g[j]) for j in range(i + 1, m_0)) % P
        pm = pow(2, m_0 - 1 - i, P)
        cst = (pm * prefix + pm * inv2 % P * suffix) % P
        return [(pm * g[i][0] + cst) % P] + [pm * c % P for c in g[i][1:]]

    r_cands = []
    for i in range(m_0):
        bl = blind_round(i)
        rt = roots([(zk[i][d] - rho * bl[d]) % P for d in range(4)])
        r_cands.append([(1 - x0) * pow((1 - 2 * x0) % P, P - 2, P) % P for x0 in rt if (1 - 2 * x0) % P])
    ce = pr["claimed_evaluations_raw"]
    f_sums = [int.from_bytes(ce[8 + 32 * j:40 + 32 * j], "little") for j in range(3)]
    target = (o.eval_univariate(zk[19], alphas[19]) - rho * pr["blind_sums"][0]) * pow((f_sums[0] * f_sums[1] - f_sums[2]) % P, P - 2, P) % P
    hits = []

    def dfs(i, prod, choice):
        if i == m_0:
            if prod == target:
                hits.append(list(choice))
            return
        for r in r_cands[i]:
            choice.append(r)
            dfs(i + 1, prod * ((r * alphas[i] + (1 - r) * (1 - alphas[i])) % P) % P, choice)
            choice.pop()

    dfs(0, 1, [])
    assert len(hits) == 1
    rand = hits[0]
    # blinding WHIR
    a_f, a_g = pr["commit_h"]["ood"]
    d = pgcd([(F[0] - a_f) % P] + F[1:], [(G[0] - a_g) % P] + G[1:])
    z_h = (-d[0]) * pow(d[1], P - 2, P) % P
    b_h, r_a = T.first_fold(pr)
    init, blk1 = h["initial_sumcheck"], r0["sumcheck"]
    s_f, s_g = pr["blind_sums"]
    gamma0 = ((init[0][0] + init[0][1]) - (a_f + b_h * a_g)) * pow((s_f + b_h * s_g) % P, P - 2, P) % P
    cp, fconst, ood0 = h["final_answers"][0], h["final_coeffs"][0], r0["ood"][0]
    w32 = pow(o.root_of_unity(9), 16, P)
    pts = [pow(w32, i, P) for i in r0["multipath"][3]]
    folds = [o.eval_univariate(cp, y) for y in pts]
    last_a = o.eval_univariate(quad_from_evals(*init[3]), r_a[3])
    z_c = roots([(cp[0] - ood0) % P] + cp[1:])
    g1_c = roots([(last_a + ood0 - (blk1[0][0] + blk1[0][1])) % P] + folds)
    sol = []
    for combo in itertools.product(*[T.quad_link(blk1[i - 1], blk1[i]) for i in range(1, 4)]):
        r123 = list(combo)
        r4 = (fconst - o.eval_multilinear_coeffs(cp[:8], r123)) * pow(o.eval_multilinear_coeffs(cp[8:], r123), P - 2, P) % P
        big_r = (r_a + r123 + [r4])[::-1]
        last = o.eval_univariate(quad_from_evals(*blk1[3]), r4)
        for zp, g1 in itertools.product(z_c, g1_c):
            value = o.eq_poly_outside(o.expand_from_univariate(z_h, 8), big_r[:8]) + gamma0 * h["deferred"][0]
            gp = 1
            for pt in [zp] + pts:
                value += gp * o.eq_poly_outside(o.expand_from_univariate(pt, 4), big_r[:4])
                gp = gp * g1 % P
            if last == value % P * fconst % P:
                sol.append((r123 + [r4], zp, g1))
    assert len(sol) == 1
    r_b, z_r0, gamma1 = sol[0]
    out = {
        "_comment": "challenge values of the reference-produced proof recovered by algebra (tests/golden/make_challenges.py); "
                    "canonical integers, hex; listed in transcript order. Absorbed inputs between them are the proof bytes "
                    "(tests/fixture_walk.py gives the layout).",
        "rand_r_1..20 (20 consecutive squeezes after commit(W))": [hx(v) for v in rand],
        "blinding_commit_ood_point (after absorbing root_H)": hx(z_h),
        "blinding_commit_batching_randomness (after absorbing the 2 OOD answers)": hx(b_h),
        "rho (after absorbing sum_g)": hx(rho),
        "zk_sumcheck_alpha_1..20 (each after absorbing the 4 coefficients of its round)": [hx(v) for v in alphas],
        "blinding_whir_initial_combination_randomness (after absorbing the 2 polynomial sums)": hx(gamma0),
        "blinding_whir_initial_folding_randomness_1..4 (each after absorbing 3 evaluations)": [hx(v) for v in r_a],
        "blinding_whir_round0_ood_point (after absorbing the round-0 root)": hx(z_r0),
        "blinding_whir_round0_combination_randomness (after the PoW and STIR squeezes)": hx(gamma1),
        "blinding_whir_round0_folding_randomness_1..4": [hx(v) for v in r_b],
    }
    path = os.path.join(HERE, "poseidon-1000.challenges.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
