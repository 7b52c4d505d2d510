"""GPU parity: every sm_100a kernel, called through the C-ABI (libpkwhir.so), against the CPU oracle on
the same seeded inputs — bit-exact (integer arithmetic).  Run on the B200 box: pytest -m gpu."""
import ctypes

import numpy as np
import pytest

from helpers import arr_to_ints, from_mont, ints_to_arr, ptr, rand_fr, to_mont
from oracle import pyref as o

pytestmark = pytest.mark.gpu
P = o.P
sz = ctypes.c_size_t


@pytest.fixture(scope="module")
def ctx():
    import provekit_b200 as pk
    c = pk.Context(0)
    yield c
    c.close()


def rng_fr(seed, n):
    return rand_fr(np.random.default_rng(seed), n)


# ---- Skyscraper ------------------------------------------------------------------------------------
def test_compress_many_kats(ctx):
    """reference.rs:153-188 KATs through the CompressManyFn seam: compress = permute(l,r).0 + l."""
    cases = [(0, 0), (50417215636675310123686652273432694184389644587803328798109154235492038730484,
                      14620920779025509970947930308416120371903474543120179490887326852503500806990)]
    exp_l = [5793276905781313965269111743763131906666794041798623267477617572701829069290,
             8412949970293910117511617126618515787729842528183672400383899220234743146062]
    msgs = b"".join(l.to_bytes(32, "little") + r.to_bytes(32, "little") for l, r in cases)
    out = ctx.compress_many(msgs)
    for i, (l, _) in enumerate(cases):
        assert int.from_bytes(out[32 * i:32 * i + 32], "little") == (exp_l[i] + l) % P


@pytest.mark.parametrize("n", [1, 3, 128, 1000, 70001])
def test_compress_many_vs_oracle(ctx, orc, n):
    rng = np.random.default_rng(n)
    msgs = rng.integers(0, 1 << 64, size=(n, 8), dtype=np.uint64)  # arbitrary 256-bit values (>= p allowed)
    edge = [0, (1 << 256) - 1, P - 1, P, 2 * P, P + 1, 5 * P, 5 * P + 12345, 1 << 255]
    # skyscraper/core/src/reduce.rs:86-96,112-124 (test_reduce_partial_max): low limbs all ones, top limb = top limb of k*p + 1
    edge += [((((k * P) >> 192) + 1) << 192) | ((1 << 192) - 1) for k in range(6)]
    for i, e in enumerate(edge[:n]):
        msgs[i, :4] = ints_to_arr([e])[0]
        msgs[i, 4:] = ints_to_arr([edge[-1 - i]])[0]
    exp = np.zeros((n, 4), np.uint64)
    orc.orc_sky_compress_many(ptr(msgs), ptr(exp), sz(n), 2)
    assert ctx.compress_many(msgs.tobytes()) == exp.tobytes()


def test_compress_many_empty_and_ragged(ctx):
    assert ctx.compress_many(b"") == b""
    with pytest.raises(ValueError):
        ctx.compress_many(b"\0" * 65)  # generic.rs:19 "Message length not a multiple of 64"


@pytest.mark.parametrize("bits", [0.0, 3.141592653589793, 10.0, 17.5])
def test_pow_solve_vs_oracle(ctx, orc, bits):
    for ch in ([(1 << 64) - 1] * 4, [5, 6, 7, 8]):  # pow.rs:105-111 uses challenge = [u64::MAX; 4]
        cha = np.array(ch, dtype=np.uint64)
        nonce = ctx.pow_solve(cha, bits)
        assert nonce == orc.orc_pow_solve(ptr(cha), bits)
        assert orc.orc_pow_verify(ptr(cha), bits, ctypes.c_uint64(nonce)) == 1


def test_pow_bits_out_of_range(ctx):
    from provekit_b200 import PkError
    with pytest.raises(PkError):
        ctx.pow_solve(np.zeros(4, np.uint64), 60.0)  # provekit/common/src/skyscraper/pow.rs:16


# ---- wavelet / RS encode -----------------------------------------------------------------------------
@pytest.mark.parametrize("log_n", [0, 1, 4, 11, 12, 16, 19])
def test_wavelet_vs_oracle(ctx, orc, log_n):
    a = rng_fr(log_n, 1 << log_n)
    exp = a.copy()
    orc.orc_evals_to_coeffs(ptr(exp), log_n)
    buf = ctx.upload(a)
    ctx.evals_to_coeffs(buf, log_n)
    got = buf.download()
    assert np.array_equal(got, exp)
    ctx.coeffs_to_evals(buf, log_n)
    assert np.array_equal(buf.download(), a)  # round trip
    buf.free()


# column length 2^L, L = log_n - 4.  kernel "r8": the TMA-staged radix-8 kernel for every L >= 7 — one pass (L <= 9: 11, 12, 13),
# two (17 -> 7+6, 18 -> 7+7, 20 -> 8+8, 21 -> 9+8, 22 -> 9+9) or three (23 -> 7+6+6); L < 7 always runs the radix-2 kernel.
# kernel "radix2": the register-staged radix-2 kernel everywhere.  The product picks r8 for L = 17, 18 and radix-2 otherwise
# (pkwhir.cu rs_encode_raw), so every shape below is covered on the kernel the product uses and on the other one.
@pytest.mark.parametrize("kernel", ["r8", "radix2"])
@pytest.mark.parametrize("log_n,rate", [(4, 1), (4, 4), (5, 13), (9, 10), (10, 3), (11, 1), (11, 0), (12, 2), (13, 7), (14, 1), (17, 4),
                                        (18, 1), (20, 1), (21, 1), (22, 1), (23, 1)])
def test_rs_encode_vs_oracle(ctx, orc, log_n, rate, kernel, monkeypatch):
    monkeypatch.setenv("PK_NTT_KERNEL", kernel)
    a = rng_fr(1000 + log_n, 1 << log_n)
    rows = 1 << (log_n + rate - 4)
    exp = np.zeros((rows * 16, 4), np.uint64)
    orc.orc_rs_encode(ptr(a), log_n, rate, 4, ptr(exp), sz(16), sz(0))
    coeffs = ctx.upload(a)
    leaves = ctx.buffer(rows * 16)
    ctx.rs_encode(coeffs, log_n, rate, leaves)
    assert np.array_equal(leaves.download(), exp)
    coeffs.free()
    leaves.free()


def test_rs_encode_reference_proof_kat(ctx):
    """SURVEY A.6 on the GPU: RS-encode the 32 coefficients recovered from the reference-produced proof
    and compare the 9 opened leaves byte for byte."""
    import fixture_walk as fw
    w = fw.walk_proof()["whir_w"]
    leaves, idx = w["final_answers"], w["final_multipath"][3]
    g16 = pow(o.root_of_unity(18), 16, P)
    y = [pow(g16, i, P) for i in idx]
    inv = pow((y[1] - y[0]) % P, -1, P)
    hi = [(leaves[1][k] - leaves[0][k]) * inv % P for k in range(16)]
    lo = [(leaves[0][k] - hi[k] * y[0]) % P for k in range(16)]
    coeffs = ctx.upload(to_mont(lo + hi))
    out = ctx.buffer((1 << 14) * 16)
    ctx.rs_encode(coeffs, 5, 13, out)
    got = out.download()
    for i, leaf in zip(idx, leaves):
        assert from_mont(got[i * 16:(i + 1) * 16]) == leaf


# ---- Merkle / commit / open ----------------------------------------------------------------------------
@pytest.mark.parametrize("L,w", [(2, 16), (8, 32), (512, 16), (1024, 32), (4096, 1), (1 << 14, 16)])
def test_merkle_vs_oracle(ctx, orc, L, w):
    a = rng_fr(L + w, L * w)
    exp = np.zeros((2 * L, 4), np.uint64)
    orc.orc_merkle_build(ptr(a), sz(L), sz(w), ptr(exp), 2)
    exp_canon = np.zeros_like(exp)
    orc.orc_from_montgomery(ptr(exp), ptr(exp_canon), sz(2 * L))
    leaves, nodes = ctx.upload(a), ctx.buffer(2 * L)
    ctx.merkle_build(leaves, L, w, nodes)
    got = nodes.download()
    assert np.array_equal(got[1:], exp_canon[1:])
    leaves.free()
    nodes.free()


def test_merkle_errors(ctx):
    from provekit_b200 import PkError
    leaves, nodes = ctx.buffer(64), ctx.buffer(16)
    with pytest.raises(PkError) as e:
        ctx.merkle_build(leaves, 4, 0, nodes)  # empty leaf: IncorrectInputLength(0), skyscraper/whir.rs:47
    assert e.value.code == -6
    with pytest.raises(PkError):
        ctx.merkle_build(leaves, 3, 16, nodes)


@pytest.mark.parametrize("log_n,rate,batch", [(8, 1, 2), (12, 1, 2), (13, 4, 1), (16, 1, 2)])
def test_commit_batch_and_open(ctx, orc, log_n, rate, batch):
    polys = [rng_fr(7 * log_n + b, 1 << log_n) for b in range(batch)]
    L = 1 << (log_n + rate - 4)
    w = 16 * batch
    leaves = np.zeros((L * w, 4), np.uint64)
    for b in range(batch):
        orc.orc_rs_encode(ptr(polys[b]), log_n, rate, 4, ptr(leaves), sz(w), sz(16 * b))
    nodes = np.zeros((2 * L, 4), np.uint64)
    orc.orc_merkle_build(ptr(leaves), sz(L), sz(w), ptr(nodes), 2)
    bufs = [ctx.upload(p) for p in polys]
    cm = ctx.commit_batch(bufs, log_n, rate)
    assert (cm.num_leaves, cm.leaf_width) == (L, w)
    assert np.array_equal(cm.root, nodes[1])  # Montgomery root
    rng = np.random.default_rng(log_n)
    idx = sorted(set(int(i) for i in rng.integers(0, L, size=40)) | {0, L - 1})
    got_leaves, sib, pre, sufs = cm.open(idx)
    nodes_int = from_mont(nodes)
    e_sib, e_pre, e_suf, _ = o.merkle_multipath(nodes_int, idx)
    assert arr_to_ints(sib) == e_sib
    assert [int(x) for x in pre] == e_pre
    assert [arr_to_ints(s) if len(s) else [] for s in sufs] == e_suf
    lv = leaves.reshape(L, w, 4)
    for j, i in enumerate(idx):
        assert np.array_equal(got_leaves[j], lv[i])
    # and the opened paths verify against the root like the Go verifier does (whir_utilities.go:13-46)
    prev = []
    root = nodes_int[1]
    for j, i in enumerate(idx[:5]):
        prev = prev[:e_pre[j]] + e_suf[j]
        assert o.merkle_verify_path(root, i, from_mont(got_leaves[j]), e_sib[j], prev)
    cm.free()
    for b in bufs:
        b.free()


@pytest.mark.parametrize("log_n,shards", [(8, 1), (12, 2), (12, 8), (16, 4)])
def test_sharded_opening_on_one_device(ctx, log_n, shards):
    """The building blocks of the sharded opening against pk_commit_open on the whole tree: every "rank" opens its
    rows on a view of its leaf block and sub-tree (pk_commit_wrap + pk_commit_open_paths, local indexes), the levels above
    the sub-trees are hashed from the sub-roots, pk_multipath_build compresses once.  Ranks are slices of one device here; the
    multi-process form over CUDA IPC is tests/test_gpu_sharded.py."""
    polys = [ctx.upload(rng_fr(31 * log_n + b, 1 << log_n)) for b in range(2)]
    cm = ctx.commit_batch(polys, log_n, 1)
    L, w = cm.num_leaves, cm.leaf_width
    per = L // shards
    rng = np.random.default_rng(shards)
    idx = np.array(sorted(set(int(i) for i in rng.integers(0, L, size=60)) | {0, 1, L // 2, L - 1}), dtype=np.uint64)
    exp = cm.open(idx)
    # one-shot: uncompressed paths of the whole tree -> the same MultiPath
    rows, paths = cm.open_paths(idx)
    assert np.array_equal(rows, exp[0])
    sib, pre, sufs = ctx.multipath_build(paths)
    assert np.array_equal(sib, exp[1]) and np.array_equal(pre, exp[2]) and all(np.array_equal(a, b) for a, b in zip(sufs, exp[3]))
    # sharded: rebuild every rank's leaf block and sub-tree, open per rank, combine
    leaves = np.zeros((L * w, 4), np.uint64).reshape(L, w, 4)
    for b, poly in enumerate(polys):
        blk = ctx.buffer(L * 16)
        ctx.rs_encode(poly, log_n, 1, blk, 16, 0)
        leaves[:, 16 * b:16 * b + 16] = blk.download().reshape(L, 16, 4)
        blk.free()
    sub_roots, all_rows, all_paths, owners = [], [], [], []
    for g in range(shards):
        lb, nb = ctx.upload(leaves[g * per:(g + 1) * per].reshape(-1, 4)), ctx.buffer(2 * per)
        ctx.merkle_build(lb, per, w, nb)
        sub_roots.append(nb.download(1, 1)[0])
        view = ctx.commit_wrap(lb, nb, per, w)
        mine = idx[(idx // np.uint64(per)) == g] - np.uint64(g * per)
        if len(mine):
            r_, p_ = view.open_paths(mine)
            all_rows.append(r_)
            all_paths.append(p_)
            owners += [g] * len(mine)
        view.free()
        lb.free()
        nb.free()
    levels = [np.array(sub_roots, dtype=np.uint64)]
    while len(levels[-1]) > 1:
        levels.append(np.frombuffer(ctx.compress_many(np.ascontiguousarray(levels[-1]).tobytes()), dtype=np.uint64).reshape(-1, 4).copy())
    top = np.array([[levels[j][(g >> j) ^ 1] for j in range(len(levels) - 1)] for g in owners], dtype=np.uint64).reshape(len(idx), -1, 4)
    full = np.concatenate([np.concatenate(all_paths, axis=0), top], axis=1)
    sib, pre, sufs = ctx.multipath_build(full)
    assert np.array_equal(np.concatenate(all_rows, axis=0), exp[0])
    assert np.array_equal(sib, exp[1]) and np.array_equal(pre, exp[2]) and all(np.array_equal(a, b) for a, b in zip(sufs, exp[3]))
    cm.free()
    for b in polys:
        b.free()


def test_commit_open_rejects_unsorted(ctx):
    from provekit_b200 import PkError
    b = ctx.upload(rng_fr(1, 256))
    cm = ctx.commit_batch([b], 8, 1)
    with pytest.raises(PkError):
        cm.open([3, 3])
    with pytest.raises(PkError):
        cm.open([cm.num_leaves])
    cm.free()
    b.free()


# ---- helpers -----------------------------------------------------------------------------------------
@pytest.mark.parametrize("log_n", [0, 1, 5, 13, 18])
def test_univariate_dot_axpy_fold_mle(ctx, orc, log_n):
    n = 1 << log_n
    a, b = rng_fr(log_n, n), rng_fr(log_n + 50, n)
    z = rng_fr(99, 1)
    exp = np.zeros(4, np.uint64)
    da, db = ctx.upload(a), ctx.upload(b)
    orc.orc_eval_univariate(ptr(a), sz(n), ptr(z), ptr(exp))
    assert np.array_equal(ctx.eval_univariate(da, n, z), exp)
    orc.orc_dot(ptr(a), ptr(b), sz(n), ptr(exp))
    assert np.array_equal(ctx.dot(da, db, n), exp)
    pt = rng_fr(7, max(log_n, 1))[:log_n]
    orc.orc_mle_eval(ptr(a), log_n, ptr(pt), ptr(exp))
    assert np.array_equal(ctx.mle_eval(da, log_n, pt), exp)
    if log_n >= 4:
        r = rng_fr(8, 4)
        ef = np.zeros((n // 16, 4), np.uint64)
        orc.orc_fold_coeffs(ptr(a), log_n, ptr(r), 4, ptr(ef))
        out = ctx.buffer(n // 16)
        ctx.fold_coeffs(da, log_n, r, out)
        assert np.array_equal(out.download(), ef)
        out.free()
    ea = a.copy()
    orc.orc_axpy(ptr(ea), ptr(b), ptr(z), sz(n))
    ctx.axpy(da, db, z, n)
    assert np.array_equal(da.download(), ea)
    da.free()
    db.free()


@pytest.mark.parametrize("n,k", [(0, 1), (1, 2), (4, 3), (9, 29), (13, 110), (17, 5)])
def test_eval_eq_batch(ctx, orc, n, k):
    N = 1 << n
    base = rng_fr(n + k, N)
    pts = rng_fr(3 * n + 1, max(k * n, 1))[:k * n]
    sc = rng_fr(5, k)
    exp = base.copy()
    for j in range(k):
        orc.orc_eval_eq_accumulate(ptr(np.ascontiguousarray(pts[j * n:(j + 1) * n])), n, ptr(np.ascontiguousarray(sc[j])), ptr(exp))
    out = ctx.upload(base)
    ctx.eval_eq(pts, sc, n, out)
    assert np.array_equal(out.download(), exp)
    out.free()


# ---- sumchecks ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("log_n", [1, 2, 3, 8, 15, 19])
def test_zk_sumcheck_all_rounds(ctx, orc, log_n):
    n = 1 << log_n
    host = [rng_fr(10 * log_n + k, n) for k in range(4)]
    dev = [ctx.upload(h) for h in host]
    fold = None
    cur = log_n
    for rnd in range(log_n):
        exp = np.zeros((3, 4), np.uint64)
        orc.orc_zk_sumcheck_round(ptr(host[0]), ptr(host[1]), ptr(host[2]), ptr(host[3]), cur, ptr(fold) if fold is not None else None, ptr(exp))
        got = ctx.sumcheck_fold_map_reduce(*dev, cur, fold)
        assert np.array_equal(got, exp), f"round {rnd}"
        if fold is not None:
            cur -= 1
        if rnd in (1, log_n - 1):
            for k in range(4):
                assert np.array_equal(dev[k].download(1 << cur), host[k][:1 << cur])
        fold = rng_fr(1000 + rnd, 1)
    for d in dev:
        d.free()


def test_zk_sumcheck_asserts(ctx):
    from provekit_b200 import PkError
    bufs = [ctx.buffer(4) for _ in range(4)]
    with pytest.raises(PkError):
        ctx.sumcheck_fold_map_reduce(*bufs, 0, None)            # sumcheck.rs:23 size >= 2
    with pytest.raises(PkError):
        ctx.sumcheck_fold_map_reduce(*bufs, 1, rng_fr(1, 1))    # sumcheck.rs:27 size >= 4 when folding
    with pytest.raises(PkError):
        ctx.sumcheck_fold_map_reduce(*bufs, 3, None)            # arrays shorter than 2^log_n
    for b in bufs:
        b.free()


@pytest.mark.parametrize("log_n", [1, 2, 5, 12, 19])
def test_whir_sumcheck_all_rounds(ctx, orc, log_n):
    n = 1 << log_n
    hp, hw = rng_fr(log_n, n), rng_fr(log_n + 77, n)
    bufs = [(ctx.upload(hp), ctx.upload(hw)), (ctx.buffer(max(n // 2, 1)), ctx.buffer(max(n // 2, 1)))]
    fold = None
    cur, which = log_n, 0
    for rnd in range(log_n):
        exp = np.zeros((3, 4), np.uint64)
        orc.orc_whir_sumcheck_round(ptr(hp), ptr(hw), cur, ptr(fold) if fold is not None else None, ptr(exp))
        src, dst = bufs[which], bufs[1 - which]
        if fold is None:
            got = ctx.whir_sumcheck_round(src[0], src[1], cur)
        else:
            got = ctx.whir_sumcheck_round(src[0], src[1], cur, fold, dst[0], dst[1])
            which = 1 - which
            cur -= 1
        assert np.array_equal(got, exp), f"round {rnd}"
        if rnd == log_n - 1 or rnd == 1:
            assert np.array_equal(bufs[which][0].download(1 << cur), hp[:1 << cur])
            assert np.array_equal(bufs[which][1].download(1 << cur), hw[:1 << cur])
        fold = rng_fr(2000 + rnd, 1)
    for a, b in bufs:
        a.free()
        b.free()


@pytest.mark.parametrize("log_n", [3, 12, 17])
def test_batched_helpers(ctx, orc, log_n):
    """pk_eval_univariate_batch / pk_multi_dot / pk_mle_eval_batch == the single-array oracle calls."""
    n = 1 << log_n
    arrs = [rng_fr(100 * log_n + k, n) for k in range(5)]
    dev = [ctx.upload(a) for a in arrs]
    z = rng_fr(5, 1)
    exp = np.zeros(4, np.uint64)
    got = ctx.eval_univariate_batch(dev[:2], n, z)
    for j in range(2):
        orc.orc_eval_univariate(ptr(arrs[j]), sz(n), ptr(z), ptr(exp))
        assert np.array_equal(got[j], exp)
    for na in (3, 1):
        got = ctx.multi_dot(dev[:na], dev[3:5], n)
        for ja in range(na):
            for jb in range(2):
                orc.orc_dot(ptr(arrs[ja]), ptr(arrs[3 + jb]), sz(n), ptr(exp))
                assert np.array_equal(got[ja, jb], exp)
    pt = rng_fr(9, log_n)
    got = ctx.mle_eval_batch(dev[:3], log_n, pt)
    for j in range(3):
        orc.orc_mle_eval(ptr(arrs[j]), log_n, ptr(pt), ptr(exp))
        assert np.array_equal(got[j], exp)
    for d in dev:
        d.free()


def test_reductions_repeatable(ctx, orc):
    """The single-launch grid reduction resets its ticket counter: many back-to-back reductions stay exact."""
    n = 1 << 16
    a, b = rng_fr(1, n), rng_fr(2, n)
    da, db = ctx.upload(a), ctx.upload(b)
    exp = np.zeros(4, np.uint64)
    orc.orc_dot(ptr(a), ptr(b), sz(n), ptr(exp))
    for _ in range(50):
        assert np.array_equal(ctx.dot(da, db, n), exp)
    da.free()
    db.free()


@pytest.mark.parametrize("log_n,rate,n_cols,n_peers", [(12, 1, 16, 1), (12, 1, 8, 2), (14, 1, 4, 4), (13, 4, 4, 8), (16, 1, 8, 4),
                                                       (18, 1, 4, 8), (8, 1, 16, 2)])
def test_rs_encode_sharded_layout(ctx, orc, log_n, rate, n_cols, n_peers):
    """Sharded commit building block on ONE GPU: column subsets + row blocks into separate 'peer' buffers must
    reassemble to exactly the single-GPU codeword (the multi-process version only swaps local pointers for IPC ones)."""
    a = rng_fr(4242 + log_n, 1 << log_n)
    rows = 1 << (log_n + rate - 4)
    w = 32  # batched layout: this polynomial occupies columns 16..31
    exp = np.zeros((rows * w, 4), np.uint64)
    orc.orc_rs_encode(ptr(a), log_n, rate, 4, ptr(exp), sz(w), sz(16))
    coeffs = ctx.upload(a)
    per = rows // n_peers
    blocks = [ctx.buffer_shared(per * w).zero() for _ in range(n_peers)]
    ptrs = [b.device_ptr for b in blocks]
    for c0 in range(0, 16, n_cols):
        ctx.rs_encode_sharded(coeffs, log_n, rate, c0, n_cols, ptrs, w, 16)
    got = np.concatenate([b.download() for b in blocks])
    assert np.array_equal(got, exp)
    for b in blocks:
        b.free()
    coeffs.free()


def test_merkle_combine_roots(ctx, orc):
    L, w, G = 1 << 10, 16, 4
    a = rng_fr(77, L * w)
    nodes = np.zeros((2 * L, 4), np.uint64)
    orc.orc_merkle_build(ptr(a), sz(L), sz(w), ptr(nodes), 2)
    canon = np.zeros_like(nodes)
    orc.orc_from_montgomery(ptr(nodes), ptr(canon), sz(2 * L))
    sub = canon[G:2 * G]          # heap order: level with G nodes = the G sub-tree roots, left to right
    assert np.array_equal(ctx.merkle_combine_roots(sub), canon[1])
    # and each sub-tree root equals a local tree over that rank's rows
    per = L // G
    for r in range(G):
        leaves, nd = ctx.upload(a[r * per * w:(r + 1) * per * w]), ctx.buffer(2 * per)
        ctx.merkle_build(leaves, per, w, nd)
        assert np.array_equal(nd.download(1, 1)[0], sub[r])
        leaves.free()
        nd.free()


@pytest.mark.parametrize("log_d,n,k", [(18, 17, 109), (14, 13, 81), (13, 12, 300), (12, 12, 401), (17, 13, 300), (14, 13, 40), (9, 5, 31)])
def test_eval_eq_roots_equals_per_point_method(ctx, log_d, n, k):
    """pk_eval_eq_roots_batch (STIR constraint weights as M^T of a sparse DFT, or its small-batch fallback) == pk_eval_eq_batch
    on the expanded points z_k = omega_D^e_k -> (z^(2^(n-1)), .., z^2, z) (ExpandFromUnivariate,
    recursive-verifier/app/utilities/utilities.go:182-190), bit for bit; duplicates among the e_k must add up."""
    ROOT28 = 19103219067921713944291392827692070036145651957329286315305642004821462161904  # arkworks TWO_ADIC_ROOT_OF_UNITY
    w = pow(ROOT28, 1 << (28 - log_d), P)
    rng = np.random.default_rng(1000 * log_d + n)
    exps = rng.integers(0, 1 << log_d, size=k, dtype=np.uint64)
    exps[1] = exps[0]  # a repeated point
    scal = rng_fr(77 + k, k)
    pts = []
    for e in exps:
        z = pow(w, int(e), P)
        col = []
        for _ in range(n):
            col.append(z)
            z = z * z % P
        pts.extend(reversed(col))
    base = rng_fr(5, 1 << n)
    a, b = ctx.upload(base), ctx.upload(base)
    ctx.eval_eq(to_mont(pts), scal, n, a)
    ctx.eval_eq_roots(exps, log_d, n, scal, b)
    assert np.array_equal(a.download(), b.download())
    a.free()
    b.free()


@pytest.mark.parametrize("log_n,n_prefix", [(12, 1), (12, 2500), (12, 4096), (17, 70001)])
def test_mle_eval_prefix_equals_full_on_zero_extended_arrays(ctx, log_n, n_prefix):
    """pk_mle_eval_batch_prefix reads only the first n_prefix elements; on arrays that are zero beyond them (the
    zero-extended R1CS weight vectors, whir_r1cs.rs:382-412) it must equal pk_mle_eval_batch; errors on a bad prefix."""
    from provekit_b200 import PkError
    n = 1 << log_n
    arrs = []
    for k in range(3):
        a = rng_fr(300 + 10 * log_n + k, n)
        a[n_prefix:] = 0
        arrs.append(a)
    dev = [ctx.upload(a) for a in arrs]
    point = rng_fr(9, log_n)
    full = ctx.mle_eval_batch(dev, log_n, point)
    assert np.array_equal(ctx.mle_eval_batch_prefix(dev, log_n, n_prefix, point), full)
    # poison the tail: the prefix variant must not look at it
    for d, a in zip(dev, arrs):
        b = a.copy()
        b[n_prefix:] = rng_fr(1, n - n_prefix) if n_prefix < n else b[n_prefix:]
        d.upload(b)
    assert np.array_equal(ctx.mle_eval_batch_prefix(dev, log_n, n_prefix, point), full)
    for bad in (0, n + 1):
        with pytest.raises(PkError):
            ctx.mle_eval_batch_prefix(dev, log_n, bad, point)
    for d in dev:
        d.free()
