"""Deterministic synthetic workloads for bench.py (harness code, not the product).

`synth_r1cs` builds a satisfiable R1CS in the reference's interned-CSR form with prescribed dimensions
and per-matrix non-zero counts (config 1/2 uses the shapes of the reference fixture
poseidon-1000.nps — 729 560 constraints, 860 637 witnesses, nnz 740 508 / 609 440 / 1 915 568, 366
interned constants).  Row i of C carries a dedicated product witness (K + i, 1) so that any assignment
of the K free witnesses extends to a satisfying one; surplus C entries point at a witness fixed to 0.
"""
import numpy as np

P = 21888242871839275222246405745257275088548364400416034343698204186575808495617
R = (1 << 256) % P

# `structure` reproduces the fixture's sparsity extremes as well (pk_nps_read_r1cs on poseidon-1000.nps, session 4): A has
# two rows of 65 536 entries and two columns used 65 537 times, B one column used 131 076 times, C one column used
# 598 479 times; 120 122 / 120 120 rows of A / B are empty.  These drive the long-row and hot-column SpMV kernels.
POSEIDON_1000 = dict(num_constraints=729_560, num_witnesses=860_637, nnz=(740_508, 609_440, 1_915_568),
                     n_interned=366,
                     structure=dict(long_rows=((2, 65_536), ()), hot_cols=(((65_537, 65_537)), (131_076,)), c_hot=598_479))


def log2ceil(n: int) -> int:
    return max(0, (n - 1).bit_length())


def ints_to_limbs(xs) -> np.ndarray:
    o = np.asarray(xs, dtype=object)
    mask = (1 << 64) - 1
    out = np.empty((len(o), 4), dtype=np.uint64)
    for k in range(4):
        out[:, k] = ((o >> (64 * k)) & mask).astype(np.uint64)
    return out


def to_mont(xs) -> np.ndarray:
    o = np.asarray(xs, dtype=object)
    return ints_to_limbs((o * R) % P)


def rand_ints(rng, n):
    """n python ints < p (object array)"""
    limbs = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64).astype(object)
    v = limbs[:, 0] | (limbs[:, 1] << 62) | (limbs[:, 2] << 124) | (limbs[:, 3] << 186)
    return v % P


def rand_fr(rng, n) -> np.ndarray:
    """n field elements < p as (n,4) uint64 (uniform enough for masks; valid Montgomery representations)"""
    a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64(0x1FFFFFFFFFFFFFFF)
    return a


def _row_counts(rng, rows, nnz, minimum=0):
    base, extra = divmod(nnz, rows)
    cnt = np.full(rows, base, dtype=np.int64)
    if extra:
        cnt[rng.choice(rows, size=extra, replace=False)] += 1
    assert cnt.min() >= minimum
    return cnt


def synth_r1cs(num_constraints, num_witnesses, nnz, n_interned=366, seed=1, structure=None):
    rng = np.random.default_rng(seed)
    nc, nw = num_constraints, num_witnesses
    K = nw - nc
    assert K >= 2
    interned = rand_ints(rng, n_interned)
    interned[0] = 1
    free = rand_ints(rng, K)
    free[0], free[1] = 1, 0
    mats = []
    az = []
    st = structure or {}
    for which, n_entries in enumerate(nnz[:2]):
        long_spec = (st.get("long_rows") or ((), ()))[which]
        n_long, long_len = long_spec if long_spec else (0, 0)
        cnt = _row_counts(rng, nc, n_entries - n_long * long_len)
        if n_long:  # a few very long rows (the fixture's lookup-sum constraints); the rest of the entries stay spread out
            rows = rng.choice(nc, size=n_long, replace=False)
            cnt[rows] += long_len
        row_start = np.concatenate([[0], np.cumsum(cnt)[:-1]]).astype(np.uint64)
        col = rng.integers(0, K, size=n_entries, dtype=np.uint32)
        hot = (st.get("hot_cols") or ((), ()))[which]
        if hot:  # columns used by tens of thousands of rows: positions drawn among the short rows' entries
            short = np.flatnonzero(np.repeat(cnt, cnt) <= 64)
            picks = rng.choice(short, size=int(sum(hot)), replace=False)
            off = 0
            for h_i, uses in enumerate(hot):
                col[picks[off:off + uses]] = 2 + h_i
                off += uses
        val = rng.integers(0, n_interned, size=n_entries, dtype=np.uint32)
        prod = interned[val] * free[col]
        sums = np.zeros(nc, dtype=object)
        nz = cnt > 0
        if n_entries:
            red = np.add.reduceat(prod, row_start[nz].astype(np.int64))
            sums[nz] = red
        az.append(sums % P)
        mats.append((row_start, col, val))
    cnt = _row_counts(rng, nc, nnz[2], minimum=1)
    row_start = np.concatenate([[0], np.cumsum(cnt)[:-1]]).astype(np.uint64)
    col = np.ones(nnz[2], dtype=np.uint32)                       # surplus entries hit the zero witness ...
    val = rng.integers(0, n_interned, size=nnz[2], dtype=np.uint32)
    first = row_start.astype(np.int64)
    surplus = np.ones(nnz[2], dtype=bool)
    surplus[first] = False
    if st.get("c_hot") is not None:                              # ... or, with `structure`, one hot column and random free ones
        idx = np.flatnonzero(surplus)
        spread = rng.choice(idx, size=max(0, len(idx) - int(st["c_hot"])), replace=False)
        col[spread] = rng.integers(0, K, size=len(spread), dtype=np.uint32)
    col[first] = (K + np.arange(nc)).astype(np.uint32)
    val[first] = 0                                               # interned[0] = 1
    mats.append((row_start, col, val))
    # the product witness of row i absorbs the surplus terms of C's row i: w = Az * Bz - sum(surplus coefficients * z)
    extra = np.where(surplus, interned[val] * free[np.minimum(col, K - 1)], 0)
    extra_row = np.add.reduceat(extra, first) % P
    witness = np.concatenate([free, (az[0] * az[1] - extra_row) % P])
    return dict(num_constraints=nc, num_witnesses=nw, interned=to_mont(interned), a=mats[0], b=mats[1], c=mats[2],
                witness=to_mont(witness))


def shapes(r1cs):
    m = log2ceil(r1cs["num_witnesses"]) + 1
    m0 = log2ceil(r1cs["num_constraints"])
    mh = log2ceil(4 * m0) + 1
    return m, m0, mh


def randomness(r1cs, seed=7):
    m, m0, mh = shapes(r1cs)
    rng = np.random.default_rng(seed)
    return dict(mask_w=rand_fr(rng, 1 << (m - 1)), g_w=rand_fr(rng, 1 << m), blind=rand_fr(rng, 4 * m0),
                mask_h=rand_fr(rng, 1 << (mh - 1)), g_h=rand_fr(rng, 1 << mh))


def whir_rounds(num_variables):
    """(domain_log, leaf_width-agnostic) list of the commitments a WHIR opening creates after the first one."""
    fsr = num_variables % 4
    n_rounds = (num_variables - fsr) // 4 - 1
    return [num_variables + 1 - 1 - r for r in range(n_rounds)]  # domain logs of round commitments


def merkle_leaf_bytes(m, mh):
    """Algorithmic bytes of all Merkle leaf-hash launches of one proof: 32*(L*w + L) per tree."""
    total = 0
    launches = 0
    for nv in (m, mh):
        L = 1 << (nv + 1 - 4)
        total += 32 * (L * 32 + L)
        launches += 1
        for dl in whir_rounds(nv):
            L = 1 << (dl - 4)
            total += 32 * (L * 16 + L)
            launches += 1
    return total, launches


def rs_encode_bytes(m, mh):
    """Algorithmic bytes of all RS-encode launches of one proof: 32*(n + D) per polynomial."""
    total = 0
    for nv in (m, mh):
        total += 2 * 32 * ((1 << nv) + (1 << (nv + 1)))
        n = nv
        for dl in whir_rounds(nv):
            n -= 4
            total += 32 * ((1 << n) + (1 << dl))
    return total
