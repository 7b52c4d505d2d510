"""Pure-Python (big-int) restatement of the reference algorithms on the WHIR hot path.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import this module.  The product path (provekit_b200/) never does.

Python integers mod p are unambiguous, so this file is the *root* of the parity chain:
    reference KATs / fixture  ->  pyref (this file)  ->  oracle C library  ->  CUDA kernels.

The conventions of the [EXT] functions below (fold order, eq / MLE index order, coefficient<->evaluation transform,
RS-encode leaf layout, univariate OOD evaluation) are additionally pinned by the reference-produced proof itself,
sponge-free: tests/test_fixture_algebra.py.

Every function cites the reference file:line it follows (paths relative to /root/reference).
Functions whose algorithm lives in an un-vendored dependency are marked [EXT] and cite the
in-tree call site / Go restatement that pins them.
"""
from __future__ import annotations

# --- BN254 scalar field --------------------------------------------------------------------
# skyscraper/block-multiplier/src/constants.rs:3-8 (U64_P)
P = 21888242871839275222246405745257275088548364400416034343698204186575808495617
R = (1 << 256) % P                      # constants.rs:18-23 (U64_R)
R2 = (R * R) % P                        # constants.rs:26-31 (U64_R2)
R_INV = pow(R, -1, P)                   # constants.rs:34-39 (U64_R_INV)
NP0 = (-pow(P, -1, 1 << 64)) % (1 << 64)  # constants.rs:1 (U64_NP0 = 0xc2e1f593efffffff)
# arkworks BN254 Fr TWO_ADIC_ROOT_OF_UNITY (order 2^28); the fixture's domain generator is
# ROOT28^(2^6) for the 2^22 domain.
ROOT28 = 19103219067921713944291392827692070036145651957329286315305642004821462161904
TWO_ADICITY = 28
# provekit/common/src/utils/mod.rs:23-25
HALF = 10944121435919637611123202872628637544274182200208017171849102093287904247809


def root_of_unity(log_n: int) -> int:
    """Generator of the arkworks Radix2EvaluationDomain of size 2^log_n [EXT ark-poly]."""
    assert 0 <= log_n <= TWO_ADICITY
    return pow(ROOT28, 1 << (TWO_ADICITY - log_n), P)


def to_limbs(x: int) -> list[int]:
    return [(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]


def from_limbs(l) -> int:
    return sum(int(v) << (64 * i) for i, v in enumerate(l))


# --- Skyscraper ------------------------------------------------------------------------------
# skyscraper/core/src/constants.rs:32-51
ROUND_CONSTANTS = [from_limbs(l) for l in [
    [0x0000000000000000, 0x0000000000000000, 0x0000000000000000, 0x0000000000000000],
    [0x903c4324270bd744, 0x873125f708a7d269, 0x081dd27906c83855, 0x276b1823ea6d7667],
    [0x7ac8edbb4b378d71, 0xe29d79f3d99e2cb7, 0x751417914c1a5a18, 0x0cf02bd758a484a6],
    [0xfa7adc6769e5bc36, 0x1c3f8e297cca387d, 0x0eb7730d63481db0, 0x25b0e03f18ede544],
    [0x57847e652f03cfb7, 0x33440b9668873404, 0x955a32e849af80bc, 0x002882fcbe14ae70],
    [0x979231396257d4d7, 0x29989c3e1b37d3c1, 0x12ef02b47f1277ba, 0x039ad8571e2b7a9c],
    [0xb5b48465abbb7887, 0xa72a6bc5e6ba2d2b, 0x4cd48043712f7b29, 0x1142d5410fc1fc1a],
    [0x7ab2c156059075d3, 0x17cb3594047999b2, 0x44f2c93598f289f7, 0x1d78439f69bc0bec],
    [0x05d7a965138b8edb, 0x36ef35a3d55c48b1, 0x8ddfb8a1ac6f1628, 0x258588a508f4ff82],
    [0x1596fb9afccb49e9, 0x9a7367d69a09a95b, 0x9bc43f6984e4c157, 0x13087879d2f514fe],
    [0x295ccd233b4109fa, 0xe1d72f89ed868012, 0x2e9e1eea4bc88a8e, 0x17dadee898c45232],
    [0x9a8590b4aa1f486f, 0xb75834b430e9130e, 0xb8e90b1034d5de31, 0x295c6d1546e7f4a6],
    [0x850adcb74c6eb892, 0x07699ef305b92fc3, 0x4ef96a2ba1720f2d, 0x1288ca0e1d3ed446],
    [0x01960f9349d1b5ee, 0x8ccad30769371c69, 0xe5c81e8991c98662, 0x17563b4d1ae023f3],
    [0x6ba01e9476b32917, 0xa1cb0a3add977bc9, 0x86815a945815f030, 0x2869043be91a1eea],
    [0x81776c885511d976, 0x7475d34f47f414e7, 0x5d090056095d96cf, 0x14941f0aff59e79a],
    [0xbc40b4fd8fc8c034, 0xbb7142c3cce4fd48, 0x318356758a39005a, 0x1ce337a190f4379f],
    [0x0000000000000000, 0x0000000000000000, 0x0000000000000000, 0x0000000000000000],
]]
# skyscraper/core/src/reference.rs:22-26 — equals 2^-256 mod p (checked in tests)
SIGMA_INV = 9915499612839321149637521777990102151350674507940716049588462388200839649614


def sbox(v: int) -> int:
    """skyscraper/core/src/reference.rs:96-98 / bar.rs:40-42."""
    def rotl(x, k):
        return ((x << k) | (x >> (8 - k))) & 0xFF
    return rotl(v ^ (rotl((~v) & 0xFF, 1) & rotl(v, 2) & rotl(v, 3)), 1)


SBOX = [sbox(v) for v in range(256)]


def bar(x: int) -> int:
    """skyscraper/core/src/reference.rs:80-94: canonical LE bytes, swap 16-byte halves,
    per-byte sbox, reinterpret LE, reduce mod p."""
    b = (x % P).to_bytes(32, "little")
    b = b[16:] + b[:16]
    return int.from_bytes(bytes(SBOX[v] for v in b), "little") % P


def _sq(x: int) -> int:
    """x^2 * sigma^-1, sigma = 2^256 (reference.rs:63-69); equals block_multiplier::scalar_sqr
    on raw integers (skyscraper/block-multiplier/src/scalar.rs:11-70)."""
    return x * x % P * SIGMA_INV % P


_BAR_ROUNDS_V2 = (6, 7, 10, 11)


def permute(l: int, r: int) -> tuple[int, int]:
    """Skyscraper-v2 permutation, skyscraper/core/src/reference.rs:49-78 (Figure 2.a):
    18 Feistel rounds (l, r) <- (r + F_i(l) + rc_i, l)."""
    l %= P
    r %= P
    for i in range(18):
        f = bar(l) if i in _BAR_ROUNDS_V2 else _sq(l)
        l, r = (r + f + ROUND_CONSTANTS[i]) % P, l
    return l, r


def compress(l: int, r: int) -> int:
    """skyscraper/core/src/reference.rs:41-46 / generic.rs:77-102: permute(l,r).0 + l."""
    return (permute(l, r)[0] + l) % P


def compress_v1(l: int, r: int) -> int:
    """Old 10-round variant, skyscraper/core/src/v1.rs:19-32.  Needed only because the checked-in
    proof fixture was produced with it."""
    l %= P
    r %= P
    t = l
    bars = (2, 3, 6, 7)
    for i in range(10):
        f = bar(l) if i in bars else _sq(l)
        rc = ROUND_CONSTANTS[i] if i < 9 else 0
        l, r = (r + f + rc) % P, l
    return (l + t) % P


def compress_many(messages: bytes, version: int = 2) -> bytes:
    """CompressManyFn contract, skyscraper/core/src/generic.rs:14-37 / lib.rs:26:
    n messages of 64 B (two LE 256-bit ints) -> n hashes of 32 B (canonical LE)."""
    assert len(messages) % 64 == 0
    f = compress if version == 2 else compress_v1
    out = bytearray()
    for i in range(0, len(messages), 64):
        l = int.from_bytes(messages[i:i + 32], "little")
        r = int.from_bytes(messages[i + 32:i + 64], "little")
        out += f(l, r).to_bytes(32, "little")
    return bytes(out)


# --- PoW -------------------------------------------------------------------------------------
def _f64_to_u256(f: float) -> int:
    """skyscraper/core/src/pow.rs:61-82."""
    import math
    import struct
    bits = struct.unpack("<Q", struct.pack("<d", f))[0]
    sign = bits >> 63
    exp_bits = (bits >> 52) & 0x7FF
    frac = bits & ((1 << 52) - 1)
    if exp_bits == 0:
        exp, sig = -1022, frac
    else:
        exp, sig = exp_bits - 1023, frac + (1 << 52)
    if sign:
        return 0
    if exp > 256:
        return (1 << 256) - 1
    shift = exp - 52
    if shift < 0:
        # f.round() as u64 (round half away from zero)
        return int(math.floor(f + 0.5))
    limb, sh = divmod(shift, 64)
    res = [0, 0, 0, 0]
    res[limb] = (sig << sh) & 0xFFFFFFFFFFFFFFFF
    if sh != 0 and limb < 3:
        res[limb + 1] = sig >> (64 - sh)
    return from_limbs(res)


def pow_threshold(difficulty: float) -> int:
    """skyscraper/core/src/pow.rs:14-22."""
    assert 0.0 <= difficulty < 80.0
    modulus = float(to_limbs(P)[3]) * 2.0 ** 192
    prob = 2.0 ** (-difficulty)
    return _f64_to_u256(prob * modulus)


PROVER_BIAS = 0.01  # pow.rs:6


def pow_verify(challenge: int, difficulty: float, nonce: int) -> bool:
    """skyscraper/core/src/pow.rs:24-26."""
    return difficulty == 0.0 or compress(challenge, nonce) < pow_threshold(difficulty)


def pow_solve(challenge: int, difficulty: float) -> int:
    """skyscraper/core/src/pow.rs:33-41 + generic.rs:42-71: smallest accepted nonce wins
    (fetch_min), with the +0.01 bit prover bias."""
    if difficulty == 0.0:
        return 0
    thr = pow_threshold(difficulty + PROVER_BIAS)
    nonce = 0
    while compress(challenge, nonce) >= thr:
        nonce += 1
    return nonce


# --- Merkle (ark-crypto-primitives 0.5 layout) [EXT], hash plug-in in-tree ---------------------
def leaf_hash(leaf: list[int], comp=compress) -> int:
    """SkyscraperCRH::evaluate, provekit/common/src/skyscraper/whir.rs:30-48: left fold of
    compress over the leaf slice."""
    assert len(leaf) > 0, "IncorrectInputLength(0)"
    d = leaf[0]
    for x in leaf[1:]:
        d = comp(d, x)
    return d


def merkle_tree(leaves: list[list[int]], comp=compress) -> list[int]:
    """Heap-ordered tree: nodes[1] = root, children of i are 2i, 2i+1, leaf digests at
    nodes[L .. 2L) (layout is ours; root + paths are what ark's MerkleTree::new pins)."""
    L = len(leaves)
    assert L >= 2 and L & (L - 1) == 0
    nodes = [0] * (2 * L)
    for i, leaf in enumerate(leaves):
        nodes[L + i] = leaf_hash(leaf, comp)
    for i in range(L - 1, 0, -1):
        nodes[i] = comp(nodes[2 * i], nodes[2 * i + 1])
    return nodes


def merkle_multipath(nodes: list[int], indexes: list[int]):
    """ark MultiPath [EXT]; encoding pinned by recursive-verifier/app/circuit/mt.go:36-50 and
    app/utilities/utilities.go:71-82 and by the fixture walk:
    per queried leaf: sibling leaf digest; auth path ordered root->leaf *excluding* the leaf
    level, prefix-compressed against the previous path."""
    L = len(nodes) // 2
    sib, prefix_lens, suffixes = [], [], []
    prev: list[int] = []
    for idx in indexes:
        pos = L + idx
        sib.append(nodes[pos ^ 1])
        path = []
        pos >>= 1
        while pos > 1:
            path.append(nodes[pos ^ 1])
            pos >>= 1
        path.reverse()  # root -> leaf
        k = 0
        while k < len(prev) and k < len(path) and prev[k] == path[k]:
            k += 1
        prefix_lens.append(k)
        suffixes.append(path[k:])
        prev = path
    return sib, prefix_lens, suffixes, list(indexes)


def merkle_verify_path(root: int, idx: int, leaf: list[int], sibling: int, path_root_to_leaf: list[int],
                       comp=compress) -> bool:
    """recursive-verifier/app/circuit/whir_utilities.go:13-46 (leaf -> root, index bits LSB first)."""
    h = leaf_hash(leaf, comp)
    h = comp(sibling, h) if idx & 1 else comp(h, sibling)
    idx >>= 1
    for s in reversed(path_root_to_leaf):
        h = comp(s, h) if idx & 1 else comp(h, s)
        idx >>= 1
    return h == root


# --- multilinear / univariate helpers [EXT whir::poly_utils], pinned by the Go verifier -----------
def eval_multilinear_coeffs(coeffs: list[int], point: list[int]) -> int:
    """MultivarPoly, recursive-verifier/app/utilities/utilities.go:15-22: the LAST variable binds
    the top half of the coefficient list, i.e. point[j] binds bit j of the coefficient index."""
    if not point:
        return coeffs[0]
    h = len(coeffs) // 2
    return (eval_multilinear_coeffs(coeffs[:h], point[:-1])
            + point[-1] * eval_multilinear_coeffs(coeffs[h:], point[:-1])) % P


def eval_coeffs_at_point(coeffs: list[int], point: list[int]) -> int:
    """[EXT] whir CoefficientList::evaluate convention: point[0] binds the MOST significant bit of the
    coefficient index (so expand_from_univariate(z, n) gives the univariate value at z).  It is the Go
    MultivarPoly with the variable list reversed (whir.go:203 reverses the folding randomness)."""
    return eval_multilinear_coeffs(coeffs, point[::-1])


def expand_from_univariate(z: int, n: int) -> list[int]:
    """utilities.go:182-190: (z^(2^(n-1)), ..., z^2, z)."""
    res = [0] * n
    acc = z
    for i in range(n):
        res[n - 1 - i] = acc
        acc = acc * acc % P
    return res


def eval_univariate(coeffs: list[int], z: int) -> int:
    """UnivarPoly, utilities.go:24-38 (Horner)."""
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * z + c) % P
    return acc


def eq_poly_outside(coords: list[int], point: list[int]) -> int:
    """utilities.go:140-146."""
    acc = 1
    for c, p_ in zip(coords, point):
        acc = acc * (c * p_ + (1 - c) * (1 - p_)) % P
    return acc


def eval_eq(point: list[int], scalar: int = 1) -> list[int]:
    """provekit/common/src/utils/sumcheck.rs:145-171: first variable <-> top half (MSB)."""
    out = [scalar % P]
    for x in point:
        nxt = []
        for s in out:
            s1 = s * x % P
            nxt += [(s - s1) % P, s1]
        out = nxt
    return out


def coeffs_to_evals(c: list[int]) -> list[int]:
    """[EXT] CoefficientList -> EvaluationsList (forward "wavelet"): evals[idx] = sum of coeffs
    whose index bits are a subset of idx's bits."""
    a = list(c)
    n = len(a)
    h = 1
    while h < n:
        for i in range(0, n, 2 * h):
            for j in range(i, i + h):
                a[j + h] = (a[j + h] + a[j]) % P
        h *= 2
    return a


def evals_to_coeffs(e: list[int]) -> list[int]:
    """[EXT] EvaluationsList::to_coeffs (inverse wavelet); call site whir_r1cs.rs:195,198."""
    a = list(e)
    n = len(a)
    h = 1
    while h < n:
        for i in range(0, n, 2 * h):
            for j in range(i, i + h):
                a[j + h] = (a[j + h] - a[j]) % P
        h *= 2
    return a


def rs_encode_leaves(coeffs: list[int], log_inv_rate: int, fold: int = 4) -> list[list[int]]:
    """[EXT] whir commit-time RS encoding in the "prover helps" coefficient layout, pinned by a
    reference-produced proof and by the Go verifier (whir.go:139-142,
    whir_utilities.go:180-186): with f(X) = sum_k X^k f_k(X^(2^fold)) and g the generator of the
    domain of size D = n << log_inv_rate, leaf i entry k = f_k((g^(2^fold))^i).  O(n^2): tiny only."""
    n = len(coeffs)
    w = 1 << fold
    D = n << log_inv_rate
    logD = D.bit_length() - 1
    g = root_of_unity(logD)
    gw = pow(g, w, P)
    L = D // w
    leaves = []
    for i in range(L):
        z = pow(gw, i, P)
        leaves.append([eval_univariate(coeffs[k::w], z) for k in range(w)])
    return leaves


# --- sumchecks ---------------------------------------------------------------------------------
def zk_sumcheck_round(a, b, c, eq, fold):
    """sumcheck_fold_map_reduce with the map of whir_r1cs.rs:284-291
    (provekit/common/src/utils/sumcheck.rs:16-104).  MSB pairing: element i with i + len/2.
    Returns ((f0, f_em1, f_inf), folded arrays)."""
    arrs = [list(a), list(b), list(c), list(eq)]
    if fold is not None:
        out = []
        for x in arrs:
            h = len(x) // 2
            out.append([(x[i] + fold * (x[i + h] - x[i])) % P for i in range(h)])
        arrs = out
    h = len(arrs[0]) // 2
    f0 = fm = fi = 0
    A, B, C, E = arrs
    for i in range(h):
        a0, a1, b0, b1, c0, c1, e0, e1 = A[i], A[i + h], B[i], B[i + h], C[i], C[i + h], E[i], E[i + h]
        f0 += e0 * (a0 * b0 - c0)
        fm += (2 * e0 - e1) * ((2 * a0 - a1) * (2 * b0 - b1) - (2 * c0 - c1))
        fi += (e1 - e0) * (a1 - a0) * (b1 - b0)
    return (f0 % P, fm % P, fi % P), arrs


def whir_sumcheck_round(p, w, fold):
    """[EXT] whir SumcheckSingle: optional fold of adjacent pairs by the previous challenge, then
    h(X) = sum_i p_i(X) w_i(X) sent as [h(0), h(1), h(2)] (Go: whir_utilities.go:107-131,
    utilities.go:148-154).  LSB pairing: elements 2i, 2i+1."""
    p = list(p)
    w = list(w)
    if fold is not None:
        p = [(p[2 * i] + fold * (p[2 * i + 1] - p[2 * i])) % P for i in range(len(p) // 2)]
        w = [(w[2 * i] + fold * (w[2 * i + 1] - w[2 * i])) % P for i in range(len(w) // 2)]
    h0 = h1 = h2 = 0
    for i in range(len(p) // 2):
        p0, p1, w0, w1 = p[2 * i], p[2 * i + 1], w[2 * i], w[2 * i + 1]
        h0 += p0 * w0
        h1 += p1 * w1
        h2 += (2 * p1 - p0) * (2 * w1 - w0)
    return (h0 % P, h1 % P, h2 % P), p, w


def fold_coeffs(coeffs: list[int], r: list[int]) -> list[int]:
    """[EXT] CoefficientList::fold: each block of 2^k consecutive coefficients -> its multilinear
    evaluation at r where r[j] binds bit j of the in-block index (computeFold = MultivarPoly(leaf, r),
    whir_utilities.go:180-186)."""
    w = 1 << len(r)
    return [eval_multilinear_coeffs(coeffs[i:i + w], r) for i in range(0, len(coeffs), w)]


# --- WHIR parameter derivation [EXT whir::parameters], pinned by the two WhirConfigs inside the
# reference fixture poseidon-1000.nps ----------------------------------------
def whir_config(num_variables: int, batch_size: int = 2, security_level: int = 128,
                folding_factor: int = 4, starting_log_inv_rate: int = 1) -> dict:
    """Restates WhirConfig::new for the only setting ProveKit uses
    (provekit/r1cs-compiler/src/whir_r1cs.rs:38-52): SoundnessType::ConjectureList, security 128,
    FoldingFactor::Constant(4), starting_log_inv_rate 1,
    pow_bits = default_max_pow(num_variables, 1) = num_variables + 1 - 3, initial_statement = true.
    For the 254-bit field the OOD-sample count is 1 and all folding PoW bits are 0 for every size
    we can reach; the remaining quantities follow
        queries(rate)  = ceil((security - max_pow_bits) / log_inv_rate)
        pow_bits(round)= max(0, security - queries * log_inv_rate)."""
    import math
    assert num_variables >= folding_factor
    max_pow_bits = num_variables + starting_log_inv_rate - 3
    protocol_security = max(0, security_level - max_pow_bits)
    final_sumcheck_rounds = num_variables % folding_factor
    n_rounds = (num_variables - final_sumcheck_rounds) // folding_factor - 1
    log_inv_rate = starting_log_inv_rate
    nv = num_variables - folding_factor
    domain_log = num_variables + starting_log_inv_rate
    rounds = []
    for _ in range(n_rounds):
        next_rate = log_inv_rate + folding_factor - 1
        q = math.ceil(protocol_security / log_inv_rate)
        rounds.append(dict(pow_bits=float(max(0, security_level - q * log_inv_rate)),
                           folding_pow_bits=0.0, num_queries=q, ood_samples=1,
                           log_inv_rate=log_inv_rate, num_variables=nv,
                           folding_factor=folding_factor, domain_log=domain_log))
        nv -= folding_factor
        log_inv_rate = next_rate
        domain_log -= 1
    fq = math.ceil(protocol_security / log_inv_rate)
    return dict(num_variables=num_variables, max_pow_bits=max_pow_bits,
                committment_ood_samples=1, starting_log_inv_rate=starting_log_inv_rate,
                starting_folding_pow_bits=0.0, folding_factor=folding_factor,
                starting_domain_log=num_variables + starting_log_inv_rate,
                rounds=rounds, final_queries=fq,
                final_pow_bits=float(max(0, security_level - fq * log_inv_rate)),
                final_log_inv_rate=log_inv_rate, final_sumcheck_rounds=final_sumcheck_rounds,
                final_folding_pow_bits=0.0, final_domain_log=domain_log, batch_size=batch_size)


# ---------------------------------------------------------------------------------------------
# Mask generator (twin of pk_rng_fill / oracle/rng.c): ChaCha12 counter stream -> uniform Fr.
# The reference draws masks with F::rand(&mut thread_rng()) (provekit/common/src/utils/zk_utils.rs:13-22):
# a ChaCha12 stream, rejection sampling of 254-bit strings below p.  Block function = RFC 8439 section 2.3.
# ---------------------------------------------------------------------------------------------
def chacha_block(state: list[int], rounds: int) -> list[int]:
    x = list(state)

    def qr(a, b, c, d):
        x[a] = (x[a] + x[b]) & 0xFFFFFFFF; x[d] ^= x[a]; x[d] = ((x[d] << 16) | (x[d] >> 16)) & 0xFFFFFFFF
        x[c] = (x[c] + x[d]) & 0xFFFFFFFF; x[b] ^= x[c]; x[b] = ((x[b] << 12) | (x[b] >> 20)) & 0xFFFFFFFF
        x[a] = (x[a] + x[b]) & 0xFFFFFFFF; x[d] ^= x[a]; x[d] = ((x[d] << 8) | (x[d] >> 24)) & 0xFFFFFFFF
        x[c] = (x[c] + x[d]) & 0xFFFFFFFF; x[b] ^= x[c]; x[b] = ((x[b] << 7) | (x[b] >> 25)) & 0xFFFFFFFF

    for _ in range(rounds // 2):
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
    return [(a + b) & 0xFFFFFFFF for a, b in zip(x, state)]


CHACHA_CONST = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574]


def rng_element(seed: bytes, stream: int, i: int) -> int:
    """Element i of stream `stream`: the integer whose 32 LE bytes are the element's in-memory (Montgomery) form."""
    key = [int.from_bytes(seed[4 * k:4 * k + 4], "little") for k in range(8)]
    attempt = 0
    while True:
        blk = chacha_block(CHACHA_CONST + key + [i & 0xFFFFFFFF, i >> 32, stream, attempt], 12)
        for c in range(2):
            w = blk[8 * c:8 * c + 8]
            w[7] &= 0x3FFFFFFF
            v = sum(x << (32 * k) for k, x in enumerate(w))
            if v < P:
                return v
        attempt += 1
