"""Pins the oracle (pyref + C) against the reference's own known-answer tests.
KAT sources: skyscraper/core/src/reference.rs:100-188, pow.rs:84-112,
skyscraper/block-multiplier proptests (scalar.rs:134-206, proptest-regressions/scalar.txt)."""
import ctypes

import numpy as np

from helpers import arr_to_ints, from_mont, ints_to_arr, ptr, rand_fr, to_mont
from oracle import pyref as o

P = o.P


def test_constants():
    assert o.SIGMA_INV == o.R_INV                       # reference.rs:22-26 == 2^-256
    assert o.NP0 == 0xc2e1f593efffffff                  # block-multiplier constants.rs:1
    assert o.to_limbs(o.R) == [0xac96341c4ffffffb, 0x36fc76959f60cd29, 0x666ea36f7879462e, 0x0e0a77c19a07df2f]
    assert o.to_limbs(o.R2) == [0x1bb8e645ae216da7, 0x53fe3ab1e35c59e3, 0x8c49833d53bb8085, 0x0216d0b17f4e44a5]
    assert o.to_limbs(o.R_INV) == [0xdc5ba0056db1194e, 0x090ef5a9e111ec87, 0xc8260de4aeb85d5d, 0x15ebf95182c5551c]
    assert o.HALF * 2 % P == 1                          # utils/mod.rs:23-25
    assert pow(o.ROOT28, 1 << 28, P) == 1 and pow(o.ROOT28, 1 << 27, P) != 1


def test_sbox_table3():  # reference.rs:123-131
    for v, e in [(0xcd, 0xd3), (0x17, 0x0e), (0x83, 0x17), (0x14, 0x28), (0x2b, 0x46), (0x1e, 0xbc)]:
        assert o.sbox(v) == e


def _ss(rnd, l, r):
    for i in (rnd, rnd + 1):
        l, r = (r + o._sq(l) + o.ROUND_CONSTANTS[i]) % P, l
    return l, r


def _bb(rnd, l, r):
    for i in (rnd, rnd + 1):
        l, r = (r + o.bar(l) + o.ROUND_CONSTANTS[i]) % P, l
    return l, r


def test_ss_2():  # reference.rs:104-121
    l, r = _ss(2, 11818428481613126259506041491792444971306025298632020312923851211664140080269,
               16089984100220651117533376273482359701319211672522891227502963383930673183481)
    assert l == 2897520731550929941842826131888578795995028656093850302425034320680216166225
    assert r == 10274752619072178425540318899508997829349102488123199431506343228471746115261


def test_bb_6():  # reference.rs:134-151
    l, r = _bb(6, 13251711941470795978907268022756015766767985221093713388330058285942871890923,
               1017722258958995329580328739423576514309327442471989504101393158056883989572)
    assert l == 3193610555912363022088172260048956988022957239290210718020144819371540058981
    assert r == 17363210535454321713488811303876243393424286347736908007836172565366081010820


KAT_PERMUTE = [  # reference.rs:153-188
    ((0, 0),
     (5793276905781313965269111743763131906666794041798623267477617572701829069290,
      12296274483727574983376829575121280934973829438414198530604912453551798647077)),
    ((50417215636675310123686652273432694184389644587803328798109154235492038730484,
      14620920779025509970947930308416120371903474543120179490887326852503500806990),
     (8412949970293910117511617126618515787729842528183672400383899220234743146062,
      11868175801025513844525564200589229804433722826344843184417708742749423276015)),
]


def test_permute_kats_pyref_and_c(orc):
    for (l, r), (el, er) in KAT_PERMUTE:
        assert o.permute(l, r) == (el, er)
        # reference.rs test_random feeds l > p: Fr::new reduces it; the C oracle must accept it too
        a, b = ints_to_arr([l % (1 << 256)]), ints_to_arr([r])
        lo, ro = np.zeros(4, np.uint64), np.zeros(4, np.uint64)
        orc.orc_sky_permute(ptr(a), ptr(b), ptr(lo), ptr(ro))
        assert arr_to_ints(lo)[0] == el and arr_to_ints(ro)[0] == er


def test_compress_c_matches_pyref(orc):
    rng = np.random.default_rng(7)
    n = 300
    msgs = rng.integers(0, 1 << 64, size=(n, 8), dtype=np.uint64)
    # edge cases: zero, all-ones (>= 5p), p-1, p, 2p, values around the reduce thresholds
    # (skyscraper/core/src/reduce.rs:97-127 test_reduce_partial_max)
    edge = [0, (1 << 256) - 1, P - 1, P, 2 * P, P + 1, 5 * P, 5 * P + 12345, (1 << 255)]
    for i, e in enumerate(edge):
        msgs[i, :4] = ints_to_arr([e])[0]
        msgs[i, 4:] = ints_to_arr([edge[-1 - i]])[0]
    for version in (1, 2):
        out = np.zeros((n, 4), np.uint64)
        assert orc.orc_sky_compress_many(ptr(msgs), ptr(out), ctypes.c_size_t(n), version) == 0
        exp = o.compress_many(msgs.tobytes(), version)
        assert out.tobytes() == exp


def test_montgomery_mul_matches_definition(orc):
    """block-multiplier proptest (scalar.rs:146-154): mul(l, r) == l*r*2^-256 mod p, incl. the
    saved regression cases proptest-regressions/scalar.txt (reduced mod p for our canonical API)."""
    rng = np.random.default_rng(0)
    a, b = rand_fr(rng, 2000), rand_fr(rng, 2000)
    reg = [([0, 0, 0, 1], [0, 0, 0, 1]),
           ([0, 887, 0, 15778841185528309819],
            [458854615557053794, 8784556235901218364, 1751211468174275388, 16873806747226852460])]
    for i, (l, r) in enumerate(reg):
        a[i] = ints_to_arr([o.from_limbs(l) % P])[0]
        b[i] = ints_to_arr([o.from_limbs(r) % P])[0]
    out = np.zeros_like(a)
    orc.orc_fr_mul(ptr(a), ptr(b), ptr(out), ctypes.c_size_t(len(a)))
    ai, bi, oi = arr_to_ints(a), arr_to_ints(b), arr_to_ints(out)
    for x, y, z in zip(ai, bi, oi):
        assert z == x * y * o.R_INV % P


def test_pow_threshold_and_solve(orc):  # pow.rs:84-112
    assert o._f64_to_u256(0.0) == 0 and o._f64_to_u256(0.49) == 0 and o._f64_to_u256(0.5) == 1
    assert o._f64_to_u256(1.0) == 1 and o._f64_to_u256(2.0 ** 128) == 1 << 128
    assert o._f64_to_u256(float("inf")) == (1 << 256) - 1 and o._f64_to_u256(-42.0) == 0
    for d in (1.0, 3.141592653589793, 10.0, 19.0, 19.01):
        t = np.zeros(4, np.uint64)
        orc.orc_pow_threshold(d, ptr(t))
        assert arr_to_ints(t)[0] == o.pow_threshold(d)
    ch = ints_to_arr([(1 << 256) - 1])  # test_solve_verify uses challenge = [u64::MAX; 4]
    for d in (0.0, 3.141592653589793, 8.0):
        n_c = orc.orc_pow_solve(ptr(ch), d)
        assert n_c == o.pow_solve((1 << 256) - 1, d)
        assert orc.orc_pow_verify(ptr(ch), d, ctypes.c_uint64(n_c)) == 1
        assert o.pow_verify((1 << 256) - 1, d, n_c)
