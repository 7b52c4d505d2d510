"""Walks the reference-produced proof transcript (tests/golden/poseidon-1000.transcript.bin)
following the layout reconstructed in SURVEY A.4.  Test helper; uses oracle/pyref only."""
import os
import struct

from oracle import pyref as o

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                      "poseidon-1000.transcript.bin")


class Reader:
    def __init__(self, b: bytes):
        self.b = b
        self.pos = 0

    def scalar(self) -> int:
        v = int.from_bytes(self.b[self.pos:self.pos + 32], "little")
        self.pos += 32
        return v

    def scalars(self, n):
        return [self.scalar() for _ in range(n)]

    def raw(self, n):
        v = self.b[self.pos:self.pos + n]
        self.pos += n
        return v

    def hint(self) -> bytes:
        (n,) = struct.unpack_from("<I", self.b, self.pos)
        self.pos += 4
        return self.raw(n)


def parse_stir_answers(h: bytes):
    """ark-serialize Vec<Vec<Fr>>: u64 count, per answer u64 len + 32 B canonical LE elems."""
    pos = 0
    (n,) = struct.unpack_from("<Q", h, pos)
    pos += 8
    out = []
    for _ in range(n):
        (m,) = struct.unpack_from("<Q", h, pos)
        pos += 8
        out.append([int.from_bytes(h[pos + 32 * i:pos + 32 * i + 32], "little") for i in range(m)])
        pos += 32 * m
    assert pos == len(h)
    return out


def parse_multipath(h: bytes):
    """ark MultiPath: leaf_sibling_hashes Vec<[u8;32]>, auth_paths_prefix_lengths Vec<u64>,
    auth_paths_suffixes Vec<Vec<[u8;32]>>, leaf_indexes Vec<u64> (recursive-verifier types.go:17-22)."""
    pos = 0

    def u64():
        nonlocal pos
        (v,) = struct.unpack_from("<Q", h, pos)
        pos += 8
        return v

    def dig():
        nonlocal pos
        v = int.from_bytes(h[pos:pos + 32], "little")
        pos += 32
        return v

    sib = [dig() for _ in range(u64())]
    pre = [u64() for _ in range(u64())]
    suf = []
    for _ in range(u64()):
        suf.append([dig() for _ in range(u64())])
    idx = [u64() for _ in range(u64())]
    assert pos == len(h)
    return sib, pre, suf, idx


def decode_paths(pre, suf):
    """utilities.go:71-82 PrefixDecodePath; returns root->leaf paths."""
    paths, prev = [], []
    for k, s in zip(pre, suf):
        prev = prev[:k] + s
        paths.append(prev)
    return paths


def walk_whir(rd: Reader, cfg: dict):
    """One WHIR opening (SURVEY A.4 'whir(C)')."""
    out = dict(rounds=[])
    ff = cfg["folding_factor"]
    out["initial_sumcheck"] = [rd.scalars(3) for _ in range(ff)]
    for r in cfg["rounds"]:
        e = dict(root=rd.scalar(), ood=rd.scalars(r["ood_samples"]))
        if r["pow_bits"] > 0:
            e["nonce"] = int.from_bytes(rd.raw(8), "big")
        e["answers"] = parse_stir_answers(rd.hint())
        e["multipath"] = parse_multipath(rd.hint())
        e["sumcheck"] = [rd.scalars(3) for _ in range(ff)]
        out["rounds"].append(e)
    out["final_coeffs"] = rd.scalars(1 << cfg["final_sumcheck_rounds"])
    if cfg["final_pow_bits"] > 0:
        out["final_nonce"] = int.from_bytes(rd.raw(8), "big")
    out["final_answers"] = parse_stir_answers(rd.hint())
    out["final_multipath"] = parse_multipath(rd.hint())
    out["final_sumcheck"] = [rd.scalars(3) for _ in range(cfg["final_sumcheck_rounds"])]
    d = rd.hint()
    (n,) = struct.unpack_from("<Q", d, 0)
    out["deferred"] = [int.from_bytes(d[8 + 32 * i:40 + 32 * i], "little") for i in range(n)]
    return out


def walk_proof(m=21, m_0=20, data=None):
    """proof := commit(W) commit(H) zk-sumcheck whir(H) hint claimed_evaluations whir(W).
    data: transcript bytes to walk (default: the reference-produced fixture)."""
    b = open(GOLDEN, "rb").read() if data is None else data
    rd = Reader(b)
    cfg_w = o.whir_config(m)
    blind_vars = (4 * m_0 - 1).bit_length() + 1
    cfg_h = o.whir_config(blind_vars)
    out = dict(cfg_w=cfg_w, cfg_h=cfg_h)
    out["commit_w"] = dict(root=rd.scalar(), ood=rd.scalars(2))
    out["commit_h"] = dict(root=rd.scalar(), ood=rd.scalars(2))
    out["sum_g"] = rd.scalar()
    out["zk_sumcheck"] = [rd.scalars(4) for _ in range(m_0)]
    out["blind_sums"] = rd.scalars(2)
    out["whir_h"] = walk_whir(rd, cfg_h)
    ce = rd.hint()
    out["claimed_evaluations_raw"] = ce
    out["whir_w"] = walk_whir(rd, cfg_w)
    assert rd.pos == len(b), (rd.pos, len(b))
    return out
