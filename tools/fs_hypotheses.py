"""Hypothesis tester for the un-vendored Fiat-Shamir layer (spongefish + whir domain separator), round-2 research aid.

tests/golden/poseidon-1000.challenges.json holds challenge VALUES of the reference-produced proof recovered by algebra.
The first of them, rand r_1, is the 2nd or 3rd squeeze of the transcript, after absorbing values that are in the proof
(root_W, two OOD answers): a candidate (domain-separator string -> IV, permutation, state layout) can therefore be
checked in milliseconds.  This script enumerates the variants of the restated domain separator (label spellings, merged
or split OOD answers, PoW before/after the STIR squeeze, hint order, byte counts in units or bytes), three ways of
hashing it into the IV (unpadded Keccak duplex, Keccak-256, SHA3-256), both Skyscraper permutations (v2 and the
10-round v1 the fixture's Merkle tree uses), both state orders and both IV endiannesses.

Also tried by hand with the same tester: a byte-oriented Keccak duplex transcript (spongefish's DefaultHash; scalars
absorbed as 32 bytes, challenges from 47/48/32 squeezed bytes, big- or little-endian), and additive instead of
overwriting absorption / eager permutation after a full rate.

Result (r01 session 4): 9 984 strings x 2 permutations x 4 state variants (+ the families above), no match -- the fixture's transcript was built
from a domain separator this restatement does not reproduce; the sources (spongefish, whir @ the fixture's revision) are
needed.  Run:  python tools/fs_hypotheses.py
"""
import hashlib
import itertools
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import pyref as o
import fixture_walk as fw
P = o.P

# ---- keccak-f[1600]
RC = [0x0000000000000001,0x0000000000008082,0x800000000000808A,0x8000000080008000,0x000000000000808B,0x0000000080000001,0x8000000080008081,0x8000000000008009,0x000000000000008A,0x0000000000000088,0x0000000080008009,0x000000008000000A,0x000000008000808B,0x800000000000008B,0x8000000000008089,0x8000000000008003,0x8000000000008002,0x8000000000000080,0x000000000000800A,0x800000008000000A,0x8000000080008081,0x8000000000008080,0x0000000080000001,0x8000000080008008]
ROT = [[0,36,3,41,18],[1,44,10,45,2],[62,6,43,15,61],[28,55,25,21,56],[27,20,39,8,14]]
M64 = (1<<64)-1
def rol(x,n): return ((x<<n)|(x>>(64-n)))&M64 if n else x
def keccak_f(s):
    for rnd in range(24):
        C=[s[x]^s[x+5]^s[x+10]^s[x+15]^s[x+20] for x in range(5)]
        D=[C[(x+4)%5]^rol(C[(x+1)%5],1) for x in range(5)]
        s=[s[i]^D[i%5] for i in range(25)]
        B=[0]*25
        for x in range(5):
            for y in range(5):
                B[y+5*((2*x+3*y)%5)]=rol(s[x+5*y],ROT[x][y])
        s=[B[x+5*y]^((~B[(x+1)%5+5*y])&M64&B[(x+2)%5+5*y]) for y in range(5) for x in range(5)]
        s[0]^=RC[rnd]
    return s
def tag_overwrite_duplex(io: bytes) -> bytes:
    st=bytearray(200); ap=0
    for ch in io:
        if ap==136:
            w=[int.from_bytes(st[8*i:8*i+8],"little") for i in range(25)]; w=keccak_f(w)
            st=bytearray(b"".join(x.to_bytes(8,"little") for x in w)); ap=0
        st[ap]=ch; ap+=1
    w=[int.from_bytes(st[8*i:8*i+8],"little") for i in range(25)]; w=keccak_f(w)
    return b"".join(x.to_bytes(8,"little") for x in w)[:32]

def permute_v1(l, r):
    l%=P; r%=P
    bars=(2,3,6,7)
    for i in range(10):
        f=o.bar(l) if i in bars else o._sq(l)
        rc=o.ROUND_CONSTANTS[i] if i<9 else 0
        l,r=(r+f+rc)%P,l
    return l,r

class Sponge:
    def __init__(self, iv_int, perm, state_order=0):
        self.st=[0, iv_int%P] if state_order==0 else [iv_int%P,0]
        self.perm=perm; self.ap=0; self.sp=1
    def absorb(self,x):
        if self.ap==1:
            self.st=list(self.perm(*self.st)); self.ap=0
        self.st[0]=x%P; self.ap=1; self.sp=1
    def squeeze(self):
        if self.sp==1:
            self.sp=0; self.ap=0; self.st=list(self.perm(*self.st))
        out=self.st[0]; self.sp=1
        return out

pr = fw.walk_proof()
gold = json.load(open(os.path.join(ROOT, "tests", "golden", "poseidon-1000.challenges.json")))
G = {k.split(" ")[0]: v for k, v in gold.items()}
RAND = [int(x,16) for x in G["rand_r_1..20"]]

def test_iv(tagbytes, perm, verbose=False):
    for order in (0,1):
        for ivmode in ("le","be"):
            iv=int.from_bytes(tagbytes,"little" if ivmode=="le" else "big")
            sp=Sponge(iv,perm,order)
            sp.absorb(pr["commit_w"]["root"]); z=sp.squeeze()
            sp.absorb(pr["commit_w"]["ood"][0]); sp.absorb(pr["commit_w"]["ood"][1])
            outs=[sp.squeeze() for _ in range(3)]
            if RAND[0] in outs:
                return (order, ivmode, outs.index(RAND[0]))
    return None


def ceil_div(a,b): return (a+b-1)//b
cw, ch = pr["cfg_w"], pr["cfg_h"]
def build(L, V):
    s = "\U0001F32A️".encode()
    def op(k,c,l): return b"\0"+k.encode()+str(c).encode()+l.encode()
    def hint(l): return b"\0H"+l.encode()
    def pow_(bits): return (op("S",ceil_div(32,15) if V["pow_units"] else 32,L["pow_q"])+op("A",8,L["pow_n"])) if bits>0 else b""
    def commit(c):
        r=op("A",1,L["digest"])+op("S",1,L["ood_q"])
        r+= op("A",c["batch_size"],L["ood_a"]) if V["ood_merged"] else b"".join(op("A",1,L["ood_a"]) for _ in range(c["batch_size"]))
        if c["batch_size"]>1 and L["batch"] is not None: r+=op("S",1,L["batch"])
        return r
    def sumcheck(n): return b"".join(op("A",3,L["sc_poly"])+op("S",1,L["fold_r"]) for _ in range(n))
    def hints(): return (hint(L["stir_a"])+hint(L["mproof"])) if V["hint_order"]==0 else (hint(L["mproof"])+hint(L["stir_a"]))
    def whir(c):
        r=op("S",1,L["init_comb"])+sumcheck(4)
        for rd in c["rounds"]:
            nb=ceil_div(rd["domain_log"]-4,8)
            r+=op("A",1,L["digest"])+op("S",1,L["ood_q"])+op("A",1,L["ood_a"])
            q=op("S",ceil_div(rd["num_queries"]*nb,15) if V["pow_units"] else rd["num_queries"]*nb,L["stir_q"])
            r+= (pow_(rd["pow_bits"])+q) if V["pow_first"] else (q+pow_(rd["pow_bits"]))
            r+=hints()+op("S",1,L["comb"])+sumcheck(4)
        nb=ceil_div(c["final_domain_log"]-4,8)
        r+=op("A",1<<c["final_sumcheck_rounds"],L["final_c"])
        q=op("S",ceil_div(c["final_queries"]*nb,15) if V["pow_units"] else c["final_queries"]*nb,L["final_q"])
        r+= (pow_(c["final_pow_bits"])+q) if V["pow_first"] else (q+pow_(c["final_pow_bits"]))
        r+=hints()+sumcheck(c["final_sumcheck_rounds"])
        if V["deferred"]: r+=hint(L["deferred"])
        return r
    s+=commit(cw)+op("S",20,"rand")+commit(ch)
    s+=op("A",1,"Sum of G over boolean hypercube")+op("S",1,"Rho")
    for _ in range(20): s+=op("A",4,"Sumcheck Polynomials")+op("S",1,"Sumcheck Random")
    s+=op("A",2,"Polynomial sums")
    s+=whir(ch)+hint("claimed_evaluations")+whir(cw)
    return s
def keccak256(data, pad):
    # standard sponge with padding byte `pad` (0x01 keccak, 0x06 sha3)
    st=bytearray(200); rate=136
    data=bytearray(data); data.append(pad); 
    while len(data)%rate: data.append(0)
    data[-1]|=0x80
    for off in range(0,len(data),rate):
        for i in range(rate): st[i]^=data[off+i]
        w=[int.from_bytes(st[8*i:8*i+8],"little") for i in range(25)]; w=keccak_f(w)
        st=bytearray(b"".join(x.to_bytes(8,"little") for x in w))
    return bytes(st[:32])
base=dict(digest="merkle_digest",ood_q="ood_query",ood_a="ood_ans",batch="batching_randomness",init_comb="initial_combination_randomness",sc_poly="sumcheck_poly",fold_r="folding_randomness",pow_q="pow_queries",pow_n="pow-nonce",stir_q="stir_queries",stir_a="stir_answers",mproof="merkle_proof",comb="combination_randomness",final_c="final_coeffs",final_q="final_queries",deferred="deferred_weight_evaluations")
batch_labels=["batching_randomness","batch_randomness","batching_rand","batching randomness","batching_challenge","batching_scalar","batching","batch_combination_randomness","combination_randomness","initial_combination_randomness","batching_coeff","batch_challenge",None]
t0=time.time(); n=0; found=[]
assert hashlib.sha3_256(b"abc").digest()==keccak256(b"abc",0x06)
for bl in batch_labels:
  for pq in ("pow_queries","pow-queries"):
    for fq in ("final_queries","stir_queries"):
      for V in itertools.product((0,1),(0,1),(0,1),(0,1),(0,1)):
        Vd=dict(pow_units=V[0],ood_merged=V[1],hint_order=V[2],pow_first=V[3],deferred=V[4])
        L=dict(base,batch=bl,pow_q=pq,final_q=fq)
        s=build(L,Vd)
        for tname,t in (("duplex",tag_overwrite_duplex(s)),("keccak",keccak256(s,0x01)),("sha3",keccak256(s,0x06))):
            for pname,perm in (("v2",o.permute),("v1",permute_v1)):
                n+=1
                r=test_iv(t,perm)
                if r: found.append((bl,pq,fq,Vd,tname,pname,r)); print("FOUND",found[-1])
print("tested",n,"in",round(time.time()-t0,1),"s; found",len(found))
