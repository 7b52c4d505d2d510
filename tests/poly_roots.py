"""Roots of small polynomials over BN254-Fr (test helper): gcd with x^p - x, then Cantor-Zassenhaus splitting.
Polynomials are coefficient lists, lowest degree first.  Used to recover verifier challenges from a reference-produced
proof WITHOUT the Fiat-Shamir sponge: a sumcheck challenge is a root of h_{i-1}(X) - (h_i(0) + h_i(1))."""
import random

from oracle import pyref as o

P = o.P

def trim(a):
    a = list(a)
    while a and a[-1] == 0: a.pop()
    return a
def pmod(a, f):
    a = trim(a); f = trim(f)
    inv = pow(f[-1], P - 2, P)
    while len(a) >= len(f):
        c = a[-1] * inv % P
        off = len(a) - len(f)
        for j in range(len(f)):
            a[off + j] = (a[off + j] - c * f[j]) % P
        a = trim(a)
    return a
def pmul(a, b):
    if not a or not b: return []
    res = [0] * (len(a) + len(b) - 1)
    for i, x in enumerate(a):
        for j, y in enumerate(b):
            res[i + j] = (res[i + j] + x * y) % P
    return res
def ppowmod(b, e, f):
    r = [1]
    b = pmod(b, f)
    while e:
        if e & 1: r = pmod(pmul(r, b), f)
        b = pmod(pmul(b, b), f)
        e >>= 1
    return r
def pgcd(a, b):
    a, b = trim(a), trim(b)
    while b:
        a, b = b, pmod(a, b)
    return a
def roots(f):
    f = trim(f)
    if len(f) <= 1: return []
    xp = ppowmod([0, 1], P, f)
    xp = xp + [0] * (2 - len(xp))
    xp[1] = (xp[1] - 1) % P
    g = pgcd(f, xp)
    out = []
    def split(g):
        g = trim(g)
        d = len(g) - 1
        if d == 0: return
        if d == 1:
            out.append((-g[0]) * pow(g[1], P - 2, P) % P); return
        while True:
            a = random.randrange(P)
            h = ppowmod([a, 1], (P - 1) // 2, g)
            h = h + [0] * (1 - len(h)) if h else [0]
            h[0] = (h[0] - 1) % P
            d1 = pgcd(g, h)
            if 0 < len(d1) - 1 < d:
                split(d1)
                # quotient
                q = []
                rem = list(g)
                inv = pow(d1[-1], P - 2, P)
                for k in range(len(g) - len(d1), -1, -1):
                    c = rem[k + len(d1) - 1] * inv % P
                    q.insert(0, c)
                    for j in range(len(d1)):
                        rem[k + j] = (rem[k + j] - c * d1[j]) % P
                split(q)
                return
    split(g)
    return sorted(out)

def quad_from_evals(h0, h1, h2):
    inv2 = pow(2, P - 2, P)
    c2 = (h2 - 2 * h1 + h0) * inv2 % P
    c1 = (h1 - h0 - c2) % P
    return [h0, c1, c2]
