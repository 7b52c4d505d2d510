// provekit_b200/csrc/fr.cuh — BN254 scalar field on the B200 integer pipe (sm_100a).
//
// Device-side replacement for the arithmetic the reference gets from
//   * ark-ff 0.5 `Fp<MontBackend<BN254Config,4>,4>` [EXT] (provekit/prover/src/whir_r1cs.rs:136-139), and
//   * skyscraper/block-multiplier `scalar_mul/scalar_sqr` (src/scalar.rs:11-132; x*y*2^-256 mod p).
// An element is 8 x 32-bit little-endian limbs, i.e. the SAME 32 bytes as arkworks' 4 x u64 limbs in
// Montgomery form (R = 2^256), so host buffers cross the C-ABI without repacking.
//
// The multiplier is a register-resident CIOS Montgomery product on 32-bit limbs.  Partial products of the
// even and the odd limbs of `a` are accumulated in two separate 8-limb windows so that every row is ONE
// uninterrupted carry chain of mad.lo.cc / madc.hi.cc (ptxas fuses each lo/hi pair into IMAD.WIDE.U32 with
// carry predicates); the 32-bit right shift of every reduction step is free (the windows swap roles).
// No shared memory, no local memory, no tensor cores (integer modular arithmetic, no dense contraction).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace pk {

struct alignas(16) fr {
    uint32_t v[8];
};

// skyscraper/block-multiplier/src/constants.rs:3-8 (U64_P) as 32-bit limbs
#define PK_P0 0xf0000001u
#define PK_P1 0x43e1f593u
#define PK_P2 0x79b97091u
#define PK_P3 0x2833e848u
#define PK_P4 0x8181585du
#define PK_P5 0xb85045b6u
#define PK_P6 0xe131a029u
#define PK_P7 0x30644e72u
// low 32 bits of U64_NP0 = -p^-1 mod 2^64 (constants.rs:1)
#define PK_NP0 0xefffffffu

__device__ __forceinline__ fr fr_zero() {
    fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = 0;
    return r;
}
// R mod p (constants.rs:18-23): Montgomery form of 1
__device__ __forceinline__ fr fr_one() {
    fr r;
    r.v[0] = 0x4ffffffbu; r.v[1] = 0xac96341cu; r.v[2] = 0x9f60cd29u; r.v[3] = 0x36fc7695u;
    r.v[4] = 0x7879462eu; r.v[5] = 0x666ea36fu; r.v[6] = 0x9a07df2fu; r.v[7] = 0x0e0a77c1u;
    return r;
}
// R^2 mod p (constants.rs:26-31)
__device__ __forceinline__ fr fr_r2() {
    fr r;
    r.v[0] = 0xae216da7u; r.v[1] = 0x1bb8e645u; r.v[2] = 0xe35c59e3u; r.v[3] = 0x53fe3ab1u;
    r.v[4] = 0x53bb8085u; r.v[5] = 0x8c49833du; r.v[6] = 0x7f4e44a5u; r.v[7] = 0x0216d0b1u;
    return r;
}

__device__ __forceinline__ fr fr_load(const void* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    fr r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ fr fr_load_nc(const void* p) {  // read-only path (ld.global.nc)
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    fr r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ void fr_store(void* p, const fr& x) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
    q[1] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
}

__device__ __forceinline__ bool fr_is_zero(const fr& a) {
    return (a.v[0] | a.v[1] | a.v[2] | a.v[3] | a.v[4] | a.v[5] | a.v[6] | a.v[7]) == 0;
}

// r = a - p, returns borrow (1 if a < p)
__device__ __forceinline__ uint32_t sub_p(fr& r, const fr& a) {
    uint32_t borrow;
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
          "=r"(r.v[7]), "=r"(borrow)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(PK_P0), "r"(PK_P1), "r"(PK_P2), "r"(PK_P3), "r"(PK_P4), "r"(PK_P5), "r"(PK_P6), "r"(PK_P7));
    return borrow;  // 0 or 0xffffffff
}
// conditional final subtraction: a in [0, 2p) -> [0, p)
__device__ __forceinline__ fr fr_reduce_once(const fr& a) {
    fr t;
    uint32_t borrow = sub_p(t, a);
    fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = borrow ? a.v[i] : t.v[i];
    return r;
}

__device__ __forceinline__ fr fr_add(const fr& a, const fr& b) {
    fr s;
    asm("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32 %7, %15, %23;"
        : "=r"(s.v[0]), "=r"(s.v[1]), "=r"(s.v[2]), "=r"(s.v[3]), "=r"(s.v[4]), "=r"(s.v[5]), "=r"(s.v[6]),
          "=r"(s.v[7])
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    return fr_reduce_once(s);  // p < 2^254: no carry out of 256 bits
}

__device__ __forceinline__ fr fr_sub(const fr& a, const fr& b) {
    fr d;
    uint32_t borrow;
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=r"(d.v[0]), "=r"(d.v[1]), "=r"(d.v[2]), "=r"(d.v[3]), "=r"(d.v[4]), "=r"(d.v[5]), "=r"(d.v[6]),
          "=r"(d.v[7]), "=r"(borrow)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    // add back p & borrow-mask
    uint32_t m0 = PK_P0 & borrow, m1 = PK_P1 & borrow, m2 = PK_P2 & borrow, m3 = PK_P3 & borrow;
    uint32_t m4 = PK_P4 & borrow, m5 = PK_P5 & borrow, m6 = PK_P6 & borrow, m7 = PK_P7 & borrow;
    asm("add.cc.u32 %0, %0, %8;\n\t"
        "addc.cc.u32 %1, %1, %9;\n\t"
        "addc.cc.u32 %2, %2, %10;\n\t"
        "addc.cc.u32 %3, %3, %11;\n\t"
        "addc.cc.u32 %4, %4, %12;\n\t"
        "addc.cc.u32 %5, %5, %13;\n\t"
        "addc.cc.u32 %6, %6, %14;\n\t"
        "addc.u32 %7, %7, %15;"
        : "+r"(d.v[0]), "+r"(d.v[1]), "+r"(d.v[2]), "+r"(d.v[3]), "+r"(d.v[4]), "+r"(d.v[5]), "+r"(d.v[6]),
          "+r"(d.v[7])
        : "r"(m0), "r"(m1), "r"(m2), "r"(m3), "r"(m4), "r"(m5), "r"(m6), "r"(m7));
    return d;
}
__device__ __forceinline__ fr fr_dbl(const fr& a) { return fr_add(a, a); }
__device__ __forceinline__ fr fr_neg(const fr& a) { return fr_sub(fr_zero(), a); }

// ---- Montgomery multiplier building blocks (each one asm statement = one carry chain) ----------
// acc[0..7] = x0,x2,x4,x6 (4 limbs) * b : lo -> acc[2k], hi -> acc[2k+1]
__device__ __forceinline__ void mul4(uint32_t* acc, uint32_t x0, uint32_t x2, uint32_t x4, uint32_t x6, uint32_t b) {
    // mul.wide keeps each product ONE IMAD.WIDE (separate mul.lo / mul.hi are not fused by ptxas)
    asm("{\n\t"
        ".reg .u64 w;\n\t"
        "mul.wide.u32 w, %8, %12;\n\t"
        "mov.b64 {%0, %1}, w;\n\t"
        "mul.wide.u32 w, %9, %12;\n\t"
        "mov.b64 {%2, %3}, w;\n\t"
        "mul.wide.u32 w, %10, %12;\n\t"
        "mov.b64 {%4, %5}, w;\n\t"
        "mul.wide.u32 w, %11, %12;\n\t"
        "mov.b64 {%6, %7}, w;\n\t"
        "}"
        : "=r"(acc[0]), "=r"(acc[1]), "=r"(acc[2]), "=r"(acc[3]), "=r"(acc[4]), "=r"(acc[5]), "=r"(acc[6]),
          "=r"(acc[7])
        : "r"(x0), "r"(x2), "r"(x4), "r"(x6), "r"(b));
}
// acc[0..7] += x0,x2,x4,x6 * b as one carry chain; returns the carry out of acc[7]
__device__ __forceinline__ uint32_t mad4(uint32_t* acc, uint32_t x0, uint32_t x2, uint32_t x4, uint32_t x6,
                                         uint32_t b) {
    uint32_t c;
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, 0, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
          "+r"(acc[7]), "=r"(c)
        : "r"(x0), "r"(x2), "r"(x4), "r"(x6), "r"(b));
    return c;
}
// acc[0..7] += x0,x2,x4,x6 * b as one carry chain whose carry out lands in `top` (the limb above acc[7]) INSIDE the same
// chain: one IADD3.X.  Returning the carry and adding it in C++ (`top += mad4(..)`) costs three instructions per use — ptxas
// materialises the predicate as an increment, a predicated move and a pair-aligning move — i.e. 32 extra instructions per
// Montgomery product.
__device__ __forceinline__ void mad4_top(uint32_t* acc, uint32_t& top, uint32_t x0, uint32_t x2, uint32_t x4, uint32_t x6,
                                         uint32_t b) {
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
          "+r"(acc[7]), "+r"(top)
        : "r"(x0), "r"(x2), "r"(x4), "r"(x6), "r"(b));
}
// e0 += o[1] (carry c); o = (o >> 64) + x1,x3,x5,x7 * b + c   (one chain; top limbs start from zero)
__device__ __forceinline__ void mad4_rshift(uint32_t& e0, uint32_t* o, uint32_t x1, uint32_t x3, uint32_t x5,
                                            uint32_t x7, uint32_t b) {
    asm("add.cc.u32 %0, %0, %2;\n\t"
        "madc.lo.cc.u32 %1, %9, %13, %3;\n\t"
        "madc.hi.cc.u32 %2, %9, %13, %4;\n\t"
        "madc.lo.cc.u32 %3, %10, %13, %5;\n\t"
        "madc.hi.cc.u32 %4, %10, %13, %6;\n\t"
        "madc.lo.cc.u32 %5, %11, %13, %7;\n\t"
        "madc.hi.cc.u32 %6, %11, %13, %8;\n\t"
        "madc.lo.cc.u32 %7, %12, %13, 0;\n\t"
        "madc.hi.u32 %8, %12, %13, 0;"
        : "+r"(e0), "+r"(o[0]), "+r"(o[1]), "+r"(o[2]), "+r"(o[3]), "+r"(o[4]), "+r"(o[5]), "+r"(o[6]), "+r"(o[7])
        : "r"(x1), "r"(x3), "r"(x5), "r"(x7), "r"(b));
}

// one CIOS row: T += a*bi ; m = T0 * np0 ; T += m*p ; (the >>32 is the role swap done by the caller)
//   T = E + O * 2^32 with E limb k <-> e[k], O limb k <-> o[k]
template <bool FIRST>
__device__ __forceinline__ void mont_row(uint32_t* e, uint32_t* o, const fr& a, uint32_t bi) {
    if (FIRST) {
        mul4(o, a.v[1], a.v[3], a.v[5], a.v[7], bi);
        mul4(e, a.v[0], a.v[2], a.v[4], a.v[6], bi);
    } else {
        mad4_rshift(e[0], o, a.v[1], a.v[3], a.v[5], a.v[7], bi);
        mad4_top(e, o[7], a.v[0], a.v[2], a.v[4], a.v[6], bi);
    }
    uint32_t m = e[0] * PK_NP0;
    (void)mad4(o, PK_P1, PK_P3, PK_P5, PK_P7, m);  // limb 8 never overflows (p < 2^254)
    mad4_top(e, o[7], PK_P0, PK_P2, PK_P4, PK_P6, m);
}

// a * b * 2^-256 mod p.  LAZY = false: inputs in [0, p) (Montgomery form or raw, see skyscraper.cuh), output in [0, p).
// LAZY = true: a in [0, 4p], b in [0, p), output (a b + m p) / 2^256 < 4p * 0.19 + p < 2p WITHOUT the final conditional
// subtraction.  The rows are the same: every row's running sum is below (a + p) * 2^32 < 2^288 because 5p < 2^256, so the
// ninth limb (o[7] after the role swap) still never overflows.
template <bool LAZY = false>
__device__ __forceinline__ fr fr_mul_t(const fr& a, const fr& b) {
    uint32_t e[8], o[8];
    mont_row<true>(e, o, a, b.v[0]);
    mont_row<false>(o, e, a, b.v[1]);
    mont_row<false>(e, o, a, b.v[2]);
    mont_row<false>(o, e, a, b.v[3]);
    mont_row<false>(e, o, a, b.v[4]);
    mont_row<false>(o, e, a, b.v[5]);
    mont_row<false>(e, o, a, b.v[6]);
    mont_row<false>(o, e, a, b.v[7]);
    // after an even number of rows: result limb k = o[k+1] + e[k]   (o plays "E", e plays "O")
    fr r;
    asm("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32 %7, %15, 0;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
          "=r"(r.v[7])
        : "r"(e[0]), "r"(e[1]), "r"(e[2]), "r"(e[3]), "r"(e[4]), "r"(e[5]), "r"(e[6]), "r"(e[7]), "r"(o[1]),
          "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]));
    return LAZY ? r : fr_reduce_once(r);
}
__device__ __forceinline__ fr fr_mul(const fr& a, const fr& b) { return fr_mul_t<false>(a, b); }
__device__ __forceinline__ fr fr_mul_lazy(const fr& a, const fr& b) { return fr_mul_t<true>(a, b); }
// ---- lazily reduced forms for butterfly networks (ntt.cu): values live in [0, 2p] between butterflies ----------------
// 2p (skyscraper/core/src/constants.rs:9-16 MODULUS[2])
#define PK_2P0 0xe0000002u
#define PK_2P1 0x87c3eb27u
#define PK_2P2 0xf372e122u
#define PK_2P3 0x5067d090u
#define PK_2P4 0x0302b0bau
#define PK_2P5 0x70a08b6du
#define PK_2P6 0xc2634053u
#define PK_2P7 0x60c89ce5u
// a in [0, 4p] -> [0, 2p]: one conditional subtraction of 2p
__device__ __forceinline__ fr fr_reduce_2p_once(const fr& a) {
    fr t;
    uint32_t borrow;
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=r"(t.v[0]), "=r"(t.v[1]), "=r"(t.v[2]), "=r"(t.v[3]), "=r"(t.v[4]), "=r"(t.v[5]), "=r"(t.v[6]), "=r"(t.v[7]), "=r"(borrow)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(PK_2P0), "r"(PK_2P1), "r"(PK_2P2), "r"(PK_2P3), "r"(PK_2P4), "r"(PK_2P5), "r"(PK_2P6), "r"(PK_2P7));
    fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = borrow ? a.v[i] : t.v[i];
    return r;
}
// a, b in [0, 2p] -> a + b in [0, 2p] (a + b <= 4p < 2^256)
__device__ __forceinline__ fr fr_add_lazy(const fr& a, const fr& b) {
    fr s;
    asm("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32 %7, %15, %23;"
        : "=r"(s.v[0]), "=r"(s.v[1]), "=r"(s.v[2]), "=r"(s.v[3]), "=r"(s.v[4]), "=r"(s.v[5]), "=r"(s.v[6]), "=r"(s.v[7])
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    return fr_reduce_2p_once(s);
}
// a, b in [0, 2p] -> a - b + 2p in [0, 4p]: NOT reduced (the twiddle product that follows accepts it)
__device__ __forceinline__ fr fr_sub_lazy(const fr& a, const fr& b) {
    fr d;
    asm("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32 %7, %15, %23;"
        : "=r"(d.v[0]), "=r"(d.v[1]), "=r"(d.v[2]), "=r"(d.v[3]), "=r"(d.v[4]), "=r"(d.v[5]), "=r"(d.v[6]), "=r"(d.v[7])
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(PK_2P0), "r"(PK_2P1), "r"(PK_2P2), "r"(PK_2P3), "r"(PK_2P4), "r"(PK_2P5), "r"(PK_2P6), "r"(PK_2P7));
    asm("sub.cc.u32 %0, %0, %8;\n\t"
        "subc.cc.u32 %1, %1, %9;\n\t"
        "subc.cc.u32 %2, %2, %10;\n\t"
        "subc.cc.u32 %3, %3, %11;\n\t"
        "subc.cc.u32 %4, %4, %12;\n\t"
        "subc.cc.u32 %5, %5, %13;\n\t"
        "subc.cc.u32 %6, %6, %14;\n\t"
        "subc.u32 %7, %7, %15;"
        : "+r"(d.v[0]), "+r"(d.v[1]), "+r"(d.v[2]), "+r"(d.v[3]), "+r"(d.v[4]), "+r"(d.v[5]), "+r"(d.v[6]), "+r"(d.v[7])
        : "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    return d;
}
// [0, 2p] -> 2p - a in [0, 2p]
__device__ __forceinline__ fr fr_neg_lazy(const fr& a) {
    fr z = fr_zero();
    return fr_reduce_2p_once(fr_sub_lazy(z, a));  // 2p - a <= 2p: the reduction only maps 2p -> 0
}
// [0, 2p] -> canonical representative in [0, p)
__device__ __forceinline__ fr fr_normalize_2p(const fr& a) { return fr_reduce_once(fr_reduce_once(a)); }
#include "fr_sqr.inc"

// canonical integer < p  ->  Montgomery form
__device__ __forceinline__ fr fr_to_mont(const fr& c) { return fr_mul(c, fr_r2()); }
// Montgomery form -> canonical integer (reduction rows only, fr_sqr.inc)
__device__ __forceinline__ fr fr_from_mont(const fr& a) { return fr_redc(a); }
// any 256-bit value -> [0, p): 2^256 / p < 5.3, so at most 5 subtractions
__device__ __forceinline__ fr fr_reduce_any(fr a) {
#pragma unroll 1
    for (int i = 0; i < 5; i++) {
        fr t;
        uint32_t borrow = sub_p(t, a);
        if (borrow) break;
        a = t;
    }
    return a;
}

}  // namespace pk
