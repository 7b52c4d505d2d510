// tools/coissue_bench.cu — how much squaring throughput does a SECOND multiplier on the FP64 pipe add on sm_100a?
//
// The integer squaring (fr_sqr, fr_sqr.inc) keeps the IMAD.WIDE pipe ~80 % busy; the reference's own trick on aarch64
// (skyscraper/block-multiplier: scalar multiplier interleaved with an FP "52-bit limb" multiplier,
// src/portable_simd.rs) suggests running a second squaring on the FP64 pipe next to it.  This
// benchmark answers "is that worth building" before anyone builds it: it co-issues the real integer squaring with an
// instruction-faithful PROXY of a 5 x 52-bit-limb FP squaring (same DFMA / DADD / 64-bit integer add / mask counts and
// dependency shape as a real one: 15 limb products + 25 reduction products, each a fma.rz hi/lo pair whose raw bit
// patterns are accumulated in 64-bit integer columns, 5 reduction rounds, re-normalisation to doubles).  The proxy's
// VALUES are not a correct square (no exactness proof, no final carries) — only its instruction mix is real.
//   modes: int2 = 2 integer chains / thread (the shipping kernel's regime)     fp2 = 2 proxy chains
//          mix11 = 1 integer + 1 proxy chain                                    mix21 = 2 integer + 1 proxy
// Output: squarings per second per mode; the gain of mixNM over int2 is the upper bound of the hybrid's benefit.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Iprovekit_b200/csrc -o build/coissue_bench tools/coissue_bench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#include "fr.cuh"

using namespace pk;

#define ITERS 512

struct fp5 {
    double l[5];
};

__device__ __forceinline__ double fma_rz(double a, double b, double c) {
    double r;
    asm("fma.rz.f64 %0, %1, %2, %3;" : "=d"(r) : "d"(a), "d"(b), "d"(c));
    return r;
}
__device__ __forceinline__ uint64_t bits(double x) { return (uint64_t)__double_as_longlong(x); }
__device__ __forceinline__ double from_limb(uint64_t v) {  // exact int < 2^52 -> double: OR into the mantissa of 2^52, subtract
    return __longlong_as_double((long long)(v | 0x4330000000000000ull)) - 4503599627370496.0;
}

// instruction-faithful proxy of a 5 x 52-bit limb Montgomery squaring on the FP64 pipe (see header)
__device__ __forceinline__ fp5 fp_sqr_proxy(const fp5& a, const double* __restrict__ p52, double np0) {
    const double C1 = 20282409603651670423947251286016.0;             // 2^104
    const double C2 = 20282409603651670423947251286016.0 + 4503599627370496.0;  // 2^104 + 2^52
    const uint64_t MASK = (1ull << 52) - 1;
    uint64_t col[10];
#pragma unroll
    for (int k = 0; k < 10; k++) col[k] = 0x10ull * k;  // stands for the pre-biased accumulators (make_initial)
    // 15 limb products, off-diagonal ones counted twice
#pragma unroll
    for (int i = 0; i < 5; i++)
#pragma unroll
        for (int j = i; j < 5; j++) {
            double hi = fma_rz(a.l[i], a.l[j], C1);
            double lo = fma_rz(a.l[i], a.l[j], C2 - hi);
            uint64_t bh = bits(hi), bl = bits(lo);
            if (i != j) {
                bh += bh;
                bl += bl;
            }
            col[i + j + 1] += bh;
            col[i + j] += bl;
        }
    // 5 reduction rounds: m = low 52 bits of col[i] * np0, col += m * p, carry into the next column
#pragma unroll
    for (int i = 0; i < 5; i++) {
        double ci = from_limb(col[i] & MASK);
        double mh = fma_rz(ci, np0, C1);
        double m = from_limb(bits(fma_rz(ci, np0, C2 - mh)) & MASK);
#pragma unroll
        for (int j = 0; j < 5; j++) {
            double hi = fma_rz(m, p52[j], C1);
            double lo = fma_rz(m, p52[j], C2 - hi);
            col[i + j + 1] += bits(hi);
            col[i + j] += bits(lo);
        }
        col[i + 1] += col[i] >> 52;
    }
    // re-normalise the high half to 52-bit limbs held as doubles
    fp5 r;
    uint64_t carry = 0;
#pragma unroll
    for (int k = 0; k < 5; k++) {
        uint64_t v = col[5 + k] + carry;
        carry = v >> 52;
        r.l[k] = from_limb(v & MASK);
    }
    return r;
}

template <int NI, int NF>
__global__ void __launch_bounds__(256) k_mix(fr* data, double* fdata, const double* __restrict__ p52, double np0, int iters) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    fr x[NI > 0 ? NI : 1];
    fp5 f[NF > 0 ? NF : 1];
#pragma unroll
    for (int k = 0; k < NI; k++) x[k] = fr_load(&data[NI * t + k]);
#pragma unroll
    for (int k = 0; k < NF; k++)
#pragma unroll
        for (int q = 0; q < 5; q++) f[k].l[q] = fdata[(NF * t + k) * 5 + q];
    double pl[5];
#pragma unroll
    for (int q = 0; q < 5; q++) pl[q] = p52[q];
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < NI; k++) x[k] = fr_sqr(x[k]);
#pragma unroll
        for (int k = 0; k < NF; k++) f[k] = fp_sqr_proxy(f[k], pl, np0);
    }
#pragma unroll
    for (int k = 0; k < NI; k++) fr_store(&data[NI * t + k], x[k]);
#pragma unroll
    for (int k = 0; k < NF; k++)
#pragma unroll
        for (int q = 0; q < 5; q++) fdata[(NF * t + k) * 5 + q] = f[k].l[q];
}

template <int NI, int NF>
static double run(const char* name, fr* data, double* fdata, const double* p52, int sms, double base) {
    int grid = sms * 8, block = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_mix<NI, NF><<<grid, block>>>(data, fdata, p52, 123457.0, 8);
    cudaDeviceSynchronize();
    float ms = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        k_mix<NI, NF><<<grid, block>>>(data, fdata, p52, 123457.0, ITERS);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float tt;
        cudaEventElapsedTime(&tt, e0, e1);
        if (tt < ms) ms = tt;
    }
    cudaFuncAttributes at;
    cudaFuncGetAttributes(&at, k_mix<NI, NF>);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_mix<NI, NF>, block, 0);
    double sq = (double)grid * block * ITERS * (NI + NF) / (ms * 1e-3);
    printf("{\"mode\": \"%s\", \"int_chains\": %d, \"fp_proxy_chains\": %d, \"regs\": %d, \"blocks_per_sm\": %d, \"ms\": %.3f, "
           "\"gsqr_s\": %.2f, \"int_gsqr_s\": %.2f, \"vs_int2\": %.3f}\n",
           name, NI, NF, at.numRegs, occ, ms, sq / 1e9, sq / 1e9 * NI / (NI + NF), base > 0 ? sq / 1e9 / base : 1.0);
    return sq / 1e9;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"note\": \"fp chains are an instruction-mix proxy, not a verified squaring\"}\n", p.name, sms);
    size_t threads = (size_t)sms * 8 * 256;
    fr* data;
    double *fdata, *p52;
    cudaMalloc(&data, threads * 2 * sizeof(fr));
    cudaMalloc(&fdata, threads * 2 * 5 * sizeof(double));
    cudaMalloc(&p52, 5 * sizeof(double));
    cudaMemset(data, 0x11, threads * 2 * sizeof(fr));
    {
        // limbs < 2^52 as doubles
        size_t n = threads * 2 * 5;
        double* h = (double*)malloc(n * sizeof(double));
        for (size_t i = 0; i < n; i++) h[i] = (double)((i * 2654435761ull) & ((1ull << 52) - 1));
        cudaMemcpy(fdata, h, n * sizeof(double), cudaMemcpyHostToDevice);
        free(h);
        // BN254-Fr modulus in 52-bit limbs
        const uint64_t P[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
        double pl[5];
        for (int k = 0; k < 5; k++) {
            int bit = 52 * k, w = bit / 64, s = bit % 64;
            uint64_t v = P[w] >> s;
            if (s > 12 && w + 1 < 4) v |= P[w + 1] << (64 - s);
            pl[k] = (double)(v & ((1ull << 52) - 1));
        }
        cudaMemcpy(p52, pl, sizeof pl, cudaMemcpyHostToDevice);
    }
    double base = run<2, 0>("int2", data, fdata, p52, sms, 0);
    run<1, 0>("int1", data, fdata, p52, sms, base);
    run<0, 2>("fp2", data, fdata, p52, sms, base);
    run<0, 1>("fp1", data, fdata, p52, sms, base);
    run<1, 1>("mix11", data, fdata, p52, sms, base);
    run<2, 1>("mix21", data, fdata, p52, sms, base);
    run<1, 2>("mix12", data, fdata, p52, sms, base);
    return 0;
}
