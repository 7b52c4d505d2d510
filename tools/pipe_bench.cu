// tools/pipe_bench.cu — issue-rate microbenchmark of the sm_100a pipes a 256-bit modular multiplier can use:
// IMAD.WIDE with and without carry predicates (fmaheavy), 32-bit IMAD, DFMA (fp64, the pipe the reference's NEON
// path uses for 52-bit limbs, skyscraper/block-multiplier/src/portable_simd.rs), IADD3 (alu), and DFMA + IMAD.WIDE
// co-issue.  Result (profiles/r01_pipe_bench.jsonl): IMAD.WIDE = 32 / clk / SM in every carry form, DFMA = 23-43 / clk / SM
// depending on operand reuse, so an FP64 multiplier would not beat the integer one; see DESIGN.md section 5.
// Rates are per REAL SM clock (clock64 / globaltimer inside the kernel), so power throttling does not distort them.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/pipe_bench tools/pipe_bench.cu && ./build/pipe_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048

__device__ __forceinline__ uint64_t gtime() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

struct Stamp {
    long long clk;
    uint64_t ns;
};

// 4 wide multiply-adds in ONE carry chain over four aligned 64-bit accumulators
#define CHAIN4(W0, W1, W2, W3, X0, X1, X2, X3, B)                                                               \
    asm volatile(                                                                                               \
        "{\n\t.reg .u32 a0,a1,a2,a3,a4,a5,a6,a7;\n\t"                                                          \
        "mov.b64 {a0,a1}, %0;\n\tmov.b64 {a2,a3}, %1;\n\tmov.b64 {a4,a5}, %2;\n\tmov.b64 {a6,a7}, %3;\n\t"       \
        "mad.lo.cc.u32 a0, %4, %8, a0;\n\tmadc.hi.cc.u32 a1, %4, %8, a1;\n\t"                                   \
        "madc.lo.cc.u32 a2, %5, %8, a2;\n\tmadc.hi.cc.u32 a3, %5, %8, a3;\n\t"                                  \
        "madc.lo.cc.u32 a4, %6, %8, a4;\n\tmadc.hi.cc.u32 a5, %6, %8, a5;\n\t"                                  \
        "madc.lo.cc.u32 a6, %7, %8, a6;\n\tmadc.hi.u32 a7, %7, %8, a7;\n\t"                                     \
        "mov.b64 %0, {a0,a1};\n\tmov.b64 %1, {a2,a3};\n\tmov.b64 %2, {a4,a5};\n\tmov.b64 %3, {a6,a7};\n\t}"      \
        : "+l"(W0), "+l"(W1), "+l"(W2), "+l"(W3)                                                                \
        : "r"(X0), "r"(X1), "r"(X2), "r"(X3), "r"(B))

template <int MODE>
__global__ void __launch_bounds__(256) k_pipe(uint64_t* out, Stamp* stamps, double da, double db, uint32_t ia,
                                              uint32_t ib) {
    const uint32_t t = threadIdx.x;
    uint32_t e[8], o[8], x[8];
    uint64_t w[8];
    uint32_t y[16];
#pragma unroll
    for (int k = 0; k < 16; k++) y[k] = ib * (k + 3) + t * 97u;
    double d[8], f[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        e[k] = ia + k * 77u + t;
        o[k] = ib + k * 131u + t * 3u;
        x[k] = (ia ^ ib) + k + t * 2654435761u;
        w[k] = ((uint64_t)e[k] << 32) | o[k];
        d[k] = da + k + t;
        f[k] = db + 2 * k + t;
    }
    long long c0 = clock64();
    uint64_t n0 = gtime();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        if (MODE == 0) {  // 8 plain IMAD.WIDE, independent 64-bit accumulators
#pragma unroll
            for (int k = 0; k < 8; k++)
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[k]) : "r"(x[k]), "r"(x[(k + 1) & 7]));
        }
        if (MODE == 2) {  // 8 x 32-bit IMAD (lo)
#pragma unroll
            for (int k = 0; k < 8; k++)
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(e[k]) : "r"(x[k]), "r"(x[(k + 1) & 7]));
        }
        if (MODE == 3) {  // 8 x IMAD.HI
#pragma unroll
            for (int k = 0; k < 8; k++)
                asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(e[k]) : "r"(x[k]), "r"(x[(k + 1) & 7]));
        }
        if (MODE == 4) {  // 8 x DFMA, addend an immediate-encodable constant (2^104)
#pragma unroll
            for (int k = 0; k < 8; k++) {
                double r;
                asm volatile("fma.rz.f64 %0, %1, %2, 0d4670000000000000;" : "=d"(r) : "d"(d[k]), "d"(f[k]));
                d[k] = __longlong_as_double((__double_as_longlong(r) & 0x000fffffffffffffll) | 0x3ff0000000000000ll);
            }
        }
        if (MODE == 5) {  // 8 x DFMA, three register operands
#pragma unroll
            for (int k = 0; k < 8; k++)
                asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(d[k]) : "d"(f[k]), "d"(f[(k + 1) & 7]));
        }
        if (MODE == 8) {  // 8 x (IADD3 + IADD3.X): 64-bit three-input adds
#pragma unroll
            for (int k = 0; k < 8; k++) w[k] += w[(k + 1) & 7] + (uint64_t)x[k];
        }
        if (MODE == 11 || MODE == 12 || MODE == 13) {
            // 16 x 8 wide multiply-adds, straight line, so loop-boundary register shuffles amortise:
            //   11 = carry-free (plain IMAD.WIDE), 12 = one carry chain per 4 (IMAD.WIDE.X), 13 = carry-out only
            //   (IMAD.WIDE with a predicate output, carries collected by IADD3.X on the alu pipe)
#pragma unroll
            for (int u = 0; u < 16; u++) {
                if (MODE == 11) {
#pragma unroll
                    for (int k = 0; k < 8; k++)
                        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[k]) : "r"(x[(k + u) & 7]), "r"(y[u]));
                }
                if (MODE == 12) {
                    CHAIN4(w[0], w[1], w[2], w[3], x[u & 7], x[(u + 1) & 7], x[(u + 2) & 7], x[(u + 3) & 7], (uint32_t)w[4]);
                    CHAIN4(w[4], w[5], w[6], w[7], x[(u + 4) & 7], x[(u + 5) & 7], x[(u + 6) & 7], x[(u + 7) & 7], (uint32_t)w[0]);
                }
                if (MODE == 13) {
#pragma unroll
                    for (int k = 0; k < 8; k++)
                        asm volatile(
                            "{\n\t.reg .u32 lo, hi;\n\tmov.b64 {lo,hi}, %0;\n\t"
                            "mad.lo.cc.u32 lo, %2, %3, lo;\n\tmadc.hi.cc.u32 hi, %2, %3, hi;\n\taddc.u32 %1, %1, 0;\n\t"
                            "mov.b64 %0, {lo,hi};\n\t}"
                            : "+l"(w[k]), "+r"(e[k])
                            : "r"(x[(k + u) & 7]), "r"((uint32_t)w[(k + 1) & 7]));
                }
            }
        }
        if (MODE == 9) {  // 8 DFMA (imm addend) + 2 integer chains (8 wide), independent: co-issue test
#pragma unroll
            for (int k = 0; k < 8; k++)
                asm volatile("fma.rz.f64 %0, %1, %2, 0d4670000000000000;" : "=d"(d[k]) : "d"(d[k]), "d"(f[k]));
            CHAIN4(w[0], w[1], w[2], w[3], x[0], x[1], x[2], x[3], x[4]);
            CHAIN4(w[4], w[5], w[6], w[7], x[4], x[5], x[6], x[7], x[0]);
        }
    }
    uint64_t n1 = gtime();
    long long c1 = clock64();
    uint64_t acc = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) acc += (uint64_t)__double_as_longlong(d[k]) + w[k] + e[k] + o[k];
    out[blockIdx.x * blockDim.x + t] = acc;
    if (t == 0) {
        stamps[blockIdx.x].clk = c1 - c0;
        stamps[blockIdx.x].ns = n1 - n0;
    }
}

template <int MODE>
static void run(const char* name, int ops_per_iter, uint64_t* out, Stamp* stamps, int sms) {
    int grid = sms * 8, block = 256;  // 8 resident CTAs of 8 warps per SM = one wave
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_pipe<MODE><<<grid, block>>>(out, stamps, 1.5, 3.0, 12345u, 7u);
    cudaDeviceSynchronize();
    float ms = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        k_pipe<MODE><<<grid, block>>>(out, stamps, 1.5, 3.0, 12345u, 7u);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float t;
        cudaEventElapsedTime(&t, e0, e1);
        if (t < ms) ms = t;
    }
    static Stamp h[4096];
    cudaMemcpy(h, stamps, sizeof(Stamp) * grid, cudaMemcpyDeviceToHost);
    double clk = 0, ns = 0;
    for (int i = 0; i < grid; i++) {
        clk += (double)h[i].clk;
        ns += (double)h[i].ns;
    }
    double mhz = clk / ns * 1e3;
    double ops = (double)grid * block * ITERS * ops_per_iter;
    double per_clk_sm = ops / (ms * 1e-3) / (mhz * 1e6) / sms;  // wall time x measured clock
    printf("{\"mode\": \"%s\", \"ms\": %.3f, \"sm_mhz\": %.0f, \"ops_per_clk_sm\": %.2f, \"gops_s\": %.1f}\n", name, ms,
           mhz, per_clk_sm, ops / ms / 1e6);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d}\n", p.name, sms);
    uint64_t* out;
    Stamp* stamps;
    cudaMalloc(&out, (size_t)sms * 8 * 256 * 8);
    cudaMalloc(&stamps, sizeof(Stamp) * sms * 8);
    run<8>("warmup", 8, out, stamps, sms);
    run<0>("imad_wide_plain", 8, out, stamps, sms);
    run<2>("imad_lo32", 8, out, stamps, sms);
    run<3>("imad_hi32", 8, out, stamps, sms);
    run<4>("dfma_imm_addend(+2lop3)", 8, out, stamps, sms);
    run<5>("dfma_3reg", 8, out, stamps, sms);
    run<8>("add64_3input", 8, out, stamps, sms);
    run<9>("dfma_imm x8 + imad_chain x8 (counted: 16)", 16, out, stamps, sms);
    run<11>("wide_plain_straightline", 128, out, stamps, sms);
    run<12>("wide_carry_chain4_straightline", 128, out, stamps, sms);
    run<13>("wide_carry_out_only+iadd3x_straightline (counted: wide)", 128, out, stamps, sms);
    return 0;
}
