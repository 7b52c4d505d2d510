"""Per-kernel timings on one B200 (device-resident inputs, CUDA events on the library's stream).
Usage: python tools/microbench.py [--m 21]   -> prints one JSON object per kernel."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import provekit_b200 as pk  # noqa: E402

HBM_PEAK = 6464.9
try:
    HBM_PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def rand_fr(rng, n):
    a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64(0x0FFFFFFFFFFFFFFF)
    return a


def timeit(ctx, stream, fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        ctx.sync()
        ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=21)
    args = ap.parse_args()
    m = args.m
    ctx = pk.Context(0)
    stream = torch.cuda.ExternalStream(ctx.stream)
    rng = np.random.default_rng(1)
    out = []

    # modmul ceiling
    nthreads = 148 * 2048 * 4
    iters = 512
    ms = min(ctx.modmul_bench(nthreads, iters) for _ in range(3))
    out.append(dict(kernel="modmul_bench", ms=ms, gmodmul_s=nthreads * iters * 2 / ms / 1e6))
    ms = min(ctx.modmul_bench(nthreads, iters, True) for _ in range(3))
    out.append(dict(kernel="modsqr_bench", ms=ms, gmodsqr_s=nthreads * iters * 2 / ms / 1e6))

    n = 1 << m
    # compress_many on device data
    nm = 1 << 22
    msgs = ctx.upload(rand_fr(rng, 2 * nm))
    hashes = ctx.buffer(nm)
    best, med = timeit(ctx, stream, lambda: ctx._chk(ctx.L.pk_skyscraper_compress_many_dev(ctx.h, msgs.h, hashes.h, nm)))
    out.append(dict(kernel="compress_many", n=nm, ms=best, ms_median=med, mcompress_s=nm / best / 1e3,
                    alg_gbs=96 * nm / best / 1e6, frac_hbm=96 * nm / best / 1e6 / HBM_PEAK))
    msgs.free()
    hashes.free()

    # RS encode of one polynomial (n coeffs, rate 1/2)
    coeffs = ctx.upload(rand_fr(rng, n))
    L = n * 2 // 16
    leaves = ctx.buffer(L * 32)
    best, med = timeit(ctx, stream, lambda: ctx.rs_encode(coeffs, m, 1, leaves, 32, 0))
    nmul = n * (m - 3)  # SURVEY 8d radix-2 count for rate 1/2
    out.append(dict(kernel="rs_encode", log_n=m, ms=best, ms_median=med, alg_gbs=96 * n / best / 1e6,
                    frac_hbm=96 * n / best / 1e6 / HBM_PEAK, gmodmul_s=nmul / best / 1e6))
    ctx.rs_encode(coeffs, m, 1, leaves, 32, 16)
    # Merkle over L leaves x 32
    nodes = ctx.buffer(2 * L)
    best, med = timeit(ctx, stream, lambda: ctx.merkle_build(leaves, L, 32, nodes), reps=3, warm=1)
    by = 32 * (L * 32 + 2 * L - 1)
    ncomp = L * 31 + L - 1
    out.append(dict(kernel="merkle", leaves=L, width=32, ms=best, ms_median=med, alg_gbs=by / best / 1e6,
                    frac_hbm=by / best / 1e6 / HBM_PEAK, mcompress_s=ncomp / best / 1e3))
    # wavelet
    best, med = timeit(ctx, stream, lambda: ctx.evals_to_coeffs(coeffs, m))
    out.append(dict(kernel="wavelet", log_n=m, ms=best, ms_median=med, alg_gbs=64 * n / best / 1e6,
                    frac_hbm=64 * n / best / 1e6 / HBM_PEAK))
    leaves.free()
    nodes.free()

    # zk sumcheck: all rounds on 4 arrays of 2^(m-1)
    m0 = m - 1
    N = 1 << m0
    arrs = [ctx.upload(rand_fr(rng, N)) for _ in range(4)]
    fold = rand_fr(rng, 1)

    def zk_all():
        cur = m0
        ctx.sumcheck_fold_map_reduce(*arrs, cur, None)
        for _ in range(m0 - 1):
            ctx.sumcheck_fold_map_reduce(*arrs, cur, fold)
            cur -= 1
    best, med = timeit(ctx, stream, zk_all, reps=3, warm=1)
    by = 128 * (N + sum(N / 2 ** (i - 1) + N / 2 ** i for i in range(1, m0)))
    out.append(dict(kernel="zk_sumcheck_all_rounds", log_n=m0, ms=best, ms_median=med, alg_gbs=by / best / 1e6,
                    frac_hbm=by / best / 1e6 / HBM_PEAK))
    # first two rounds alone (the HBM-relevant ones)
    best, _ = timeit(ctx, stream, lambda: ctx.sumcheck_fold_map_reduce(*arrs, m0, None), reps=5, warm=1)
    out.append(dict(kernel="zk_sumcheck_round0", log_n=m0, ms=best, alg_gbs=128 * N / best / 1e6,
                    frac_hbm=128 * N / best / 1e6 / HBM_PEAK))
    best, _ = timeit(ctx, stream, lambda: ctx.sumcheck_fold_map_reduce(*arrs, m0, fold), reps=5, warm=1)
    out.append(dict(kernel="zk_sumcheck_round1_fold", log_n=m0, ms=best, alg_gbs=128 * 1.5 * N / best / 1e6,
                    frac_hbm=128 * 1.5 * N / best / 1e6 / HBM_PEAK))
    for a in arrs:
        a.free()

    # whir sumcheck: 4 rounds on p, w of 2^m
    p, w = ctx.upload(rand_fr(rng, n)), ctx.upload(rand_fr(rng, n))
    p2, w2 = ctx.buffer(n // 2), ctx.buffer(n // 2)

    def whir4():
        ctx.whir_sumcheck_round(p, w, m)
        ctx.whir_sumcheck_round(p, w, m, fold, p2, w2)
        ctx.whir_sumcheck_round(p2, w2, m - 1, fold, p, w)
        ctx.whir_sumcheck_round(p, w, m - 2, fold, p2, w2)
    best, med = timeit(ctx, stream, whir4, reps=3, warm=1)
    by = 64 * (n + sum(n / 2 ** (j - 1) + n / 2 ** j for j in range(1, 4)))
    out.append(dict(kernel="whir_sumcheck_4rounds", log_n=m, ms=best, ms_median=med, alg_gbs=by / best / 1e6,
                    frac_hbm=by / best / 1e6 / HBM_PEAK))
    for r in out:
        print(json.dumps(r))
    ctx.close()


if __name__ == "__main__":
    main()
