#!/bin/bash
# Round-end evidence on ONE B200 (run under gpurun): bench lines, microbenchmarks, ncu launch list + full capture of the
# dominant kernels, compute-sanitizer passes.  Everything lands in gpurun_out/f_*; tools/summarize_ncu.py and the commit
# copy the tracked summaries into profiles/.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err
python bench.py --workload sha256-substitute --no-cpu-baseline > gpurun_out/f_bench_sha.json 2> gpurun_out/f_bench_sha.err
python bench.py --workload synthetic-2p22 --no-cpu-baseline --steps 8 --in-flight 8 > gpurun_out/f_bench_2p22.json 2> gpurun_out/f_bench_2p22.err
python tools/microbench.py > gpurun_out/f_micro.jsonl 2> gpurun_out/f_micro.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/f_launches.csv python tools/profile_step.py > gpurun_out/f_prof.log 2>&1
# gpurun brings back at most 64 MiB: the 23-kernel capture goes without the source pages, the dominant kernel gets its own
# capture with --import-source on
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:"k_merkle_leaves|k_ntt_r8|k_ntt_pass|k_merkle_reduce|k_zk_sumcheck|k_whir_sumcheck|k_wavelet_tile" -c 23 -f -o gpurun_out/f_full python tools/profile_step.py > gpurun_out/f_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_merkle_leaves" -c 1 -f -o gpurun_out/f_full_merkle_src python tools/profile_step.py > gpurun_out/f_full_merkle.log 2>&1
ncu -i gpurun_out/f_full.ncu-rep --page raw --csv > gpurun_out/f_full_raw.csv 2> /dev/null
rm -f gpurun_out/f_full.ncu-rep  # 59 MB; the raw page and the single-kernel report with sources travel instead
ls -la gpurun_out/
if [ -z "$SKIP_SANITIZER" ]; then
{
echo "== memcheck: kernels (small and medium shapes) + prover + transcripts"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_prover.py tests/test_gpu_transcript.py -q -m gpu -x \
  -k "not (20-1 or 21-1 or 22-1 or 23-1 or 16-1-2 or 16-4 or 18-1 or full_size or 15 or 5000 or 1000 or hot_column or sharded_layout or exhaust)" 2>&1 | tail -4
echo "rc=$?"
echo "== racecheck: TMA radix-8 NTT, radix-2 NTT, Merkle reduce, wavelet, commit/open"
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x \
  -k "(rs_encode_vs_oracle and (11-1 or 12-2 or 13-7 or 14-1 or 9-10)) or (commit_batch_and_open and (8-1 or 12-1)) or (sharded_opening and 8-1) or merkle_vs_oracle" 2>&1 | tail -4
echo "== synccheck: one full proof with the device transcript"
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_prover.py -q -m gpu -x -k "100-300 or enqueue_collect" 2>&1 | tail -4
} > gpurun_out/f_sanitizer.txt 2>&1
fi
tail -1 gpurun_out/f_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],2), round(d['e2e']['value'],2), round(d['ms_per_step_one_in_flight'],2), d['roofline']['frac'], d['cpu_baseline'])"
[ -z "$SKIP_SANITIZER" ] && cat gpurun_out/f_sanitizer.txt
du -sh gpurun_out
