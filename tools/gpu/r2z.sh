#!/bin/bash
# GPU session Z (final evidence of round 2): full GPU suite (parity numbers are written by the tests), default bench, ncu launch list of
# one step, ncu --set full of every GEMM launch of one step (raw CSV), kernel micro-benchmarks
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > $O/r2z_tests.log 2>&1; echo "rc=$?" >> $O/r2z_tests.log
tail -4 $O/r2z_tests.log
timeout 900 python bench.py > $O/r2z_bench.json 2> $O/r2z_bench.err; echo "bench rc=$?" >> $O/r2z_bench.err
tail -2 $O/r2z_bench.err; cut -c1-300 $O/r2z_bench.json
ARGS="--steps 1 --warmup 3 --no-attribution --no-cpu-baseline --no-gpu-baseline"
timeout 300 python bench.py $ARGS > $O/r2z_bench_step_only.json 2> $O/r2z_bench_step_only.err &&
CTC_BENCH_PROFILE_RANGE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off \
    --csv --log-file $O/r2z_launches.csv python bench.py $ARGS > $O/r2z_ncu1.log 2>&1
timeout 200 python tools/prof_step.py 8 > $O/r2z_prof_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:gemm_tcgen05 -c 90 \
    -o /tmp/r2z_gemm python tools/prof_step.py 8 > $O/r2z_ncu2.log 2>&1
ncu -i /tmp/r2z_gemm.ncu-rep --page raw --csv > $O/r2z_gemm_raw.csv 2> $O/r2z_gemm_raw.err
timeout 300 python tools/kernel_bench.py all > $O/r2z_kbench.log 2>&1
du -sh $O; tail -3 $O/r2z_ncu1.log $O/r2z_ncu2.log 2>/dev/null | tail -8
