#!/bin/bash
# GPU session C: parity with the hi+lo latent weight, launch list + ncu --set full of the GEMM family (raw CSV only: the
# .ncu-rep of 82 launches is 130 MB, gpurun_out is capped at 64 MiB) and of the one-pass attention backward, bench
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "latent" > $O/r2c_ops.log 2>&1; echo "rc=$?" >> $O/r2c_ops.log
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -s > $O/r2c_parity.log 2>&1; echo "parity rc=$?" >> $O/r2c_parity.log
ARGS="--steps 1 --warmup 3 --no-attribution --no-cpu-baseline --no-gpu-baseline"
timeout 300 python bench.py $ARGS > $O/r2c_bench_step_only.json 2> $O/r2c_bench_step_only.err &&
CTC_BENCH_PROFILE_RANGE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off \
    --csv --log-file $O/r2c_launches.csv python bench.py $ARGS > $O/r2c_ncu1.log 2>&1
timeout 200 python tools/prof_step.py 8 > $O/r2c_prof_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:gemm_tcgen05 -c 90 \
    -o /tmp/r2c_gemm python tools/prof_step.py 8 > $O/r2c_ncu2.log 2>&1
ncu -i /tmp/r2c_gemm.ncu-rep --page raw --csv > $O/r2c_gemm_raw.csv 2> $O/r2c_gemm_raw.err
CTC_ATTN_BWD=2 timeout 200 python tools/prof_step.py 8 > $O/r2c_prof_plain2.log 2>&1 &&
CTC_ATTN_BWD=2 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:onepass -c 1 \
    -o $O/r2c_onepass python tools/prof_step.py 8 > $O/r2c_ncu3.log 2>&1
timeout 700 python bench.py --steps 10 --warmup 3 > $O/r2c_bench.json 2> $O/r2c_bench.err; echo "bench rc=$?" >> $O/r2c_bench.err
du -sh $O; ls -la $O | head -30
tail -3 $O/r2c_ops.log; tail -4 $O/r2c_parity.log; tail -2 $O/r2c_bench.err
