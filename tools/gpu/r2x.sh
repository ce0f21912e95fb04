#!/bin/bash
# GPU session X: bench sanity run (step + both e2e arms)
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
timeout 150 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-attribution > $O/r2x_bench.json 2> $O/r2x_bench.err; echo "bench rc=$?" >> $O/r2x_bench.err
tail -2 $O/r2x_bench.err; python -c "
import json; d=json.loads(open('$O/r2x_bench.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['fp32_host_tensors']['value'], d['e2e']['logit_max_abs_diff_vs_device_resident_step'], d['data'])"
