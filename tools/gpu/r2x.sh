#!/bin/bash
# GPU session X: bench with the int16-ingest e2e arm; preprocess tests
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
timeout 300 python -m pytest tests/test_preprocess.py -m gpu -q -x > $O/r2x_pre.log 2>&1; echo "rc=$?" >> $O/r2x_pre.log; tail -3 $O/r2x_pre.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-attribution > $O/r2x_bench.json 2> $O/r2x_bench.err; echo "bench rc=$?" >> $O/r2x_bench.err
tail -3 $O/r2x_bench.err; python -c "
import json; d=json.loads(open('$O/r2x_bench.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value']); print(json.dumps(d['e2e'], indent=1)[:1800])"
