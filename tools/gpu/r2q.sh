#!/bin/bash
# GPU session Q: does micro-batching the 8-volume step (activations nearer the 126 MB L2) pay?  fwd+bwd device time at batch 8 / 4 / 2
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
for b in 8 4 2 8 4; do timeout 200 python tools/time_engine.py $b 2>&1 | grep -E "^B=" ; done > $O/r2q_microbatch.log
cat $O/r2q_microbatch.log
