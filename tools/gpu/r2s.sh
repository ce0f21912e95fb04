#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
(echo "== default"; timeout 200 python tools/ig_batch_check.py 2>&1 | tail -6
echo "== CTC_ATTN_BWD=0"; CTC_ATTN_BWD=0 timeout 200 python tools/ig_batch_check.py 2>&1 | tail -6
echo "== CTC_GEMM_DIRECT=0"; CTC_GEMM_DIRECT=0 timeout 200 python tools/ig_batch_check.py 2>&1 | tail -6) > $O/r2s.log
cat $O/r2s.log
