#!/bin/bash
# GPU session O: ncu --set full with source of the attention backward kernels (one-pass tcgen05 v2 and the two mma.sync kernels)
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
CTC_ATTN_BWD=2 timeout 200 python tools/prof_step.py 8 > $O/r2o_plain.log 2>&1 || { tail -5 $O/r2o_plain.log; exit 1; }
CTC_ATTN_BWD=2 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:onepass -c 1 \
    -o $O/r2o_onepass python tools/prof_step.py 8 > $O/r2o_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attn_bwd_d -c 2 \
    -o $O/r2o_mmasync python tools/prof_step.py 8 > $O/r2o_ncu2.log 2>&1
tail -2 $O/r2o_ncu1.log $O/r2o_ncu2.log; ls -la $O/*.ncu-rep
