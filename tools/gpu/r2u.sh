#!/bin/bash
# GPU session U: ncu --set full of the temporal attention kernels (n = 24)
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
timeout 200 python tools/prof_step.py 8 > $O/r2u_plain.log 2>&1 || { tail -5 $O/r2u_plain.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attn_small_bwd -c 1 \
    -o $O/r2u_small_bwd python tools/prof_step.py 8 > $O/r2u_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attn_small_fwd -c 1 \
    -o $O/r2u_small_fwd python tools/prof_step.py 8 > $O/r2u_ncu2.log 2>&1
ls -la $O/r2u_*.ncu-rep
