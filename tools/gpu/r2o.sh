#!/bin/bash
# GPU session O: ncu --set full with source of the spatial attention kernels (one-pass tcgen05 backward, tcgen05 forward)
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
timeout 200 python tools/prof_step.py 8 > $O/r2o_plain.log 2>&1 || { tail -5 $O/r2o_plain.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:onepass -c 1 \
    -o $O/r2o_attn python tools/prof_step.py 8 > $O/r2o_ncu1.log 2>&1
tail -n 2 $O/r2o_ncu1.log; ls -la $O/r2o_attn.ncu-rep
