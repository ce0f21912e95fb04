#!/bin/bash
# GPU session H: ncu --set full with source of the epilogue-bound GEMM launches (FF1 + GEGLU + factors, dh + adjoint, out + residual, FF2 + residual)
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
timeout 200 python tools/prof_gemm.py all > $O/r2h_plain.log 2>&1 || { tail -20 $O/r2h_plain.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tcgen05 -c 4 \
    -o $O/r2h_gemm python tools/prof_gemm.py all > $O/r2h_ncu.log 2>&1
cat $O/r2h_plain.log; tail -3 $O/r2h_ncu.log; ls -la $O/r2h_gemm.ncu-rep
