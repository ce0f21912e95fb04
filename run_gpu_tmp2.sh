cd $GRAFT_REPO_ROOT
timeout 120 python tools/kernel_bench.py attn_s > gpurun_out/r33_plain.log 2>&1 || { echo plain failed; exit 1; }
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_bwd -s 6 -c 2 -f -o gpurun_out/r33_bwd python tools/kernel_bench.py attn_s > gpurun_out/r33_ncu.log 2>&1
echo done
