cd $GRAFT_REPO_ROOT
timeout 120 python tools/kernel_bench.py attn_s > gpurun_out/r31_plain.log 2>&1 || { echo plain failed; exit 1; }
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 3 -c 1 -f -o gpurun_out/r31_tc python tools/kernel_bench.py attn_s > gpurun_out/r31_ncu.log 2>&1
echo done
