cd $GRAFT_REPO_ROOT
timeout 500 python bench.py --steps 5 --warmup 3 > gpurun_out/r44_bench.json 2> gpurun_out/r44_bench.err
echo done
