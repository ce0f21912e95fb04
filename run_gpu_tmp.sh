cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | grep -v Warning | tail -n 30 > gpurun_out/r17_tests.log
timeout 600 python tools/time_engine.py 8 > gpurun_out/r17_time_b8.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r17_bench.json 2> gpurun_out/r17_bench.err
echo done
