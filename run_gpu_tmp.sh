cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -q -m gpu -x -s 2>&1 | grep -v Warning | tail -n 60 > gpurun_out/r11_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r11_bench.json 2> gpurun_out/r11_bench.err
timeout 600 python tools/time_engine.py 8 > gpurun_out/r11_time_b8.log 2>&1
echo done
