cd $GRAFT_REPO_ROOT
timeout 400 python -m pytest tests -q -m gpu -x 2>&1 | grep -v Warning | tail -n 12 > gpurun_out/r32_tests.log
timeout 200 python tools/time_engine.py 8 > gpurun_out/r32_time_b8.log 2>&1
timeout 200 python tools/time_occlusion.py 32 > gpurun_out/r32_occ.log 2>&1
echo done
