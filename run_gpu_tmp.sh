cd $GRAFT_REPO_ROOT
timeout 200 python -m pytest tests/test_gpu_dropin.py -q -m gpu -x 2>&1 | grep -v Warning | tail -n 30 > gpurun_out/r37_dropin.log
echo done
