cd $GRAFT_REPO_ROOT
timeout 200 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "attention" 2>&1 | grep -v Warning | tail -n 5 > gpurun_out/r36_ops.log
echo done
