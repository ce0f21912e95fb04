cd $GRAFT_REPO_ROOT
timeout 300 python tools/grad_err_probe.py > gpurun_out/r40_probe.log 2>&1
echo done
