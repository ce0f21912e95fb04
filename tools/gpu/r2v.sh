#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q -x > $O/r2v_model.log 2>&1; echo "rc=$?" >> $O/r2v_model.log; tail -8 $O/r2v_model.log
