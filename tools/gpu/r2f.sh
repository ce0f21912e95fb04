#!/bin/bash
# GPU session F: full GPU suite after the single-M-tile lookahead fix of the one-pass attention backward
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2f_tests.log 2>&1; echo "rc=$?" >> $O/r2f_tests.log
tail -5 $O/r2f_tests.log
