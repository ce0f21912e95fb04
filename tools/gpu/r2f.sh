#!/bin/bash
# GPU session F: full GPU suite + smoke() on the final tree
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > $O/r2f_tests.log 2>&1; echo "rc=$?" >> $O/r2f_tests.log
tail -3 $O/r2f_tests.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
