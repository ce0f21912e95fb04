#!/bin/bash
# GPU session T: smoke() of the final code + kernel table of the occlusion fast path (32-window batch)
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
timeout 300 python __graft_entry__.py smoke > $O/r2t_smoke.log 2>&1; echo "rc=$?" >> $O/r2t_smoke.log; tail -3 $O/r2t_smoke.log
timeout 300 python tools/time_occlusion.py 32 > $O/r2t_occ.log 2>&1; head -45 $O/r2t_occ.log | cut -c1-150
