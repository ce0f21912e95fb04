#!/bin/bash
# GPU session M: BASELINE configs[4] - the full attribution suite over 64 volumes through the drop-in entry point, N GPUs
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
N=${1:-1}
free -g | head -2 > $O/r2m_suite_${N}gpu.err; df -h /tmp | tail -1 >> $O/r2m_suite_${N}gpu.err
if [ "$N" = "1" ]; then
  timeout 2400 python tools/suite.py --volumes 64 > $O/r2m_suite_${N}gpu.json 2>> $O/r2m_suite_${N}gpu.err
else
  timeout 2400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      tools/suite.py --volumes 64 > $O/r2m_suite_${N}gpu.json 2>> $O/r2m_suite_${N}gpu.err
fi
echo "rc=$?" >> $O/r2m_suite_${N}gpu.err
tail -5 $O/r2m_suite_${N}gpu.err; cat $O/r2m_suite_${N}gpu.json
