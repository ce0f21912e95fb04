#!/bin/bash
# retry_n.sh <gpus> <timeout_s> <script>
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --gpus "$1" --timeout "$2" -- "bash $3"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
