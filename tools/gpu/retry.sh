#!/bin/bash
# retry.sh <timeout_s> <script>: keep asking for a GPU box until the pod answers something other than "busy" (rc 3)
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$1" -- "bash $2"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
