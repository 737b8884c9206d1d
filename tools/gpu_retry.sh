#!/bin/bash
# tools/gpu_retry.sh LOGFILE TIMEOUT 'command' [GPUS] -- submit one gpurun call, retrying while the pod answers busy (rc 3).
LOG=$1; TMO=$2; CMD=$3; GPUS=${4:-1}
for i in $(seq 1 40); do
  if [ "$GPUS" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $TMO -- "$CMD" > $LOG 2>&1
  else /usr/local/graft/bin/gpurun --gpus $GPUS --timeout $TMO -- "$CMD" > $LOG 2>&1; fi
  rc=$?
  if [ $rc -ne 3 ] && [ $rc -ne 2 ]; then exit $rc; fi
  sleep 45
done
exit 3
