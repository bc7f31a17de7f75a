#!/usr/bin/env bash
# retry.sh <gpus> <timeout> <script>: call gpurun until it is not refused as transient (exit code 3)
G=$1; T=$2; S=$3
for i in $(seq 1 20); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $T -- "bash $S" > /tmp/gpucall.out 2>&1; else /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "bash $S" > /tmp/gpucall.out 2>&1; fi
  rc=$?
  if grep -q "status=transient\|status=busy" /tmp/gpucall.out || [ $rc -eq 3 ]; then sleep 150; continue; fi
  break
done
echo "attempts=$i rc=$rc" >> /tmp/gpucall.out
