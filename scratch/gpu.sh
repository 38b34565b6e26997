#!/bin/bash
# build in-tree, then run a command on the GPU box, retrying while the pod has no free slot:
#   scratch/gpu.sh <timeout_s> '<command>'
cd /root/repo
python -c "import __graft_entry__ as g; g.build()" > /tmp/build.log 2>&1 || { tail -30 /tmp/build.log; exit 1; }
T=$1; shift
for try in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[gpu.sh] no slot (try $try), retrying in 45 s"
  sleep 45
done
exit 3
