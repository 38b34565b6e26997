#!/bin/bash
# scratch/gpu_n.sh <ngpus> <timeout_s> '<command>' : build, then gpurun --gpus N with retries
cd /root/repo
python -c "import __graft_entry__ as g; g.build()" > /tmp/build.log 2>&1 || { tail -30 /tmp/build.log; exit 1; }
N=$1; T=$2; shift; shift
for try in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun --gpus "$N" --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[gpu_n.sh] no slot (try $try), retrying in 60 s"
  sleep 60
done
exit 3
