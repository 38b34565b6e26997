#!/bin/bash
# build in-tree, then run a command on the GPU box:  scratch/gpu.sh <timeout_s> '<command>'
set -e
cd /root/repo
python -c "import __graft_entry__ as g; g.build()" > /tmp/build.log 2>&1 || { tail -30 /tmp/build.log; exit 1; }
T=$1; shift
exec /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"
