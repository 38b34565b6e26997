#!/bin/bash
# ncu --set full of the BN / PReLU passes on one layer-sized tensor (64 x 152 x 152 x 128): direct vs TMA row pipe
cd "$(dirname "$0")/.."
RYOLO_BN_PIPE=7 timeout 600 ncu --set full --clock-control none -k regex:bn_ -c 12 -o gpurun_out/r02_bn_pipe -f python scratch/bn_sweep.py 64,152,152,128 > gpurun_out/ncu_bn_pipe.log 2>&1; tail -2 gpurun_out/ncu_bn_pipe.log | cut -c1-150
RYOLO_BN_PIPE=0 timeout 600 ncu --set full --clock-control none -k regex:bn_ -c 12 -o gpurun_out/r02_bn_direct -f python scratch/bn_sweep.py 64,152,152,128 > gpurun_out/ncu_bn_direct.log 2>&1; tail -2 gpurun_out/ncu_bn_direct.log | cut -c1-150
