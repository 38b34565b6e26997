#!/bin/bash
# ncu --set full of one launch per tile shape of the eval plan (batch 32): evidence for the tensor-pipe activity per shape
cd "$(dirname "$0")/.."
for key in 128,256,3,76 256,512,3,38 512,1024,3,19 512,256,1,38 256,128,1,76 32,64,3,304 64,32,1,304; do
  tag=$(echo $key | tr ',' '_')
  timeout 300 ncu --set full --import-source on --clock-control none --profile-from-start off -c 1 -o gpurun_out/r02_conv_$tag -f python scratch/one_layer.py $key 3 > gpurun_out/ncu_conv_$tag.log 2>&1
  tail -1 gpurun_out/ncu_conv_$tag.log | cut -c1-120
done
