#!/bin/bash
# usage: tools/gpu_batch.sh <tag> ; runs a batch on the GPU box, everything into gpurun_out/<tag>/
tag=${1:-b}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $out/box.txt 2>&1
nproc >> $out/box.txt; free -g | head -2 >> $out/box.txt
