#!/bin/bash
# library baseline (nccl_torch) for the flagship, and configs 3-5, one GPU
mkdir -p gpurun_out
run() { tag=$1; shift; timeout ${TMO:-900} python bench.py "$@" > gpurun_out/cfg_$tag.json 2> gpurun_out/cfg_$tag.err; echo "== $tag rc=$?"; tail -2 gpurun_out/cfg_$tag.err | cut -c1-300; cut -c1-420 gpurun_out/cfg_$tag.json; echo; }
run c2_nccl_torch --impl nccl_torch --steps 2 --warmup 3 --no-e2e
run c4_ours --config 4 --steps 3 --warmup 3
run c4_nccl_torch --config 4 --impl nccl_torch --steps 1 --warmup 3 --no-e2e
run c5_ours --config 5 --steps 1 --warmup 3 --no-e2e
run c3_ours --config 3 --steps 1 --warmup 3 --no-e2e
