#!/bin/bash
# compute-sanitizer over the batched GEMM / conv family and the nn support kernels (tiny shapes: the tools slow kernels 10-100x)
mkdir -p gpurun_out
run() { tool=$1; sel=$2; tag=$3
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python -m pytest tests/test_gpu_nn.py -m gpu -q -x -k "$sel" > gpurun_out/san_nn_${tool}_$tag.log 2>&1
  echo "$tool $tag rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' gpurun_out/san_nn_${tool}_$tag.log | tr '\n' ' ')"; }
run memcheck "linear and 128-256-256" linear
run memcheck "attention" attention
run memcheck "conv and 4-512-512" conv4
run memcheck "conv and 16-128-256-3-2" conv_s2
run memcheck "conv and 32-64-64-3-1" conv_halo
run memcheck "conv and 8-256-256-3-1" conv_wgrad_t
run memcheck "groupnorm and 32-64" groupnorm_32x32
run memcheck "groupnorm and 4-512" groupnorm
run memcheck "layernorm" layernorm
run memcheck "flow_ticks" flow
run racecheck "groupnorm and 4-512" groupnorm
run racecheck "layernorm" layernorm
run synccheck "linear and 128-256-256" linear
run synccheck "conv and 32-64-64-3-1" conv_halo
run racecheck "groupnorm and 32-64" groupnorm_32x32
