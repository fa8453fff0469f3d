#!/bin/bash
# round-2 loop C: op-level roofline table + ncu --set full captures of the main bgemm instantiations
mkdir -p gpurun_out
timeout 900 python tools/op_bench.py --clients 64 --json gpurun_out/op_bench.json > gpurun_out/op_bench.txt 2>&1; cat gpurun_out/op_bench.txt
for op in linear_fwd_ff1_gelu linear_dgrad_ff2_gelubwd linear_wgrad_ff1 attn_scores conv_fwd_32x32_64-64_s1 conv_dgrad_32x32_64-64_s1 conv_wgrad_32x32_64-64_s1 conv_wgrad_8x8; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:bgemm_kernel --launch-skip 2 --launch-count 1 -f -o gpurun_out/ncu_$op python tools/op_bench.py --clients 32 --once --only $op > gpurun_out/ncu_$op.log 2>&1
  tail -1 gpurun_out/ncu_$op.log
done
timeout 600 python -m pytest tests/test_gpu_nn.py -m gpu -q -k "fedprox or resnet18_native" 2>&1 | tail -3
