#!/bin/bash
# ncu --set full of every kernel variant (one launch each) -> gpurun_out/ncuF_*.ncu-rep
mkdir -p gpurun_out
cap() { name=$1; regex=$2; skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$regex --launch-skip $skip --launch-count 1 -f -o gpurun_out/ncuF_$name "$@" > gpurun_out/ncuF_$name.log 2>&1; tail -1 gpurun_out/ncuF_$name.log | cut -c1-120; }
for cl in 1024 128; do
  for k in conv1_window_kernel conv_fwd_fold_kernel conv_fwd_shift_kernel head_fc_xent_kernel conv_dgrad_fold_kernel conv_dgrad_pair_kernel "conv_wgrad_kernel<64" "conv_wgrad_kernel<32" bcast_scatter_cnn2 fedavg_partial fedavg_finish deviceflow_mask round_stats; do
    n=$(echo $k | tr -cd 'a-z0-9_'); cap cnn2_${cl}_$n "$k" 12 python bench.py --config 2 --steps 1 --warmup 3 --no-e2e --no-baseline --clients $cl
  done
done
for op in linear_fwd_qkv linear_fwd_ff1_gelu linear_dgrad_ff2_gelubwd linear_dgrad_qkv linear_wgrad_ff1 attn_scores attn_pv conv_fwd_32x32_64-64_s1 conv_fwd_8x8 conv_dgrad_16x16 conv_dgrad_32x32_64-128_s2 conv_wgrad_32x32_64-64_s1 conv_wgrad_4x4; do
  cap op_$op bgemm_kernel 2 python tools/op_bench.py --clients 32 --once --only $op
done
for k in gn_apply_kernel gn_bwd_reduce_kernel gn_bwd_apply_kernel pool_fc_xent_kernel; do cap nn_$k $k 6 python bench.py --config 3 --steps 1 --warmup 3 --no-e2e --no-baseline --clients 64; done
for k in ln_fwd_kernel ln_bwd_kernel colsum_kernel emb_bwd_kernel cls_head_kernel bcast_scatter_generic fedavg_partial fedavg_finish; do cap nnb_$k $k 4 python bench.py --config 5 --steps 1 --warmup 3 --no-e2e --no-baseline --clients 64; done
ls gpurun_out/ncuF_*.ncu-rep | wc -l
