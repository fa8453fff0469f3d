#!/bin/bash
# ncu --set full of every kernel variant at the benchmarked client counts -> gpurun_out/ncuF_*.ncu-rep (a handful of
# processes: every capture window covers one launch of each kernel of a phase)
mkdir -p gpurun_out
cap() { name=$1; regex=$2; skip=$3; count=$4; src=$5; shift 5
  timeout 900 ncu --set full --clock-control none $src -k "regex:$regex" --launch-skip $skip --launch-count $count -f -o gpurun_out/ncuF_$name "$@" > gpurun_out/ncuF_$name.log 2>&1
  tail -1 gpurun_out/ncuF_$name.log | cut -c1-120; }
SRC="--import-source on"
for cl in 1024 128; do
  # one local step = 6 launches (conv1 fwd, conv2 fwd, head, conv2 dgrad, conv2 wgrad, conv1 wgrad); 30 per round
  cap cnn2_${cl}_step "conv1_window|conv_fwd|head_fc|conv_dgrad|conv_wgrad" 96 6 "$SRC" python bench.py --config 2 --steps 1 --warmup 3 --no-e2e --no-baseline --clients $cl
  cap cnn2_${cl}_fl "bcast_scatter|fedavg_partial|fedavg_finish|deviceflow_mask|round_stats" 10 5 "$SRC" python bench.py --config 2 --steps 1 --warmup 3 --no-e2e --no-baseline --clients $cl
done
# every op of the ResNet-18 / BERT-mini steps exactly once (tools/op_bench.py --nowarm: ncu replays with its own cache control)
cap ops_bgemm "bgemm_kernel" 0 60 "" python tools/op_bench.py --clients 32 --nowarm
cap ops_top "bgemm_kernel" 0 3 "$SRC" python tools/op_bench.py --clients 32 --nowarm --only conv_fwd_32x32_64-64
cap ops_top2 "bgemm_kernel" 0 2 "$SRC" python tools/op_bench.py --clients 32 --nowarm --only linear_fwd_ff1
cap nn_resnet "gn_apply|gn_bwd|pool_fc_xent|stem_im2col|slice_sgd|bcast_scatter_generic" 40 40 "" python bench.py --config 3 --steps 1 --warmup 1 --no-e2e --no-baseline --clients 32
cap nn_bert "ln_fwd|ln_bwd|colsum|emb_bwd|cls_head|slice_sgd|fedavg_partial|fedavg_finish|bcast_scatter_generic" 30 40 "" python bench.py --config 5 --steps 1 --warmup 1 --no-e2e --no-baseline --clients 32
ls -la gpurun_out/ncuF_*.ncu-rep | awk '{s+=$5} END {print NR " reports, " s/1e6 " MB"}'
