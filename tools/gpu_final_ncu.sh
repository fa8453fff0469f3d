#!/bin/bash
# ncu --set full of every kernel variant at the benchmarked client counts.  The .ncu-rep files are exported to small CSVs ON
# THE BOX (tools/ncu_export.py -> gpurun_out/ncu_csv/) and deleted: gpurun brings back at most 64 MiB.
mkdir -p gpurun_out/ncu_csv /tmp/ncu
cap() { name=$1; regex=$2; skip=$3; count=$4; src=$5; shift 5
  timeout 900 ncu --set full --clock-control none $src -k "regex:$regex" --launch-skip $skip --launch-count $count -f -o /tmp/ncu/ncuF_$name "$@" > gpurun_out/ncu_csv/ncuF_$name.log 2>&1
  tail -1 gpurun_out/ncu_csv/ncuF_$name.log | cut -c1-120
  if [ -f /tmp/ncu/ncuF_$name.ncu-rep ]; then
    OLSB_PROFILE_DIR=gpurun_out/ncu_csv python tools/ncu_export.py r2 /tmp/ncu/ncuF_$name.ncu-rep > /dev/null 2>&1
    sz=$(stat -c %s /tmp/ncu/ncuF_$name.ncu-rep); echo "  $name: report $sz bytes"
    # keep the (small) report itself only for the two headline captures
    case $name in cnn2_1024_step|ops_top) [ $sz -lt 12000000 ] && cp /tmp/ncu/ncuF_$name.ncu-rep gpurun_out/ncu_csv/ ;; esac
    rm -f /tmp/ncu/ncuF_$name.ncu-rep
  fi; }
SRC="--import-source on"
for cl in 1024 128; do
  # one local step = 6 launches (conv1 fwd, conv2 fwd, head, conv2 dgrad, conv2 wgrad, conv1 wgrad); 30 per round
  cap cnn2_${cl}_step "conv1_window|conv_fwd|head_fc|conv_dgrad|conv_wgrad" 96 6 "$SRC" python bench.py --config 2 --steps 1 --warmup 3 --no-e2e --no-baseline --clients $cl
  cap cnn2_${cl}_fl "bcast_scatter|fedavg_partial|fedavg_finish|deviceflow_mask|round_stats" 10 5 "" python bench.py --config 2 --steps 1 --warmup 3 --no-e2e --no-baseline --clients $cl
done
# every op of the ResNet-18 / BERT-mini steps exactly once (tools/op_bench.py --nowarm: ncu replays with its own cache control)
cap ops_bgemm "bgemm_kernel" 0 60 "" python tools/op_bench.py --clients 32 --nowarm
cap ops_top "bgemm_kernel" 0 3 "$SRC" python tools/op_bench.py --clients 32 --nowarm --only conv_fwd_32x32_64-64
cap ops_top2 "bgemm_kernel" 0 1 "$SRC" python tools/op_bench.py --clients 32 --nowarm --only linear_fwd_ff1
cap nn_resnet "gn_apply|gn_bwd|pool_fc_xent|stem_im2col|slice_sgd|bcast_scatter_generic" 40 24 "" python bench.py --config 3 --steps 1 --warmup 1 --no-e2e --no-baseline --clients 32
cap nn_bert "ln_fwd|ln_bwd|colsum|emb_bwd|cls_head|slice_sgd|fedavg_partial|fedavg_finish|bcast_scatter_generic" 30 24 "" python bench.py --config 5 --steps 1 --warmup 1 --no-e2e --no-baseline --clients 32
du -sh gpurun_out/ncu_csv
