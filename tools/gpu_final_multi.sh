#!/bin/bash
# multi-GPU evidence at N ranks: NCCL-oracle correctness of every native family, then the flagship scaling bench
# (ours + in-process library arm, phases) — results under gpurun_out/, copied to profiles/ by the caller
N=${N:-2}
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -12 > gpurun_out/topo_n$N.txt
if [ -z "$SKIP_ORACLE" ]; then
: > gpurun_out/mgpu_oracle_n$N${TAG:-}.jsonl
IFS=";" read -ra SPECS_A <<< "${SPECS:-cnn2 avg;cnn2 adam;mlp2 momentum;resnet18 avg;bert_mini avg}"
for spec in "${SPECS_A[@]}"; do
  set -- $spec
  MODEL=$1 OPT=$2 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/mgpu_oracle.py 2>gpurun_out/mgpu_oracle_${1}_$2.err | grep -E "^\{" >> gpurun_out/mgpu_oracle_n$N${TAG:-}.jsonl || tail -5 gpurun_out/mgpu_oracle_${1}_$2.err
done
# the same comm check with the multicast path disabled (CUDA-IPC peer pointers): timing comparison + fallback correctness
IFS=";" read -ra P2P_A <<< "${P2P_SPECS:-cnn2 avg;bert_mini avg}"
for spec in "${P2P_A[@]}"; do
  set -- $spec
  OLSB_NVLS=0 MODEL=$1 OPT=$2 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 tools/mgpu_oracle.py 2>gpurun_out/mgpu_oracle_${1}_$2_p2p.err | grep -E "^\{" >> gpurun_out/mgpu_oracle_n$N${TAG:-}.jsonl || tail -5 gpurun_out/mgpu_oracle_${1}_$2_p2p.err
done
cat gpurun_out/mgpu_oracle_n$N${TAG:-}.jsonl | cut -c1-600
fi
[ -z "$SKIP_PEER" ] && timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 tools/peer_loss_check.py 2>gpurun_out/peer_loss_n$N.err | grep -E "^\{" > gpurun_out/peer_loss_n$N.jsonl; cat gpurun_out/peer_loss_n$N.jsonl
[ -n "$SKIP_BENCH" ] && exit 0
for c in ${CONFIGS:-2}; do
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2962$N bench.py --gpus $N --config $c --steps ${STEPS:-10} --warmup 3 --phases ${BENCH_FLAGS:-} > gpurun_out/bench_c${c}_n$N.json 2>gpurun_out/bench_c${c}_n$N.err
  tail -2 gpurun_out/bench_c${c}_n$N.err; grep -h '^{"metric"' gpurun_out/bench_c${c}_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); b=d.get('baseline') or {}
print('cfg',d['config']['id'],'N',d['n_gpus'],'value',round(d['value']),'ms',round(d['ms_per_step'],3),'vs_baseline',d['vs_baseline'],'e2e',d['e2e'] and round(d['e2e']['value']),'lib',b.get('value') and round(b['value']),'lib_ms',b.get('ms_per_step'),'err',b.get('error'))
print('PHASES',d.get('phases',{}).get('ms_max'),d.get('phases',{}).get('ms_min'))"
done
