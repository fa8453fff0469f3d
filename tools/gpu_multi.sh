#!/bin/bash
# multi-GPU: P2P correctness vs single rank, then the scaling bench at N ranks
N=${N:-2}
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -12
for OPT in avg adam; do
OPT=$OPT timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py 2>&1 | grep -E "^\{|Error|error|Traceback" | head -8
done
MODEL=mlp2 BACKEND=hybrid OPT=adam CLIENTS=48 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/mgpu_check.py 2>&1 | grep -E "^\{|Error|error|Traceback" | head -8
for n in ${SCALE:-1 $N}; do
  if [ "$n" = "1" ]; then timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2>gpurun_out/bench_n1.err; cut -c1-330 gpurun_out/bench_n1.json; echo
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 5 --warmup 3 --phases > gpurun_out/bench_n$n.json 2>gpurun_out/bench_n$n.err; tail -2 gpurun_out/bench_n$n.err; grep -h '^{"metric"' gpurun_out/bench_n$n.json | cut -c1-330; grep -h '^{"metric"' gpurun_out/bench_n$n.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('E2E', d['e2e'], 'launches', d['gpu_launches']); print('PHASES', d.get('phases'))"
  fi
done
