#!/bin/bash
# Runs every kernel self-check in its own process (a trapped kernel poisons the CUDA context) and collects JSON.
mkdir -p gpurun_out
CLIENTS="${CLIENTS:-3 160}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tail -1
for chk in ${CHECKS:-ingest bcast reduce mask conv1_fwd conv2_fwd head conv2_dgrad conv2_wgrad conv1_wgrad local_step}; do
  timeout 180 python -m olearning_sim_b200.ops.selfcheck --clients $CLIENTS --only $chk --json gpurun_out/selfcheck_$chk.json 2>&1 | grep -v "^$" | tail -12
  echo "== $chk exit=$?"
done
