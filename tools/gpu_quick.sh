#!/bin/bash
# quick loop: selected selfchecks + bench + ncu launch list
mkdir -p gpurun_out
CHECKS="${CHECKS:-head conv2_wgrad conv1_wgrad local_step}" CLIENTS="${CLIENTS:-3 160}" bash tools/gpu_selfcheck.sh 2>&1 | grep -E "ok\"|exception|exit=|selfcheck:" | cut -c1-400
timeout 900 python bench.py --steps ${STEPS:-5} --warmup 3 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -3 gpurun_out/bench_1gpu.err; cut -c1-420 gpurun_out/bench_1gpu.json; echo
python - <<'PY'
import json; d=json.load(open('gpurun_out/bench_1gpu.json')); print('E2E', d['e2e'])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-e2e --clients ${NCU_CLIENTS:-592} > gpurun_out/ncu_launch_run.log 2>&1
python tools/ncu_launch_summary.py gpurun_out/launches.csv | head -10
