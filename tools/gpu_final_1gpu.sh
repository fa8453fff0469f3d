#!/bin/bash
# single-GPU evidence: full GPU test-suite, every BASELINE config with the in-process library arm, op-level roofline table,
# launch lists and ncu --set full captures of every kernel variant at the benchmarked client counts
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/gpu_tests_final.log; tail -3 gpurun_out/gpu_tests_final.log
for c in 2 4 3 5; do
  timeout 1500 python bench.py --config $c --steps ${STEPS:-10} --warmup 3 > gpurun_out/final_c${c}_n1.json 2> gpurun_out/final_c${c}_n1.err
  tail -2 gpurun_out/final_c${c}_n1.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/final_c${c}_n1.json')); b=d.get('baseline') or {}
    print('cfg',$c,'value',round(d['value']),'ms',round(d['ms_per_step'],2),'vs_baseline',d['vs_baseline'],'e2e',d['e2e'] and round(d['e2e']['value']),'lib',b.get('value') and round(b['value']),'e2e_ratio',b.get('e2e_ratio'),'clocks',d['clocks'])
except Exception as e: print('cfg',$c,'parse failed',e)
PY
done
timeout 900 python bench.py --impl nccl_torch --baseline-dtype tf32 --steps 5 --warmup 3 > gpurun_out/final_lib_c2_tf32.json 2>/dev/null; cut -c1-200 gpurun_out/final_lib_c2_tf32.json
timeout 900 python tools/op_bench.py --clients 64 --json gpurun_out/final_op_bench.json > gpurun_out/final_op_bench.txt 2>&1; tail -30 gpurun_out/final_op_bench.txt
# launch lists (shares) at the two client counts of the scaling run, and for the ResNet / BERT steps
for spec in "2 1024" "2 128" "3 96" "5 96"; do set -- $spec
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-120} -c 700 --csv --log-file gpurun_out/final_launches_c$1_$2.csv python bench.py --config $1 --steps 1 --warmup 3 --no-e2e --no-baseline --clients $2 > /dev/null 2>&1
  echo "== launches config $1, $2 clients"; python tools/ncu_launch_summary.py gpurun_out/final_launches_c$1_$2.csv | head -14
done
