#!/bin/bash
# round-2 loop B: fedprox diagnostic, every config with the in-process library arm, library arm alone (bf16 + tf32)
mkdir -p gpurun_out
timeout 600 python tools/diag_fedprox.py > gpurun_out/diag_fedprox.log 2>&1; tail -5 gpurun_out/diag_fedprox.log
for c in 2 4 3 5; do
  timeout 1200 python bench.py --config $c --steps ${STEPS:-5} --warmup 3 > gpurun_out/benchb_c$c.json 2> gpurun_out/benchb_c$c.err
  tail -3 gpurun_out/benchb_c$c.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/benchb_c$c.json'))
    print('cfg',$c,'value',round(d['value']),'ms',round(d['ms_per_step'],2),'vs_baseline',d['vs_baseline'],'e2e',d['e2e'] and round(d['e2e']['value']),'base',{k:(round(v,2) if isinstance(v,float) else v) for k,v in (d.get('baseline') or {}).items() if k in ('value','ms_per_step','chunk','cuda_graph','error','e2e_ratio')})
except Exception as e: print('cfg',$c,'parse failed',e)
PY
done
timeout 900 python bench.py --impl nccl_torch --baseline-dtype tf32 --steps 5 --warmup 3 > gpurun_out/lib_c2_tf32.json 2> gpurun_out/lib_c2_tf32.err; cut -c1-250 gpurun_out/lib_c2_tf32.json; tail -2 gpurun_out/lib_c2_tf32.err
