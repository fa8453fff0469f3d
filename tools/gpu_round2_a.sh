#!/bin/bash
# round-2 loop A: full GPU test-suite, benches of configs 2/3/5, launch lists of the ResNet / BERT steps
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/gpu_tests.log
for c in 3 5 2; do
  timeout 900 python bench.py --config $c --steps ${STEPS:-5} --warmup 3 > gpurun_out/bench_c$c.json 2> gpurun_out/bench_c$c.err
  tail -2 gpurun_out/bench_c$c.err; cut -c1-300 gpurun_out/bench_c$c.json; echo
done
for c in 3 5; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 600 --csv --log-file gpurun_out/launches_c$c.csv python bench.py --config $c --steps 1 --warmup 3 --no-e2e --clients 96 > gpurun_out/ncu_c$c.log 2>&1
  python tools/ncu_launch_summary.py gpurun_out/launches_c$c.csv | head -24
done
