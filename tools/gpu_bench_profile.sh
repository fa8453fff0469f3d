#!/bin/bash
# pytest (gpu) + 1-GPU bench + ncu launch list + one full ncu capture of the hot kernels
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "=== bench 1 GPU"; timeout 900 python bench.py --steps ${STEPS:-5} --warmup 3 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -3 gpurun_out/bench_1gpu.err; cat gpurun_out/bench_1gpu.json
if [ "${NCU:-1}" = "1" ]; then
echo "=== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-e2e --clients ${NCU_CLIENTS:-592} > gpurun_out/ncu_launch_run.log 2>&1
tail -2 gpurun_out/ncu_launch_run.log
echo "=== ncu full (conv kernels + head)"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'conv|head_fc|fedavg|bcast' -s 33 -c 12 -o gpurun_out/prof_cnn2 -f python bench.py --steps 1 --warmup 3 --no-e2e --clients ${NCU_CLIENTS:-592} > gpurun_out/ncu_full_run.log 2>&1
tail -2 gpurun_out/ncu_full_run.log; ls -la gpurun_out/
fi
