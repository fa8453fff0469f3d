#!/bin/bash
# everything the single-GPU evidence needs, in one gpurun call: tests, benches (both arms), op table, launch lists, ncu, sanitizer
bash tools/gpu_final_1gpu.sh > gpurun_out/final_1gpu.log 2>&1
bash tools/gpu_final_ncu.sh > gpurun_out/final_ncu.log 2>&1
bash tools/sanitize_nn.sh > gpurun_out/final_sanitize_nn.log 2>&1
tail -5 gpurun_out/final_1gpu.log gpurun_out/final_ncu.log gpurun_out/final_sanitize_nn.log
