#!/bin/bash
# single-GPU evidence in separate steps (each leaves only small files in gpurun_out/): STEP=bench | ncu | sanitize
case ${STEP:-bench} in
  bench) bash tools/gpu_final_1gpu.sh > gpurun_out/final_1gpu.log 2>&1; tail -30 gpurun_out/final_1gpu.log ;;
  ncu) bash tools/gpu_final_ncu.sh > gpurun_out/final_ncu.log 2>&1; tail -30 gpurun_out/final_ncu.log ;;
  sanitize) bash tools/sanitize_nn.sh > gpurun_out/final_sanitize_nn.log 2>&1; cat gpurun_out/final_sanitize_nn.log ;;
esac
du -sh gpurun_out
