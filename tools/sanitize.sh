#!/bin/bash
# compute-sanitizer passes over the kernel self-checks (tiny client counts: the tools slow kernels 10-100x)
mkdir -p gpurun_out
for tool in ${TOOLS:-memcheck racecheck synccheck}; do
  for chk in ${CHECKS:-reduce bcast mask head gemm conv2_fwd conv2_wgrad}; do
    timeout 600 compute-sanitizer --tool $tool --print-limit 5 python -m olearning_sim_b200.ops.selfcheck --clients 2 --only $chk > gpurun_out/san_${tool}_${chk}.log 2>&1
    echo "$tool $chk rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|selfcheck:' gpurun_out/san_${tool}_${chk}.log | tr '\n' ' ')"
  done
done
