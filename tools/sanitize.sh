#!/bin/bash
# compute-sanitizer passes over the kernel self-checks (tiny client counts: the tools slow kernels 10-100x).
# memcheck over everything; racecheck over the SIMT kernels that use shared memory; synccheck over the tcgen05 kernels.
mkdir -p gpurun_out
run() { tool=$1; chk=$2
  timeout 600 compute-sanitizer --tool $tool --print-limit 5 python -m olearning_sim_b200.ops.selfcheck --clients 2 --only $chk > gpurun_out/san_${tool}_${chk}.log 2>&1
  echo "$tool $chk rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|selfcheck:' gpurun_out/san_${tool}_${chk}.log | tr '\n' ' ')"; }
for chk in ${MEMCHECK:-reduce bcast mask head gemm conv1_fwd conv2_fwd conv2_dgrad conv2_wgrad conv1_wgrad}; do run memcheck $chk; done
for chk in ${RACECHECK:-reduce head}; do run racecheck $chk; done
for chk in ${SYNCCHECK:-conv2_fwd conv2_dgrad}; do run synccheck $chk; done
