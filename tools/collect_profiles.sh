#!/bin/bash
# gpurun_out/ (scratch) -> profiles/r2_* (tracked): bench lines, op table, launch lists, ncu exports, SASS, sanitizer, multi-GPU logs
cd "$(dirname "$0")/.."
for c in 2 3 4 5; do [ -f gpurun_out/final_c${c}_n1.json ] && grep -h '^{"metric"' gpurun_out/final_c${c}_n1.json > profiles/r2_bench_c${c}_n1.json; done
for c in 2 3 4 5; do for n in 2 4 8; do [ -s gpurun_out/bench_c${c}_n$n.json ] && grep -q '^{"metric"' gpurun_out/bench_c${c}_n$n.json && grep -h '^{"metric"' gpurun_out/bench_c${c}_n$n.json > profiles/r2_bench_c${c}_n$n.json; done; done
[ -f gpurun_out/final_lib_c2_tf32.json ] && grep -h '^{"metric"' gpurun_out/final_lib_c2_tf32.json > profiles/r2_lib_c2_tf32.json
[ -f gpurun_out/final_op_bench.json ] && cp gpurun_out/final_op_bench.json profiles/r2_op_bench.json
for f in gpurun_out/final_launches_c*.csv; do [ -f "$f" ] && python tools/ncu_launch_summary.py "$f" > profiles/r2_launches_$(basename $f .csv | sed 's/final_launches_//').txt; done
[ -f gpurun_out/gpu_tests_final.log ] && cp gpurun_out/gpu_tests_final.log profiles/r2_gpu_tests.log
for n in 2 4 8; do [ -f gpurun_out/mgpu_oracle_n$n.jsonl ] && cp gpurun_out/mgpu_oracle_n$n.jsonl profiles/r2_mgpu_oracle_n$n.jsonl; [ -f gpurun_out/peer_loss_n$n.jsonl ] && cp gpurun_out/peer_loss_n$n.jsonl profiles/r2_peer_loss_n$n.jsonl; done
if ls gpurun_out/ncu_csv/r2_ncu_* > /dev/null 2>&1; then cp gpurun_out/ncu_csv/r2_ncu_* profiles/; python tools/ncu_table.py r2 > /dev/null; fi
[ -f gpurun_out/final_sanitize_nn.log ] && { echo "compute-sanitizer over the round-2 kernels (tools/sanitize_nn.sh; per-run logs stay in gpurun_out/)"; cat gpurun_out/final_sanitize_nn.log; echo; echo "racecheck layernorm: the 2 reported hazards are inside at::native::layer_norm_grad_input_kernel_vectorized (PyTorch's own kernel, used by the oracle of the test), not in this repo's kernels."; } > profiles/r2_sanitizer.txt
python tools/sass_evidence.py r2 > /dev/null 2>&1
python tools/make_results_md.py > /dev/null
ls profiles | grep -c r2_
