#!/bin/bash
# compute-sanitizer (memcheck + racecheck) over a small slice of the parity tests covering every kernel family
# (SURVEY.md section 5), incl. the round-2 kernels: fused selection, streamed batches, Philox source, pass-batched
# small path, 128x128 GEMM tiles, look-ahead Cholesky (two streams), tiled LML gradient, WhiteKernel.
tag=${1:-san}
out=gpurun_out/$tag
mkdir -p $out
SEL1="test_c2s_predict_and_acq_vs_golden or test_constrained_acquisition_vs_golden or test_kernel_families_vs_golden or (test_small_batch_path_vs_tiled_and_oracle and 40) or (test_fp32_mode_tcgen05_vs_oracle and 300) or test_incremental_append_equals_full_fit or (test_predict_return_cov_vs_sklearn and 50)"
SEL2="(test_fused_selection_equals_numpy_on_the_same_values and 129) or test_fused_selection_nan_semantics or (test_philox_rows_match_oracle_bitwise and 5) or (test_throughput_mode_equals_host_evaluation_of_the_same_rows and 300) or (test_lockstep_refinement_bit_identical_to_sequential_small_n and 100) or (test_lookahead_cholesky_and_tiled_gemm_vs_serial_and_sklearn and 700) or test_tiled_lml_gradient_kernel_families or (test_white_kernel_term_vs_sklearn and kernel_first)"
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 99 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "$SEL1" > $out/sanitizer_${tool}_parity.log 2>&1
  echo "$tool parity exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $out/sanitizer_${tool}_parity.log | tail -3
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 99 python -m pytest tests/test_gpu_round2.py -q -x -p no:cacheprovider -k "$SEL2" > $out/sanitizer_${tool}_round2.log 2>&1
  echo "$tool round2 exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $out/sanitizer_${tool}_round2.log | tail -3
done
