#!/bin/bash
# compute-sanitizer (memcheck + racecheck) over a small slice of the parity tests.
mkdir -p gpurun_out
T="tests/test_gpu_parity.py"
SEL="test_c2s_predict_and_acq_vs_golden or test_constrained_acquisition_vs_golden or test_kernel_families_vs_golden or (test_small_batch_path_vs_tiled_and_oracle and 40) or (test_fp32_mode_tcgen05_vs_oracle and 300) or test_incremental_append_equals_full_fit or (test_predict_return_cov_vs_sklearn and 50) or test_mixed_int_space_round_transform_and_de_branch or test_categorical_parameter_host_transform"
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 99 python -m pytest $T -q -x -p no:cacheprovider -k "$SEL" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_$tool.log | tail -3
done
