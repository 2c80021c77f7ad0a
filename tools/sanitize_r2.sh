#!/bin/bash
# compute-sanitizer over the kernels added or changed in round 2 (small meshes; outputs land in gpurun_out/, the
# summaries are committed under profiles/).  memcheck: out-of-bounds / misaligned accesses; racecheck: shared-memory hazards.
set -x
T="python -m pytest -x -q -m gpu -p no:cacheprovider"
SEL1="tests/test_gpu_eas.py::test_displacement_gradient_matrix_vector_and_alpha_update"
SEL2="tests/test_gpu_parity.py::test_gather_variants_bit_identical"
SEL3="tests/test_gpu_results.py::test_results_match_oracle"
SEL4="tests/test_gpu_parity.py::test_pattern_bit_exact_and_values"
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 5 $T $SEL1 $SEL2 $SEL3 $SEL4 > gpurun_out/sanitizer_$tool.log 2>&1
  echo "exit $?" >> gpurun_out/sanitizer_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit " gpurun_out/sanitizer_$tool.log | tail -5
done
