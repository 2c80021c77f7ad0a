#!/bin/bash
# compute-sanitizer over the kernels changed late in round 2: the strain-EAS kernel (register LDL^T with shuffles, record
# tails reused as scratch, write-out through the record area) and the generalised-tangent kernel with its strain mode and
# the principal-stretch laws.  Outputs land in gpurun_out/, the summaries are committed under profiles/.
set -x
T="python -m pytest -x -q -m gpu -p no:cacheprovider"
SEL1="tests/test_gpu_eas.py::test_eas_matrix_vector_and_alpha_update"
SEL2="tests/test_gpu_hyperelastic.py::test_blatzko_matrix_vector_energy_and_alpha_update"
SEL3="tests/test_gpu_plane_stress.py"
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 5 $T $SEL1 $SEL2 $SEL3 > gpurun_out/sanitizer_r2b_$tool.log 2>&1
  echo "exit $?" >> gpurun_out/sanitizer_r2b_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit " gpurun_out/sanitizer_r2b_$tool.log | tail -5
done
