#!/bin/bash
# compute-sanitizer over the generalised-tangent kernel after its rework (register tiles, elimination in registers).
set -x
T="python -m pytest -x -q -m gpu -p no:cacheprovider"
SEL1="tests/test_gpu_eas.py::test_displacement_gradient_matrix_vector_and_alpha_update"
SEL2="tests/test_gpu_hyperelastic.py::test_blatzko_matrix_vector_energy_and_alpha_update"
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 5 $T $SEL1 $SEL2 > gpurun_out/sanitizer_r2c_$tool.log 2>&1
  echo "exit $?" >> gpurun_out/sanitizer_r2c_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit " gpurun_out/sanitizer_r2c_$tool.log | tail -4
done
