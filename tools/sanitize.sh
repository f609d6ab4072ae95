#!/bin/bash
# compute-sanitizer passes over a few small parity tests (memcheck: out-of-bounds / misaligned;
# racecheck: shared-memory hazards in the collectives; synccheck: barrier misuse)
T="tests/test_gpu_parity.py::test_ragged_and_tiny_sizes_parity tests/test_gpu_parity.py::test_results_do_not_depend_on_launch_geometry tests/test_gpu_postcorrect.py::test_weights_only_is2 tests/test_gpu_nlg_sde.py::test_mng_psi_apf_parity tests/test_gpu_parity.py::test_latency_kernel_equals_general_kernel tests/test_gpu_parset.py tests/test_gpu_general_m.py::test_mng_general_m tests/test_gpu_general_m.py::test_post_correction_with_a_seasonal_component tests/test_gpu_user_models.py::test_user_gbm_equals_registry_gbm tests/test_gpu_postcorrect.py::test_device_approximation_equals_host_approximation tests/test_gpu_postcorrect.py::test_bsm_ng_post_correction"
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest $T -x -q -m gpu 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|Error|error:" | head -12
done
