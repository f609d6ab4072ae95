#!/bin/bash
# compute-sanitizer passes over a few small parity tests (memcheck: out-of-bounds / misaligned;
# racecheck: shared-memory hazards in the collectives; synccheck: barrier misuse)
T="tests/test_gpu_parity.py::test_ragged_and_tiny_sizes_parity tests/test_gpu_parity.py::test_results_do_not_depend_on_launch_geometry tests/test_gpu_postcorrect.py::test_weights_only_is2 tests/test_gpu_nlg_sde.py::test_mng_psi_apf_parity"
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest $T -x -q -m gpu 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|Error|error:" | head -12
done
