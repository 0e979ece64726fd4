#!/bin/bash
# compute-sanitizer passes on the small fixtures (the reference has no race / memory checking at all).
#   tools/sanitize.sh            memcheck on the small parity cases + the reverse-mode tests
#   tools/sanitize.sh racecheck  shared-memory hazards of the same cases (slow: the tcgen05 kernels synchronise through mbarriers
#                                and TMEM, which racecheck does not model -- read its report for the FFT / SIMT kernels only)
# Run on the GPU box (gpurun); writes gpurun_out/sanitize_<tool>.log.
TOOL=${1:-memcheck}
mkdir -p gpurun_out
export TFPNP_TEST_GRAD=1
timeout 1700 compute-sanitizer --tool "$TOOL" --error-exitcode 9 --launch-timeout 120 \
  python -m pytest tests/test_gpu_parity.py tests/test_grad.py tests/test_csmri_variants.py -m gpu -q -x --timeout 1500 -p no:cacheprovider \
  -k "small or golden or grad or backward or vjp or variants" > "gpurun_out/sanitize_${TOOL}.log" 2>&1
echo "exit $?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|Race" "gpurun_out/sanitize_${TOOL}.log" | tail -12
