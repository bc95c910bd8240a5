#!/bin/bash
# compute-sanitizer passes over the kernel tests with hand-rolled mbarrier pipelines / global reductions
# (missing 7).  usage:  gpurun --timeout 2400 -- 'bash scripts/sanitize.sh'
# memcheck: out-of-bounds / misaligned accesses;  racecheck: shared-memory hazards between warps of a CTA.
# Logs -> gpurun_out/r2_sanitizer_<tool>.log; the summaries (tool banner, test result, ERROR SUMMARY) are copied to profiles/.
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
run() {   # tool, seconds, pytest args...
  tool=$1; secs=$2; shift; shift
  echo "== $tool: pytest $*"
  timeout $secs $CS --tool $tool --print-limit 3 --error-exitcode 0 python -m pytest -m gpu -q -x -p no:cacheprovider "$@" 2>&1 | tail -12
  echo "== exit: ${PIPESTATUS[0]}"
}
for tool in memcheck racecheck; do
  {
    run $tool 420 tests/test_conv_tc_gpu.py -k "(cta_pairs or wgrad_halo or dgrad or tiled or test_conv3x3_tc_forward or stride2) and (case0 or case1 or shape0)"
    run $tool 420 tests/test_conv_nc_gpu.py -k "(twelve or bulk or wgrad or forward) and case0"
    run $tool 300 tests/test_kernels_gpu.py -k "tps or bn_ or softmax or film"
  } > gpurun_out/r2_sanitizer_$tool.log 2>&1
  tail -40 gpurun_out/r2_sanitizer_$tool.log
done
