#!/bin/bash
# compute-sanitizer passes over the kernel tests with hand-rolled mbarrier pipelines / global reductions
# (missing 7).  usage:  gpurun --timeout 1500 -- 'bash scripts/sanitize.sh'
# memcheck: out-of-bounds / misaligned accesses;  racecheck: shared-memory hazards between warps of a CTA.
# Logs -> gpurun_out/r2_sanitizer_*.log (summaries are copied to profiles/ by hand).
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
K_TC="tests/test_conv_tc_gpu.py -k forward_cta_pairs or wgrad_halo or tc_dgrad"
for tool in memcheck racecheck; do
  echo "== $tool: tests/test_conv_tc_gpu.py (subset: haloed / CTA-pair forward, dgrad, wgrad)" 
  timeout 600 $CS --tool $tool --print-limit 5 --error-exitcode 0 python -m pytest tests/test_conv_tc_gpu.py -m gpu -q -x \
      -k "cta_pairs or wgrad_halo or dgrad or tiled or test_conv3x3_tc_forward" 2>&1 | tail -25
  echo "== $tool: tests/test_conv_nc_gpu.py (first cases)"
  timeout 600 $CS --tool $tool --print-limit 5 --error-exitcode 0 python -m pytest tests/test_conv_nc_gpu.py -m gpu -q -x \
      -k "twelve or bulk or wgrad" 2>&1 | tail -25
  echo "== $tool: tests/test_kernels_gpu.py -k tps"
  timeout 600 $CS --tool $tool --print-limit 5 --error-exitcode 0 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "tps" 2>&1 | tail -25
done
