#!/bin/bash
# First GPU call of the next round: run every opt-in path that round 1 left compiled-but-untimed,
# each under its own timeout (the bounded mbarrier wait turns a protocol bug into a trapped launch, not a hang), and
# leave the logs in gpurun_out/.   usage:  gpurun --timeout 600 -- 'bash scripts/round2_first_call.sh'
mkdir -p gpurun_out
export DAFK_TEST_EXPERIMENTAL=1
{
  echo "== cta_group::2 haloed-tile kernel (DAFK_CONV_HALO2)"
  timeout 150 python -m pytest tests/test_conv_tc_gpu.py -m gpu -q -k cta_pairs 2>&1 | tail -15
  echo "== tiled-transpose weight packing (DAFK_PACK_TILED)"
  timeout 100 python -m pytest tests/test_conv_tc_gpu.py -m gpu -q -k tiled_transpose 2>&1 | tail -5
  echo "== 128 x 256 weight-gradient tile (DAFK_WGRAD_BN256)"
  timeout 100 python -m pytest tests/test_conv_tc_gpu.py -m gpu -q -k wide_tile 2>&1 | tail -5
  echo "== 12-warp raster-strip layout (DAFK_NC_L12)"
  timeout 150 python -m pytest tests/test_conv_nc_gpu.py -m gpu -q -k twelve 2>&1 | tail -5
  echo "== CUDA components against the reference builders' outputs"
  timeout 200 python -m pytest tests/test_zz_reference_builders_gpu.py -m gpu -q 2>&1 | tail -12
  echo "== setmaxnreg register split of the raster-strip kernel (DAFK_NC_L16)"
  timeout 150 python -m pytest tests/test_conv_nc_gpu.py -m gpu -q -k register_split 2>&1 | tail -5
  echo "== costs.py functional helpers"
  timeout 100 python -m pytest tests/test_costs_gpu.py -m gpu -q 2>&1 | tail -8
  echo "== bf16 decoder (DAFK_DEC_BF16), decoder-level comparison"
  timeout 100 python -m pytest tests/test_decoder_bf16_gpu.py -m gpu -q 2>&1 | tail -5
} > gpurun_out/r2_experimental_tests.log 2>&1
{
  for v in 0 1; do
    echo "== DAFK_CONV_HALO2=$v"
    DAFK_CONV_HALO2=$v timeout 120 python scripts/bench_tc.py "@224" 20 2>&1 | tail -6
    DAFK_CONV_HALO2=$v timeout 120 python scripts/bench_tc.py "@112" 20 2>&1 | tail -6
  done
  for v in 0 1; do
    echo "== DAFK_WGRAD_BN256=$v"
    DAFK_WGRAD_BN256=$v timeout 120 python scripts/bench_tc.py "@28" 20 2>&1 | tail -6
    DAFK_WGRAD_BN256=$v timeout 120 python scripts/bench_tc.py "@14" 20 2>&1 | tail -4
  done
  echo "== DAFK_NC_L16=1"
  DAFK_NC_L16=1 timeout 60 python scripts/bench_nc.py film8x8 50 2>&1 | tail -3
  DAFK_NC_L16=1 timeout 60 python scripts/bench_nc.py seg8x64 50 2>&1 | tail -3
  for v in 0 1; do
    echo "== DAFK_NC_L12=$v"
    DAFK_NC_L12=$v timeout 60 python scripts/bench_nc.py film8x8 50 2>&1 | tail -3
    DAFK_NC_L12=$v timeout 60 python scripts/bench_nc.py seg8x64 50 2>&1 | tail -3
  done
} > gpurun_out/r2_experimental_bench.log 2>&1
for v in 0 1; do
  echo "== whole step, DAFK_PACK_TILED=$v" >> gpurun_out/r2_experimental_bench.log
  DAFK_PACK_TILED=$v timeout 150 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200 >> gpurun_out/r2_experimental_bench.log
done
tail -40 gpurun_out/r2_experimental_tests.log
tail -40 gpurun_out/r2_experimental_bench.log
