set -x
python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py tests/test_mmsdnet_gpu.py -x -q > gpurun_out/s3_tests_a.log 2>&1; echo "rc=$?" >> gpurun_out/s3_tests_a.log; tail -12 gpurun_out/s3_tests_a.log
python bench.py --no-cpu-baseline > gpurun_out/s3_bench7.json 2> gpurun_out/s3_bench7.err; echo rc=$?; tail -3 gpurun_out/s3_bench7.err
DAFK_PROFILE_SHAPES=1 python bench.py --steps 1 --warmup 2 --profile-all --no-e2e --no-cpu-baseline > gpurun_out/s3_all_shapes2.json 2> gpurun_out/s3_all_shapes2.err; echo rc=$?
