set -x
python -m pytest tests/test_conv1x1_gpu.py tests/test_kernels_gpu.py -x -q > gpurun_out/s3_tests_a.log 2>&1; echo "rc=$?" >> gpurun_out/s3_tests_a.log; tail -5 gpurun_out/s3_tests_a.log
python scripts/bench_bw.py > gpurun_out/s3_bw2.txt 2> gpurun_out/s3_bw2.err; echo rc=$?; tail -3 gpurun_out/s3_bw2.err
python -m pytest tests -m gpu -x -q > gpurun_out/s3_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s3_tests.log; tail -5 gpurun_out/s3_tests.log
python bench.py --no-cpu-baseline > gpurun_out/s3_bench2.json 2> gpurun_out/s3_bench2.err; echo rc=$?; tail -3 gpurun_out/s3_bench2.err
