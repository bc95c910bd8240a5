set -x
python -m pytest tests -m gpu -x -q > gpurun_out/s3_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s3_tests.log; tail -12 gpurun_out/s3_tests.log
python bench.py --no-cpu-baseline > gpurun_out/s3_bench6.json 2> gpurun_out/s3_bench6.err; echo rc=$?; tail -3 gpurun_out/s3_bench6.err
