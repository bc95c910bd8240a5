set -x
python -m pytest tests -m gpu -x -q > gpurun_out/s3_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s3_tests.log; tail -4 gpurun_out/s3_tests.log
python bench.py --no-cpu-baseline --no-e2e > gpurun_out/s3_bench10.json 2> gpurun_out/s3_bench10.err; echo rc=$?
