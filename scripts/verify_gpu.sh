set -x
python -m pytest tests -m gpu -x -q > gpurun_out/s3_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s3_tests.log; tail -3 gpurun_out/s3_tests.log
python bench.py > gpurun_out/s3_bench12.json 2> gpurun_out/s3_bench12.err; echo rc=$?
