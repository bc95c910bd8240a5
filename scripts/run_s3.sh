set -x
python -m pytest tests -m gpu -x -q > gpurun_out/s3_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s3_tests.log; tail -12 gpurun_out/s3_tests.log
python bench.py --no-cpu-baseline --no-e2e > gpurun_out/s3_bench8.json 2> gpurun_out/s3_bench8.err; echo rc=$?
DAFK_CONV_WRES=0 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/s3_bench8b.json 2> gpurun_out/s3_bench8b.err; echo rc=$?
