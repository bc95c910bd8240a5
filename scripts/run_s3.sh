set -x
python -m pytest tests -m gpu -x -q > gpurun_out/s3_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s3_tests.log; tail -4 gpurun_out/s3_tests.log
python bench.py > gpurun_out/s3_bench11.json 2> gpurun_out/s3_bench11.err; echo rc=$?
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s3_ref2.json 2> gpurun_out/s3_ref2.err; echo rc=$?
