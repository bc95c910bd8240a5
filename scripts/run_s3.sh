set -x
python bench.py --no-cpu-baseline --no-e2e > gpurun_out/s3_bench8.json 2> gpurun_out/s3_bench8.err; echo rc=$?
DAFK_CONV_WRES=0 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/s3_bench8b.json 2> gpurun_out/s3_bench8b.err; echo rc=$?
