set -x
ncu --metrics gpu__time_duration.sum --clock-control none -s 2600 -c 2390 --csv --log-file gpurun_out/s3_launches2.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/s3_ncu_launch2.log 2>&1; echo rc=$?
python scripts/bench_bw.py > gpurun_out/s3_bw3.txt 2>&1; tail -3 gpurun_out/s3_bw3.txt
