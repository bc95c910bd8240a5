set -x
python bench.py > gpurun_out/s3_bench3.json 2> gpurun_out/s3_bench3.err; echo rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -s 2600 -c 2390 --csv --log-file gpurun_out/s3_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/s3_ncu_launch.log 2>&1; echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:conv_tc_ -s 200 -c 12 -o gpurun_out/s3_conv_tc python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/s3_ncu_full.log 2>&1; echo rc=$?
ls -la gpurun_out/*.ncu-rep
