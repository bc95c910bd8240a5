timeout 300 python -m pytest tests/test_conv_nc_gpu.py -x -q 2>&1 | tail -3
NC_B=192 timeout 200 python scripts/bench_nc.py film8x8 10 2>&1 | tail -3
timeout 200 python scripts/bench_nc.py "" 10 2>&1 | tail -12
