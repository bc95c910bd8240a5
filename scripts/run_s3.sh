timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_golden_gpu.py -x -q -k "tps or spline or golden" 2>&1 | tail -4
python scripts/bench_bw.py --only tps
