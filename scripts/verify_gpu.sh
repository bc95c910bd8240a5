#!/bin/bash
# what the driver runs at round end, in one go on a GPU box:  bash scripts/verify_gpu.sh
set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py | tee gpurun_out/bench_verify.json | cut -c1-400
