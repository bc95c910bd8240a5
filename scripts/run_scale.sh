#!/bin/bash
# usage (inside gpurun --gpus N): bash scripts/run_scale.sh N [extra bench.py flags...]   -> gpurun_out/r2_scale_n<N><tag>.json
N=$1; shift
tag=$(echo "$*" | tr -d ' -' | tr '.' 'p')
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/r2_scale_n${N}${tag}.json 2> gpurun_out/r2_scale_n${N}${tag}.err
echo rc=$?
grep -c "NCCL INFO" gpurun_out/r2_scale_n${N}${tag}.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_scale_n${N}${tag}.json').read().strip().splitlines()[-1]);print('N=%d'%d['n_gpus'],d['config']['workload'][:60],d['ms_per_step'],d['value'],d['e2e']['value'],d['clocks'])"
