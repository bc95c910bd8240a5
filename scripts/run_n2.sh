set -x
# N=2 bench with the overlapped all-reduce (default) and with one all-reduce per bucket after the backward pass
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_n2_overlap.json 2> gpurun_out/r2_n2_overlap.err; echo rc=$?
DAFK_AR_OVERLAP=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_n2_serial.json 2> gpurun_out/r2_n2_serial.err; echo rc=$?
tail -n 3 gpurun_out/r2_n2_overlap.err
for f in gpurun_out/r2_n2_overlap.json gpurun_out/r2_n2_serial.json; do python -c "
import json,sys;d=json.loads(open('$f').read().strip().splitlines()[-1]);print('$f',d['n_gpus'],d['ms_per_step'],d['value'],d['e2e']['value'])"; done
