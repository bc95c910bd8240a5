set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/s3_n2.json 2> gpurun_out/s3_n2.err; echo rc=$?
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --l_mix 0.5 --no-e2e > gpurun_out/s3_n2_lmix05.json 2> gpurun_out/s3_n2_lmix05.err; echo rc=$?
tail -2 gpurun_out/s3_n2.err gpurun_out/s3_n2_lmix05.err
