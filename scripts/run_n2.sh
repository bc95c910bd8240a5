set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/s3_n2_final.json 2> gpurun_out/s3_n2_final.err; echo rc=$?
tail -n 2 gpurun_out/s3_n2_final.err
