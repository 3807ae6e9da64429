mkdir -p gpurun_out/r2
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 100 --warmup 3 > gpurun_out/r2/b9_bonsai_n$N.json 2> gpurun_out/r2/b9_bonsai_n$N.err
tail -2 gpurun_out/r2/b9_bonsai_n$N.err
