mkdir -p gpurun_out/r2
N=$1
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -x -rs 2>&1 | tail -8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus $N --steps 200 --warmup 3 --no-cpu-baseline > gpurun_out/r2/b21_bonsai_n$N.json 2> gpurun_out/r2/b21_bonsai_n$N.err
GS_PEER_DOUBLE=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 200 --warmup 3 --no-cpu-baseline > gpurun_out/r2/b21_bonsai_single_n$N.json 2> gpurun_out/r2/b21_bonsai_single_n$N.err
tail -3 gpurun_out/r2/b21_bonsai_n$N.err
