mkdir -p gpurun_out/r2
N=$1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 20 --warmup 3 --workload synth16m > gpurun_out/r2/b19_16m_n$N.json 2> gpurun_out/r2/b19_16m_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 tools/sort_sweep.py --sizes 1,4,16 > gpurun_out/r2/sweep19_n$N.jsonl 2> gpurun_out/r2/sweep19_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus $N --steps 200 --warmup 3 > gpurun_out/r2/b19_bonsai_n$N.json 2> gpurun_out/r2/b19_bonsai_n$N.err
tail -2 gpurun_out/r2/b19_16m_n$N.err; tail -2 gpurun_out/r2/sweep19_n$N.err; tail -2 gpurun_out/r2/b19_bonsai_n$N.err
