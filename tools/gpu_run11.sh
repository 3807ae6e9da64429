mkdir -p gpurun_out/r2
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 20 --warmup 3 --workload synth16m > gpurun_out/r2/b11_16m_n$N.json 2> gpurun_out/r2/b11_16m_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 tools/sort_sweep.py --sizes 1,4,16 > gpurun_out/r2/sweep11_n$N.jsonl 2> gpurun_out/r2/sweep11_n$N.err
tail -2 gpurun_out/r2/b11_16m_n$N.err; tail -2 gpurun_out/r2/sweep11_n$N.err
