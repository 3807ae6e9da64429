mkdir -p gpurun_out/r2
python -m pytest tests/test_multi_gpu.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r2/t4.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 3 > gpurun_out/r2/b4_n2.json 2> gpurun_out/r2/b4_n2.err
python bench.py --gpus 1 --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/r2/b4_n1.json 2> gpurun_out/r2/b4_n1.err
cat gpurun_out/r2/t4.log
