mkdir -p gpurun_out/r2
timeout 150 python -m pytest tests/test_raster_gpu.py -m gpu -q -x -k "pipelined or prefetch" 2>&1 | tail -3
timeout 100 python bench.py --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/r2/b26_bonsai.json 2> gpurun_out/r2/b26.err; tail -1 gpurun_out/r2/b26.err
