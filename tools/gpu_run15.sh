mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_raster_gpu.py tests/test_multi_gpu.py -m gpu -q -x 2>&1 | tail -5
timeout 300 python bench.py --steps 200 --warmup 3 --no-cpu-baseline > gpurun_out/r2/b15_bonsai.json 2> gpurun_out/r2/b15_bonsai.err
GS_BLEND_ROUNDS=2 timeout 300 python bench.py --steps 200 --warmup 3 --no-cpu-baseline > gpurun_out/r2/b15_bonsai_r2.json 2> gpurun_out/r2/b15_bonsai_r2.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload garden > gpurun_out/r2/b15_garden.json 2> gpurun_out/r2/b15_garden.err
GS_BLEND_ROUNDS=2 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload garden > gpurun_out/r2/b15_garden_r2.json 2> gpurun_out/r2/b15_garden_r2.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload synth16m > gpurun_out/r2/b15_16m.json 2> gpurun_out/r2/b15_16m.err
