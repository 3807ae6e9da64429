mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_raster_gpu.py tests/test_multi_gpu.py tests/test_sort_gpu.py -m gpu -q -x 2>&1 | tail -5
timeout 300 python bench.py --steps 200 --warmup 3 --no-cpu-baseline > gpurun_out/r2/b17_bonsai.json 2> gpurun_out/r2/b17_bonsai.err
timeout 300 python bench.py --steps 120 --warmup 3 --no-cpu-baseline --workload garden > gpurun_out/r2/b17_garden.json 2> gpurun_out/r2/b17_garden.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload synth16m > gpurun_out/r2/b17_16m.json 2> gpurun_out/r2/b17_16m.err
tail -3 gpurun_out/r2/b17_bonsai.err
