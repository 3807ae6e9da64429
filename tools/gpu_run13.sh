mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_baseline_configs_gpu.py 2>&1 | tail -8 > gpurun_out/r2/t13.log
timeout 300 python bench.py --steps 200 --warmup 3 --no-cpu-baseline > gpurun_out/r2/b13_bonsai.json 2> gpurun_out/r2/b13_bonsai.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload garden > gpurun_out/r2/b13_garden.json 2> gpurun_out/r2/b13_garden.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload synth16m > gpurun_out/r2/b13_16m.json 2> gpurun_out/r2/b13_16m.err
cat gpurun_out/r2/t13.log
