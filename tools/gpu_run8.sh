mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -q -x --deselect tests/test_baseline_configs_gpu.py 2>&1 | tail -8 > gpurun_out/r2/t8.log
python bench.py --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/r2/b8_bonsai.json 2> gpurun_out/r2/b8_bonsai.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload garden > gpurun_out/r2/b8_garden.json 2> gpurun_out/r2/b8_garden.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload synth16m > gpurun_out/r2/b8_16m.json 2> gpurun_out/r2/b8_16m.err
cat gpurun_out/r2/t8.log
