mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r2/t23.log
timeout 300 python bench.py --steps 200 --warmup 3 --no-cpu-baseline > gpurun_out/r2/b23_bonsai.json 2> gpurun_out/r2/b23_bonsai.err
timeout 300 python bench.py --steps 120 --warmup 3 --no-cpu-baseline --workload garden > gpurun_out/r2/b23_garden.json 2> gpurun_out/r2/b23_garden.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload synth16m > gpurun_out/r2/b23_16m.json 2> gpurun_out/r2/b23_16m.err
GS_BIN_COMPACT=0 timeout 300 python bench.py --steps 200 --warmup 3 --no-cpu-baseline > gpurun_out/r2/b23_bonsai_nocompact.json 2> gpurun_out/r2/b23_bonsai_nocompact.err
cat gpurun_out/r2/t23.log
