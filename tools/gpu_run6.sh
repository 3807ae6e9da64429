mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -q -x 2>&1 | tail -12 > gpurun_out/r2/t6.log
python bench.py --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/r2/b6_bonsai.json 2> gpurun_out/r2/b6_bonsai.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload garden > gpurun_out/r2/b6_garden.json 2> gpurun_out/r2/b6_garden.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload synth16m > gpurun_out/r2/b6_16m.json 2> gpurun_out/r2/b6_16m.err
python tools/sort_sweep.py --sizes 1,4,16 > gpurun_out/r2/sweep6.jsonl 2> gpurun_out/r2/sweep6.err
cat gpurun_out/r2/t6.log
