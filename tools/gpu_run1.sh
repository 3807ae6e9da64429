mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2/t1.log
for cfg in "1 1" "2 1" "2 2"; do set -- $cfg; GS_BIN=$1 GS_BLEND=$2 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2/b1_bonsai_$1$2.json 2> gpurun_out/r2/b1_bonsai_$1$2.err; done
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload synth16m > gpurun_out/r2/b1_16m.json 2> gpurun_out/r2/b1_16m.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload garden > gpurun_out/r2/b1_garden.json 2> gpurun_out/r2/b1_garden.err
cat gpurun_out/r2/t1.log
