mkdir -p gpurun_out/r2
( time python -m pytest tests -m gpu -q -x --durations=8 ) 2>&1 | tail -25 > gpurun_out/r2/t5.log
for cfg in "1 1" "1 2" "0 1"; do set -- $cfg; GS_FUSED_SORT=$1 GS_FUSED_CTAS=$2 python bench.py --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/r2/b5_bonsai_f$1_c$2.json 2> gpurun_out/r2/b5_bonsai_f$1_c$2.err; done
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload garden > gpurun_out/r2/b5_garden.json 2> gpurun_out/r2/b5_garden.err
GS_FUSED_MAX=8000000 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload garden > gpurun_out/r2/b5_garden_fused.json 2> gpurun_out/r2/b5_garden_fused.err
cat gpurun_out/r2/t5.log
