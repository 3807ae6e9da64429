mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2/t3.log
for cfg in "1 0" "0 0" "1 1"; do set -- $cfg; GS_PDL=$1 GS_BINCFG=$2 python bench.py --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/r2/b3_bonsai_pdl$1_cfg$2.json 2> gpurun_out/r2/b3_bonsai_pdl$1_cfg$2.err; done
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload garden > gpurun_out/r2/b3_garden.json 2> gpurun_out/r2/b3_garden.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload synth16m > gpurun_out/r2/b3_16m.json 2> gpurun_out/r2/b3_16m.err
cat gpurun_out/r2/t3.log
