mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2/t2.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2/b2_bonsai.json 2> gpurun_out/r2/b2_bonsai.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload garden > gpurun_out/r2/b2_garden.json 2> gpurun_out/r2/b2_garden.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload synth16m > gpurun_out/r2/b2_16m.json 2> gpurun_out/r2/b2_16m.err
GS_BLEND=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload garden > gpurun_out/r2/b2_garden_blend1.json 2> gpurun_out/r2/b2_garden_blend1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_blend2|k_bin_|k_project|k_radix_scatter|k_bucket|k_depth" -s 60 -c 14 -o gpurun_out/r2/prof2_bonsai -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2/ncu2.log 2>&1
cat gpurun_out/r2/t2.log
