mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -q -rs 2>&1 | tail -12 > gpurun_out/r2/t22.log
timeout 600 python bench.py > gpurun_out/r2/b22_bonsai.json 2> gpurun_out/r2/b22_bonsai.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_(depth|bucket|radix|raster|project|bin|blend)" -s 39 -c 13 -o gpurun_out/r2/prof22_garden -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload garden > gpurun_out/r2/ncu22g.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_(depth|bucket|radix|raster|project|bin|blend)" -s 39 -c 13 -o gpurun_out/r2/prof22_16m -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload synth16m > gpurun_out/r2/ncu22m.log 2>&1
cat gpurun_out/r2/t22.log
