mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -q -rs 2>&1 | tail -25 > gpurun_out/r2/t20.log
timeout 600 python bench.py --steps 200 --warmup 3 > gpurun_out/r2/b20_bonsai.json 2> gpurun_out/r2/b20_bonsai.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2/b20_ref.json 2> gpurun_out/r2/b20_ref.err
timeout 300 python bench.py --steps 120 --warmup 3 --no-cpu-baseline --workload garden > gpurun_out/r2/b20_garden.json 2> gpurun_out/r2/b20_garden.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload synth16m > gpurun_out/r2/b20_16m.json 2> gpurun_out/r2/b20_16m.err
timeout 300 python tools/sort_sweep.py --sizes 1,2,4,8,16 > gpurun_out/r2/sweep20.jsonl 2> gpurun_out/r2/sweep20.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/r2/launches20.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2/ncu20a.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_(depth|bucket|radix|raster|project|bin|blend)" -s 39 -c 13 -o gpurun_out/r2/prof20_bonsai -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2/ncu20.log 2>&1
cat gpurun_out/r2/t20.log
