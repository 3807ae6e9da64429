mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2/t12.log
python bench.py --steps 200 --warmup 3 > gpurun_out/r2/b12_bonsai.json 2> gpurun_out/r2/b12_bonsai.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2/b12_ref.json 2> gpurun_out/r2/b12_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/r2/launches12.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2/ncu12a.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_(depth|bucket|radix|raster|project|bin|blend)" -s 39 -c 13 -o gpurun_out/r2/prof12_bonsai -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2/ncu12.log 2>&1
cat gpurun_out/r2/t12.log
