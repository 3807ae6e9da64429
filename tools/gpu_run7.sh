mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2/t7.log
python bench.py --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/r2/b7_bonsai.json 2> gpurun_out/r2/b7_bonsai.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_blend2|k_bin_" -s 40 -c 4 -o gpurun_out/r2/prof7_bonsai -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2/ncu7.log 2>&1
cat gpurun_out/r2/t7.log
