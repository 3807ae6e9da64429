mkdir -p gpurun_out/r2
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 200 python bench.py --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/r2/b25_bonsai.json 2> gpurun_out/r2/b25.err; tail -1 gpurun_out/r2/b25.err
