mkdir -p gpurun_out/r2
for i in 1 2; do
timeout 200 python bench.py --steps 200 --warmup 3 --no-cpu-baseline > gpurun_out/r2/b24_bonsai_c$i.json 2> gpurun_out/r2/b24_c$i.err
GS_BIN_COMPACT=0 timeout 200 python bench.py --steps 200 --warmup 3 --no-cpu-baseline > gpurun_out/r2/b24_bonsai_n$i.json 2> gpurun_out/r2/b24_n$i.err
done
