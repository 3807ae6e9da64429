mkdir -p gpurun_out/r2
GS_BLEND_WIDE=2 timeout 600 python -m pytest tests/test_raster_gpu.py tests/test_multi_gpu.py -m gpu -q -x 2>&1 | tail -5
for rd in 4 2; do
GS_BLEND_WIDE=2 GS_BLEND_ROUNDS=$rd timeout 300 python bench.py --steps 200 --warmup 3 --no-cpu-baseline > gpurun_out/r2/b16_bonsai_w2_r$rd.json 2> gpurun_out/r2/b16_bonsai_w2_r$rd.err
GS_BLEND_WIDE=2 GS_BLEND_ROUNDS=$rd timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload garden > gpurun_out/r2/b16_garden_w2_r$rd.json 2> gpurun_out/r2/b16_garden_w2_r$rd.err
done
timeout 300 python bench.py --steps 200 --warmup 3 --no-cpu-baseline > gpurun_out/r2/b16_bonsai_w1.json 2> gpurun_out/r2/b16_bonsai_w1.err
