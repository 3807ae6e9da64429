mkdir -p gpurun_out/r2
for tma in 0 1; do for rd in 4 2; do
GS_BLEND_TMA=$tma GS_BLEND_ROUNDS=$rd timeout 300 python bench.py --steps 200 --warmup 3 --no-cpu-baseline > gpurun_out/r2/b14_bonsai_t${tma}_r${rd}.json 2> gpurun_out/r2/b14_bonsai_t${tma}_r${rd}.err
done; done
for tma in 0 1; do
GS_BLEND_TMA=$tma timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload synth16m > gpurun_out/r2/b14_16m_t${tma}.json 2> gpurun_out/r2/b14_16m_t${tma}.err
GS_BLEND_TMA=$tma GS_BLEND_ROUNDS=2 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload garden > gpurun_out/r2/b14_garden_t${tma}_r2.json 2> gpurun_out/r2/b14_garden_t${tma}_r2.err
done
timeout 600 python -m pytest tests/test_raster_gpu.py -m gpu -q -x 2>&1 | tail -3
