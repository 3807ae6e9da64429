mkdir -p gpurun_out/r2
timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_baseline_configs_gpu.py 2>&1 | tail -6
timeout 300 python bench.py --steps 200 --warmup 3 --no-cpu-baseline > gpurun_out/r2/b18_bonsai.json 2> gpurun_out/r2/b18_bonsai.err
GS_SORT_FUSE_TILES=0 timeout 300 python bench.py --steps 200 --warmup 3 --no-cpu-baseline > gpurun_out/r2/b18_bonsai_nofuse.json 2> gpurun_out/r2/b18_bonsai_nofuse.err
timeout 300 python tools/sort_sweep.py --sizes 1,2 > gpurun_out/r2/sweep18.jsonl 2> gpurun_out/r2/sweep18.err
GS_SORT_FUSE_TILES=0 timeout 300 python tools/sort_sweep.py --sizes 1,2 > gpurun_out/r2/sweep18_nofuse.jsonl 2> gpurun_out/r2/sweep18_nofuse.err
GS_SORT_FUSE_TILES=512 timeout 300 python tools/sort_sweep.py --sizes 1,2 > gpurun_out/r2/sweep18_512.jsonl 2> gpurun_out/r2/sweep18_512.err
