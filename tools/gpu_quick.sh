cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_reference_kernel.py -m gpu -q 2>&1 | tail -8
python tools/ref_gpu_bench.py 2>&1 | tail -2 | tee gpurun_out/ref_gpu_bench_r01.json
python tools/calib_bench.py 2>&1 | tail -1 | tee gpurun_out/calib_bench_r01.json
