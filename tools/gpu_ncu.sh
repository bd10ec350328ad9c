cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --warp-sampling-interval 0 -k regex:antq_stream -s 12 -c 1 -f -o gpurun_out/stream_${TAG:-x} python tools/quick_bench.py --reps 1 > gpurun_out/ncu_${TAG:-x}.log 2>&1
tail -3 gpurun_out/ncu_${TAG:-x}.log
ls -la gpurun_out/*.ncu-rep
