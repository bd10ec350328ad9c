# Full GPU validation for one round: smoke, parity tests, bench, ncu launch list + full capture of the hot kernel.
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
R=${ROUND:-r01}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -12
timeout 600 python bench.py --steps 50 --warmup 5 2>&1 | tail -2 | tee gpurun_out/bench_$R.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ref_$R.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/launches_$R.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:antq_stream_kernel -s 30 -c 2 -f -o gpurun_out/stream_$R python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_$R.log 2>&1
ls -la gpurun_out | tail -8
