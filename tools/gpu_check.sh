set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60
timeout 600 python bench.py --steps 50 --warmup 5 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1
tail -5 gpurun_out/launches_r01.csv
ncu --set full --clock-control none --import-source on -k regex:antq_rows_kernel -s 30 -c 3 -f -o gpurun_out/rows_r01 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
