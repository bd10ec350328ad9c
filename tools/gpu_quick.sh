cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python tools/quick_bench.py --tag full
ANTQ_DEBUG=7 python tools/quick_bench.py --tag d7
ANTQ_DEBUG=4 python tools/quick_bench.py --tag d4_noprologue
ANTQ_DEBUG=2 python tools/quick_bench.py --tag d2_nochain
ANTQ_DEBUG=3 python tools/quick_bench.py --tag d3_nochain_nostore
python tools/quick_bench.py --per-tensor --tag pertensor
ncu --set full --clock-control none --import-source on -k regex:antq_rows_kernel -s 20 -c 1 -f -o gpurun_out/rows_persist python tools/quick_bench.py --reps 2 > gpurun_out/ncu_p.log 2>&1
