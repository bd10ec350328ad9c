cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out/ab_bench.jsonl; : > $O
qb() { timeout 120 python tools/quick_bench.py "$@" 2>&1 | tail -1 | tee -a $O; }
for i in 1 2 3; do qb --tag new$i; ANTQ_LIB_SUFFIX=_old qb --tag old$i; done
qb --per-tensor --tag pt_new; ANTQ_LIB_SUFFIX=_old qb --per-tensor --tag pt_old
qb --rows 8192 --cols 8192 --nb 4 --tag 8k_new; ANTQ_LIB_SUFFIX=_old qb --rows 8192 --cols 8192 --nb 4 --tag 8k_old
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
