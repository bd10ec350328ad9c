cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out/r2s_bench.jsonl; : > $O
qb() { timeout 120 python tools/quick_bench.py "$@" 2>&1 | tail -1 | tee -a $O; }
qb --tag th12
for t in 0 6 24 55; do ANTQ_TAILHALVES=$t qb --tag th$t; done
qb --tag th12_again
ANTQ_TAILHALVES=0 qb --tag th0_again
qb --per-tensor --tag pt_th12
ANTQ_TAILHALVES=0 qb --per-tensor --tag pt_th0
qb --rows 8192 --cols 8192 --nb 4 --tag 8k_th12
ANTQ_TAILHALVES=0 qb --rows 8192 --cols 8192 --nb 4 --tag 8k_th0
qb --rows 2048 --cols 2048 --nb 16 --tag 2k_th12
ANTQ_TAILHALVES=0 qb --rows 2048 --cols 2048 --nb 16 --tag 2k_th0
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
