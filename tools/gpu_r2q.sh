cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out/r2q_bench.jsonl; : > $O
qb() { timeout 120 python tools/quick_bench.py "$@" 2>&1 | tail -1 | tee -a $O; }
qb --tag vec2
ANTQ_DEBUG=64 qb --tag vec1
qb --tag vec2_again
ANTQ_DEBUG=64 qb --tag vec1_again
qb --per-tensor --tag pt_vec2
ANTQ_DEBUG=64 qb --per-tensor --tag pt_vec1
qb --rows 8192 --cols 8192 --nb 4 --tag 8k_vec2
ANTQ_DEBUG=64 qb --rows 8192 --cols 8192 --nb 4 --tag 8k_vec1
qb --kind int --tag int4_vec2
ANTQ_DEBUG=64 qb --kind int --tag int4_vec1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
