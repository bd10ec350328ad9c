cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out/r2o_bench.jsonl; : > $O
qb() { timeout 120 python tools/quick_bench.py "$@" 2>&1 | tail -1 | tee -a $O; }
ANTQ_CHUNK=4096 qb --tag c4k
for v in _k4c12 _k4c14b2 _k4c16b2 _k4c16b1 _k4c20b1; do
  ANTQ_LIB_SUFFIX=$v qb --tag hl$v
  ANTQ_LIB_SUFFIX=$v qb --rows 8192 --cols 8192 --nb 4 --tag 8k$v
  ANTQ_LIB_SUFFIX=$v qb --per-tensor --tag pt$v
done
ANTQ_CHUNK=4096 qb --rows 8192 --cols 8192 --nb 4 --tag 8k_c4k
qb --rows 8192 --cols 8192 --nb 4 --tag 8k_c8k
ANTQ_CHUNK=4096 qb --per-tensor --tag pt_c4k
ANTQ_CHUNK=4096 ANTQ_DEBUG=16 qb --tag c4k_nofma
ANTQ_CHUNK=4096 ANTQ_DEBUG=2 qb --tag c4k_copy
