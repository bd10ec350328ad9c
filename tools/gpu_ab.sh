# Same-box A/B of the working tree (libantq.so) against a variant library built from another revision / define set
# (tools/build_variant.py <suffix> ...):   OLD=_base bash tools/gpu_ab.sh
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out/ab_bench.jsonl; : > $O
OLD=${OLD:-_old}
qb() { timeout 120 python tools/quick_bench.py "$@" 2>&1 | tail -1 | tee -a $O; }
for i in 1 2 3; do qb --tag new$i; ANTQ_LIB_SUFFIX=$OLD qb --tag old$i; done
qb --per-tensor --tag pt_new; ANTQ_LIB_SUFFIX=$OLD qb --per-tensor --tag pt_old
qb --rows 8192 --cols 8192 --nb 4 --tag 8k_new; ANTQ_LIB_SUFFIX=$OLD qb --rows 8192 --cols 8192 --nb 4 --tag 8k_old
qb --olive --tag olive_new; ANTQ_LIB_SUFFIX=$OLD qb --olive --tag olive_old
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
