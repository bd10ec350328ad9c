cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for sfx in _r2c2 _r2c4 _r4c1 _r2c8; do
ANTQ_LIB_SUFFIX=$sfx python tools/quick_bench.py --tag full$sfx
ANTQ_LIB_SUFFIX=$sfx python tools/quick_bench.py --per-tensor --tag pertensor$sfx
ANTQ_DEBUG=7 ANTQ_LIB_SUFFIX=$sfx python tools/quick_bench.py --per-tensor --tag d7$sfx
done
