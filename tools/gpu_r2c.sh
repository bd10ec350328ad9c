cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out/r2c_bench.jsonl; : > $O
qb() { timeout 120 python tools/quick_bench.py "$@" 2>&1 | tail -1 | tee -a $O; }
for s in 24 20 16 12 8 4; do ANTQ_STAGES=$s qb --tag st$s; done
for s in 24 16 8; do ANTQ_STAGES=$s ANTQ_DEBUG=2 qb --tag copy_st$s; done
ANTQ_LIB_SUFFIX=_trace timeout 120 python tools/trace_stream.py 2>&1 | tail -30 | tee gpurun_out/trace_s24.txt
ANTQ_STAGES=12 ANTQ_LIB_SUFFIX=_trace timeout 120 python tools/trace_stream.py 2>&1 | tail -30 | tee gpurun_out/trace_s12.txt
TRACE_N=8192 ANTQ_LIB_SUFFIX=_trace timeout 120 python tools/trace_stream.py 2>&1 | tail -30 | tee gpurun_out/trace_8k.txt
