cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out/r2l_bench.jsonl; : > $O
qb() { timeout 120 python tools/quick_bench.py "$@" 2>&1 | tail -1 | tee -a $O; }
qb --tag new
for v in _c8 _c13 _k4r4 _k4r3 _k4r3c16; do ANTQ_LIB_SUFFIX=$v qb --tag new$v; done
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -4
ANTQ_CHUNK=4096 qb --tag c4k
ANTQ_PDL=0 qb --tag nopdl
ANTQ_DEBUG=2 qb --tag new_copy
ANTQ_LIB_SUFFIX=_trace ANTQ_PDL=0 timeout 120 python tools/trace_stream.py 2>&1 | tail -32 | tee gpurun_out/trace_l.txt
qb --alpha-mult 1.0 --tag new_a1.0
qb --per-tensor --tag new_pt
qb --dtype f32 --tag new_f32
qb --dtype bf16 --tag new_bf16
qb --kind int --tag new_int
qb --olive --tag new_olive
qb --rows 8192 --cols 8192 --nb 4 --tag new_8k
qb --rows 1024 --cols 1024 --nb 16 --tag new_1k
qb --rows 16384 --cols 1024 --nb 8 --tag new_16kx1k
qb --rows 2048 --cols 2048 --nb 16 --tag new_2k
