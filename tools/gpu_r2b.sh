cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out/r2b_bench.jsonl; : > $O
qb() { timeout 120 python tools/quick_bench.py "$@" 2>&1 | tail -1 | tee -a $O; }
qb --tag new
ANTQ_DEBUG=2 qb --tag new_copy
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -4
for v in _p1 _p2 _c8 _k4 _c16 _c12s16; do ANTQ_LIB_SUFFIX=$v qb --tag new$v; ANTQ_LIB_SUFFIX=$v ANTQ_DEBUG=2 qb --tag copy$v; done
ANTQ_CHUNK=4096 qb --tag new_c4k
ANTQ_DEBUG=16 qb --tag new_nofma
qb --per-tensor --tag new_pt
qb --dtype f32 --tag new_f32
qb --kind int --tag new_int
qb --olive --tag new_olive
qb --rows 8192 --cols 8192 --nb 4 --tag new_8k
qb --rows 1024 --cols 1024 --nb 16 --tag new_1k
timeout 300 ncu --set full --clock-control none --import-source on -k regex:antq_stream -s 12 -c 1 -f -o gpurun_out/stream_r2b python tools/quick_bench.py --reps 1 > gpurun_out/ncu_r2b.log 2>&1
ls -la gpurun_out | tail -5
