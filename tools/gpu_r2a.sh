# Stream kernel bring-up: parity tests, A/B against the row kernel, tuning variants.
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out/r2a_bench.jsonl; : > $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
qb() { timeout 120 python tools/quick_bench.py "$@" 2>&1 | tail -1 | tee -a $O; }
qb --tag new
ANTQ_KERNEL=rows qb --tag old
ANTQ_DEBUG=2 qb --tag new_copy
qb --torch-copy --tag torch_copy
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -15
ANTQ_CHUNK=4096 qb --tag new_c4k
ANTQ_CHUNK=2048 qb --tag new_c2k
ANTQ_DEBUG=16 qb --tag new_nofma
for v in _c15 _c8 _c10s20 _c12s16; do ANTQ_LIB_SUFFIX=$v qb --tag new$v; ANTQ_LIB_SUFFIX=$v ANTQ_DEBUG=2 qb --tag copy$v; done
ANTQ_LIB_SUFFIX=_c15 ANTQ_CHUNK=4096 qb --tag new_c15_c4k
qb --per-tensor --tag new_pt
ANTQ_KERNEL=rows qb --per-tensor --tag old_pt
qb --dtype f32 --tag new_f32
ANTQ_KERNEL=rows qb --dtype f32 --tag old_f32
qb --dtype bf16 --tag new_bf16
qb --kind int --tag new_int
ANTQ_KERNEL=rows qb --kind int --tag old_int
qb --olive --tag new_olive
ANTQ_KERNEL=rows qb --olive --tag old_olive
qb --rows 8192 --cols 8192 --nb 4 --tag new_8k
qb --rows 1024 --cols 1024 --nb 16 --tag new_1k
ANTQ_KERNEL=rows qb --rows 1024 --cols 1024 --nb 16 --tag old_1k
qb --rows 16384 --cols 4096 --nb 4 --tag new_16kx4k
