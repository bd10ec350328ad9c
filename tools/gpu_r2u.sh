cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out/r2u_bench.jsonl; : > $O
qb() { timeout 120 python tools/quick_bench.py "$@" 2>&1 | tail -1 | tee -a $O; }
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "short" 2>&1 | tail -5
for c in 8 16 32 64 256; do r=$((16777216 / c)); qb --rows $r --cols $c --nb 4 --tag group$c; done
qb --rows 524288 --cols 32 --nb 4 --flat --tag group32_flat
qb --rows 524288 --cols 32 --nb 4 --kind int --tag group32_int
qb --rows 524288 --cols 32 --nb 4 --unsigned --tag group32_flint_u
qb --rows 524288 --cols 32 --nb 4 --olive --tag group32_olive
qb --rows 524288 --cols 32 --nb 4 --dtype f32 --tag group32_f32
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3
