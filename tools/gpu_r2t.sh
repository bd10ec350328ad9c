cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out/r2t_bench.jsonl; : > $O
qb() { timeout 120 python tools/quick_bench.py "$@" 2>&1 | tail -1 | tee -a $O; }
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "olive" 2>&1 | tail -3
qb --olive --tag olive_2phase
ANTQ_DEBUG=128 qb --olive --tag olive_1phase
qb --olive --per-tensor --tag olive_pt_2phase
ANTQ_DEBUG=128 qb --olive --per-tensor --tag olive_pt_1phase
qb --olive --kind int --tag oliveint_2phase
ANTQ_DEBUG=128 qb --olive --kind int --tag oliveint_1phase
qb --olive --dtype bf16 --tag olive_bf16
qb --olive --rows 8192 --cols 8192 --nb 4 --tag olive_8k
qb --tag flint
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_modules.py -m gpu -q 2>&1 | tail -3
