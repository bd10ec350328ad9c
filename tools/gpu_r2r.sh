cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out/r2r_bench.jsonl; : > $O
qb() { timeout 120 python tools/quick_bench.py "$@" 2>&1 | tail -1 | tee -a $O; }
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "dependent" 2>&1 | tail -3
for shape in "65536 512" "32768 1024" "16384 2048" "131072 256"; do set -- $shape
  qb --rows $1 --cols $2 --nb 4 --tag stream_$1x$2
  qb --rows $1 --cols $2 --nb 4 --flat --tag flat_$1x$2
done
qb --rows 1024 --cols 1024 --nb 16 --tag stream_1k
qb --rows 1024 --cols 1024 --nb 16 --flat --tag flat_1k
qb --rows 512 --cols 512 --nb 16 --tag stream_512
qb --rows 512 --cols 512 --nb 16 --flat --tag flat_512
qb --rows 256 --cols 65536 --nb 4 --tag stream_256x64k
