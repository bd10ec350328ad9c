cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out/r2p_bench.jsonl; : > $O
qb() { timeout 120 python tools/quick_bench.py "$@" 2>&1 | tail -1 | tee -a $O; }
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -6
qb --tag new
qb --kind int --tag int4
ANTQ_DEBUG=32 qb --kind int --tag int4_asym
qb --kind int --bit 3 --tag int3
qb --kind int --bit 5 --tag int5
qb --kind int --bit 6 --tag int6
qb --kind int --dtype f32 --tag int4_f32
qb --kind pot --tag pot4
qb --kind flint --unsigned --tag flint4u
qb --olive --tag olive
python - <<'PY'
import torch, time
n = 256 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(f, reps=5):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
a = t(lambda: d.copy_(h, non_blocking=True)); b = t(lambda: h2.copy_(d2, non_blocking=True))
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
c = t(both)
print("PCIe GB/s: H2D %.1f  D2H %.1f  concurrent %.1f + %.1f" % (n / a / 1e9, n / b / 1e9, n / c / 1e9, n / c / 1e9))
PY
