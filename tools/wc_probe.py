"""H2D bandwidth from default pinned host memory vs write-combined pinned host memory (cudaHostAllocWriteCombined)."""
import ctypes, glob, os, sys, json
import torch
torch.cuda.init()
paths = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*"))
rt = ctypes.CDLL(paths[0] if paths else "libcudart.so")
rt.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
n = 256 << 20
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
res = {}
for name, flags in (("default", 0), ("write_combined", 4), ("default_again", 0)):
    p = ctypes.c_void_p()
    assert rt.cudaHostAlloc(ctypes.byref(p), n, flags) == 0
    ctypes.memset(p, 1, n)
    st = torch.cuda.current_stream().cuda_stream
    for kind, (dst, src, k) in (("h2d", (dev.data_ptr(), p.value, 1)), ("d2h", (p.value, dev.data_ptr(), 2))):
        for _ in range(2):
            rt.cudaMemcpyAsync(dst, src, n, k, st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            rt.cudaMemcpyAsync(dst, src, n, k, st)
        e1.record(); torch.cuda.synchronize()
        res["%s_%s_GBps" % (name, kind)] = round(10 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9, 1)
    rt.cudaFreeHost(p)
print(json.dumps(res))
# the same through torch's pinned allocator (what bench.py's e2e leg and probe use)
hb = torch.empty(n, dtype=torch.uint8).pin_memory()
hb2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
out = {}
for name, h in (("torch_pin_memory()", hb), ("torch_empty(pin_memory=True)", hb2)):
    for kind, fn in (("h2d", lambda: dev.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(dev, non_blocking=True))):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record(); torch.cuda.synchronize()
        out["%s %s" % (name, kind)] = round(10 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9, 1)
print(json.dumps(out))
# both directions at once, on two streams (the full-duplex rate the e2e leg can hope for)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
for chunk in (n, 16 << 20):
    def both():
        for off in range(0, n, chunk):
            with torch.cuda.stream(s1):
                d_in[off:off + chunk].copy_(h_in[off:off + chunk], non_blocking=True)
            with torch.cuda.stream(s2):
                h_out[off:off + chunk].copy_(d_out[off:off + chunk], non_blocking=True)
    both(); torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    for _ in range(10):
        both()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps({"both_directions_chunk_MiB": chunk >> 20, "GBps_each_way": round(10 * n / dt / 1e9, 1), "GBps_total": round(20 * n / dt / 1e9, 1)}))
