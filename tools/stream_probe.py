"""Streaming micro-benchmarks (antq_debug_stream): which access shape reaches the HBM roofline at 33.5 MB?"""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ant-quantization_b200"))
import torch
from antq import _lib
L = _lib.lib
L.antq_debug_stream.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong] + [ctypes.c_int] * 6 + [ctypes.c_void_p]
L.antq_debug_stream.restype = ctypes.c_int
dev = torch.device("cuda:0")
NB = 8
nbytes = 4096 * 4096 * 2
xs = [torch.randn(4096, 4096, device=dev).to(torch.float16) for _ in range(NB)]
outs = [torch.empty_like(x) for x in xs]

def run(mode, threads, unroll, span=512, grid=148 * 8, ro=0, reps=20):
    def step():
        for i in range(NB):
            rc = L.antq_debug_stream(xs[i].data_ptr(), outs[i].data_ptr(), nbytes, mode, threads, unroll, span, grid, ro,
                                     torch.cuda.current_stream().cuda_stream)
            assert rc == 0, rc
    step(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): step()
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * NB)
    moved = nbytes * (1 if ro else 2)
    print(json.dumps({"mode": mode, "threads": threads, "unroll": unroll, "span_vecs": span, "grid": grid if mode == 3 else None,
                      "read_only": ro, "us": round(us, 2), "GBps_moved": round(moved / us / 1e3, 1)}))

for ro in (0, 1):
    run(0, 256, 4, ro=ro)
    run(2, 32, 4, span=512, ro=ro)
    run(2, 128, 4, span=512, ro=ro)
    for mode in (4, 5):
        for t, u, span in ((32, 4, 512), (128, 4, 512), (32, 4, 128), (128, 4, 128), (64, 4, 256), (32, 8, 512), (256, 4, 128)):
            run(mode, t, u, span=span, ro=ro)
