"""End-to-end (host pinned buffers -> H2D -> kernel -> D2H) throughput of the host pipeline vs chunk size / stages."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ant-quantization_b200"))
import torch, antq
from antq import codebooks
N, NB = 4096, 8
g = torch.Generator().manual_seed(0)
xs = [(torch.randn(N, N, generator=g) * 0.02).half().pin_memory() for _ in range(NB)]
outs = [torch.empty_like(x).pin_memory() for x in xs]
als = [(x.float().abs().amax(1) * 0.9).contiguous() for x in xs]
grid = codebooks.ant_grid("flint", 4, True)
for sync in (True, False):
    for chunk_mb, stages in [(8, 3), (8, 4), (8, 6), (4, 4), (4, 8), (16, 3), (16, 4), (32, 3), (2, 8)]:
        hp = antq.HostPipeline(device=0, chunk_bytes=int(chunk_mb * (1 << 20)), n_stages=stages)
        def step():
            for i in range(NB): hp.fakequant(xs[i], outs[i], als[i], grid, per_row=True, sync=sync)
            hp.synchronize()
        step(); torch.cuda.synchronize()
        t0 = time.perf_counter(); reps = 5
        for _ in range(reps): step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(json.dumps({"sync_per_tensor": sync, "chunk_MiB": chunk_mb, "stages": stages,
                          "GBps_algorithmic": round(reps * NB * N * N * 4 / dt / 1e9, 1),
                          "ms_per_tensor": round(dt / reps / NB * 1e3, 3)}))
        hp.close()
