"""Per-CTA timeline of antq_stream_kernel (needs a -DANTQS_TRACE build: ANTQ_LIB_SUFFIX=_trace).
Prints when CTAs start, when the first chunk is issued / published / ready, how the chunk-ready times
progress and when CTAs finish -- all in microseconds relative to the earliest CTA start."""
import ctypes, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ant-quantization_b200"))
import numpy as np, torch, antq
from antq import codebooks, _lib

N = int(os.environ.get("TRACE_N", "4096"))
dev = torch.device("cuda:0")
cb = antq.prepare_codebook(codebooks.ant_grid("flint", 4, True).to(dev))
g = torch.Generator().manual_seed(0)
NB = 6
xs = [(torch.randn(N, N, generator=g) * 0.02).half().to(dev) for _ in range(NB)]
als = [(x.float().abs().amax(1) * 0.9).contiguous() for x in xs]
outs = [torch.empty_like(x) for x in xs]
STRIDE, NCH = 8 + 4 * 64, 64
buf = torch.zeros(148 * STRIDE, dtype=torch.int64, device=dev)
_lib.lib.antq_debug_stream_trace.argtypes = [ctypes.c_void_p]
_lib.lib.antq_debug_stream_trace.restype = None
for i in range(NB):
    antq.fakequant(xs[i], als[i], cb, True, out=outs[i])
torch.cuda.synchronize()
_lib.lib.antq_debug_stream_trace(buf.data_ptr())
res = []
for i in range(NB):                       # back-to-back launches; the last one is analysed
    antq.fakequant(xs[i], als[i], cb, True, out=outs[i])
torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(148, STRIDE).astype(np.int64)
t0 = t[:, 0].min()
rel = lambda a: (a - t0) / 1e3
start, synced, end = rel(t[:, 0]), rel(t[:, 1]), rel(t[:, 2])
ch = t[:, 8:].reshape(148, NCH, 4)
valid = ch[:, :, 2] > 0
def stat(a): return "min %.2f  p50 %.2f  max %.2f" % (np.min(a), np.median(a), np.max(a))
print("CTA start      ", stat(start))
print("after sync     ", stat(synced))
print("CTA end        ", stat(end), "  span %.2f us" % (end.max()))
for slot, name in ((3, "b0 build enter"), (5, "b0 scale known"), (6, "b0 thresholds "), (4, "b0 build done ")):
    print(name, stat(rel(t[:, slot])))
for k in (0, 1, 2, 4, 8, 11, 12, 16, 23, 24, 32, 40, 48, 52, 53, 54, 55):
    v = valid[:, k]
    if v.sum() == 0: continue
    c = ch[v, k, :]
    print("chunk %2d (n=%3d): issue %s | tables %s | ready %s | done %s" % (
        k, v.sum(), *["%.2f/%.2f/%.2f" % (np.min(rel(c[:, i])), np.median(rel(c[:, i])), np.max(rel(c[:, i]))) for i in range(4)]))
dur = (ch[:, :, 3] - ch[:, :, 2])[valid] / 1e3
print("consumer busy per chunk (ready->done) us:", stat(dur))
nper = valid.sum(1)
print("chunks per CTA:", nper.min(), nper.max())
# slowest / fastest CTAs
order = np.argsort(end)
print("earliest-finishing CTAs:", [(int(i), round(float(end[i]), 2)) for i in order[:5]])
print("latest-finishing CTAs:  ", [(int(i), round(float(end[i]), 2)) for i in order[-5:]])
np.save(os.path.join(ROOT, "gpurun_out", "stream_trace.npy"), t)
