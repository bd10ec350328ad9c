"""world_size = 2 on CPU (gloo): the path's only exchange step -- the post-calibration sync of a
quantizer (alpha averaged over ranks, rank 0's grid broadcast; A/antquant/quant_modules.py:517-531) --
and the data-parallel sharding rule bench.py uses (independent tensors per rank, no collective)."""
import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json, types
import torch, torch.distributed as dist
sys.path.append(os.path.join(%(root)r, "ant-quantization_b200", %(flavor)r, "antquant"))
from quant_modules import TensorQuantizer
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=rank, world_size=world)
args = types.SimpleNamespace(w_up=150, a_up=150, w_low=75, a_low=75, percent=100, search=False, no_outlier=False)
q = TensorQuantizer(mode="ant-int-flint", bit=4, is_signed=True, is_enable=True, is_input=False, args=args)
# what each rank would hold after calibrating on ITS shard of the data
q.alpha.data = torch.full((6, 1), 1.0 + rank)                         # rank 0: 1.0, rank 1: 2.0
q.quant_grid.data = q.int_value() if rank == 0 else q.flint_value()   # ranks disagree on the numeric type
q._sync_after_calibration()
out = {"rank": rank, "alpha": q.alpha.data.flatten().tolist(), "grid": q.quant_grid.tolist(),
       "int": q.int_value().tolist()}
# bench.py sharding rule: rank r owns tensors seeded 1234 + r; nothing is exchanged
g = torch.Generator().manual_seed(1234 + rank)
mine = torch.randn(4, generator=g)
allv = [torch.zeros(4) for _ in range(world)]
dist.all_gather(allv, mine)
out["distinct_shards"] = not torch.equal(allv[0], allv[1])
print("OUT=" + json.dumps(out))
dist.barrier()
dist.destroy_process_group()
'''


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _launch(flavor):
    port = _free_port()
    code = WORKER % dict(root=ROOT, flavor=flavor, port=port)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2")
        procs.append(subprocess.Popen([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = []
    for p in procs:
        so, se = p.communicate(timeout=300)
        assert p.returncode == 0, se[-3000:]
        import json
        outs.append(json.loads([l for l in so.splitlines() if l.startswith("OUT=")][-1][4:]))
    return sorted(outs, key=lambda o: o["rank"])


def test_ant_calibration_sync_two_ranks():
    r0, r1 = _launch("ant")
    assert r0["alpha"] == [1.5] * 6 and r1["alpha"] == [1.5] * 6            # SUM / world
    assert r0["grid"] == r0["int"] and r1["grid"] == r0["int"]              # rank 0's type choice wins
    assert r0["distinct_shards"] and r1["distinct_shards"]


def test_olive_never_touches_the_process_group():
    r0, r1 = _launch("olive")
    assert r0["alpha"] == [1.0] * 6 and r1["alpha"] == [2.0] * 6            # the OliVe reference has no dist calls
    assert r0["grid"] != r1["grid"]


SHARD_WORKER = r'''
import os, sys, json
import torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
import bench
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=rank, world_size=world)
# batch-sharded step of bench.py: scatter one sample per rank from rank 0, run the layer, all_gather the outputs
S, H = 5, 8
x_all = torch.arange(world * S * H, dtype=torch.float32).view(world, S, H) if rank == 0 else None
x_loc = torch.empty(1, S, H)
y_all = torch.empty(world, S, H)
fn = lambda t: t * 2.0 + rank                                  # stands in for the quantized layer forward
bench.scatter_forward_gather(dist, fn, x_all, x_loc, y_all, rank)
expect = torch.arange(world * S * H, dtype=torch.float32).view(world, S, H) * 2.0
for r in range(world):
    expect[r] += r
out = {"rank": rank, "gathered_ok": bool(torch.equal(y_all, expect)),
       "rows": [bench.shard_rows(16384, world, r) for r in range(world)], "odd": [bench.shard_rows(10, 4, r) for r in range(4)]}
print("OUT=" + json.dumps(out))
dist.barrier()
dist.destroy_process_group()
'''


def test_bench_batch_shard_and_row_shard_two_ranks():
    """bench.py's N > 1 plumbing on CPU: scatter(inputs) -> per-rank forward -> all_gather(outputs), and the row ranges
    of the strong-scaling shard (disjoint, covering, ragged tails)."""
    import json
    port = _free_port()
    code = SHARD_WORKER % dict(root=ROOT, port=port)
    procs = [subprocess.Popen([sys.executable, "-c", code], env=dict(os.environ, RANK=str(r), WORLD_SIZE="2"),
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = []
    for p in procs:
        so, se = p.communicate(timeout=300)
        assert p.returncode == 0, se[-3000:]
        outs.append(json.loads([l for l in so.splitlines() if l.startswith("OUT=")][-1][4:]))
    for o in outs:
        assert o["gathered_ok"]
        assert o["rows"] == [[0, 8192], [8192, 16384]]
        assert o["odd"] == [[0, 3], [3, 6], [6, 9], [9, 10]]
