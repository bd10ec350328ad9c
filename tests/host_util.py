"""Run a snippet with one of the two antquant mirrors on sys.path (their module names collide,
exactly like the reference's two trees, so each flavour gets its own process)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATHS = {"ant": os.path.join(ROOT, "ant-quantization_b200", "ant", "antquant"),
         "olive": os.path.join(ROOT, "ant-quantization_b200", "olive", "antquant")}

PRELUDE = r'''
import sys, os, json, types
import numpy as np, torch, torch.nn as nn
sys.path.append(%(path)r)
from quant_model import *
from quant_utils import *
def mkargs(mode, **kw):
    d = dict(mode=mode, wbit=4, abit=4, w_up=150, a_up=150, w_low=75, a_low=75, percent=100, search=False, no_outlier=False)
    d.update(kw)
    return types.SimpleNamespace(**d)
class Net(nn.Module):
    def __init__(self):
        super().__init__()
        self.features = nn.Sequential(nn.Conv2d(3, 8, 3, padding=1), nn.ReLU(), nn.Conv2d(8, 8, 3, padding=1, bias=False), nn.ReLU())
        self.pool = nn.AdaptiveAvgPool2d(2)
        self.blocks = nn.ModuleList([nn.Linear(32, 32), nn.Linear(32, 32)])
        self.head = nn.Linear(32, 10)
    def forward(self, x):
        x = self.pool(self.features(x)).flatten(1)
        for b in self.blocks:
            x = torch.relu(b(x))
        return self.head(x)
RESULT = {}
'''


def run(flavor, body, timeout=600):
    code = PRELUDE % dict(path=PATHS[flavor]) + body + "\nprint('RESULT=' + json.dumps(RESULT))\n"
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    if r.returncode != 0:
        raise AssertionError("subprocess failed:\n" + r.stdout[-2000:] + "\n" + r.stderr[-4000:])
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT=")][-1]
    return json.loads(line[len("RESULT="):]), r.stdout
