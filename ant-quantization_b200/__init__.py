"""ant-quantization_b200: B200-native fake-quant forward for ANT / OliVe.

Layout
  csrc/            hand-written sm_100a CUDA + the C ABI (include/antq.h) -> libantq.so
  antq/            ctypes binding and torch-facing ops (device pointers + streams only)
  ant/antquant/    drop-in mirror of ant_quantization/antquant   (quant_modules, quant_model, ...)
  olive/antquant/  drop-in mirror of olive_quantization/antquant

The directory name contains a hyphen, as the reference's own trees do; use it the
way the reference is used -- `sys.path.append(".../ant-quantization_b200/ant/antquant")`
then `from quant_model import *` -- or `importlib.import_module("ant-quantization_b200")`.
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)

ANT_PATH = os.path.join(_HERE, "ant", "antquant")
OLIVE_PATH = os.path.join(_HERE, "olive", "antquant")
