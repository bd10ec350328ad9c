"""Make the shared core package `antq` importable when this directory is used the way the
reference's antquant/ is used: sys.path.append(".../antquant"); from quant_model import *"""
import os
import sys

_PKG = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # .../ant-quantization_b200
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)
