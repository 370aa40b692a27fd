"""shared helpers for the tests: golden-fixture loading and model construction"""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        return json.load(f)


def unhex(v):
    if isinstance(v, str):
        return float.fromhex(v)
    return np.array([float.fromhex(s) for s in v], dtype=np.float64)


def scalar(v):
    a = unhex(v)
    return float(a[0]) if isinstance(a, np.ndarray) else a


def qm_model(po, c):
    kind, ip, dp = c["kind"], c["ip"], c["dp"]
    if kind == po.HO:
        return po.ho(ip[0], dp[0], dp[1], dp[2])
    if kind == po.QUARTIC:
        return po.quartic(ip[0], *dp)
    return po.rotor(ip[0], dp[0], dp[1])


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = max(np.max(np.abs(b)), 1e-300)
    return float(np.max(np.abs(a - b)) / scale)
