"""Shared builders for the parity tests: BASELINE.json's five configurations and oracle runs."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle.oracle import OracleSpace, oracle_from_config  # noqa: E402
from transrot_b200 import Config, WORKLOADS, system_from_config, workload_config  # noqa: E402,F401


def oracle_start(dbase, cfg: Config, seed: int, pair_table=None) -> OracleSpace:
    return oracle_from_config(dbase, cfg, seed, pair_table)


def rel_err(a, b, scale=None):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    s = np.maximum(np.abs(a), np.abs(b)) if scale is None else np.asarray(scale, dtype=np.float64)
    s = np.where(s == 0, 1.0, s)
    return np.abs(a - b) / s


def compare_traces(gpu, orc, abs_scale=None):
    """Decision parity + energy parity between a device trace and an oracle trace.
    Returns (first mismatching step or -1, err_rel, err_cond) over the steps before the first mismatch:
      err_rel  = max |gpu - oracle| / max(|eStart|, |eEnd|)  over eStart, eEnd, dE
      err_cond = the same differences over abs_scale = sum of |pair term| of the move (the conditioning
                 of the sum: what "every pair term is right to 1e-12" can promise about a cancelling sum)."""
    n = min(len(gpu), len(orc))
    gpu, orc = gpu[:n], orc[:n]
    same = (
        (gpu["mol"] == orc["mol"]) & (gpu["kind"] == orc["kind"]) & (gpu["magwalk"] == orc["magwalk"])
        & (gpu["bounds"] == orc["bounds"]) & (gpu["accepted"] == orc["accepted"]) & (gpu["axis"] == orc["axis"])
        & (gpu["drew_rand"] == orc["drew_rand"])
    )
    bad = np.nonzero(~same)[0]
    first = int(bad[0]) if len(bad) else -1
    upto = n if first < 0 else first
    if not upto:
        return first, 0.0, 0.0
    diff = np.maximum.reduce([
        np.abs(gpu["e_start"][:upto] - orc["e_start"][:upto]),
        np.abs(gpu["e_end"][:upto] - orc["e_end"][:upto]),
        np.abs(gpu["dE"][:upto] - orc["dE"][:upto]),
    ])
    ev = orc["bounds"][:upto] == 0
    scale = np.maximum(np.abs(orc["e_start"][:upto]), np.abs(orc["e_end"][:upto]))
    err_rel = float(np.max(diff[ev] / scale[ev])) if ev.any() else 0.0
    err_cond = 0.0
    if abs_scale is not None and ev.any():
        err_cond = float(np.max(diff[ev] / abs_scale[:upto][ev]))
    return first, err_rel, err_cond
