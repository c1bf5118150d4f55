"""Shared inputs for the risk-path tests (oracle and GPU)."""
import json
import os

import numpy as np

from oracle import cref

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def c1_inputs():
    """BASELINE config C1: the shipped sample.  E is derived per DESIGN.md 3.0: E[p][f] = sum over the
    portfolio's positions (file order, fma) of weight * B[instrument][f], with B the 10000 x 10 loading
    table from the documented generator (seed 0x5EED0001, tensor 3, scale 1e-2); S = the macro.json rows."""
    d = json.load(open(os.path.join(GOLD, "c1_sample.json")))
    F = 10
    B = cref.gen_normal(10000, F, 0x5EED0001, 3, 1e-2, 0.0)
    inst, w, off = c1_positions(d)
    E = cref.build_exposures(inst, w, off, B)  # fma chain in file order; the GPU ingest kernel must match it bit for bit
    S = np.array(d["macro"], np.float64)
    return d, E, S


def c1_positions(d):
    inst = np.concatenate([np.asarray(p["instruments"], np.int32) for p in d["portfolios"]])
    w = np.concatenate([np.asarray(p["weights"], np.float64) for p in d["portfolios"]])
    off = np.zeros(len(d["portfolios"]) + 1, np.int64)
    np.cumsum([len(p["instruments"]) for p in d["portfolios"]], out=off[1:])
    return inst, w, off


def synthetic(P, N, F, seed, scale_e=100.0, drift=0.05):
    E = cref.gen_normal(P, F, seed, 1, scale_e, 0.0)
    S = cref.gen_normal(N, F, seed, 2, 1.0, drift)
    return E, S


def with_ties(P, N, F, seed):
    """Integer-valued inputs so that many trials share exactly the same loss (tie-break by trial index)."""
    rng = np.random.default_rng(seed)
    E = rng.integers(-3, 4, size=(P, F)).astype(np.float64)
    S = rng.integers(-2, 3, size=(N, F)).astype(np.float64)
    return E, S
