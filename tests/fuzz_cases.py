"""Seeded random scenario generators shared by tests/test_gpu_fuzz.py and the ad-hoc tests/try_*_fuzz.py drivers."""
import importlib

import numpy as np

synth = importlib.import_module("open-symuvia_b200.synth")


def corridor_case(seed: int, n_steps: int = 250):
    r = np.random.default_rng(seed)
    kw = dict(n_links=int(r.integers(3, 9)), length=float(r.choice([250.0, 400.0, 600.0, 900.0])), n_lanes=int(r.integers(2, 5)),
              demand=float(r.uniform(0.5, 1.8)), offramp_every=int(r.integers(1, 4)), offramp_share=float(r.uniform(0.05, 0.4)),
              preseed_density=float(r.choice([0.0, 0.03, 0.06, 0.09])), shuffle=bool(r.integers(0, 2)), seed=seed,
              dst_chgtvoie=float(r.choice([100.0, 200.0, 350.0])), dst_force=float(r.choice([0.0, 60.0, 120.0])), phi_force=float(r.choice([1.0, 2.5])),
              accbornee=bool(r.integers(0, 2)), relax=float(r.choice([1.0, 0.6])), n_steps=n_steps)
    if r.random() < 0.4:
        kw["signal"] = (int(r.integers(15, 40)), 3, int(r.integers(10, 30)))
    if r.random() < 0.3:
        kw["exponential"] = True
    return kw


def grid_case(seed: int, n_steps: int = 200):
    r = np.random.default_rng(1000 + seed)
    kw = dict(n=int(r.integers(2, 7)), ny=int(r.integers(2, 6)), block=float(r.choice([120.0, 200.0, 300.0])), n_lanes=int(r.integers(1, 3)),
              vit_reg=float(r.choice([8.33, 13.89, 19.44])), green=float(r.integers(15, 45)), amber=3.0, seed=seed,
              demand_per_lane=float(r.choice([0.0, 0.15, 0.3, 0.45])), demand_duration=float(r.choice([150.0, 1e12])),
              preseed_per_lane=float(r.choice([0.0, 4.0, 9.0, 13.0])), shuffle=bool(r.integers(0, 2)), offsets=bool(r.integers(0, 2)), n_steps=n_steps)
    if kw["demand_per_lane"] == 0.0 and kw["preseed_per_lane"] == 0.0:
        kw["demand_per_lane"] = 0.25
    if r.random() < 0.35:
        kw["classes"] = [synth.VL, synth.PL]
    return kw
