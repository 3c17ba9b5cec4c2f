#!/usr/bin/env python
"""Time bisip_ensemble_run alone for a given shape (developer tool, not the bench).
   python tools/kernel_time.py --model decomp --B 296 --W 256 --T 500 --N 64 --S 64 [--reps 3]"""
import argparse, os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bisip_b200 import _lib, engine, synthetic
from bisip_b200.batch import BatchInversion

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="decomp"); ap.add_argument("--B", type=int, default=296)
ap.add_argument("--W", type=int, default=256); ap.add_argument("--T", type=int, default=500)
ap.add_argument("--N", type=int, default=64); ap.add_argument("--S", type=int, default=64)
ap.add_argument("--P", type=int, default=4); ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--K", type=int, default=1)
ap.add_argument("--precision", default="fp64"); ap.add_argument("--c_exp", type=float, default=1.0)
a = ap.parse_args()
dev = torch.device("cuda:0")
_, w = synthetic.frequencies(a.N)
kw = dict(poly_deg=a.P, n_tau=a.S, c_exp=a.c_exp, precision=a.precision) if a.model == "decomp" else (dict(n_modes=a.K) if a.model == "colecole" else {})
probe = BatchInversion(a.model, w, np.zeros((1, 2, a.N)), np.ones((1, 2, a.N)), device=dev, **kw)
fwd = lambda th, ww: engine.forward(probe._spec(), _lib.dev_f64(th[:, None, :], dev), _lib.dev_f64(ww, dev))[:, 0].cpu().numpy()
syn = synthetic.make(a.model, 0, a.B, fwd, N=a.N, poly_deg=a.P, n_modes=a.K, n_tau=a.S)
inv = BatchInversion(a.model, w, syn["zn"], syn["zn_err"], nwalkers=a.W, nsteps=a.T, seed=1, device=dev, **kw)
p0 = _lib.dev_f64(inv.draw_p0(0, a.B), dev)
y, ye = _lib.dev_f64(syn["zn"], dev), _lib.dev_f64(syn["zn_err"], dev)
wd, bd = _lib.dev_f64(w, dev), _lib.dev_f64(inv.param_bounds, dev)
spec = inv._spec()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e30
for r in range(a.reps + 1):
    c = p0.clone()
    e0.record()
    res = engine.ensemble_run(spec, c, wd, y, ye, bd, nsteps=a.T, seed=1, discard=a.T // 2, thin=10, store_logp=False)
    e1.record(); torch.cuda.synchronize()
    if r: best = min(best, e0.elapsed_time(e1))
evals = a.B * a.W * (a.T + 1)
flop = 2 * (a.P + 1) * a.S + 2 * a.S * 2 * a.N + 12 * a.N
print(json.dumps({"model": a.model, "B": a.B, "W": a.W, "T": a.T, "N": a.N, "S": a.S, "ms": best,
                  "us_per_step": 1e3 * best / a.T, "evals_per_s": evals / best * 1e3,
                  "tflops_alg": evals * flop / best / 1e9 if a.model == "decomp" else None,
                  "acc": float(res["accepted"].double().mean() / a.T), "flags": int(res["flags"].sum())}))
