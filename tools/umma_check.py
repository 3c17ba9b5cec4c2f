#!/usr/bin/env python
"""Developer check of the tcgen05 decomposition path against the FP64 DMMA path and the mma.sync TF32 tiles:
forward error (norm-wise), log-probability error at posterior-like thetas, then sampler throughput.
   python tools/umma_check.py [--quick]"""
import argparse, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bisip_b200 import _lib, engine, synthetic
from bisip_b200.batch import BatchInversion

ap = argparse.ArgumentParser()
ap.add_argument("--quick", action="store_true")
ap.add_argument("--B", type=int, default=296)
ap.add_argument("--T", type=int, default=300)
a = ap.parse_args()
dev = torch.device("cuda:0")
P = 4


def build(N, S, B, W, prec, c_exp=1.0, syn=None):
    _, w = synthetic.frequencies(N)
    kw = dict(poly_deg=P, n_tau=S, c_exp=c_exp)
    if syn is None:
        probe = BatchInversion("decomp", w, np.zeros((1, 2, N)), np.ones((1, 2, N)), device=dev, precision="fp64", **kw)
        fwd = lambda th, ww: engine.forward(probe._spec(), _lib.dev_f64(th[:, None, :], dev), _lib.dev_f64(ww, dev))[:, 0].cpu().numpy()
        syn = synthetic.make("decomp", 0, B, fwd, N=N, poly_deg=P, n_tau=S)
    inv = BatchInversion("decomp", w, syn["zn"], syn["zn_err"], nwalkers=W, nsteps=a.T, seed=1, device=dev, precision=prec, **kw)
    return inv, syn, w


def accuracy(N, S, c_exp=1.0):
    B, n = 4, 300
    inv64, syn, w = build(N, S, B, 64, "fp64", c_exp)
    rng = np.random.default_rng(0)
    truth = syn["theta_true"].copy()
    truth[:, 0] /= syn["norm_factor"]
    th = truth[:, None, :] * (1 + 1e-3 * rng.standard_normal((B, n, P + 2)))              # near the truth
    thd = _lib.dev_f64(th, dev)
    wd, bd = _lib.dev_f64(w, dev), _lib.dev_f64(inv64.param_bounds, dev)
    y, ye = _lib.dev_f64(syn["zn"], dev), _lib.dev_f64(syn["zn_err"], dev)
    Z64 = engine.forward(inv64._spec(), thd, wd)
    lp64 = engine.log_probability(inv64._spec(), thd, wd, y, ye, bd)
    out = {}
    for prec in ("tf32", "tf32-mma", "3xtf32", "3xtf32-mma"):
        try:
            inv, _, _ = build(N, S, B, 64, prec, c_exp, syn)
            Z = engine.forward(inv._spec(), thd, wd)
            lp = engine.log_probability(inv._spec(), thd, wd, y, ye, bd)
            torch.cuda.synchronize()
            zerr = float(((Z - Z64).abs().amax((2, 3)) / Z64.abs().amax((2, 3))).max())
            out[prec] = {"Z_err": zerr, "lp_abs": float((lp - lp64).abs().max()),
                         "lp_rel": float(((lp - lp64).abs() / lp64.abs().clamp(min=1)).max())}
        except _lib.BisipError as e:
            out[prec] = {"error": str(e)[:80]}
    print(json.dumps({"accuracy": {"N": N, "S": S, "c_exp": c_exp}, **out}), flush=True)


def speed(N, S, B, W, precs):
    syn = None
    for prec in precs:
        try:
            inv, syn, w = build(N, S, B, W, prec, 1.0, syn)
        except _lib.BisipError as e:
            print(json.dumps({"speed": prec, "error": str(e)[:80]})); continue
        p0 = _lib.dev_f64(inv.draw_p0(0, B), dev)
        y, ye = _lib.dev_f64(syn["zn"], dev), _lib.dev_f64(syn["zn_err"], dev)
        wd, bd = _lib.dev_f64(w, dev), _lib.dev_f64(inv.param_bounds, dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e30
        try:
            for r in range(3):
                c = p0.clone()
                e0.record()
                res = engine.ensemble_run(inv._spec(), c, wd, y, ye, bd, nsteps=a.T, seed=1, discard=a.T // 2, thin=10, store_logp=False)
                e1.record(); torch.cuda.synchronize()
                if r: best = min(best, e0.elapsed_time(e1))
        except _lib.BisipError as e:
            print(json.dumps({"speed": prec, "error": str(e)[:80]})); continue
        evals = B * W * (a.T + 1)
        st = engine.column_stats(res["chain"].reshape(B, -1, P + 2), p=[50.0], want_mean=True, want_std=True)
        print(json.dumps({"speed": prec, "N": N, "S": S, "B": B, "W": W, "ms": best, "evals_per_s": evals / best * 1e3,
                          "acc": float(res["accepted"].double().mean() / a.T), "flags": int(res["flags"].sum()),
                          "median_r0": float(st["pct"][:, 0, 0].mean()), "std_r0": float(st["std"][:, 0].mean())}), flush=True)


accuracy(20, 40)
accuracy(64, 64)
if not a.quick:
    accuracy(64, 64, 0.5)
    accuracy(33, 50)
    accuracy(64, 128)
    accuracy(64, 256)
speed(64, 64, a.B, 256, ("fp64", "3xtf32-mma", "3xtf32", "tf32"))
if not a.quick:
    speed(64, 64, a.B, 64, ("fp64", "3xtf32"))
    speed(64, 128, a.B, 256, ("fp64", "3xtf32-mma", "3xtf32", "tf32"))
    speed(64, 256, a.B, 256, ("fp64", "tf32-mma", "tf32"))
