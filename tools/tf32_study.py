#!/usr/bin/env python
"""FP64 DMMA vs TF32 vs 3xTF32 (tcgen05 and mma.sync kernels) tolerance study for the decomposition (BASELINE config 4).

    python tools/tf32_study.py --out profiles/r01_tf32_study [--spectra 64] [--steps 1500]

For each (n_tau, c_exp) case, on synthetic 64-frequency spectra:
  * forward error  max|Z - Z_fp64| / max|Z|           at theta_true and at random prior draws
  * log-prob error |lp - lp_fp64| (absolute) and relative, at theta within the posterior bulk
  * sampler: posterior median / 2.5 / 97.5 percentiles and acceptance vs the FP64 run of the same
    Philox stream, in units of the FP64 posterior standard deviation
  * throughput of the ensemble kernel in each mode
Reference for all comparisons is this library's own FP64 path, which the parity tests tie to the
reference implementation at 1e-12.
"""
import argparse, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bisip_b200 import _lib, engine, synthetic
from bisip_b200.batch import BatchInversion

ap = argparse.ArgumentParser()
ap.add_argument("--out", default="profiles/tf32_study")
ap.add_argument("--spectra", type=int, default=296)
ap.add_argument("--walkers", type=int, default=256)
ap.add_argument("--steps", type=int, default=1500)
a = ap.parse_args()
dev = torch.device("cuda:0")
N, P = 64, 4
_, w = synthetic.frequencies(N)
cases = [dict(n_tau=64, c_exp=1.0), dict(n_tau=128, c_exp=1.0), dict(n_tau=256, c_exp=1.0), dict(n_tau=256, c_exp=0.5)]
MODES = ("tf32", "3xtf32", "tf32-mma", "3xtf32-mma")
report = {"spectra": a.spectra, "walkers": a.walkers, "steps": a.steps, "cases": []}


def evs(e0, e1):
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


for case in cases:
    B = a.spectra
    mk = lambda prec: BatchInversion('decomp', w, syn["zn"], syn["zn_err"], nwalkers=a.walkers, nsteps=a.steps,
                                     poly_deg=P, seed=5, precision=prec, device=dev, **case)
    probe = BatchInversion('decomp', w, np.zeros((1, 2, N)), np.ones((1, 2, N)), poly_deg=P, device=dev, **case)
    fwd = lambda th, ww: engine.forward(probe._spec(), _lib.dev_f64(th[:, None, :], dev), _lib.dev_f64(ww, dev))[:, 0].cpu().numpy()
    syn = synthetic.make('decomp', 0, B, fwd, N=N, poly_deg=P, n_tau=case["n_tau"])
    invs = {p: mk(p) for p in ("fp64",) + MODES}
    rng = np.random.default_rng(0)
    lo, hi = invs["fp64"].param_bounds
    th_true = syn["theta_true"][:, None, :]                                          # (B,1,ndim)
    th_prior = rng.uniform(lo, hi, (B, 32, lo.shape[0]))
    wd = _lib.dev_f64(w, dev)
    y, ye, bd = _lib.dev_f64(syn["zn"], dev), _lib.dev_f64(syn["zn_err"], dev), _lib.dev_f64(invs["fp64"].param_bounds, dev)
    entry = dict(case=case, forward={}, logprob={}, sampler={}, throughput={},
                 kernel={p: engine.decomp_kernel_kind(invs[p]._spec(), N, a.walkers) for p in invs})
    # ---- FP64 sampler first: its posterior defines the "bulk" thetas for the log-prob comparison
    runs = {}
    for prec, inv in invs.items():
        p0 = _lib.dev_f64(invs["fp64"].draw_p0(0, B), dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ms = 1e30
        for rep in range(4 if not report["cases"] and prec == "fp64" else 2):   # best of the runs; the very first mode also ramps the clocks
            c0 = p0.clone()
            e0.record()
            res = engine.ensemble_run(inv._spec(), c0, wd, y, ye, bd, nsteps=a.steps, seed=5, discard=a.steps // 2, thin=5)
            e1.record()
            ms = min(ms, evs(e0, e1))
        ch = res["chain"].reshape(B, -1, lo.shape[0])
        st = engine.column_stats(ch, p=[2.5, 50, 97.5], want_mean=True, want_std=True)
        runs[prec] = dict(pct=st["pct"].cpu().numpy(), std=st["std"].cpu().numpy(), mean=st["mean"].cpu().numpy(),
                          acc=float(res["accepted"].double().mean() / a.steps), chain=ch)
        entry["throughput"][prec] = dict(ms=ms, evals_per_s=B * a.walkers * a.steps / ms * 1e3)
    sd = runs["fp64"]["std"]
    for prec in MODES:
        d = np.abs(runs[prec]["pct"] - runs["fp64"]["pct"]) / sd[:, None, :]
        entry["sampler"][prec] = dict(acc=runs[prec]["acc"], acc_fp64=runs["fp64"]["acc"],
                                      pct_shift_in_sd_median=float(np.median(d)), pct_shift_in_sd_max=float(d.max()),
                                      std_ratio_median=float(np.median(runs[prec]["std"] / sd)))
    # thetas in the posterior bulk: 64 draws per spectrum from the FP64 kept chain
    idx = torch.from_numpy(rng.integers(0, runs["fp64"]["chain"].shape[1], 64)).to(dev)
    th_bulk = runs["fp64"]["chain"][:, idx, :].contiguous()
    sets = {"theta_true": _lib.dev_f64(th_true, dev), "prior": _lib.dev_f64(th_prior, dev), "posterior_bulk": th_bulk}
    ref = {}
    for prec, inv in invs.items():
        spec = inv._spec()
        for name, th in sets.items():
            Z = engine.forward(spec, th, wd).cpu().numpy()
            lp = engine.log_probability(spec, th, wd, y, ye, bd).cpu().numpy()
            if prec == "fp64":
                ref[name] = (Z, lp)
                continue
            Z0, lp0 = ref[name]
            zerr = np.max(np.abs(Z - Z0), axis=(2, 3)) / np.max(np.abs(Z0), axis=(2, 3))
            fin = np.isfinite(lp0)
            dl = np.abs(lp - lp0)[fin]
            entry["forward"].setdefault(prec, {})[name] = dict(normwise_max=float(zerr.max()), normwise_median=float(np.median(zerr)))
            entry["logprob"].setdefault(prec, {})[name] = dict(abs_max=float(dl.max()), abs_median=float(np.median(dl)),
                                                               rel_max=float((dl / np.maximum(1, np.abs(lp0[fin]))).max()))
    report["cases"].append(entry)
    print(json.dumps(entry)[:1500])

os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
json.dump(report, open(a.out + ".json", "w"), indent=1)
with open(a.out + ".md", "w") as f:
    f.write("# FP64 DMMA vs TF32 vs 3xTF32 — decomposition tolerance study (tcgen05 and mma.sync kernels)\n\n")
    f.write(f"{a.spectra} synthetic 64-frequency spectra, poly_deg 4, {a.walkers} walkers x {a.steps} steps (second half kept, thin 5). "
            "`tf32` / `3xtf32` run on the tcgen05 kernel (csrc/decomp_umma.cuh: both stages on the 5th-generation tensor cores, operands and FP32 accumulators in tensor memory) whenever the problem fits it, `*-mma` always on the mma.sync tile kernel (csrc/decomp_tf32.cuh: FP64 stage 1, TF32 stage 2); the `kernel` column says which one ran. "
            "Errors are against this library's FP64 path. theta_true is the un-normalised truth (a forward-only probe); the \"posterior bulk\" thetas are draws from the kept FP64 chain, i.e. where the sampler actually evaluates. evals/s is the ensemble kernel alone (CUDA events), one wave of spectra.\n\n")
    for e in report["cases"]:
        f.write(f"## n_tau = {e['case']['n_tau']}, c_exp = {e['case']['c_exp']}\n\n")
        f.write("| mode | kernel | forward err (theta_true) | forward err (prior) | abs lp err, posterior bulk (median / max) | rel lp err, prior (max) | "
                "percentile shift / posterior sd (median / max) | sd ratio | acceptance (mode / fp64) | evals/s | vs fp64 |\n|---|---|---|---|---|---|---|---|---|---|---|\n")
        for prec in MODES:
            fw, lp, sm = e["forward"][prec], e["logprob"][prec], e["sampler"][prec]
            f.write(f"| {prec} | {e['kernel'][prec]} | {fw['theta_true']['normwise_max']:.2e} | {fw['prior']['normwise_max']:.2e} | "
                    f"{lp['posterior_bulk']['abs_median']:.2e} / {lp['posterior_bulk']['abs_max']:.2e} | "
                    f"{lp['prior']['rel_max']:.2e} | {sm['pct_shift_in_sd_median']:.3f} / {sm['pct_shift_in_sd_max']:.3f} | "
                    f"{sm['std_ratio_median']:.3f} | {sm['acc']:.3f} / {sm['acc_fp64']:.3f} | {e['throughput'][prec]['evals_per_s']:.3e} | {e['throughput'][prec]['evals_per_s'] / e['throughput']['fp64']['evals_per_s']:.2f}x |\n")
        f.write(f"| fp64 | {e['kernel']['fp64']} | 0 | 0 | 0 | 0 | 0 | 1 | {e['sampler']['tf32']['acc_fp64']:.3f} | {e['throughput']['fp64']['evals_per_s']:.3e} | 1.00x |\n\n")
print("written", a.out)
