#!/usr/bin/env python
"""z-scores of a sampler against the real-emcee numbers in the reference's notebooks (tests/anchors.py).

    python tools/emcee_anchor_report.py --backend oracle --seeds 16 [--out profiles/r02_emcee_anchors_oracle.json]
    python tools/emcee_anchor_report.py --backend gpu --seeds 64   [--out profiles/r02_emcee_anchors_gpu.json]

backend `oracle`: oracle/bisip_oracle.c (CPU restatement, the same Philox stream as the CUDA kernel);
backend `gpu`:    bisip_ensemble_run through BatchInversion (one launch per anchor: the seeds are batch entries
                  that hold the same spectrum and differ in their Philox spectrum index and p0).
Developer / evidence tool: imports oracle/ and tests/, never imported by the package.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import anchors as A   # noqa: E402


def chains_oracle(anchor, pr, seeds):
    from oracle import oracle
    kw = dict(n_modes=anchor.get("n_modes", 1))
    if anchor["model"] == "decomp":
        kw.update(taus=pr["taus"], log_taus=pr["log_taus"], c_exp=anchor["c_exp"])
    prob = oracle.Problem(anchor["model"], pr["data"]["w"], pr["data"]["zn"], pr["data"]["zn_err"], pr["bounds"], **kw)
    out, acc = [], []
    for s in seeds:
        r = prob.run(A.p0_for(anchor, pr["bounds"], s), anchor["nsteps"], seed=0xE3CEE, spectrum=s)
        out.append(r["chain"])
        acc.append(r["accepted"].mean() / anchor["nsteps"])
    return out, acc


def chains_gpu(anchor, pr, seeds):
    from bisip_b200.batch import BatchInversion
    n = len(seeds)
    d = pr["data"]
    inv = BatchInversion(anchor["model"], d["w"], np.repeat(d["zn"][None], n, 0), np.repeat(d["zn_err"][None], n, 0),
                         nwalkers=anchor["nwalkers"], nsteps=anchor["nsteps"], bounds=pr["bounds"],
                         poly_deg=anchor.get("poly_deg", 5), c_exp=anchor.get("c_exp", 1.0), n_modes=anchor.get("n_modes", 1),
                         seed=0xE3CEE, spectrum_offset=seeds[0])
    assert list(seeds) == list(range(seeds[0], seeds[0] + n))
    p0 = np.stack([A.p0_for(anchor, pr["bounds"], s) for s in seeds])
    res = inv.fit(p0=p0, keep_chain=True)
    assert np.all(res["flags"] == 0)
    return list(res["chain"]), list(res["acceptance_fraction"])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", choices=["oracle", "gpu"], default="oracle")
    ap.add_argument("--seeds", type=int, default=16)
    ap.add_argument("--only", default="")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    seeds = list(range(args.seeds))
    report = {"backend": args.backend, "seeds": args.seeds, "anchors": {}}
    worst = worst_c = 0.0
    for name, anchor in A.load().items():
        if args.only and args.only not in name:
            continue
        pr = A.problem(anchor)
        t0 = time.time()
        chains, acc = (chains_oracle if args.backend == "oracle" else chains_gpu)(anchor, pr, seeds)
        runs = [A.run_stats(anchor, c, pr["log_taus"]) for c in chains]
        z = A.zscores(anchor, runs)                           # measured seed scatter only (strict)
        tau = A.autocorr_time(chains[:8], anchor)
        zc = A.zscores(anchor, runs, tau=tau)                 # max(measured scatter, sigma / sqrt(ESS)): what the tests assert
        entry = {"source": anchor["source"], "acceptance_fraction": float(np.mean(acc)), "seconds": time.time() - t0,
                 "autocorr_time_steps": tau,
                 "z": {k: np.round(v, 2).tolist() for k, v in z.items()},
                 "max_abs_z": float(max(np.max(np.abs(v)) for v in z.values())),
                 "z_combined": {k: np.round(v, 2).tolist() for k, v in zc.items()},
                 "max_abs_z_combined": float(max(np.max(np.abs(v)) for v in zc.values()))}
        for k in z:
            v = np.array([r[k] for r in runs])
            entry[f"{k}_median"] = np.median(v, 0).tolist()
            entry[f"{k}_notebook"] = anchor[k]
        report["anchors"][name] = entry
        worst = max(worst, entry["max_abs_z"])
        worst_c = max(worst_c, entry["max_abs_z_combined"])
        print(f"{name:28s} acc {entry['acceptance_fraction']:.3f} tau {tau:.0f}  max|z| {entry['max_abs_z']:.2f} "
              f"(combined {entry['max_abs_z_combined']:.2f})  "
              + "  ".join(f"{k}: {np.round(v, 2).tolist()}" for k, v in z.items()), flush=True)
    report["max_abs_z"] = worst
    report["max_abs_z_combined"] = worst_c
    if args.out:
        with open(args.out, "w") as f:
            json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()
