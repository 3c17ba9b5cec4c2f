"""z-tests of a stretch-move sampler against the REAL-emcee numbers stored in the reference's notebooks
(tests/golden/emcee_anchors.json, extracted by tests/golden/make_emcee_anchors.py).

emcee is third-party, unpinned and not installable here (reference requirements.txt:2), so the sampler parity cannot
be pinned on a shared random stream; what the reference tree does hold is the OUTPUT of real emcee runs for four
tutorial configurations.  Each stored number is one draw of "statistic of one emcee run"; a sampler with the same
transition kernel produces, over independent seeds, draws of the same distribution.  For every stored statistic s_nb

    z = (s_nb - centre) / sqrt( scatter^2 * (1 + 1/n_eff) + rounding^2 / 12 )

with centre / scatter the median and 1.4826 MAD of the statistic over the sampler's seeds (robust: a run in which a
walker is still stuck far from the mode after burn-in — it happens in a few percent of the runs, for emcee as well —
moves a tail percentile by many scatter units), n_eff = n_seeds / (pi/2) (efficiency of the median), and `rounding`
the last printed digit of the notebook value.  `scatter` is the Monte-Carlo error of a single run, measured over the
seeds; the textbook estimate sigma / sqrt(ESS) (`ess_standard_errors`) agrees with it for means and is ~2x larger for
the 2.5 / 97.5 percentiles.  `zscores(..., tau=...)` uses the larger of the two (the "combined standard error" the
tests assert |z| <= 3 on); without `tau` it is the measured scatter alone (reported next to it, stricter).
"""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ANCHOR_FILE = os.path.join(HERE, "golden", "emcee_anchors.json")


def load():
    with open(ANCHOR_FILE) as f:
        return {a["name"]: a for a in json.load(f)["anchors"]}


def problem(anchor):
    """Everything a run needs: data dict, (2, ndim) bounds, decomposition tables, model kwargs."""
    from bisip_b200.batch import default_bounds, tau_grid
    from bisip_b200.data import example_tables
    from bisip_b200.utils import prepare_data
    tab = example_tables()[anchor["file"]][anchor["headers"] - 1:]     # headers=1 skips only the title line
    d = prepare_data(tab, "mrad")
    names, b = default_bounds(anchor["model"], anchor.get("poly_deg", 5), anchor.get("n_modes", 1))
    assert names == anchor["param_names"]
    for k, v in anchor["bounds_edits"].items():
        b[:, names.index(k)] = v
    out = dict(data=d, bounds=b, names=names, taus=None, log_taus=None, log_tau=None)
    if anchor["model"] == "decomp":
        out["log_tau"], out["taus"], out["log_taus"] = tau_grid(d["w"], None, anchor["poly_deg"])
    return out


def p0_for(anchor, bounds, seed):
    """Inversion.fit's p0 (reference models.py:104-106): uniform inside the bounds."""
    rng = np.random.default_rng([20261017, seed])
    return rng.uniform(bounds[0], bounds[1], (anchor["nwalkers"], bounds.shape[1]))


def run_stats(anchor, chain, log_taus=None):
    """The statistics the notebook cell computed, from one run's full chain (nsteps, W, ndim)."""
    d, t = anchor["discard"], anchor["thin"]
    flat = chain[d + t - 1::t].reshape(-1, chain.shape[-1])
    out = {"mean": flat.mean(0), "std": flat.std(0)}
    if "pct" in anchor:
        out["pct"] = np.percentile(flat, anchor["p"], axis=0)
    if "total_m" in anchor:
        m = sum(out["mean"][1 + i] * log_taus[i] for i in range(log_taus.shape[0]))     # the tutorial's get_m
        out["total_m"] = np.array([m.sum()])
    return out


def _norm_ppf(q):
    from math import erf, sqrt
    lo, hi = -8.0, 8.0                                       # inverse normal cdf by bisection
    for _ in range(80):
        mid = 0.5 * (lo + hi)
        if 0.5 * (1 + erf(mid / sqrt(2))) < q:
            lo = mid
        else:
            hi = mid
    return lo


def ess_standard_errors(anchor, post_sd, tau):
    """Textbook Monte-Carlo standard errors of ONE run's statistics from sigma / sqrt(ESS), ESS = kept steps x walkers
    / tau (tau = integrated autocorrelation time in steps): mean sigma/sqrt(ESS); std sigma/sqrt(2 ESS); q-quantile
    sqrt(q(1-q)/ESS)/pdf(z_q) sigma (Gaussian posterior).  For the 2.5 / 97.5 percentiles this is ~2x the scatter
    actually measured over seeds (tau of the mean over-states the correlation of tail indicators)."""
    from math import exp, pi, sqrt
    kept = len(range(anchor["discard"] + anchor["thin"] - 1, anchor["nsteps"], anchor["thin"])) * anchor["thin"]
    ess = kept * anchor["nwalkers"] / float(tau)
    post_sd = np.asarray(post_sd, dtype=float)
    out = {"mean": post_sd / sqrt(ess), "std": post_sd / sqrt(2 * ess)}
    if "pct" in anchor:
        rows = []
        for p in anchor["p"]:
            q = p / 100.0
            z = _norm_ppf(q)
            rows.append(sqrt(q * (1 - q) / ess) / (exp(-0.5 * z * z) / sqrt(2 * pi)) * post_sd)
        out["pct"] = np.array(rows)
    return out


def zscores(anchor, runs, tau=None):
    """runs: list of run_stats dicts (one per seed) -> dict stat -> z array (same shape as the stored statistic).
    With `tau` (autocorrelation time measured on the sampler's own chains) the single-run error is the LARGER of
    the measured seed scatter and the sigma/sqrt(ESS) figure ("combined standard error"); without it, the measured
    scatter alone (stricter for tail percentiles)."""
    n = len(runs)
    z = {}
    ess = None
    if tau is not None:
        ess = ess_standard_errors(anchor, np.median(np.array([r["std"] for r in runs]), axis=0), tau)
    for key in ("mean", "std", "pct", "total_m"):
        if key not in anchor:
            continue
        nb = np.atleast_1d(np.asarray(anchor[key], dtype=float))
        v = np.array([r[key] for r in runs])
        centre = np.median(v, axis=0)
        scatter = 1.4826 * np.median(np.abs(v - centre), axis=0)
        if ess is not None and key in ess:
            scatter = np.maximum(scatter, ess[key])
        dec = anchor.get(f"{key}_decimals")
        rounding = 10.0 ** (-np.asarray(dec, dtype=float)) if dec is not None else 0.0
        var = scatter ** 2 * (1 + (np.pi / 2) / n) + np.asarray(rounding) ** 2 / 12.0
        z[key] = (nb - centre) / np.sqrt(var)
    return z


def autocorr_time(chains, anchor):
    """Mean integrated autocorrelation time (steps) of the post-burn-in chains, emcee's estimator."""
    from bisip_b200.sampler import integrated_time
    return float(np.mean([integrated_time(c[anchor["discard"]:], quiet=True) for c in chains]))
