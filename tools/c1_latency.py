#!/usr/bin/env python
"""Wall time of ONE drop-in `model.fit()` call on BASELINE configs 1 and 2 (a single spectrum = one CTA of the GPU) next to the
unmodified reference's `fit()` on one host core (reference models.py + Cython under oracle/emcee_restatement.py):

    python tools/c1_latency.py [--no-reference]
"""
import json
import os
import sys
import time
import warnings

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter('ignore')
import torch  # noqa: E402
import bisip_b200 as bb  # noqa: E402

fp = bb.DataFiles()['SIP-K389175']
cases = (('C1 PolynomialDecomposition poly_deg=4, 32 walkers x 1000 steps', 'PolynomialDecomposition', dict(nwalkers=32, poly_deg=4, nsteps=1000)),
         ('C2 PeltonColeCole n_modes=2, 64 walkers x 2000 steps', 'PeltonColeCole', dict(nwalkers=64, n_modes=2, nsteps=2000)))
for tag, cls, kw in cases:
    m = getattr(bb, cls)(fp, **kw)
    ts = []
    for i in range(5):
        np.random.seed(i)
        t0 = time.perf_counter()
        m.fit()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    t0 = time.perf_counter()
    pct = m.get_param_percentile([2.5, 50, 97.5], discard=kw['nsteps'] // 2)
    t_pct = time.perf_counter() - t0
    print(json.dumps({"case": tag, "impl": "bisip_b200", "fit_s_first": ts[0], "fit_s_best": min(ts[1:]),
                      "get_param_percentile_s": t_pct, "acceptance": float(np.mean(m.sampler.acceptance_fraction)),
                      "evals": kw['nwalkers'] * (kw['nsteps'] + 1)}), flush=True)
if '--no-reference' not in sys.argv:
    from oracle import refload
    ref = refload.load()
    for tag, cls, kw in cases:
        r = getattr(ref, cls)(refload.data_file('SIP-K389175'), **kw)
        np.random.seed(0)
        t0 = time.perf_counter()
        r.fit()
        print(json.dumps({"case": tag, "impl": "reference (1 core)", "fit_s": time.perf_counter() - t0,
                          "evals": kw['nwalkers'] * (kw['nsteps'] + 1)}), flush=True)
