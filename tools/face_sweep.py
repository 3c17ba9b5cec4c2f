#!/usr/bin/env python
"""forward() of the drop-in classes against the LIVE reference (oracle/_ref) with one parameter at a time (and two at a time:
corners) on a face of the default box.  The prior is strict, so the sampler never evaluates these points, but forward() can be
called there, and the reference's C complex arithmetic returns finite values on every face (1/R = inf, 1/delta = inf ...).

    python tools/face_sweep.py          # one JSON line per disagreement, then a summary line

`sweep()` is also what tests/test_gpu_parity.py::test_forward_on_box_faces_matches_live_reference runs."""
import json
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CONFIGS = (('PolynomialDecomposition', dict(poly_deg=4), 'SIP-K389175'),
           ('PolynomialDecomposition', dict(poly_deg=3, c_exp=0.5), 'SIP-K389170'),
           ('PeltonColeCole', dict(n_modes=1), 'SIP-K389172'), ('PeltonColeCole', dict(n_modes=2), 'SIP-K389174'),
           ('PeltonColeCole', dict(n_modes=3), 'SIP-K389173'), ('Dias2000', {}, 'SIP-K389172'),
           ('Shin2015', {}, 'SIP-K389176'))


def sweep(tol=1e-12):
    """Returns (list of disagreement records, worst norm-wise error among the agreeing points, number of points)."""
    import bisip_b200 as bb
    from oracle import refload
    ref = refload.load()
    files = bb.DataFiles()
    rng = np.random.default_rng(5)
    bad, worst, npts = [], 0.0, 0
    for cls, kw, name in CONFIGS:
        r = getattr(ref, cls)(refload.data_file(name), **kw)
        m = getattr(bb, cls)(files[name], **kw)
        B = r.param_bounds
        nd = B.shape[1]
        th, tags = [], []
        for i in range(nd):
            for side in (0, 1):
                for rep in range(3):
                    t = rng.uniform(B[0], B[1])
                    if cls == 'PolynomialDecomposition':
                        t[1:] *= 0.05
                    t[i] = B[side, i]
                    th.append(t)
                    tags.append((r.param_names[i], 'lo' if side == 0 else 'hi'))
        for rep in range(12):                                     # two faces at once
            t = rng.uniform(B[0], B[1])
            i, j = rng.choice(nd, 2, replace=False)
            t[i], t[j] = B[rng.integers(2), i], B[rng.integers(2), j]
            th.append(t)
            tags.append((f'{r.param_names[i]}+{r.param_names[j]}', 'corner'))
        th = np.array(th)
        Zr = np.stack([r.forward(t, r.data['w']) for t in th])
        Z = m.forward(th, m.data['w'])
        npts += len(tags)
        # the prior-free likelihood at the same points (the fused log-probability kernel's forward, not the batched one)
        llr = np.array([r._log_likelihood(t, r.forward, r.data['w'], r.data['zn'], r.data['zn_err']) for t in th])
        ll = m._log_likelihood(th, m.forward, m.data['w'], m.data['zn'], m.data['zn_err'])
        with np.errstate(invalid='ignore'):
            ll_err = np.abs(ll - llr) / np.maximum(1.0, np.abs(llr))
        for k in np.nonzero(~(ll_err <= tol))[0]:
            bad.append({"model": cls, "kw": kw, "param": tags[k][0], "side": tags[k][1], "status": "log-likelihood",
                        "err": float(ll_err[k]), "theta": th[k].tolist(), "ref0": float(llr[k]), "cuda0": float(ll[k])})
        for k, tag in enumerate(tags):
            fr, fg = np.isfinite(Zr[k]).all(), np.isfinite(Z[k]).all()
            err = None
            if fr and fg:
                scale = np.max(np.abs(Zr[k]))
                err = float(np.max(np.abs(Z[k] - Zr[k])) / (scale if scale > 0 else 1.0))
                if err <= tol:
                    worst = max(worst, err)
                    continue
                status = 'mismatch'
            else:
                status = 'nonfinite-both' if not (fr or fg) else ('reference finite, cuda not' if fr else 'cuda finite, reference not')
            bad.append({"model": cls, "kw": kw, "param": tag[0], "side": tag[1], "status": status, "err": err,
                        "theta": th[k].tolist(), "ref0": Zr[k][:, 0].tolist(), "cuda0": Z[k][:, 0].tolist()})
    return bad, worst, npts


if __name__ == '__main__':
    sys.path.insert(0, ROOT)
    warnings.simplefilter('ignore')
    bad, worst, npts = sweep()
    for b in bad:
        print(json.dumps(b))
    print(json.dumps({"points": npts, "disagreements": len(bad), "worst_ok_err": worst}))
