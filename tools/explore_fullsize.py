"""Exploration: coverage of theta_true by the 95 % posterior interval at BASELINE sizes (developer tool)."""
import sys, os, time, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bisip_b200 import _lib, engine, synthetic
from bisip_b200.batch import BatchInversion

dev = torch.device("cuda:0")
def run(model, B, W, T, discard, thin, **kw):
    _, w = synthetic.frequencies(64)
    probe = BatchInversion(model, w, np.zeros((1, 2, 64)), np.ones((1, 2, 64)), device=dev, **kw)
    fwd = lambda th, ww: engine.forward(probe._spec(), _lib.dev_f64(th[:, None, :], dev), _lib.dev_f64(ww, dev))[:, 0].cpu().numpy()
    syn = synthetic.make(model, 0, B, fwd, N=64, poly_deg=kw.get('poly_deg', 4), n_tau=kw.get('n_tau'))
    inv = BatchInversion(model, w, syn['zn'], syn['zn_err'], nwalkers=W, nsteps=T, seed=5, device=dev, **kw)
    t0 = time.time(); r = inv.fit(discard=discard, thin=thin); dt = time.time() - t0
    truth = syn['theta_true'].copy()
    if model in ('decomp', 'colecole', 'dias'):
        truth[:, 0] /= syn['norm_factor']
    else:
        truth[:, :2] /= syn['norm_factor'][:, None]
    lo, hi = r['percentiles'][:, 0], r['percentiles'][:, 2]
    cov = ((lo <= truth) & (truth <= hi)).mean(0)
    z = (r['mean'] - truth) / r['std']
    print(json.dumps(dict(model=model, B=B, s=round(dt, 2), coverage=[round(c, 4) for c in cov], acc=[float(r['acceptance_fraction'].min()), float(r['acceptance_fraction'].mean()), float(r['acceptance_fraction'].max())],
                          flags=int((r['flags'] != 0).sum()), zmean=[round(v, 3) for v in z.mean(0)], zstd=[round(v, 3) for v in z.std(0)])))
run('decomp', 12500, 256, 2000, 1000, 10, poly_deg=4, n_tau=64)
run('dias', 1024, 128, 2000, 1000, 5)
run('shin', 1024, 128, 2000, 1000, 5)
run('colecole', 1024, 128, 2000, 1000, 5, n_modes=1)
run('decomp', 10000, 256, 300, 150, 5, poly_deg=4, n_tau=256)
