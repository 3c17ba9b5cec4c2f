#!/usr/bin/env python
"""Time bisip_column_stats on a kept chain of the C5 shard shape: [B][n_keep*W][ndim] (developer tool)."""
import argparse
import json
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bisip_b200 import engine, _lib

ap = argparse.ArgumentParser()
ap.add_argument("--spectra", type=int, default=12500)
ap.add_argument("--n", type=int, default=25600)
ap.add_argument("--ncol", type=int, default=6)
a = ap.parse_args()
dev = _lib.require_cuda()
g = torch.Generator(device=dev).manual_seed(1)
x = torch.randn((a.spectra, a.n, a.ncol), dtype=torch.float64, device=dev, generator=g)
x[..., 0] = 1.0 + 0.005 * x[..., 0]          # r0-like: narrow, straddles 1.0
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
best = 1e30
for rep in range(4):
    ev[0].record()
    st = engine.column_stats(x, p=[2.5, 50, 97.5], want_mean=True, want_std=True)
    ev[1].record()
    torch.cuda.synchronize()
    best = min(best, ev[0].elapsed_time(ev[1]))
gb = x.numel() * 8 / 1e9
print(json.dumps({"spectra": a.spectra, "n": a.n, "ncol": a.ncol, "ms": best, "GB": gb, "GB_per_s": gb / (best * 1e-3)}))
chk = torch.quantile(x[:3, :, 0], torch.tensor([0.025, 0.5, 0.975], dtype=torch.float64, device=dev), dim=1).T
print("max |pct - torch.quantile| (first 3 spectra, col 0):", float((st['pct'][:3, :, 0] - chk).abs().max()))
