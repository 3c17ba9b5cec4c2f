"""Shared helpers of the parity tests: build oracle Problems from the golden fixtures."""
import numpy as np

from oracle import oracle

CASES = {   # golden case -> (oracle model, ctor kwargs)
    'decomp_p4_debye': ('decomp', dict(poly_deg=4, c_exp=1.0)),
    'decomp_p4_warburg': ('decomp', dict(poly_deg=4, c_exp=0.5)),
    'decomp_p5_debye': ('decomp', dict(poly_deg=5, c_exp=1.0)),
    'decomp_p3_c07': ('decomp', dict(poly_deg=3, c_exp=0.7)),
    'colecole_k1': ('colecole', dict(n_modes=1)),
    'colecole_k2': ('colecole', dict(n_modes=2)),
    'colecole_k3': ('colecole', dict(n_modes=3)),
    'dias': ('dias', {}),
    'shin': ('shin', {}),
}


def normwise(a, b):
    """max|a-b| / max|b| per leading item (norm-wise relative error)."""
    a, b = np.asarray(a), np.asarray(b)
    ax = tuple(range(1, a.ndim))
    return np.max(np.abs(a - b), axis=ax) / np.max(np.abs(b), axis=ax)


def lp_err(a, b):
    """|a-b| / max(1,|b|) with -inf == -inf counted as exact."""
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    same_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    with np.errstate(invalid='ignore'):
        e = np.abs(a - b) / np.maximum(1.0, np.abs(b))
    return np.where(same_inf, 0.0, np.where(np.isnan(e), np.inf, e))


def oracle_problem(case, gold_fl, gold_ld, name='SIP-K389175'):
    model, kw = CASES[case]
    return oracle.Problem(model, gold_ld[f'{name}/w'], gold_ld[f'{name}/zn'], gold_ld[f'{name}/zn_err'],
                          gold_fl[f'{case}/bounds'], n_modes=kw.get('n_modes', 1),
                          taus=gold_fl.get(f'{case}/taus'), log_taus=gold_fl.get(f'{case}/log_taus'),
                          c_exp=kw.get('c_exp', 1.0))


def make_model(case, fp, **extra):
    import bisip_b200 as bb
    model, kw = CASES[case]
    cls = {'decomp': bb.PolynomialDecomposition, 'colecole': bb.PeltonColeCole, 'dias': bb.Dias2000,
           'shin': bb.Shin2015}[model]
    return cls(fp, **kw, **extra)
