"""The reference's native names (cython_funcs.pyx:49-108) served by the CUDA kernels: bisip_b200.cython_funcs.

* the SURVEY App. C.1 raw-kernel vectors (golden `raw/*`, generated from the reference build);
* keyword call signatures, fresh (2, N) float64 result, Cython's typed-buffer errors;
* the UNMODIFIED reference models.py (oracle/_ref) running with `bisip.cython_funcs` replaced by the shim:
  its forward / _log_probability must reproduce the golden vectors that the real Cython module produced.
"""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from helpers import normwise

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-12


def test_raw_kernel_vectors(gold_fl):
    from bisip_b200 import cython_funcs as cf
    w = gold_fl['raw/w']
    got = cf.ColeCole_cyth(w, 1.0, np.array([0.3]), np.array([-2.0]), np.array([0.5]))
    assert got.shape == (2, 4) and got.dtype == np.float64
    assert normwise(got[None], gold_fl['raw/colecole'][None]).max() <= TOL
    got = cf.Dias2000_cyth(w, 1.0, 0.25, -10.0, 5.0, 0.5)
    assert normwise(got[None], gold_fl['raw/dias'][None]).max() <= TOL
    got = cf.Shin2015_cyth(w, np.array([0.5, 0.5]), np.array([-14.0, -6.0]), np.array([0.5, 0.5]))
    assert normwise(got[None], gold_fl['raw/shin'][None]).max() <= TOL


def test_keyword_calls_like_models_py(gold_fl):
    """models.py:228-229, 267-271, 305, 345-349 call with keywords R0=, a=, m=, lt=, c=, R=, log_Q=, n=."""
    from bisip_b200 import cython_funcs as cf
    case = 'decomp_p4_debye'
    w = gold_fl['raw/w']
    th = gold_fl[f'{case}/theta'][0]
    taus, log_taus = gold_fl[f'{case}/taus'], gold_fl[f'{case}/log_taus']
    a = cf.Decomp_cyth(w, taus, log_taus, 1.0, R0=th[0], a=th[1:])
    b = cf.Decomp_cyth(w=w, taus=taus, log_taus=log_taus, c_exp=1.0, R0=th[0], a=th[1:])
    np.testing.assert_array_equal(a, b)
    assert a is not b
    a = cf.ColeCole_cyth(w, R0=1.0, m=np.array([0.3, 0.2]), lt=np.array([-2.0, -9.0]), c=np.array([0.5, 0.7]))
    s = (cf.ColeCole_cyth(w, 1.0, np.array([0.3]), np.array([-2.0]), np.array([0.5]))
         + cf.ColeCole_cyth(w, 1.0, np.array([0.2]), np.array([-9.0]), np.array([0.7])))
    s[0] -= 1.0                                   # R0 (1 - z1 - z2) = [R0 (1 - z1)] + [R0 (1 - z2)] - R0
    np.testing.assert_allclose(a, s, rtol=0, atol=1e-14)
    z3 = cf.Shin2015_cyth(w, R=np.array([0.5, 0.4, 0.3]), log_Q=np.array([-14.0, -6.0, -9.0]), n=np.array([0.5, 0.6, 0.7]))
    z2 = cf.Shin2015_cyth(w, np.array([0.5, 0.4]), np.array([-14.0, -6.0]), np.array([0.5, 0.6]))
    z1 = cf.Shin2015_cyth(w, np.array([0.3]), np.array([-9.0]), np.array([0.7]))
    np.testing.assert_allclose(z3, z2 + z1, rtol=0, atol=1e-15)
    assert np.array_equal(cf.ColeCole_cyth(w, 1.05, np.empty(0), np.empty(0), np.empty(0)),
                          np.array([np.full(4, 1.05), np.zeros(4)]))


def test_typed_buffer_errors(gold_fl):
    from bisip_b200 import cython_funcs as cf
    w = gold_fl['raw/w']
    with pytest.raises(ValueError, match='Buffer dtype mismatch'):
        cf.Dias2000_cyth(w.astype(np.float32), 1.0, 0.25, -10.0, 5.0, 0.5)
    with pytest.raises(ValueError, match='Buffer dtype mismatch'):
        cf.ColeCole_cyth(w, 1.0, np.array([1]), np.array([-2.0]), np.array([0.5]))
    with pytest.raises(ValueError, match='wrong number of dimensions'):
        cf.Decomp_cyth(w, np.ones(3), np.ones(3), 1.0, 1.0, np.ones(1))
    with pytest.raises(TypeError):
        cf.Shin2015_cyth(list(w), np.array([0.5, 0.5]), np.array([-14.0, -6.0]), np.array([0.5, 0.5]))
    with pytest.raises(TypeError):
        cf.Dias2000_cyth(w, 'one', 0.25, -10.0, 5.0, 0.5)


def test_unmodified_reference_models_run_on_the_shim(tmp_path):
    """`sys.modules['bisip.cython_funcs'] = bisip_b200.cython_funcs`, then import the reference package from
    oracle/_ref: its own models.py / utils.py drive the CUDA forward kernels.  In a subprocess, so that the other
    tests of this session keep the real Cython module."""
    if not os.path.isdir(os.path.join(ROOT, 'oracle', '_ref', 'bisip')):
        pytest.skip('oracle/_ref not built (needs /root/reference at build time)')
    code = textwrap.dedent('''
        import sys, numpy as np
        sys.path.insert(0, %r); sys.path.insert(0, %r)
        import bisip_b200.cython_funcs as shim
        from bisip_b200 import _lib
        sys.modules['bisip.cython_funcs'] = shim
        from oracle import refload
        bisip = refload.load()
        assert bisip.models.Decomp_cyth is shim.Decomp_cyth and bisip.models.ColeCole_cyth is shim.ColeCole_cyth
        from helpers import CASES, lp_err, normwise
        gold = dict(np.load(%r))
        fp = refload.data_file('SIP-K389175')
        n0 = _lib.launch_count()
        for case, (model, kw) in CASES.items():
            cls = {'decomp': bisip.PolynomialDecomposition, 'colecole': bisip.PeltonColeCole, 'dias': bisip.Dias2000,
                   'shin': bisip.Shin2015}[model]
            m = cls(fp, **kw)
            th = gold[case + '/theta'][:12]
            Z = np.array([m.forward(t, m.data['w']) for t in th])
            assert normwise(Z, gold[case + '/Z'][:12]).max() <= 1e-12, case
            lp = np.array([m._log_probability(t, m.forward, m.param_bounds, m.data['w'], m.data['zn'], m.data['zn_err'])
                           for t in th])
            assert lp_err(lp, gold[case + '/lp'][:12]).max() <= 1e-12, case
        assert _lib.launch_count() - n0 >= 2 * 12 * len(CASES) - 40      # out-of-box thetas skip the forward
        print('ok', _lib.launch_count() - n0)
    ''') % (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'tests', 'golden', 'forward_logprob.npz'))
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.strip().startswith('ok')
