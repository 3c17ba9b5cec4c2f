"""Product-level multi-GPU entry points on real devices: fit_sharded without a process group (one GPU), a world-size-2
NCCL run (skipped on single-GPU boxes), and the `device=` argument with a device that is not the current one."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _survey(B=12, N=16, seed=3):
    from bisip_b200 import _lib, engine, synthetic
    from bisip_b200.batch import BatchInversion
    dev = _lib.require_cuda()
    probe = BatchInversion('dias', synthetic.frequencies(N)[1], np.zeros((1, 2, N)), np.ones((1, 2, N)), device=dev)
    fwd = lambda th, ww: engine.forward(probe._spec(), _lib.dev_f64(th[:, None, :], dev), _lib.dev_f64(ww, dev))[:, 0].cpu().numpy()
    return synthetic.make('dias', 0, B, fwd, N=N)


def test_fit_sharded_single_process_equals_batch_fit():
    import bisip_b200 as bb
    syn = _survey()
    kw = dict(nwalkers=32, nsteps=120, seed=11)
    ref = bb.BatchInversion('dias', syn['w'], syn['zn'], syn['zn_err'], **kw).fit(discard=40, thin=2, keep_chain=True)
    res = bb.fit_sharded('dias', syn['w'], syn['zn'], syn['zn_err'], discard=40, thin=2, gather_chain=True, **kw)
    for k in ('percentiles', 'mean', 'std', 'acceptance_fraction', 'flags', 'chain'):
        np.testing.assert_array_equal(res[k], ref[k])
    assert res['shard'] == (0, 12)
    inv = bb.BatchInversion('dias', syn['w'], syn['zn'], syn['zn_err'], **kw)
    got = inv.fit_gathered(12, discard=40, thin=2)                  # no process group: equals fit()
    for k in ('percentiles', 'mean', 'std', 'acceptance_fraction', 'flags'):
        np.testing.assert_array_equal(got[k], ref[k])


_NCCL_WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["BISIP_ROOT"]); sys.path.insert(0, os.path.join(os.environ["BISIP_ROOT"], "tests"))
import numpy as np, torch, torch.distributed as dist
import bisip_b200 as bb
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", init_method="tcp://127.0.0.1:" + os.environ["PORT"], rank=rank, world_size=world,
                        device_id=torch.device("cuda", rank))
from test_gpu_multi import _survey
syn = _survey(B=13)
kw = dict(nwalkers=32, nsteps=120, seed=11)
res = bb.fit_sharded("dias", syn["w"], syn["zn"], syn["zn_err"], discard=40, thin=2, gather_chain=True, **kw)
ref = bb.BatchInversion("dias", syn["w"], syn["zn"], syn["zn_err"], **kw).fit(discard=40, thin=2, keep_chain=True)
for k in ("percentiles", "mean", "std", "acceptance_fraction", "flags"):
    np.testing.assert_array_equal(res[k], ref[k])            # independent of the number of ranks
if rank == 0:
    np.testing.assert_array_equal(res["chain"], ref["chain"])
else:
    assert res["chain"] is None
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_fit_sharded_two_ranks_nccl(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "worker_nccl.py"
    script.write_text(_NCCL_WORKER)
    port = str(29700 + os.getpid() % 200)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", PORT=port, BISIP_ROOT=ROOT)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=600)
        assert p.returncode == 0, out
        assert "ok" in out


def test_non_current_device_argument():
    """`device='cuda:1'` while device 0 is current: the C ABI switches to the device that owns the buffers."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import bisip_b200 as bb
    torch.cuda.set_device(0)
    syn = _survey(B=4)
    kw = dict(nwalkers=32, nsteps=60, seed=5)
    r0 = bb.BatchInversion('dias', syn['w'], syn['zn'], syn['zn_err'], device='cuda:0', **kw).fit(discard=20)
    r1 = bb.BatchInversion('dias', syn['w'], syn['zn'], syn['zn_err'], device='cuda:1', **kw).fit(discard=20)
    assert torch.cuda.current_device() == 0
    np.testing.assert_array_equal(r0['percentiles'], r1['percentiles'])
    m = bb.Dias2000(bb.DataFiles()['SIP-K389172'], nwalkers=32, nsteps=50, seed=1, device='cuda:1')
    m.fit()
    assert m.get_chain().shape == (50, 32, 5) and torch.cuda.current_device() == 0
