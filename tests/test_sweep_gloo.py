"""Host logic of the sharded sweep (gptest_b200/sweep.py) on CPU: world_size 2 over gloo.

The evaluator is injected (the CPU oracle stands in for the CUDA library, which is legitimate here:
this is a test of the partition / gather / multi-start logic, not of the product arithmetic)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_eval(X, y, log_hyp, want_grad):
    from oracle import gpr_oracle
    vals = np.array([gpr_oracle.nlml_chol(l, X, y) for l in log_hyp])
    grads = np.array([gpr_oracle.nlml_grad(l, X, y) for l in log_hyp]) if want_grad else None
    return vals, grads


def _data():
    rng = np.random.default_rng(0)
    X = rng.random((40, 2))
    y = np.sin(3 * X[:, 0]) + 0.1 * rng.standard_normal(40)
    lh = np.log([0.5, 0.5, 1.0, 0.1]) + 0.2 * rng.standard_normal((7, 4))       # 7 problems: ragged over 2 ranks
    return X, y, lh


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from gptest_b200 import sweep
    dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world)
    X, y, lh = _data()
    vals, grads = sweep.sweep_nlml(X, y, lh, want_grad=True, evaluate=_oracle_eval)
    v2 = sweep.sweep_nlml(X, y, lh[:1], evaluate=_oracle_eval)                    # fewer problems than ranks
    best, fbest, xs, fs = sweep.multistart_fit(X, y, lh[0], n_restarts=3, n_iter=6, evaluate=_oracle_eval)
    np.savez(os.path.join(out, 'r%d.npz' % rank), vals=vals, grads=grads, v2=v2, best=best, fbest=fbest, fs=fs)
    dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    from gptest_b200.sweep import shard_bounds
    for B in (1, 7, 1024):
        for G in (1, 2, 3, 8):
            spans = [shard_bounds(B, G, r) for r in range(G)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(G - 1))
    assert shard_bounds(1024, 8, 3) == (384, 512)


def test_sharded_sweep_equals_unsharded(tmp_path):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / 'r0.npz'), np.load(tmp_path / 'r1.npz')
    X, y, lh = _data()
    sys.path.insert(0, ROOT)
    from gptest_b200 import sweep
    vals, grads = sweep.sweep_nlml(X, y, lh, want_grad=True, evaluate=_oracle_eval)   # no process group: unsharded
    for r in (r0, r1):
        assert np.array_equal(r['vals'], vals) and np.array_equal(r['grads'], grads)
        assert r['v2'].shape == (1,) and r['v2'][0] == vals[0]
    assert np.array_equal(r0['best'], r1['best']) and r0['fbest'] == r1['fbest']
    assert r0['fbest'] <= vals[0] + 1e-12                   # the multi-start never ends above its first start
    assert np.isfinite(r0['fs']).all()
