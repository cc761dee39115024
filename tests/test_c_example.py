"""The C ABI from plain C (examples/gpr_fit.c): compiles and links against libgpb200.so with gcc on the CPU;
on a GPU it runs and its numbers are compared with the oracle."""
import os
import struct
import subprocess
import sys

import numpy as np
import pytest

from oracle import gpr_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build(tmp_path):
    from gptest_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    exe = str(tmp_path / 'gpr_fit')
    libdir = os.path.dirname(_lib.LIB_PATH)
    cmd = ['gcc', '-O2', '-Wall', '-Werror', '-I' + os.path.join(ROOT, 'include'), os.path.join(ROOT, 'examples', 'gpr_fit.c'),
           '-o', exe, '-L' + libdir, '-lgpb200', '-Wl,-rpath,' + libdir, '-lm']
    subprocess.run(cmd, check=True, capture_output=True)
    return exe


def test_c_example_compiles_and_links(tmp_path):
    exe = build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and 'usage' in r.stderr


@pytest.mark.gpu
def test_c_example_matches_oracle(tmp_path):
    exe = build(tmp_path)
    rng = np.random.default_rng(3)
    n, d, m = 600, 3, 40
    X = rng.random((n, d))
    y = np.sin(X.sum(1)) + 0.1 * rng.standard_normal(n)
    Z = rng.random((m, d))
    lh = np.log([0.6, 0.5, 0.7, 1.1, 0.15])
    khyp = np.r_[np.exp(lh[:d]), np.exp(lh[d]) ** 2, np.exp(lh[d + 1]) ** 2]
    path = str(tmp_path / 'data.bin')
    with open(path, 'wb') as f:
        f.write(struct.pack('<qqq', n, d, m))
        for a in (X, y, Z, khyp):
            f.write(np.ascontiguousarray(a, dtype='<f8').tobytes())
    r = subprocess.run([exe, path], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = {'grad': {}, 'pred': {}}
    for line in r.stdout.splitlines():
        p = line.split()
        if p[0] == 'nlml':
            out['nlml'] = float(p[1])
        elif p[0] == 'grad':
            out['grad'][int(p[1])] = float(p[2])
        elif p[0] == 'pred':
            out['pred'][int(p[1])] = (float(p[2]), float(p[3]))
        elif p[0] == 'launches':
            out['launches'] = int(p[1])
    ref = gpr_oracle.nlml_chol(lh, X, y)
    assert abs(out['nlml'] - ref) <= 1e-8 * abs(ref)
    g = np.array([out['grad'][k] for k in range(d + 2)])
    rg = gpr_oracle.nlml_grad(lh, X, y)
    assert np.abs(g - rg).max() <= 1e-7 * max(1.0, np.abs(rg).max())
    fz = np.array([out['pred'][i][0] for i in range(m)])
    cov = np.array([out['pred'][i][1] for i in range(m)])
    rf, rc = gpr_oracle.predict_chol(lh, X, y, Z)
    assert np.abs(fz - rf).max() <= 1e-9 * max(1.0, np.abs(rf).max()) and np.abs(cov - rc).max() <= 1e-9
    assert out['launches'] > 0
