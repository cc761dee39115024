"""Hyper-parameter sweeps / multi-start fits sharded over the GPUs of one box.

What GP_parameter_fit.py:30-33 does one evaluation at a time through GPy -
``GPRegression(X, Y, RBF(...))``, ``optimize()``, ``optimize_restarts(num_restarts=10)`` - is a set of
independent likelihood evaluations of one data set.  They shard naturally: one process per GPU,
rank r owns the contiguous slice ``[r*B/G, (r+1)*B/G)`` of the B hyper-parameter vectors, evaluates
it with the batched kernels (libgpb200 ``gpb_gpr_nlml_batched``), and only the scalar likelihoods
(and the D+2 gradients) are all-gathered - over NCCL/NVLink on GPUs, over gloo in the CPU tests
of this host logic.  There is no data-path collective: (X, y) is replicated.

``evaluate`` is injectable so that the partition / gather logic can be tested on CPU
(tests/test_sweep_gloo.py) without a device; the default goes to the CUDA library and fails
loudly if it is not there.
"""
import numpy as np

from . import _lib


def shard_bounds(B, world, rank):
    """Contiguous slice of problems owned by ``rank``."""
    return rank * B // world, (rank + 1) * B // world


def natural_params(log_hyp):
    """[l_1..l_D, sf2, sn2] per row from log hyper-parameters (GPr.py:93-97)."""
    h = np.exp(np.atleast_2d(np.asarray(log_hyp, dtype=float)))
    return np.concatenate([h[:, :-2], h[:, -2:-1] ** 2, h[:, -1:] ** 2], axis=1)


def _cuda_evaluate(X, y, log_hyp, want_grad):
    h = _lib.default_handle()
    h.set_train(X, y)
    kh = natural_params(log_hyp)
    if want_grad:
        vals, grads, info = h.gpr_nlml_batched(kh, want_grad=True)
    else:
        (vals, info), grads = h.gpr_nlml_batched(kh), None
    vals = np.where(info == 0, vals, np.inf)       # not positive definite -> +inf objective, like a failed restart
    return vals, grads


def sweep_nlml(X, y, log_hyp, want_grad=False, group=None, evaluate=None):
    """Negative log marginal likelihood (and gradient) of every row of ``log_hyp`` (B, D+2).

    With an initialised ``torch.distributed`` process group the rows are sharded over the ranks
    and the results all-gathered, so every rank returns the full (B,) / (B, D+2) arrays.
    """
    log_hyp = np.atleast_2d(np.asarray(log_hyp, dtype=float))
    B, P = log_hyp.shape
    evaluate = evaluate or _cuda_evaluate
    dist = None
    world, rank = 1, 0
    try:
        import torch.distributed as dist_mod
        if dist_mod.is_available() and dist_mod.is_initialized():
            dist = dist_mod
            world, rank = dist.get_world_size(group), dist.get_rank(group)
    except ImportError:
        pass
    lo, hi = shard_bounds(B, world, rank)
    if hi > lo:
        vals, grads = evaluate(X, y, log_hyp[lo:hi], want_grad)
    else:
        vals, grads = np.empty(0), (np.empty((0, P)) if want_grad else None)
    if dist is None or world == 1:
        return (vals, grads) if want_grad else vals

    import torch
    backend = dist.get_backend(group)
    dev = torch.device('cuda', torch.cuda.current_device()) if backend == 'nccl' else torch.device('cpu')
    width = 1 + (P if want_grad else 0)
    cap = max(shard_bounds(B, world, r)[1] - shard_bounds(B, world, r)[0] for r in range(world))
    mine = torch.zeros((cap, width), dtype=torch.float64, device=dev)
    if hi > lo:
        block = vals[:, None] if not want_grad else np.concatenate([vals[:, None], grads], axis=1)
        mine[:hi - lo] = torch.from_numpy(np.ascontiguousarray(block)).to(dev)
    if backend == 'nccl':
        # one collective into one buffer and ONE device-to-host copy (a copy per rank costs more than the gather)
        flat = torch.empty((world, cap, width), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(flat, mine, group=group)
        host = flat.cpu().numpy()
        parts = [host[r] for r in range(world)]
    else:
        plist = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(plist, mine, group=group)
        parts = [p.numpy() for p in plist]
    out = np.empty((B, width))
    for r in range(world):
        a, b = shard_bounds(B, world, r)
        out[a:b] = parts[r][:b - a]
    return (out[:, 0], out[:, 1:]) if want_grad else out[:, 0]


def multistart_fit(X, y, log_hyp0, n_restarts=10, n_iter=50, seed=0, group=None, evaluate=None):
    """GPy-style ``optimize`` + ``optimize_restarts`` (GP_parameter_fit.py:32-33) in lock step.

    All restarts advance together: each iteration evaluates value + gradient of every live start
    with one sharded batched call, then takes one L-BFGS step per start on the host (scipy's
    two-loop recursion state per start, history 8, backtracking on the batched value).
    Returns (best_log_hyp, best_nlml, all_final_log_hyp, all_final_nlml).
    """
    rng = np.random.default_rng(seed)
    log_hyp0 = np.asarray(log_hyp0, dtype=float)
    P = log_hyp0.size
    starts = np.vstack([log_hyp0[None, :], log_hyp0[None, :] + rng.standard_normal((n_restarts, P))])
    S = len(starts)
    x = starts.copy()
    f, g = sweep_nlml(X, y, x, want_grad=True, group=group, evaluate=evaluate)
    hist = [([], []) for _ in range(S)]          # (s, y) pairs per start
    step0 = np.full(S, 1.0)
    for _ in range(n_iter):
        d = np.empty_like(x)
        for i in range(S):
            q = g[i].copy()
            s_l, y_l = hist[i]
            al = []
            for s_k, y_k in zip(reversed(s_l), reversed(y_l)):
                a = (s_k @ q) / (y_k @ s_k)
                al.append(a)
                q -= a * y_k
            gamma = (s_l[-1] @ y_l[-1]) / (y_l[-1] @ y_l[-1]) if s_l else 1.0 / max(np.linalg.norm(g[i]), 1.0)
            r = gamma * q
            for (s_k, y_k), a in zip(zip(s_l, y_l), reversed(al)):
                b = (y_k @ r) / (y_k @ s_k)
                r += s_k * (a - b)
            d[i] = -r
        t = step0.copy()
        done = ~np.isfinite(f)
        x_new, f_new, g_new = x.copy(), f.copy(), g.copy()
        for _ls in range(8):                      # batched backtracking: every trial is one sharded call
            todo = np.where(~done)[0]
            if todo.size == 0:
                break
            cand = x[todo] + t[todo, None] * d[todo]
            fc, gc = sweep_nlml(X, y, cand, want_grad=True, group=group, evaluate=evaluate)
            ok = np.isfinite(fc) & (fc <= f[todo] + 1e-4 * t[todo] * np.einsum('ij,ij->i', g[todo], d[todo]))
            for j, i in enumerate(todo):
                if ok[j]:
                    x_new[i], f_new[i], g_new[i] = cand[j], fc[j], gc[j]
                    done[i] = True
                else:
                    t[i] *= 0.5
        for i in range(S):
            s_k, y_k = x_new[i] - x[i], g_new[i] - g[i]
            if s_k @ y_k > 1e-12:
                hist[i][0].append(s_k)
                hist[i][1].append(y_k)
                if len(hist[i][0]) > 8:
                    hist[i][0].pop(0)
                    hist[i][1].pop(0)
        if np.max(np.abs(x_new - x)) < 1e-9:
            x, f, g = x_new, f_new, g_new
            break
        x, f, g = x_new, f_new, g_new
    best = int(np.nanargmin(np.where(np.isfinite(f), f, np.inf)))
    return x[best], float(f[best]), x, f
