"""Worker of test_sweep_nlml_two_ranks_equal_one_rank_bitwise: launched by torch.distributed.run with two
ranks, one per GPU.  Every rank calls the product's sweep_nlml (default CUDA evaluator, NCCL all-gather);
rank 0 also evaluates all rows alone and compares bit for bit."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import bench_configs as cfg
    from gptest_b200 import _lib, sweep

    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    _lib.set_default_device(local)
    X, Y, lhs = cfg.make_c5()
    sel = lhs[::8][:96]                                   # 96 problems spread over the grid, 48 per rank
    vals = sweep.sweep_nlml(X, Y, sel)
    vg, gg = sweep.sweep_nlml(X, Y, sel[:10], want_grad=True)
    ok = True
    if dist.get_rank() == 0:
        alone, _ = sweep._cuda_evaluate(X, Y, sel, False)
        ok = np.array_equal(alone, vals) and np.isfinite(vals).all()
        a2, g2 = sweep._cuda_evaluate(X, Y, sel[:10], True)
        ok = ok and np.array_equal(a2, vg) and np.allclose(g2, gg, rtol=1e-12, atol=0)
        if not ok:
            print("mismatch", np.abs(alone - vals).max(), np.abs(a2 - vg).max(), np.abs(g2 - gg).max())
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    if dist_rank0_print(flag):
        print("SWEEP_NCCL_OK")
    sys.exit(0 if flag.item() == 1 else 1)


def dist_rank0_print(flag):
    return int(os.environ.get("RANK", "0")) == 0 and flag.item() == 1


if __name__ == "__main__":
    main()
