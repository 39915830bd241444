"""Golden answers of the reference's OWN hyper-parameter search (lib.metrics.utils.maximize_metric :151-196 with
its scipy, LBFGS and grid stages, optim_func_torch :123-127, calc_scores_given_hparams_vectorized(torch_arr=True)
:47-61) run live on a small LEMoN-like validation frame.  Builder container only:
    python tests/golden/make_golden_hparam_search.py
tests/test_hparam_search.py pins the port (oracle/hparam_search_port.py) to the file on the CPU and the GPU drop-ins
(patch_reference_metrics / patch_reference_hparam_search applied to the port module) on the B200."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_live  # noqa: E402
from tests.helpers import hparam_search_case  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    mu = ref_live.import_reference_metrics()
    out = {}
    for tag in ("plain", "ablate"):
        df, grid, x0s, fz, fo = hparam_search_case(tag)
        bx, bv, bt = mu.maximize_metric(df, grid, x0s, mu.optimize_f1_efficient, {}, force_zero=fz, force_one=fo)
        out[f"{tag}_best_x"], out[f"{tag}_best_val"], out[f"{tag}_best_thr"] = np.asarray(bx, np.float64), bv, bt
        # the differentiable stage alone: loss and gradient of optim_func_torch at a fixed point
        x = torch.tensor([1.5, 0.5, 0.3, 2.0, 0.7, 1.0], dtype=torch.float64, requires_grad=True)
        loss = mu.optim_func_torch(x, df, force_zero=fz, force_one=fo)
        loss.backward()
        out[f"{tag}_loss"], out[f"{tag}_grad"] = loss.item(), x.grad.numpy()
        lb = mu.maximize_metric_torch(df, x0s[1], mu.optimize_f1_efficient, {}, force_zero=fz, force_one=fo)
        out[f"{tag}_lbfgs_x"], out[f"{tag}_lbfgs_fun"] = lb["x"], lb["fun"]
        print(tag, bx, bv, bt, loss.item())
    np.savez_compressed(os.path.join(HERE, "hparam_search.npz"), **out)


if __name__ == "__main__":
    main()
