"""How many query rows lose the exactness certificate of the fp16 tensor-core pass, and how many are left after the
split-precision second pass, on hard distributions at production scale.
    python tools/uncert_rate.py [M] [D] [NQ]
Distributions: iid-Gaussian unit vectors (SURVEY.md 8d worst case), narrow-cone embeddings (a shared component
compresses every similarity and gap by 1 - shared), and a degenerate one (random-init encoder: all rows within 1e-3)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lemon_b200
from lemon_b200.scoring import _slice_prepared, count_uncertified

m = int(sys.argv[1]) if len(sys.argv) > 1 else 3_300_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 768
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 16_384
dev = torch.device("cuda", 0)
sc = lemon_b200.get_scorer(0)
g = torch.Generator(device=dev).manual_seed(1)
nrm = torch.nn.functional.normalize


def cone(shared):
    c0 = nrm(torch.randn(1, d, generator=g, device=dev), dim=1)
    out = torch.empty((m, d), device=dev)
    for s in range(0, m, 1 << 18):
        z = nrm(torch.randn(min(1 << 18, m - s), d, generator=g, device=dev), dim=1)
        out[s:s + z.shape[0]] = nrm(shared ** 0.5 * c0 + (1 - shared) ** 0.5 * z, dim=1)
    return out


res = []
shares = [float(a) for a in sys.argv[4:]] or [0.0, 0.7, 0.9, 0.98, 0.9999]
for shared in shares:
    name = "iid-gaussian" if shared == 0 else ("degenerate (cone %g)" % shared if shared > 0.999 else "cone %g" % shared)
    x = cone(shared)
    dbp = sc.prepare(x, True)
    del x
    qp = _slice_prepared(dbp, 0, nq)
    row = {"distribution": name, "db_rows": m, "dim": d, "queries": nq}
    for second in (False, True):
        sc.second_pass_enabled = second
        sc.knn(qp, dbp, 31, 0, mode="tc")
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sc.knn(qp, dbp, 31, 0, mode="tc")
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        info = sc.last_info
        row["first_pass_uncertified"] = info["n_uncertified_first_pass"]
        row["with_second_pass" if second else "without_second_pass"] = {"to_exact_kernel": count_uncertified(info), "ms": dt * 1e3}
    sc.second_pass_enabled = True
    res.append(row)
    print(json.dumps(row), flush=True)
    del dbp, qp
