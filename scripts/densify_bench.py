"""Times one densify_and_prune / prune_only at P points: fused path (gaussianip_b200.densify) vs the reference's
tensor-op sequence (oracle/densify_torch.py, boolean indexing + torch.cat) on the same GPU."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from oracle import densify_torch as OD            # measurement script: reference-sequence arm
from tests.test_gpu_densify import big_state, make_model, run_product

dev = torch.device("cuda", 0)
for P, deg in ((1_000_000, 0), (1_000_000, 3), (3_000_000, 0)):
    st = big_state(P, deg, 1)
    for op in ("prune_only", "densify_and_prune"):
        res = {}
        for arm in ("reference-sequence", "fused"):
            ts = []
            for it in range(4):
                if arm == "fused":
                    m = make_model(st, fused=True)
                else:
                    s = {k: v.to(dev) for k, v in st.items()}
                torch.manual_seed(3)
                torch.cuda.synchronize(); t0 = time.perf_counter()
                if arm == "fused":
                    run_product(op, m, None)
                else:
                    OD.run_case(op, s)
                torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
            res[arm] = min(ts[1:]) * 1e3
        print({"P": P, "sh_degree": deg, "op": op, **{k: round(v, 3) for k, v in res.items()},
               "speedup": round(res["reference-sequence"] / res["fused"], 2)}, flush=True)
