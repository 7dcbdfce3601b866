"""Times distCUDA2 (csrc/knn.cu) against the reference's own simple-knn build (oracle/_ref) on one GPU.
python scripts/knn_bench.py [P ...]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from gaussianip_b200 import synthetic  # noqa: E402
from gaussianip_b200.knn import distCUDA2  # noqa: E402
from oracle import knn as oknn  # noqa: E402  (measurement script: reference arm only)


def timed(fn, n=10):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    dev = torch.device("cuda", 0)
    sizes = [int(x) for x in sys.argv[1:]] or [100_000, 1_000_000, 3_000_000]
    for P in sizes:
        pts = synthetic.make_cloud(P, sh_degree=0).xyz.detach().float().contiguous().to(dev)
        ms = timed(lambda: distCUDA2(pts))
        line = {"P": P, "distCUDA2_ms": round(ms, 3)}
        if oknn.reference_available():
            t = time.perf_counter()
            for _ in range(3):
                ref = oknn.reference_dist2(pts)
            line["reference_simple_knn_ms"] = round((time.perf_counter() - t) / 3 * 1e3, 3)   # it synchronises itself
            line["bit_equal"] = bool(torch.equal(ref, distCUDA2(pts)))
        print(line, flush=True)


if __name__ == "__main__":
    main()
