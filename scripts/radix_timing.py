"""Diagnostic (needs a library built with -DGSB_RADIX_TIMING): where one radix pass spends its time."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from gaussianip_b200 import _lib
lib = _lib.load(); dev = torch.device("cuda", 0); st = torch.cuda.current_stream().cuda_stream
raw = C.CDLL(str(_lib.LIB_PATH))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
bits = int(sys.argv[2]) if len(sys.argv) > 2 else 32
keys = (torch.randint(0, 2 ** 31 - 1, (n,), device=dev, dtype=torch.int64) & ((1 << bits) - 1)).to(torch.int32)
vals = torch.arange(n, device=dev, dtype=torch.int32)
ko, vo = torch.empty_like(keys), torch.empty_like(vals)
tmp = torch.empty(lib.gsb_radix_tmp_bytes(n, 4), dtype=torch.uint8, device=dev)
for _ in range(4):
    lib.gsb_radix_sort_pairs_u32(n, keys.data_ptr(), vals.data_ptr(), ko.data_ptr(), vo.data_ptr(), bits, tmp.data_ptr(), st)
torch.cuda.synchronize()
nb = (n + 2047) // 2048
buf = np.zeros(8 * 8192, dtype=np.int64)
raw.gsb_debug_radix_timing.argtypes = [C.c_void_p, C.c_int]
rc = raw.gsb_debug_radix_timing(buf.ctypes.data, buf.size)
t = buf.reshape(8192, 8)[:nb]
names = ["load+hist", "rank", "barrier", "scan+scatter", "lookback", "writeout"]
d = np.diff(t[:, :7], axis=1) / 1.965e3          # us at 1965 MHz
print(f"n={n} blocks={nb} rc={rc}; last pass of the sort; per-block phase durations (us): mean / p50 / max")
for i, nm in enumerate(names):
    print(f"  {nm:14s} {d[:, i].mean():7.2f} {np.median(d[:, i]):7.2f} {d[:, i].max():7.2f}")
tot = (t[:, 6] - t[:, 0]) / 1.965e3
print(f"  {'block total':14s} {tot.mean():7.2f} {np.median(tot):7.2f} {tot.max():7.2f}")
g0 = t[:, 7] - t[:, 7].min()
print(f"  block start skew (globaltimer ns): p50 {np.median(g0):.0f} max {g0.max():.0f}")
for lo in range(0, nb, max(1, nb // 8)):
    sl = slice(lo, min(nb, lo + max(1, nb // 8)))
    print(f"  blocks {lo:5d}+: lookback mean {d[sl, 4].mean():6.2f}  total {tot[sl].mean():6.2f}  start {np.median(g0[sl]):8.0f} ns")
