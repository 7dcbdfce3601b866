import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussianip_b200 import _lib
lib = _lib.load(); dev = torch.device("cuda", 0); st = torch.cuda.current_stream().cuda_stream
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
keys = torch.randint(0, 2 ** 31 - 1, (n,), device=dev, dtype=torch.int64).to(torch.int32)
vals = torch.arange(n, device=dev, dtype=torch.int32)
ko, vo = torch.empty_like(keys), torch.empty_like(vals)
tmp = torch.empty(lib.gsb_radix_tmp_bytes(n, 4), dtype=torch.uint8, device=dev)
for _ in range(3):
    lib.gsb_radix_sort_pairs_u32(n, keys.data_ptr(), vals.data_ptr(), ko.data_ptr(), vo.data_ptr(), 32, tmp.data_ptr(), st)
torch.cuda.synchronize()
