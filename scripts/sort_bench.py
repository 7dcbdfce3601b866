"""Sort keys/s of the hand-written radix sort (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussianip_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
for key_bytes, end_bit in ((4, 32), (4, 12), (8, 45)):
    for n in (1 << 10, 1 << 17, 1 << 20, 1 << 21, 1 << 23, 1 << 25):
        kd = torch.int32 if key_bytes == 4 else torch.int64
        keys = torch.randint(0, 2 ** 31 - 1, (n,), device=dev, dtype=torch.int64)
        if end_bit < 63:
            keys = keys & ((1 << min(end_bit, 62)) - 1)
        keys = keys.to(kd)
        vals = torch.arange(n, device=dev, dtype=torch.int32)
        ko, vo = torch.empty_like(keys), torch.empty_like(vals)
        tmp = torch.empty(lib.gsb_radix_tmp_bytes(n, key_bytes), dtype=torch.uint8, device=dev)
        def run():
            _lib.check(lib.gsb_radix_sort_pairs_u32(n, keys.data_ptr(), vals.data_ptr(), ko.data_ptr(), vo.data_ptr(), end_bit, tmp.data_ptr(), st) if key_bytes == 4 else
                       lib.gsb_radix_sort_pairs_u64(n, keys.data_ptr(), vals.data_ptr(), ko.data_ptr(), vo.data_ptr(), end_bit, tmp.data_ptr(), st), "sort")
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        it = 20
        e0.record()
        for _ in range(it): run()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / it * 1e3
        ref = torch.sort(keys.to(torch.int64), stable=True)
        ok = bool(torch.equal(ko.to(torch.int64), ref.values)) and bool(torch.equal(vo.long(), ref.indices))
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(it): torch.sort(keys, stable=True)
        t1.record(); torch.cuda.synchronize()
        print(f"key{key_bytes*8} bits{end_bit:2d} n={n:9d}  {us:9.1f} us  {n/us:8.1f} Mkeys/s  passes={(end_bit+7)//8}  correct={ok}   torch.sort {t0.elapsed_time(t1)/it*1e3:9.1f} us")
