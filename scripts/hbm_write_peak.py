"""Micro-benchmark: what a pure-write (store-only) stream reaches on this GPU next to the copy figure of
MEASURED_PEAKS.json.  K2 is store-only, so this is the practical ceiling its roofline fraction should be read against."""
import json
import torch

def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best

nbytes = 16 * 2**30
x = torch.empty(nbytes // 8, dtype=torch.float64, device="cuda")
y = torch.empty(nbytes // 8, dtype=torch.float64, device="cuda")
out = {}
out["fill_f64_gbs"] = nbytes / timeit(lambda: x.fill_(1.5)) / 1e6
out["memset_gbs"] = nbytes / timeit(lambda: x.zero_()) / 1e6
out["copy_gbs_read_plus_write"] = 2 * nbytes / timeit(lambda: y.copy_(x)) / 1e6
out["read_sum_gbs"] = nbytes / timeit(lambda: x.sum()) / 1e6
out["axpy_gbs_2r1w"] = 3 * nbytes / timeit(lambda: torch.add(x, y, out=y)) / 1e6
print(json.dumps(out))
