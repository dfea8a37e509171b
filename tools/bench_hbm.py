#!/usr/bin/env python
"""HBM bandwidth by direction on this GPU: write-only (fill), read-only (sum), copy (read + write)."""
import json, torch
n = 1 << 30   # 4 GiB of fp32
x = torch.empty(n, dtype=torch.float32, device="cuda")
y = torch.empty(n, dtype=torch.float32, device="cuda")
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3
tw = t(lambda: x.fill_(1.0)); tr = t(lambda: x.sum()); tc = t(lambda: y.copy_(x))
print(json.dumps({"write_only_GBs": 4 * n / tw / 1e9, "read_only_GBs": 4 * n / tr / 1e9, "copy_GBs_total": 8 * n / tc / 1e9}))
