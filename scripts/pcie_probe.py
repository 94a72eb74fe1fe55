"""Host<->device copy rates of this box with pinned memory (the ceiling of the e2e number of bench.py):
H2D alone, D2H alone, and both directions at once on two streams.  Prints one JSON line."""
import json
import time

import torch

n = 2_150_000_000 // 8
h_in = torch.empty(n, dtype=torch.float64).pin_memory()
h_out = torch.empty(n, dtype=torch.float64).pin_memory()
d_a = torch.empty(n, dtype=torch.float64, device='cuda')
d_b = torch.empty(n, dtype=torch.float64, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


def both():
    h2d()
    d2h()


gb = n * 8 / 1e9
out = {'bytes_each_way': n * 8, 'h2d_GBs': gb / timed(h2d), 'd2h_GBs': gb / timed(d2h)}
t = timed(both)
out['bidir_ms'] = 1e3 * t
out['bidir_GBs_each_way'] = gb / t
print(json.dumps(out))
