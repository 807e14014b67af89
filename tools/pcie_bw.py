#!/usr/bin/env python
"""Host<->device copy bandwidth of this box with pinned memory (what bounds the end-to-end path):
H2D alone, D2H alone, both directions at once; with and without binding the process to the GPU's NUMA node."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

def run(tag):
    n = 256 << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(n, dtype=torch.uint8, device="cuda"); d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def t(fn, reps=5):
        fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(reps): fn()
        torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
    def h2d():
        with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
    def d2h():
        with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
    def both():
        h2d(); d2h()
    r = {"h2d_GBps": n / t(h2d) / 1e9, "d2h_GBps": n / t(d2h) / 1e9}
    tb = t(both); r["both_each_GBps"] = n / tb / 1e9
    return {tag: r}
out = run("unbound")
out["numa"] = bench.bind_to_gpu_numa_node(torch, 0)
out.update(run("bound_to_gpu_node"))
print(json.dumps(out))
