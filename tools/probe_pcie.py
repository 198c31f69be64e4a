"""PCIe probe: pinned H2D / D2H of 4 GiB, alone and concurrently."""
import torch
n = 32768
h1 = torch.ones((n, n), dtype=torch.float32).pin_memory()
h2 = torch.empty((n, n), dtype=torch.float32).pin_memory()
d1 = torch.empty((n, n), dtype=torch.float32, device="cuda")
d2 = torch.zeros((n, n), dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timed(fn, label):
    for rep in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"{label} rep{rep}: {ms:.1f} ms  {h1.numel() * 4 / ms / 1e6:.1f} GB/s per direction", flush=True)
timed(lambda: d1.copy_(h1, non_blocking=True), "H2D 4 GiB")
timed(lambda: h2.copy_(d2, non_blocking=True), "D2H 4 GiB")
def both():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    cur.wait_stream(s1); cur.wait_stream(s2)
timed(both, "H2D + D2H concurrently")
