"""Dev tool: pinned-memory H2D / D2H bandwidth of the box, alone and concurrently (the floor of the e2e metric)."""
import torch, time
dev = torch.device("cuda:0")
n = 256 << 20
h_a = torch.empty(n, dtype=torch.uint8).pin_memory()
h_b = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device=dev)
d_b = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
def h2d():
    with torch.cuda.stream(s1): d_a.copy_(h_a, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_b.copy_(d_b, non_blocking=True)
def both():
    h2d(); d2h()
print(f"H2D {n/t(h2d)/1e9:.1f} GB/s  D2H {n/t(d2h)/1e9:.1f} GB/s  concurrent each {n/t(both)/1e9:.1f} GB/s")
# many small pieces like the composer's per-image copies (36 MB each)
m = 36 << 20
def h2d_pieces():
    with torch.cuda.stream(s1):
        for i in range(7): d_a[i*m:(i+1)*m].copy_(h_a[i*m:(i+1)*m], non_blocking=True)
print(f"H2D in 36 MB pieces {7*m/t(h2d_pieces)/1e9:.1f} GB/s")
