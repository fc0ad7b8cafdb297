import torch, time
n = 530841600
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
h_in = torch.empty(302096360, dtype=torch.uint8).pin_memory()
d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
d_in = torch.empty(302096360, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best
def d2h():
    with torch.cuda.stream(s1): h_out.copy_(d_out, non_blocking=True)
def h2d():
    with torch.cuda.stream(s2): d_in.copy_(h_in, non_blocking=True)
def both():
    d2h(); h2d()
def d2h_chunks(k=8):
    with torch.cuda.stream(s1):
        c = n // k
        for i in range(k): h_out[i*c:(i+1)*c].copy_(d_out[i*c:(i+1)*c], non_blocking=True)
a = t(d2h); b = t(h2d); c = t(both); d = t(d2h_chunks)
print("D2H alone %.2f ms %.1f GB/s | H2D alone %.2f ms %.1f GB/s | both %.2f ms | D2H 8 chunks %.2f ms" % (a*1e3, n/a/1e9, b*1e3, 302096360/b/1e9, c*1e3, d*1e3))
