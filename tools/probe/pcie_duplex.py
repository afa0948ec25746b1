import torch, time
n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=4, chunk=None):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps):
        if chunk is None:
            if h2d:
                with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
        else:
            for o in range(0, n, chunk):
                if h2d:
                    with torch.cuda.stream(s1): d_in[o:o+chunk].copy_(h_in[o:o+chunk], non_blocking=True)
                if d2h:
                    with torch.cuda.stream(s2): h_out[o:o+chunk].copy_(d_out[o:o+chunk], non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    return reps * n / dt / 1e9
run(True, True, 1)
print("h2d only %.1f GB/s" % run(True, False))
print("d2h only %.1f GB/s" % run(False, True))
print("both     %.1f GB/s each" % run(True, True))
print("both, 200 MB chunks %.1f GB/s each" % run(True, True, chunk=200 << 20))
print("both, 50 MB chunks %.1f GB/s each" % run(True, True, chunk=50 << 20))
