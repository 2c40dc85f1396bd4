"""PCIe copy bandwidth on the GPU box: H2D alone, D2H alone, both at once on two streams (pinned host memory)."""
import torch, time
n = 192 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
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
    h2d(); d2h()


def chunks(k=8):
    c = n // k
    for i in range(k):
        with torch.cuda.stream(s1):
            d_a[i * c:(i + 1) * c].copy_(h_in[i * c:(i + 1) * c], non_blocking=True)
        with torch.cuda.stream(s2):
            h_out[i * c:(i + 1) * c].copy_(d_b[i * c:(i + 1) * c], non_blocking=True)


for name, fn, nbytes in (("H2D", h2d, n), ("D2H", d2h, n), ("H2D+D2H concurrent", both, 2 * n), ("8 chunks each way", chunks, 2 * n)):
    t = timed(fn)
    print(f"{name:22s} {t * 1e3:7.2f} ms  {nbytes / t / 1e9:6.1f} GB/s total")
