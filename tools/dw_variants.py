import torch, time
dev = torch.device("cuda:0")
n, d = 200000, 300
dH = torch.randn(n, d, device=dev); x = torch.randn(n, d, device=dev)
ref = (dH.double().t() @ x.double())
def t(f, name):
    for _ in range(3): r = f()
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): r = f()
    b.record(); torch.cuda.synchronize()
    err = float((r.double() - ref).abs().max() / ref.abs().max())
    print("%-34s %.3f ms  relerr %.1e" % (name, a.elapsed_time(b) / 10, err))
t(lambda: dH.t() @ x, "dH.t() @ x")
t(lambda: (x.t() @ dH).t(), "(x.t() @ dH).t()")
for B in (8, 16, 32, 64, 128):
    if n % B == 0:
        t(lambda: torch.bmm(dH.view(B, n // B, d).transpose(1, 2), x.view(B, n // B, d)).sum(0), "bmm split-K B=%d" % B)
t(lambda: torch.einsum("no,ni->oi", dH, x), "einsum")
