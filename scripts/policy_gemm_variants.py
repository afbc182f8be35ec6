"""One-off timing of ways to run the policy's encoder + shared layers on bfloat16 inputs (B200, CUDA events).

A: two encoder GEMMs (bias + ReLU epilogue), shared layer as addmm + addmm_ + relu_ (what PolicyMLP does);
B: the encoders write the two halves of one [n, 512] buffer (out= views), the shared layer is ONE fused GEMM;
C: one block-diagonal encoder GEMM over a combined [n, 640] input row, then the fused shared GEMM.
"""
import torch

n, hidden = 65536, 256
dev, bf = "cuda", torch.bfloat16
torch.manual_seed(0)
xm = (torch.rand(n, 512, device=dev) < 0.02).to(bf)
xf = torch.randn(n, 128, device=dev).to(bf)
wm, wf = torch.randn(512, hidden, device=dev).to(bf) * 0.05, torch.randn(128, hidden, device=dev).to(bf) * 0.05
bm, bfe, bs = (torch.randn(hidden, device=dev).to(bf) for _ in range(3))
ws = torch.randn(2 * hidden, hidden, device=dev).to(bf) * 0.05
ws_m, ws_f = ws[:hidden].contiguous(), ws[hidden:].contiguous()
fused = torch._addmm_activation


def a():
    hm, hf = fused(bm, xm, wm), fused(bfe, xf, wf)
    return torch.addmm(bs, hm, ws_m).addmm_(hf, ws_f).relu_()


cat = torch.empty(n, 2 * hidden, dtype=bf, device=dev)
b_cat = torch.cat([bm, bfe])


def b():
    fused(bm, xm, wm, out=cat[:, :hidden])
    fused(bfe, xf, wf, out=cat[:, hidden:])
    return fused(bs, cat, ws)


x_all = torch.cat([xm, xf], dim=1).contiguous()
w_diag = torch.zeros(640, 2 * hidden, dtype=bf, device=dev)
w_diag[:512, :hidden], w_diag[512:, hidden:] = wm, wf


def c():
    return fused(bs, fused(b_cat, x_all, w_diag), ws)


def timeit(fn, iters=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


ref = a().float()
for name, fn in (("A", a), ("B", b), ("C", c)):
    try:
        err = (fn().float() - ref).abs().max().item()
        print(f"{name}: {timeit(fn):7.1f} us per step, max |diff| vs A {err:.3g}")
    except Exception as exc:  # noqa: BLE001
        print(f"{name}: failed: {exc}")
