import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from waveflow_b200 import _ffi
from waveflow_b200._ffi import lib, ptr, stream_ptr, check
dev = torch.device("cuda:0")

def split(x):
    hi = torch.empty_like(x); lo = torch.empty_like(x)
    check(lib.wf_tf32_split(ptr(x), x.numel(), ptr(hi), ptr(lo), stream_ptr()))
    return hi, lo

def dense(a_hi, a_lo, w_hi, w_lo, bias, mode):
    M, K = a_hi.shape; N = w_hi.shape[0]
    oh = torch.empty(M, N, device=dev); ol = torch.empty(M, N, device=dev) if mode == 1 else None
    check(lib.wf_tc_dense(ptr(a_hi), ptr(a_lo), M, K, ptr(w_hi), ptr(w_lo), N, ptr(bias), mode, ptr(oh), ptr(ol), stream_ptr()))
    return oh, ol

g = torch.Generator(device=dev); g.manual_seed(0)
for (M, K, N) in [(128, 32, 128), (128, 64, 256), (300, 512, 512), (1000, 512, 192 * 3), (4096, 512, 6144)]:
    A = torch.randn(M, K, device=dev, generator=g); W = torch.randn(N, K, device=dev, generator=g) / K ** 0.5
    b = torch.randn(N, device=dev, generator=g)
    ah, al = split(A); wh, wl = split(W)
    assert (A - ah - al).abs().max() < 3e-7 * A.abs().max()
    out, _ = dense(ah, al, wh, wl, b, 0)
    torch.cuda.synchronize()
    ref = (A.double() @ W.double().T + b.double())
    err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
    ref32 = (A @ W.T + b)
    e32 = (ref32.double() - ref).abs().max().item() / ref.abs().max().item()
    print(f"M={M} K={K} N={N}: rel err 3xTF32 {err:.2e}   torch fp32 (TF32 off) {e32:.2e}")
    oh, ol = dense(ah, al, wh, wl, b, 1)
    t = torch.tanh(ref)
    print("   tanh+split err", ((oh + ol).double() - t).abs().max().item())
# timing: config-5 layer shapes
M = 1 << 16
for (K, N) in [(512, 512), (512, 6144)]:
    A = torch.randn(M, K, device=dev, generator=g); W = torch.randn(N, K, device=dev, generator=g) / K ** 0.5
    ah, al = split(A); wh, wl = split(W)
    oh = torch.empty(M, N, device=dev)
    for _ in range(2):
        check(lib.wf_tc_dense(ptr(ah), ptr(al), M, K, ptr(wh), ptr(wl), N, None, 0, ptr(oh), None, stream_ptr()))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        check(lib.wf_tc_dense(ptr(ah), ptr(al), M, K, ptr(wh), ptr(wl), N, None, 0, ptr(oh), None, stream_ptr()))
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print(f"M={M} K={K} N={N}: {ms:.3f} ms  {2*M*K*N/ms/1e9:.1f} TFLOP/s (useful fp32-grade), {3*2*M*K*N/ms/1e9:.1f} TF32 TFLOP/s issued")
