"""ViT-shape GEMM (fc1 with quick-GELU: M = 80 crops x 577, N = 4096, K = 1024) for ncu + timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from slime_b200 import _lib as L
lib = L.load()
M, N, K = 46160, 4096, 1024
a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
w = (torch.randn(N, K, device="cuda") * 0.02).to(torch.bfloat16)
b = torch.randn(N, device="cuda").to(torch.bfloat16)
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
f = lambda: lib.slime_op_gemm(L.ptr(a), K, L.ptr(w), K, M, N, K, L.ptr(b), None, 0, 0, None, L.EPI_QUICK_GELU, L.ptr(out), None, N, L.stream_ptr())
for _ in range(6): f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): f()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"vit fc1 {M}x{N}x{K}: {ms:.3f} ms = {2*M*N*K/ms/1e9:.0f} TF/s")
