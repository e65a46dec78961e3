"""Runs the dominant GEMM (Llama-3 gate/up with the SwiGLU epilogue at M = 22059 packed tokens) a few times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from slime_b200 import _lib as L
lib = L.load()
M, N, K = 22059, 28672, 4096
a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
w = (torch.randn(N, K, device="cuda") * 0.02).to(torch.bfloat16)
out = torch.empty(M, N // 2, device="cuda", dtype=torch.bfloat16)
for _ in range(6):
    rc = lib.slime_op_gemm(L.ptr(a), K, L.ptr(w), K, M, N, K, None, None, 0, 0, None, L.EPI_SWIGLU, L.ptr(out), None, N // 2, L.stream_ptr())
    assert rc == 0, L.last_error()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    lib.slime_op_gemm(L.ptr(a), K, L.ptr(w), K, M, N, K, None, None, 0, 0, None, L.EPI_SWIGLU, L.ptr(out), None, N // 2, L.stream_ptr())
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"gate_up {M}x{N}x{K}: {ms:.3f} ms = {2*M*N*K/ms/1e9:.0f} TF/s (group rows env {os.environ.get('SLIME_GEMM_GROUP_ROWS','default')})")
