"""Runs one of the decoder's GEMMs at M = 22059 packed tokens a few times (target of ncu captures):
gate_up (default; SwiGLU epilogue, the dominant kernel), down (K = 14336, residual in place), qkv, o."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from slime_b200 import _lib as L
lib = L.load()
which = sys.argv[1] if len(sys.argv) > 1 else "gate_up"
M = 22059
N, K = {"gate_up": (28672, 4096), "down": (4096, 14336), "qkv": (6144, 4096), "o": (4096, 4096)}[which]
a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
w = (torch.randn(N, K, device="cuda") * 0.02).to(torch.bfloat16)
swiglu = which == "gate_up"
out = torch.zeros(M, N // 2 if swiglu else N, device="cuda", dtype=torch.bfloat16)
res = L.ptr(out) if which in ("down", "o") else None  # residual in place, as the decoder runs them


def run():
    return lib.slime_op_gemm(L.ptr(a), K, L.ptr(w), K, M, N, K, None, res, out.shape[1] if res else 0, 0, None,
                             L.EPI_SWIGLU if swiglu else 0, L.ptr(out), None, out.shape[1], L.stream_ptr())


for _ in range(6):
    rc = run()
    assert rc == 0, L.last_error()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"{which} {M}x{N}x{K}: {ms:.3f} ms = {2*M*N*K/ms/1e9:.0f} TF/s (group rows env {os.environ.get('SLIME_GEMM_GROUP_ROWS','default')})")
