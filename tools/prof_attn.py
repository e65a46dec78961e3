"""Runs the decoder-shape (and ViT-shape) tcgen05 attention a few times (target of ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from slime_b200 import _lib as L
lib = L.load()
which = sys.argv[1] if len(sys.argv) > 1 else "decoder"
IMPL = 0
if which == "decoder":
    L_, B, h, kvh, d = 1380, 16, 32, 8, 128
    W = (h + 2 * kvh) * d
    qkv = torch.randn(B * L_, W, device="cuda").to(torch.bfloat16)
    o = torch.empty(B * L_, h * d, device="cuda", dtype=torch.bfloat16)
    cu = torch.arange(0, (B + 1) * L_, L_, device="cuda", dtype=torch.int32)
    args = (L.ptr(qkv), L.ptr(qkv[:, h * d:]), L.ptr(qkv[:, (h + kvh) * d:]), L.ptr(o), W, W, W, h * d, L.ptr(cu), L.ptr(cu), L_, L_,
            0, 0, 0, B, h, kvh, d, d ** -0.5, 1, B * L_, B * L_, IMPL, L.stream_ptr())
else:
    S, B, h, d = 577, 80, 16, 64
    D = h * d
    qkv = torch.randn(B * S, 3 * D, device="cuda").to(torch.bfloat16)
    o = torch.empty(B * S, D, device="cuda", dtype=torch.bfloat16)
    args = (L.ptr(qkv), L.ptr(qkv[:, D:]), L.ptr(qkv[:, 2 * D:]), L.ptr(o), 3 * D, 3 * D, 3 * D, D, None, None, S, S, S, S, S, B, h, h, d,
            d ** -0.5, 0, 0, 0, IMPL, L.stream_ptr())
for _ in range(6):
    rc = lib.slime_op_attention(*args)
    assert rc == 0, L.last_error()
torch.cuda.synchronize()
print("done")
