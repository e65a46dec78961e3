"""Per-tile clock64() trace of the tcgen05 attention kernel (CTA 0, first 64 tiles): where does a tile's time go?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from slime_b200 import _lib as L
lib = L.load()
L_, B, h, kvh, d = 1380, 16, 32, 8, 128
W = (h + 2 * kvh) * d
qkv = torch.randn(B * L_, W, device="cuda").to(torch.bfloat16)
o = torch.empty(B * L_, h * d, device="cuda", dtype=torch.bfloat16)
cu = torch.arange(0, (B + 1) * L_, L_, device="cuda", dtype=torch.int32)
args = (L.ptr(qkv), L.ptr(qkv[:, h * d:]), L.ptr(qkv[:, (h + kvh) * d:]), L.ptr(o), W, W, W, h * d, L.ptr(cu), L.ptr(cu), L_, L_,
        0, 0, 0, B, h, kvh, d, d ** -0.5, 1, B * L_, B * L_, 2, L.stream_ptr())
for _ in range(3):
    lib.slime_op_attention(*args)
torch.cuda.synchronize()
tr = torch.zeros(64, 16, dtype=torch.int64, device="cuda")
lib.slime_attention_set_trace(L.ptr(tr))
lib.slime_op_attention(*args)
torch.cuda.synchronize()
lib.slime_attention_set_trace(None)
t = tr.cpu()
t0 = int(t[0, 0])
names = ["sm:wait_s", "sm:s_ready", "sm:ld_done", "sm:max_xchg", "sm:exp_done", "sm:st_done", "", "", "mma:wait_p", "mma:p_ready", "mma:v_ready", "mma:pv_issued"]
print("tile | " + " | ".join(f"{n:>12s}" for n in names if n))
for g in range(40):
    row = [int(t[g, k]) - t0 if int(t[g, k]) else -1 for k in range(12)]
    print(f"{g:4d} | " + " | ".join(f"{row[k]:12d}" for k in range(12) if names[k]))
print("per-tile deltas (softmax thread): wait_s->s_ready, ->ld_done, ->max_xchg, ->exp_done, ->st_done, next wait_s")
for g in range(2, 30):
    a = [int(t[g, k]) for k in range(6)]
    nxt = int(t[g + 1, 0])
    print(g, a[1] - a[0], a[2] - a[1], a[3] - a[2], a[4] - a[3], a[5] - a[4], nxt - a[5], "| mma: p_ready-wait", int(t[g, 9]) - int(t[g, 8]),
          "v", int(t[g, 10]) - int(t[g, 9]), "issue", int(t[g, 11]) - int(t[g, 10]), "| period", int(t[g + 1, 1]) - a[1])
print("deferred epilogues (row = first tile of the next item): start->o_done wait, ->first tcgen05.ld, ->first chunk stored, ->end")
for g in range(1, 40):
    if int(t[g, 6]):
        e = [int(t[g, k]) for k in (6, 7, 12, 13, 14)]
        print(g, e[1] - e[0], e[2] - e[1], e[3] - e[2], e[4] - e[3], "total", e[4] - e[0], "| st_done->epi start", e[0] - int(t[g, 5]))
