"""First light of the 2-CTA GEMM: correctness with block error maps, then throughput vs the 1-CTA kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from slime_b200 import _lib as L

lib = L.load()


def run(M, N, K, mode):
    lib.slime_gemm_set_2cta_mode(mode)
    torch.manual_seed(0)
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    rc = lib.slime_op_gemm(L.ptr(a), K, L.ptr(w), K, M, N, K, None, None, 0, 0, None, 0, L.ptr(out), None, N, L.stream_ptr())
    if rc != 0:
        print("launch failed", L.last_error()); return
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t()
    o = torch.nan_to_num(out.float())
    err = ((o - ref).norm() / ref.norm()).item()
    print(f"[mode {mode}] {M}x{N}x{K}: rel-L2 {err:.3e} nan-frac {torch.isnan(out.float()).float().mean().item():.3f}")
    if err > 1e-2:
        for i in range(0, min(M, 256), 64):
            row = []
            for j in range(0, min(N, 256), 64):
                r = ref[i:i + 64, j:j + 64]
                row.append(f"{((o[i:i+64, j:j+64] - r).norm() / r.norm()).item():5.2f}")
            print("   rows", i, " ".join(row))


if __name__ == "__main__":
    for shape in [(256, 256, 64), (256, 256, 256), (512, 512, 128), (300, 520, 200), (4096, 4096, 4096)]:
        run(*shape, 1)
    for (M, N, K) in [(8192, 8192, 8192), (22059, 6144, 4096), (22059, 28672, 4096), (22059, 4096, 14336), (22059, 4096, 4096), (46160, 3072, 1024), (46160, 4096, 1024)]:
        a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        w = torch.randn(N, K, device="cuda").to(torch.bfloat16)
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        res = {}
        for mode in (0, 1):
            lib.slime_gemm_set_2cta_mode(mode)
            f = lambda: lib.slime_op_gemm(L.ptr(a), K, L.ptr(w), K, M, N, K, None, None, 0, 0, None, 0, L.ptr(out), None, N, L.stream_ptr())
            for _ in range(3): f()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): f()
            e1.record(); torch.cuda.synchronize()
            res[mode] = e0.elapsed_time(e1) / 20
        e0.record()
        for _ in range(20): torch.matmul(a, w.t())
        e1.record(); torch.cuda.synchronize()
        mt = e0.elapsed_time(e1) / 20
        fl = 2 * M * N * K / 1e9
        print(f"perf {M}x{N}x{K}: 1cta {res[0]:.3f} ms = {fl/res[0]:.0f} TF/s | 2cta {res[1]:.3f} ms = {fl/res[1]:.0f} TF/s | cuBLAS {mt:.3f} ms = {fl/mt:.0f} TF/s")
    lib.slime_gemm_set_2cta_mode(-1)
