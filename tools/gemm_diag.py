"""First-light diagnostics for the tcgen05 GEMM: tiny problems, block-wise error maps."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from slime_b200 import _lib as L

lib = L.load()
torch.manual_seed(0)


def run(M, N, K, a=None, w=None):
    a = (torch.randn(M, K, device="cuda")).to(torch.bfloat16) if a is None else a
    w = (torch.randn(N, K, device="cuda")).to(torch.bfloat16) if w is None else w
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    rc = lib.slime_op_gemm(L.ptr(a), K, L.ptr(w), K, M, N, K, None, None, 0, 0, None, 0, L.ptr(out), None, N,
                           L.stream_ptr())
    if rc != 0:
        print(f"[{M}x{N}x{K}] launch failed rc={rc}: {L.last_error()}")
        return None
    try:
        torch.cuda.synchronize()
    except Exception as e:  # noqa
        print(f"[{M}x{N}x{K}] CUDA error: {e}")
        raise
    ref = a.float() @ w.float().t()
    o = out.float()
    nan = torch.isnan(o).float().mean().item()
    o = torch.nan_to_num(o)
    err = ((o - ref).norm() / ref.norm()).item()
    print(f"[{M}x{N}x{K}] rel-L2 {err:.3e}  nan-frac {nan:.3f}")
    if err > 1e-2:
        bm, bn = 32, 32
        print("  block rel-err map (rows = 32-row blocks, cols = 32-col blocks):")
        for i in range(0, min(M, 256), bm):
            row = []
            for j in range(0, min(N, 512), bn):
                r = ref[i:i + bm, j:j + bn]
                d = o[i:i + bm, j:j + bn] - r
                row.append(f"{(d.norm() / r.norm().clamp_min(1e-9)).item():5.2f}")
            print("   ", " ".join(row))
        # hypotheses
        for kk in (16, 32, 48):
            if kk < K:
                part = a[:, :kk].float() @ w[:, :kk].float().t()
                print(f"  vs first-{kk}-of-K partial: {((o - part).norm() / part.norm()).item():.3e}")
        print("  vs transposed:", ((o[:min(M, N), :min(M, N)] - ref[:min(M, N), :min(M, N)].t()).norm() / ref.norm()).item())
        print("  out[0,:8]", o[0, :8].tolist())
        print("  ref[0,:8]", ref[0, :8].tolist())
    return err


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    for shape in [(128, 128, 16), (128, 128, 64), (128, 128, 128), (128, 256, 64), (256, 128, 64), (128, 128, 512),
                  (64, 64, 64), (1000, 1000, 1000 // 8 * 8), (4096, 4096, 4096)]:
        run(*shape)
    # throughput first look
    for (M, N, K) in [(8192, 8192, 8192), (11264, 6144, 4096), (11264, 28672, 4096), (11264, 4096, 14336), (2885, 1024, 1024), (2885, 4096, 1024)]:
        a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        w = torch.randn(N, K, device="cuda").to(torch.bfloat16)
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        def f():
            lib.slime_op_gemm(L.ptr(a), K, L.ptr(w), K, M, N, K, None, None, 0, 0, None, 0, L.ptr(out), None, N, L.stream_ptr())
        for _ in range(3): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): f()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        e0.record()
        for _ in range(10): torch.matmul(a, w.t())
        e1.record(); torch.cuda.synchronize()
        ms_t = e0.elapsed_time(e1) / 10
        print(f"perf {M}x{N}x{K}: ours {ms:.3f} ms = {2*M*N*K/ms/1e9:.0f} TF/s | cuBLAS {ms_t:.3f} ms = {2*M*N*K/ms_t/1e9:.0f} TF/s")
