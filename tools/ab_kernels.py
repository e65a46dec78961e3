"""A/B on the GPU box: GEMM epilogue modes (direct vs staged) on the GEMM shapes of the path.  The staged output is first
compared with the direct one (bit-identical), then both are timed with CUDA events (L2 flushed between repetitions by
the working set itself: every shape moves > 126 MB).  (Attention: tools/attn_bench.py.)

    python tools/ab_kernels.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from slime_b200 import _lib as L

lib = L.load()
dev = "cuda"


def timeit(f, n=10, warm=3):
    for _ in range(warm):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / b.norm().clamp_min(1e-20))


def gemm_ab():
    torch.manual_seed(0)
    T = 22059  # decoder rows of the headline batch
    V = 80 * 577  # ViT rows
    shapes = [
        # name, M, N, K, epi, bias, residual(in place)
        ("vit qkv", V, 3072, 1024, L.EPI_NONE, True, False),
        ("vit o-proj+res", V, 1024, 1024, L.EPI_NONE, True, True),
        ("vit fc1 qgelu", V, 4096, 1024, L.EPI_QUICK_GELU, True, False),
        ("vit fc2+res", V, 1024, 4096, L.EPI_NONE, True, True),
        ("llm qkv", T, 6144, 4096, L.EPI_NONE, False, False),
        ("llm o-proj+res", T, 4096, 4096, L.EPI_NONE, False, True),
        ("llm gate_up swiglu", T, 28672, 4096, L.EPI_SWIGLU, False, False),
        ("llm down+res", T, 4096, 14336, L.EPI_NONE, False, True),
        ("proj fc1 gelu", 9216, 4096, 1024, L.EPI_GELU_ERF, True, False),
        ("ragged M, N=8008", 12345, 8008, 1024, L.EPI_NONE, True, True),
    ]
    for name, M, N, K, epi, use_bias, use_res in shapes:
        a = torch.randn(M, K, device=dev).to(torch.bfloat16)
        w = (torch.randn(N, K, device=dev) * K ** -0.5).to(torch.bfloat16)
        b = torch.randn(N, device=dev).to(torch.bfloat16) if use_bias else None
        n_out = N // 2 if epi == L.EPI_SWIGLU else N
        res0 = torch.randn(M, n_out, device=dev).to(torch.bfloat16) if use_res else None
        outs, times = [], []
        order = (1, 0) if os.environ.get("AB_REVERSE") else (0, 1)
        for mode in order:
            assert lib.slime_gemm_set_epi_mode(mode) == 0
            out = res0.clone() if use_res else torch.full((M, n_out), 7.0, device=dev, dtype=torch.bfloat16)

            def f():
                rc = lib.slime_op_gemm(L.ptr(a), K, L.ptr(w), K, M, N, K, L.ptr(b) if b is not None else None,
                                       L.ptr(out) if use_res else None, n_out if use_res else 0, 0, None, epi,
                                       L.ptr(out), None, n_out, L.stream_ptr())
                assert rc == 0, L.last_error()

            f()  # one application on the fresh residual: this is the output that is compared
            torch.cuda.synchronize()
            outs.append(out.clone())
            times.append(timeit(f, n=30))
        if order[0] == 1:
            outs.reverse()
            times.reverse()
        same = torch.equal(outs[0], outs[1])
        fl = 2.0 * M * N * K
        print(f"gemm {name:22s} M={M:6d} N={N:6d} K={K:6d}  direct {times[0]:7.3f} ms ({fl / times[0] / 1e9:6.0f} TF/s)  "
              f"staged {times[1]:7.3f} ms ({fl / times[1] / 1e9:6.0f} TF/s)  x{times[0] / times[1]:.3f}  "
              f"bit-identical={same} rel={rel(outs[1], outs[0]):.2e}", flush=True)
    lib.slime_gemm_set_epi_mode(0)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    gemm_ab()
