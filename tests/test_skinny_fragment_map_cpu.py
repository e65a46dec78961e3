"""CPU model of the operand mapping of the weight-streaming decode GEMM (slime_b200/csrc/gemm_skinny.cu).

The kernel feeds mma.sync.m16n8k16 (row.col, fp32 accumulate) with the WEIGHT rows as the n = 8 operand and the
activation rows as the m = 16 operand, and it never shuffles or transposes anything: lane (g = lane / 4, c = lane % 4)
loads 16 contiguous bytes - k = kb + 8c .. kb + 8c + 7 - of weight row n0 + g and of activation rows g and g + 8, and hands
the eight values to TWO MMAs as the k-slots that lane owns in the PTX fragment layout.  That works because a dot product
does not care about the order of k as long as A and B use the same order.  This test spells the PTX fragment layout out
(PTX ISA, "Matrix Fragments for mma.m16n8k16") and checks, with exact integer arithmetic, that the mapping coded in the
kernel yields C = X W^T, including the accumulator fragment the epilogue reads (c0, c1 = row g, columns 2c, 2c + 1;
c2, c3 = row g + 8)."""
import numpy as np


def mma_m16n8k16(a_frag, b_frag):
    """a_frag[lane] = (a0..a7), b_frag[lane] = (b0..b3) as the PTX ISA distributes a 16x16 row-major A and a 16x8
    column-major B over the 32 lanes; returns c_frag[lane] = (c0..c3) of the 16x8 product."""
    A = np.zeros((16, 16), dtype=np.int64)
    B = np.zeros((16, 8), dtype=np.int64)
    for lane in range(32):
        g, c = lane // 4, lane % 4
        a, b = a_frag[lane], b_frag[lane]
        A[g, 2 * c], A[g, 2 * c + 1] = a[0], a[1]                  # a0 a1
        A[g + 8, 2 * c], A[g + 8, 2 * c + 1] = a[2], a[3]          # a2 a3
        A[g, 2 * c + 8], A[g, 2 * c + 9] = a[4], a[5]              # a4 a5
        A[g + 8, 2 * c + 8], A[g + 8, 2 * c + 9] = a[6], a[7]      # a6 a7
        B[2 * c, g], B[2 * c + 1, g] = b[0], b[1]                  # b0 b1
        B[2 * c + 8, g], B[2 * c + 9, g] = b[2], b[3]              # b2 b3
    C = A @ B
    return {lane: (C[lane // 4, 2 * (lane % 4)], C[lane // 4, 2 * (lane % 4) + 1],
                   C[lane // 4 + 8, 2 * (lane % 4)], C[lane // 4 + 8, 2 * (lane % 4) + 1]) for lane in range(32)}


def skinny_item(X, W, n0, ks, klen):
    """One work item of gemm_skinny_kernel<MT = 1>: 8 weight rows n0..n0+7, k range [ks, ks + klen): returns the
    accumulator fragments after all k-steps (acc[lane] = (c0, c1, c2, c3))."""
    acc = {lane: np.zeros(4, dtype=np.int64) for lane in range(32)}
    for s in range(klen // 32):                      # one k-step = 32 elements = one 16-byte load per lane
        kb = ks + s * 32
        a1, b1, a2, b2 = {}, {}, {}, {}
        for lane in range(32):
            g, c = lane // 4, lane % 4
            w = W[n0 + g, kb + 8 * c: kb + 8 * c + 8]        # uint4 w: (w.x, w.y, w.z, w.w) = 4 pairs
            xa = X[g, kb + 8 * c: kb + 8 * c + 8]            # uint4 xa: activation row g
            xb = X[g + 8, kb + 8 * c: kb + 8 * c + 8]        # uint4 xb: activation row g + 8
            # mma_16816(acc, xa.x, xb.x, xa.y, xb.y, w.x, w.y): registers A[0..3] = (a0a1, a2a3, a4a5, a6a7)
            a1[lane] = (xa[0], xa[1], xb[0], xb[1], xa[2], xa[3], xb[2], xb[3])
            b1[lane] = (w[0], w[1], w[2], w[3])
            # mma_16816(acc, xa.z, xb.z, xa.w, xb.w, w.z, w.w)
            a2[lane] = (xa[4], xa[5], xb[4], xb[5], xa[6], xa[7], xb[6], xb[7])
            b2[lane] = (w[4], w[5], w[6], w[7])
        for af, bf_ in ((a1, b1), (a2, b2)):
            cf = mma_m16n8k16(af, bf_)
            for lane in range(32):
                acc[lane] += np.array(cf[lane], dtype=np.int64)
    return acc


def test_fragment_mapping_computes_x_times_w_transposed():
    rng = np.random.default_rng(0)
    M, N, K = 16, 24, 96
    X = rng.integers(-9, 10, (M, K))
    W = rng.integers(-9, 10, (N, K))
    ref = X @ W.T
    out = np.zeros((M, N), dtype=np.int64)
    for n0 in range(0, N, 8):
        # two k-splits (32 + 64 elements) summed like the split-K finishing kernel does, in split order
        for ks, klen in ((0, 32), (32, 64)):
            acc = skinny_item(X, W, n0, ks, klen)
            for lane in range(32):
                g, c = lane // 4, lane % 4
                n = n0 + 2 * c                                  # epilogue: columns n, n + 1 of rows g and g + 8
                out[g, n] += acc[lane][0]
                out[g, n + 1] += acc[lane][1]
                out[g + 8, n] += acc[lane][2]
                out[g + 8, n + 1] += acc[lane][3]
    assert np.array_equal(out, ref)


def test_rows_beyond_m_contribute_nothing():
    """M <= 8 stages only the real rows; the fragment rows g >= M (and all of g + 8) are fed as zero registers."""
    rng = np.random.default_rng(1)
    M, K = 3, 64
    X = np.zeros((16, K), dtype=np.int64)
    X[:M] = rng.integers(-9, 10, (M, K))
    W = rng.integers(-9, 10, (8, K))
    acc = skinny_item(X, W, 0, 0, K)
    ref = X @ W.T
    for lane in range(32):
        g, c = lane // 4, lane % 4
        assert acc[lane][0] == ref[g, 2 * c] and acc[lane][1] == ref[g, 2 * c + 1]
        assert acc[lane][2] == 0 and acc[lane][3] == 0
        if g >= M:
            assert acc[lane][0] == 0 and acc[lane][1] == 0
