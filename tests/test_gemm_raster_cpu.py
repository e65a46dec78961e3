"""CPU mirrors of two pieces of host / device integer logic of the cta_group::2 GEMM (csrc/gemm2_common.cuh tile_coord,
csrc/gemm2_sm100.cu slime_gemm2_block_n): the serpentine rasterisation must visit every tile exactly once for any
problem and group size, and the tile-width rule must keep the 256-column throughput shape for every GEMM of the headline
step while picking 192 columns where the last 256-wide wave would be nearly empty (batch-1 prefill)."""
import math

import pytest


def tile_coord(t, num_m, num_n, group_m_in):
    snake = group_m_in < 0
    group_m = abs(group_m_in)
    per_group = group_m * num_n
    group = t // per_group
    first_m = group * group_m
    gsize = min(group_m, num_m - first_m)
    i = t - group * per_group
    m, n = first_m + i % gsize, i // gsize
    if snake and group & 1:
        n = num_n - 1 - n
    return m, n


@pytest.mark.parametrize("num_m,num_n", [(87, 112), (6, 16), (1, 1), (181, 12), (5, 7), (87, 16), (12, 4)])
@pytest.mark.parametrize("g", [1, 4, 8, 16, 64])
@pytest.mark.parametrize("sign", [1, -1])
def test_rasterisation_visits_every_tile_once(num_m, num_n, g, sign):
    seen = [tile_coord(t, num_m, num_n, sign * g) for t in range(num_m * num_n)]
    assert len(set(seen)) == num_m * num_n
    assert all(0 <= m < num_m and 0 <= n < num_n for m, n in seen)
    if sign < 0 and num_m > g:
        # serpentine: the last tile column of a group is the first one of the next group
        last_of_group0 = seen[min(g, num_m) * num_n - 1][1]
        first_of_group1 = seen[min(g, num_m) * num_n][1]
        assert last_of_group0 == first_of_group1 == num_n - 1


def block_n(M, N, num_sms=148):
    clusters = num_sms // 2
    m_tiles = (M + 255) // 256
    t256, t192 = m_tiles * ((N + 255) // 256), m_tiles * ((N + 191) // 192)
    w256 = float((t256 + clusters - 1) // clusters)
    w192 = float((t192 + clusters - 1) // clusters) * 0.75 * 1.04
    return 192 if (N >= 192 and w192 < w256) else 256


def test_tile_width_rule():
    # headline step (batch 16: 22059 packed decoder rows, 46160 ViT rows): everything stays on the 256-column shape
    for M, N in [(22059, 28672), (22059, 6144), (22059, 4096), (46160, 3072), (46160, 4096), (46160, 1024)]:
        assert block_n(M, N) == 256, (M, N)
    # batch-1 prefill (1379 / 1251 rows): the N = 4096 projections are two waves with a nearly empty second one -> 192
    assert block_n(1379, 4096) == 192 and block_n(1251, 4096) == 192
    assert block_n(1379, 28672) == 256 and block_n(1379, 6144) == 256
    # the rule only ever picks 192 when it needs fewer width-weighted waves
    for M in (300, 1251, 1379, 2885, 7152, 22059):
        for N in (1024, 3072, 4096, 5120, 6144, 12288, 13824 * 2, 28672):
            bn = block_n(M, N)
            waves = lambda w: math.ceil(((M + 255) // 256) * math.ceil(N / w) / 74) * w / 256  # noqa: E731
            if bn == 192:
                assert waves(192) < waves(256)
