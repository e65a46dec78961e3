"""CPU model of the work schedule of the opt-in tail-split 2-CTA GEMM (slime_b200/csrc/gemm2_tail_sm100.cu: get_item
and the host-side choice of full / R / S in slime_launch_gemm_2cta_tail).  The kernel itself has not run on hardware yet;
what can be proven without a GPU is that the schedule it walks is a partition: every (tile, k-block) pair is computed by
exactly one cluster, every tail tile has exactly one finisher (part 0) and S - 1 writers with distinct workspace slots,
no cluster gets more than one tail item, the tail item is always a cluster's LAST item (the no-deadlock argument), and the
workspace bound the host checks is the one the kernel indexes."""
import itertools

import pytest

BLOCK_K = 64


def host_plan(M, N, K, num_sms=148):
    """mirror of slime_launch_gemm_2cta_tail: returns None when the tail split does not apply"""
    num_m, num_n = (M + 255) // 256, (N + 255) // 256
    tiles, num_kb, clusters = num_m * num_n, (K + BLOCK_K - 1) // BLOCK_K, num_sms // 2
    if tiles <= clusters:
        return None
    R = tiles % clusters
    full = tiles - R
    if R == 0:
        return None
    S = clusters // R
    S = min(S, num_kb // 4)
    if S < 2:
        return None
    return dict(tiles=tiles, num_kb=num_kb, clusters=clusters, full=full, R=R, S=S)


def get_item(i, cluster, clusters, num_kb, full, R, S):
    """mirror of the device function"""
    waves = full // clusters
    if i < waves:
        return dict(tile=cluster + i * clusters, kb0=0, kb1=num_kb, part=0, parts=1, tail=-1)
    if i > waves or cluster >= R * S:
        return None
    r, s = cluster % R, cluster // R
    return dict(tile=full + r, kb0=s * num_kb // S, kb1=(s + 1) * num_kb // S, part=s, parts=S, tail=r)


SHAPES = [(1379, 4096, 4096), (1379, 4096, 14336), (1379, 28672, 4096), (22059, 6144, 4096), (2885, 4096, 1024),
          (300, 20480, 512), (1251, 4096, 11008), (3392 * 8, 5120, 13824), (1379, 6144, 4096), (22059, 4096, 4096)]
SHAPES += [(m, n, k) for m, n, k in itertools.product((257, 700, 5000), (768, 5120, 9999 // 8 * 8), (256, 1024, 8192))]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_schedule_is_a_partition(M, N, K):
    pl = host_plan(M, N, K)
    if pl is None:
        pytest.skip("tail split does not apply to this shape (single wave, no remainder, or K too short)")
    C, nkb, full, R, S = pl["clusters"], pl["num_kb"], pl["full"], pl["R"], pl["S"]
    assert full % C == 0 and 0 < R < C and R * S <= C and S >= 2
    covered = {}
    finishers, writer_slots = {}, set()
    for c in range(C):
        items = []
        for i in itertools.count():
            w = get_item(i, c, C, nkb, full, R, S)
            if w is None:
                break
            items.append(w)
        # all roles of a cluster walk the same list; a tail item, if any, is the last one
        assert sum(1 for w in items if w["tail"] >= 0) <= 1
        if any(w["tail"] >= 0 for w in items):
            assert items[-1]["tail"] >= 0
        for w in items:
            assert 0 <= w["tile"] < pl["tiles"] and 0 <= w["kb0"] < w["kb1"] <= nkb
            assert w["kb1"] - w["kb0"] >= 4 or w["parts"] == 1
            for kb in range(w["kb0"], w["kb1"]):
                key = (w["tile"], kb)
                assert key not in covered, f"{key} computed twice"
                covered[key] = c
            if w["tail"] >= 0:
                if w["part"] == 0:
                    assert w["tail"] not in finishers
                    finishers[w["tail"]] = c
                else:
                    slot = w["tail"] * (w["parts"] - 1) + (w["part"] - 1)  # workspace tile index used by the kernel
                    assert slot not in writer_slots
                    writer_slots.add(slot)
                    assert slot < R * (S - 1)  # the bound the host checks against splitk_ws_floats
    assert len(covered) == pl["tiles"] * nkb, "some (tile, k-block) is never computed"
    assert sorted(finishers) == list(range(R)) and len(writer_slots) == R * (S - 1)


def test_batch1_decoder_shapes_get_the_expected_split():
    assert host_plan(1379, 4096, 4096) == dict(tiles=96, num_kb=64, clusters=74, full=74, R=22, S=3)      # o-proj
    assert host_plan(1379, 4096, 14336)["S"] == 3                                                        # down-proj
    assert host_plan(1379, 28672, 4096) == dict(tiles=672, num_kb=64, clusters=74, full=666, R=6, S=12)  # gate / up
    assert host_plan(1379, 6144, 4096) is None   # QKV: 144 tiles = 74 + 70 -> S = 1, nothing to split
    assert host_plan(64, 4096, 4096) is None     # a single partial wave
