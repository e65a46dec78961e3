"""Bank-conflict models of the shared-memory layouts of the decode-step kernels, replaying the device address formulas:

  * gemm_skinny.cu - the staged activation rows are read with one 16-byte load per lane at
    row g * stride + 16 c bytes; the plan pads the row stride to 64 (mod 128) bytes;
  * decode_attn.cu (decode_attn_mma_kernel) - K / V tiles of 16 rows x 256 bytes with chunk ch of row r stored at
    position ch ^ (r & 7) (dm_off); written by cp.async (lanes 0..15 one row, 16..31 the next) and read by
    ldmatrix.x4 (lanes 8i..8i+7 give the 8 row addresses of matrix i).

Model: shared memory has 32 four-byte banks; a 16-byte-per-lane access is served in phases of 8 lanes (a quarter warp),
an ldmatrix in phases of one 8x8 matrix; a phase is conflict-free when its 8 sixteen-byte pieces fall into 8 different
16-byte bank groups ((address / 16) % 8)."""
import pytest


def groups_distinct(addrs):
    g = [(a // 16) % 8 for a in addrs]
    return len(set(g)) == len(g)


def skinny_stride_bytes(kc):
    """make_plan(): xs_stride = kc + (kc % 64 == 0 ? 32 : 0) elements of 2 bytes"""
    return (kc + (32 if kc % 64 == 0 else 0)) * 2


@pytest.mark.parametrize("kc", [32, 64, 96, 512, 1024, 1792, 2048, 3456, 3584, 4096])
def test_skinny_activation_reads_are_conflict_free(kc):
    stride = skinny_stride_bytes(kc)
    assert stride % 128 == 64
    for k_step in range(0, min(kc // 32, 4)):
        for half in (0, 8):  # rows g (xa) and g + 8 (xb) are separate load instructions
            lane_addr = [((lane // 4) + half) * stride + (lane % 4) * 16 + k_step * 64 for lane in range(32)]
            for phase in range(4):
                assert groups_distinct(lane_addr[phase * 8: phase * 8 + 8]), (kc, k_step, half, phase)
    # the unpadded layout would put all 8 rows of a phase pair on the same banks
    bad = [((lane // 4)) * (kc * 2) + (lane % 4) * 16 for lane in range(8)]
    if (kc * 2) % 128 == 0:
        assert not groups_distinct(bad)


def dm_off(row, chunk):
    return row * 256 + ((chunk ^ (row & 7)) << 4)


def test_decode_attention_tile_layout_is_a_bijection():
    seen = {dm_off(r, ch) for r in range(16) for ch in range(16)}
    assert len(seen) == 256 and min(seen) == 0 and max(seen) == 16 * 256 - 16


def test_decode_attention_cp_async_writes_are_conflict_free():
    for j in range(8):  # instruction j of issue(): rows 2j and 2j + 1
        lane_addr = [dm_off(j * 2 + (lane >> 4), lane & 15) for lane in range(32)]
        for phase in range(4):
            assert groups_distinct(lane_addr[phase * 8: phase * 8 + 8])


def test_decode_attention_ldmatrix_reads_are_conflict_free():
    for ks in range(8):  # K: rows ((lane >> 4) << 3) + (lane & 7), chunk ks * 2 + ((lane >> 3) & 1)
        lane_addr = [dm_off(((lane >> 4) << 3) + (lane & 7), ks * 2 + ((lane >> 3) & 1)) for lane in range(32)]
        for m in range(4):
            assert groups_distinct(lane_addr[m * 8: m * 8 + 8]), ("K", ks, m)
    for dp in range(8):  # V (.trans): rows (((lane >> 3) & 1) << 3) + (lane & 7), chunk dp * 2 + (lane >> 4)
        lane_addr = [dm_off((((lane >> 3) & 1) << 3) + (lane & 7), dp * 2 + (lane >> 4)) for lane in range(32)]
        for m in range(4):
            assert groups_distinct(lane_addr[m * 8: m * 8 + 8]), ("V", dp, m)


def test_decode_attention_ldmatrix_fetches_the_intended_matrices():
    """x4 matrices for S = Q K^T: (positions 0-7 | features lo), (0-7 | hi), (8-15 | lo), (8-15 | hi) of k-step ks;
    for O += P V: (positions 0-7 | features 16 dp .. +8), (8-15 | same), (0-7 | next 8 features), (8-15 | next 8)."""
    inv = {dm_off(r, ch): (r, ch) for r in range(16) for ch in range(16)}
    for ks in range(8):
        rows_chunks = [inv[dm_off(((lane >> 4) << 3) + (lane & 7), ks * 2 + ((lane >> 3) & 1))] for lane in range(32)]
        assert rows_chunks[0:8] == [(r, 2 * ks) for r in range(8)]
        assert rows_chunks[8:16] == [(r, 2 * ks + 1) for r in range(8)]
        assert rows_chunks[16:24] == [(r, 2 * ks) for r in range(8, 16)]
        assert rows_chunks[24:32] == [(r, 2 * ks + 1) for r in range(8, 16)]
    for dp in range(8):
        rows_chunks = [inv[dm_off((((lane >> 3) & 1) << 3) + (lane & 7), dp * 2 + (lane >> 4))] for lane in range(32)]
        assert rows_chunks[0:8] == [(r, 2 * dp) for r in range(8)]
        assert rows_chunks[8:16] == [(r, 2 * dp) for r in range(8, 16)]
        assert rows_chunks[16:24] == [(r, 2 * dp + 1) for r in range(8)]
        assert rows_chunks[24:32] == [(r, 2 * dp + 1) for r in range(8, 16)]
