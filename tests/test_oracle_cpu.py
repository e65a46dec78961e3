"""Pins the CPU oracle (oracle/slime_oracle.py) to the golden vectors produced by the unmodified
reference (oracle/gen_golden.py -> tests/golden/*.npz).  fp32 vs fp32: only summation order differs,
so the float tolerance is 2e-4 relative; every integer result (lengths, masks, labels, selected
indices) must match exactly."""
import os

import numpy as np
import pytest
import torch

from oracle import slime_oracle as O
from oracle.gen_golden import CASES
from slime_b200.config import preset
from slime_b200.synth import synth_inputs, synth_state_dict

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
FLOAT_TOL = 2e-4


def rel(a, b):
    a, b = torch.as_tensor(a).float(), torch.as_tensor(b).float()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


def load_case(name):
    pname, over, B, n, T, ipos, ragged, isize, with_labels = CASES[name]
    cfg = preset(pname, **over)
    gold = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLDEN, name + ".npz")).items()}
    px, ids, mask = synth_inputs(cfg, B, n, T, image_pos=ipos, ragged=ragged)
    labels = None
    if with_labels:
        labels = ids.clone()
        labels[:, : ipos + 2] = -100
        labels[labels == -200] = -100
    return cfg, gold, px, ids, mask, labels, isize


_cache = {}


def run_oracle(name):
    if name not in _cache:
        cfg, gold, px, ids, mask, labels, isize = load_case(name)
        sd = synth_state_dict(cfg)
        grids = [O.grid_shape(isize, cfg.vit_image)] * px.shape[0]
        with torch.no_grad():
            res = O.prefill(sd, cfg, px, ids, mask, grids, labels=labels)
        _cache[name] = (cfg, gold, res, (px, ids, mask, labels))
    return _cache[name]


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_golden(name):
    cfg, gold, res, (px, ids, mask, labels) = run_oracle(name)
    B = px.shape[0]
    # --- integer results: exact ---
    assert res["lengths"] == gold["lengths"].tolist()
    assert torch.equal(res["attention_mask"], gold["attention_mask"].bool())
    if labels is not None:
        assert torch.equal(res["labels"], gold["labels"])
    if "sel_idx" in gold:
        for b in range(B):
            k = int(gold["sel_count"][b])
            assert res["sel"][b].tolist() == gold["sel_idx"][b, :k].tolist(), f"selection differs for sample {b}"
        if cfg.mm_patch_merge_type == "spatial":
            assert list(O.grid_shape(CASES[name][7], cfg.vit_image)) == gold["grid_wh"].tolist()
    # --- floating-point stages ---
    assert rel(torch.stack(res["vit"])[:, :, ::16, :], gold["vit"]) < FLOAT_TOL
    assert rel(torch.stack(res["glob"])[:, ::8, :], gold["glob"]) < FLOAT_TOL
    if "local_c" in gold:
        assert rel(torch.stack(res["local_c"])[:, :, ::8, :], gold["local_c"]) < FLOAT_TOL
        assert rel(res["local_m"][0][::8], gold["local_m"]) < FLOAT_TOL
        assert rel(torch.stack(res["probs"]), gold["probs"]) < FLOAT_TOL
    assert rel(res["inputs_embeds"][:, ::8, :], gold["embeds_rows"]) < FLOAT_TOL
    assert rel(res["inputs_embeds"].norm(dim=-1), gold["embeds_norm"]) < FLOAT_TOL
    last = torch.stack([lg[-1] for lg in res["logits"]])
    assert rel(last, gold["logits_last"]) < 5 * FLOAT_TOL


def test_grid_shape_matches_reference_examples():
    # values produced by the reference's get_anyres_image_grid_shape (llava/mm_utils.py:156-174)
    assert O.grid_shape((672, 672)) == (2, 2)
    assert O.grid_shape((336, 336)) in ((1, 2), (2, 1))
    assert O.grid_shape((1008, 672)) == (3, 2)
    assert O.grid_shape((4000, 300))[0] * O.grid_shape((4000, 300))[1] <= 7


def test_top_p_rule_edges():
    p = torch.tensor([0.5, 0.3, 0.1, 0.1])
    assert O.top_p_select(p, 0.95).tolist() == [0, 1, 2, 3]   # cum = .5,.8,.9,1.0 -> count 3 -> keep 4
    assert O.top_p_select(p, 0.5).tolist() == [0, 1]          # count 1 -> keep 2
    assert O.top_p_select(p, 0.05).tolist() == [0]            # count 0 -> keep 1
    assert O.top_p_select(p, 2.0).tolist() == [0, 1, 2, 3]    # everything under the threshold
    tie = torch.tensor([0.25, 0.25, 0.25, 0.25])
    assert O.top_p_select(tie, 0.3).tolist() == [0, 1]        # stable: lower index first
