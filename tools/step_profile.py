"""One prefill step (or one decode step) bracketed by cudaProfilerStart/Stop, for ncu launch lists:

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches_b1.csv python tools/step_profile.py --model llama3-8b --batch 1
    python tools/launch_summary.py gpurun_out/launches_b1.csv profiles/r02_launch_list_b1.txt "<header>"
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from slime_b200.config import preset
from slime_b200.engine import SlimeEngine
from slime_b200.synth import grid_for_crops, synth_inputs, synth_tensor, weight_specs

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="llama3-8b")
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--crops", type=int, default=5)
ap.add_argument("--prompt-len", type=int, default=256)
ap.add_argument("--decode", action="store_true", help="profile one decode step after the prefill instead")
ap.add_argument("--layers", type=int, default=None)
a = ap.parse_args()
cfg = preset(a.model) if a.layers is None else preset(a.model, num_hidden_layers=a.layers)
dev = torch.device("cuda", 0)
specs = {n: (s, k) for n, s, k in weight_specs(cfg)}
eng = SlimeEngine(cfg, 0, max_pos=4096)
eng.load_weights(lambda n: synth_tensor(n, specs[n][0], specs[n][1], 3407, device=dev, dtype=torch.bfloat16))
px, ids, mask = synth_inputs(cfg, a.batch, a.crops, a.prompt_len)
px, ids, mask = px.to(dev).to(torch.bfloat16), ids.to(dev), mask.to(dev)
grids = [grid_for_crops(a.crops - 1)] * a.batch
for _ in range(3):
    res = eng.prefill(px, ids, mask, grids=grids)
torch.cuda.synchronize()
if a.decode:
    eng.attach_kv_cache(a.batch, max(res.lengths) + 16)
    res = eng.prefill(px, ids, mask, grids=grids)
    lens = torch.tensor(res.lengths, dtype=torch.int32, device=dev)
    x = eng.weights["llm.embed"][res.logits_last.argmax(-1)]
    for _ in range(3):
        eng.decode_step(x, lens)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    eng.decode_step(x, lens)
else:
    torch.cuda.cudart().cudaProfilerStart()
    eng.prefill(px, ids, mask, grids=grids)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done", res.lengths[:4])
