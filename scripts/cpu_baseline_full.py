"""Full-depth CPU run of the reference path (SURVEY.md §8d C1 / BASELINE.md §4), to check bench.py's bounded sample:
the oracle port of `tranformer_forward` (19 double + 38 single blocks at FLUX.1-dev width, fp32, B = 1, 512x512 + image
condition, S = 2560) with ONE double-block and ONE single-block weight set aliased across all positions (47.6 GB of fp32
weights do not fit next to everything else; the FLOPs are identical), against bench.py's extrapolation from one
double + one single block forward.  Prints one JSON line.

  python scripts/cpu_baseline_full.py [steps]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import flux_dit as O

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
torch.set_num_threads(os.cpu_count() or 1)
full = O.FluxConfig()
one = O.FluxConfig(num_layers=1, num_single_layers=1)
P1 = O.init_params(one, seed=1234, dtype=torch.float32)
P = {}
for k, v in P1.items():
    if k.startswith("transformer_blocks.0."):
        for i in range(full.num_layers):
            P[k.replace("transformer_blocks.0.", f"transformer_blocks.{i}.", 1)] = v
    elif k.startswith("single_transformer_blocks.0."):
        for i in range(full.num_single_layers):
            P[k.replace("single_transformer_blocks.0.", f"single_transformer_blocks.{i}.", 1)] = v
    else:
        P[k] = v
side, nt = 32, 512
ni = side * side
g = torch.Generator().manual_seed(42)
lat = torch.randn(1, ni, 64, generator=g)
cond = torch.randn(1, ni, 64, generator=torch.Generator().manual_seed(43))
g2 = torch.Generator().manual_seed(44)
pe, pooled = torch.randn(1, nt, 4096, generator=g2) * 0.1, torch.randn(1, 768, generator=g2)
ids = torch.zeros(side, side, 3)
ids[..., 1] += torch.arange(side)[:, None]
ids[..., 2] += torch.arange(side)[None, :]
ids = ids.reshape(-1, 3)
cids = ids.clone()
cids[:, 2] -= side


def forward(t):
    with torch.no_grad():
        return O.tranformer_forward(P, full, cond, cids, None, {}, 0, hidden_states=lat, encoder_hidden_states=pe,
                                    pooled_projections=pooled, timestep=torch.full((1,), t), img_ids=ids,
                                    txt_ids=torch.zeros(nt, 3), guidance=torch.full((1,), 3.5))


t0 = time.perf_counter()
forward(1.0)
warm = time.perf_counter() - t0
times = []
for s in range(steps):
    t0 = time.perf_counter()
    forward(1.0 - 0.2 * (s + 1))
    times.append(time.perf_counter() - t0)
full_step = sum(times) / len(times)

# bench.py's bounded sample on the same box, same process
import bench  # noqa: E402

cs = bench.CpuSample()
cs.run()
reps = [cs.run() for _ in range(6)]
td, ts = sum(r[0] for r in reps) / len(reps), sum(r[1] for r in reps) / len(reps)
extrap_step = 19 * td + 38 * ts
flop = 37.66e12
print(json.dumps({
    "what": "oracle port, full-depth DiT forward (57 blocks, aliased weights), fp32, B=1, S=2560",
    "cores": torch.get_num_threads(), "steps_timed": steps, "warmup_s": warm, "s_per_step_full_depth": full_step,
    "s_per_step_each": times, "tflops": flop / full_step / 1e12,
    "s_per_step_extrapolated_from_2_blocks": extrap_step, "ratio_full_over_extrapolated": full_step / extrap_step,
    "s_per_28_step_edit_full_depth": 28 * full_step, "s_per_28_step_edit_extrapolated": 28 * extrap_step}))
