"""BASELINE config 2 (per-GPU batch 512 x 512, the reference's VA pre-training shape): a few training steps through the public
API, to be run under `ncu --metrics gpu__time_duration.sum` for the per-kernel times of the small-batch path.
    ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pair_kernel|finalize|merge_stats|normalize_pair|pack_stats|colsum" -s 21 -c 14 --csv --log-file gpurun_out/c2_launches.csv python scripts/c2_launches.py
"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vipant_b200 as vb  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(1213)
x1 = torch.randn(512, 512, device="cuda", generator=g).requires_grad_(True)
x2 = (0.3 * x1.detach() + torch.randn(512, 512, device="cuda", generator=g)).requires_grad_(True)
ls = torch.tensor(math.log(1 / 0.07), device="cuda", requires_grad=True)
for _ in range(6):
    loss = vb.infonce_loss(x1, x2, ls, scale_max=100.0)
    loss.backward()
torch.cuda.synchronize()
print("loss", loss.item())
