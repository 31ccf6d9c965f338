"""Race detector for the launch-tuning knobs (programmatic dependent launch, carve-out): a training-like loop whose data
changes EVERY step, enqueued without host synchronisation, against the same steps run one by one with the knobs off.
The kernels and their reduction orders are the same in both runs, so every result must be bitwise equal.
    python scripts/pdl_race_check.py [B] [steps]"""
import json, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vipant_b200 as vb
from vipant_b200 import _cabi

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
lib = _cabi.lib()
gen = torch.Generator(device="cuda").manual_seed(3)
data = [(torch.randn(B, 512, device="cuda", generator=gen), torch.randn(B, 512, device="cuda", generator=gen)) for _ in range(steps)]


def run(sync_each):
    out = []
    ls = torch.tensor(math.log(1 / 0.07), device="cuda", requires_grad=True)
    for x1, x2 in data:
        a, b = x1.clone().requires_grad_(True), (0.3 * x1 + x2).requires_grad_(True)
        ls.grad = None
        loss = vb.infonce_loss(a, b, ls)
        loss.backward()
        with torch.no_grad():
            ls -= 0.01 * ls.grad          # the temperature moves too: the next step's kernels read what this one wrote
        out.append((loss.detach().clone(), a.grad.clone(), b.grad.clone(), ls.grad.clone()))
        if sync_each:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    return out


res = {}
lib.vpa_launch_tuning(0, 0)
ref = run(True)
for pdl, carve in ((1, 0), (0, 1), (1, 1)):
    lib.vpa_launch_tuning(pdl, carve)
    got = run(False)
    bad = [i for i, (r, g) in enumerate(zip(ref, got)) if not all(torch.equal(x, y) for x, y in zip(r, g))]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(); run(False); e1.record(); torch.cuda.synchronize()
    res[f"pdl={pdl},carveout={carve}"] = {"steps_that_differ": bad, "ms_per_step": e0.elapsed_time(e1) / steps}
lib.vpa_launch_tuning(0, 0)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record(); run(False); e1.record(); torch.cuda.synchronize()
res["pdl=0,carveout=0"] = {"steps_that_differ": [], "ms_per_step": e0.elapsed_time(e1) / steps}
print(json.dumps({"B": B, "steps": steps, **res}, indent=1))
