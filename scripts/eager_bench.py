"""Same-box comparison: the reference's formula (loss_head.py:271-283) executed by PyTorch eager on the B200, fp32 (TF32 off,
torch's default for matmul) and under bf16 autocast, forward + backward, against vipant_b200 on the same inputs.
SURVEY.md 8(d) asks for this number beside the CPU baseline.  B x B logits are materialised by eager (2 x 4 GiB fp32 at 32768
plus softmax buffers): needs ~40 GB at B = 32768.
    python scripts/eager_bench.py [B] [steps] > gpurun_out/eager.json"""
import json, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vipant_b200 as vb

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
D = 512
torch.backends.cuda.matmul.allow_tf32 = False
g = torch.Generator().manual_seed(1213)
x1h = torch.randn(B, D, generator=g)
x2h = 0.3 * x1h + 0.7 * torch.randn(B, D, generator=g)
x1 = x1h.cuda().requires_grad_(True)
x2 = x2h.cuda().requires_grad_(True)
ls = torch.tensor(math.log(1 / 0.07), device="cuda", requires_grad=True)
labels = torch.arange(B, device="cuda")
ce = torch.nn.CrossEntropyLoss()


def eager():
    a = x1 / x1.norm(dim=-1, keepdim=True)
    t = x2 / x2.norm(dim=-1, keepdim=True)
    s = ls.exp()
    loss = ce(s * a @ t.t(), labels) + ce(s * t @ a.t(), labels)      # both logits matrices, as the reference does
    loss.backward()
    return loss


def eager_amp():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        return eager()


def ours():
    loss = vb.infonce_loss(x1, x2, ls, precision="bf16")
    loss.backward()
    return loss


def timed(fn):
    for _ in range(3):
        x1.grad = x2.grad = ls.grad = None
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        x1.grad = x2.grad = ls.grad = None
        loss = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, float(loss)


out = {"B": B, "D": D, "steps": steps}
for name, fn in (("vipant_b200_bf16", ours), ("torch_eager_fp32", eager), ("torch_eager_bf16_autocast", eager_amp)):
    try:
        torch.cuda.reset_peak_memory_stats()
        ms, loss = timed(fn)
        out[name] = {"ms_per_step": ms, "pairs_per_s": B / ms * 1e3, "loss": loss, "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}
    except Exception as exc:      # e.g. out of memory for the materialised logits
        out[name] = {"error": str(exc)[:200]}
        torch.cuda.empty_cache()
print(json.dumps(out, indent=1))
