"""One launch (after a warm-up launch) of every kernel that needs an ncu table in profiles/: fused scoring, normalise,
multi-label scoring, multi-pair finalize.  Run under
    ncu --set full --clock-control none --import-source on -k regex:"<names>" -f -o gpurun_out/prof_targets python scripts/ncu_targets.py"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import vipant_b200 as vb
from vipant_b200 import _cabi, functional as F_
from vipant_b200.loss_more import multilabel_scores, similarity_matrix
from oracle.make_golden import retrieval_inputs_1v5

a, t = retrieval_inputs_1v5(n=975, seed=1213)
an, tn = vb.l2_normalize(torch.from_numpy(a).cuda()), vb.l2_normalize(torch.from_numpy(t).cuda())
gt12 = torch.arange(4875, device="cuda").view(975, 5); gt21 = torch.arange(4875, device="cuda") // 5
x1, x2 = torch.randn(32768, 512, device="cuda"), torch.randn(32768, 512, device="cuda")
rng = np.random.default_rng(1213)
NA, CA = 20371, 527
text_a = rng.standard_normal((CA, 512)).astype(np.float32)
Ya = np.zeros((NA, CA), np.float32)
for _ in range(2):
    Ya[np.arange(NA), rng.integers(0, CA, NA)] = 1.0
aud_a = (Ya @ text_a * 0.08 + rng.standard_normal((NA, 512))).astype(np.float32)
Sg = similarity_matrix(vb.l2_normalize(torch.from_numpy(aud_a).cuda()), vb.l2_normalize(torch.from_numpy(text_a).cuda()))
yg = torch.from_numpy(Ya).cuda()
feats = [torch.randn(4096, 512, device="cuda", requires_grad=True) for _ in range(3)]
lss = [torch.tensor(v, device="cuda", requires_grad=True) for v in (2.0, 2.66, 3.2)]
for _ in range(2):
    F_.sim_rank_fused(an, tn, gt_q=gt12, gt_k=gt21)
    F_._KERNELS.normalize_pair(x1, x2, False, _cabi.PREC_BF16_TC)
    multilabel_scores(Sg, yg)
    vb.infonce_multi_loss(feats, [(0, 1), (0, 2), (1, 2)], lss).sum().backward()
torch.cuda.synchronize()
