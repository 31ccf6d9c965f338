"""AudioSet tagging head: the reference's ``BCELossHead`` API with the zero-shot / multi-label SCORING on the B200 kernels.

Mirrors ``/root/reference/cvap/module/decoder/loss_more.py:29-146`` (SURVEY.md section 8f row 3 = Z2, the step right after
the similarity path for the AudioSet monitor, ``cvap/monitor/audioset_clf.py:377-404``):
  * ``zero_shot(text, gold_file)`` :77-84   normalise audio and label-text embeddings, ``S = audios @ text.t()`` -- here the
    fused normalise kernel + the fp32 similarity kernel, S stays on the device;
  * ``report(...)``               :86-131  micro / macro / weighted AP, per-class AP, ROC-AUC and the middle of the PR curve --
    the reference moves S to the host and calls scikit-learn once per class; here ``vpa_multilabel_scores`` (one CTA per
    class: shared-memory sort + fixed-order fp64 sums) and only 527 x 4 numbers travel to the host.  Same report string.
The supervised probe (LayerNorm + Linear + BCEWithLogitsLoss, :29-56, :132-146) is not on the similarity path and stays the
plain PyTorch module it is in the reference.

``cfg.sklearn_pr_truncate`` (default True) selects scikit-learn 1.0.1's precision_recall_curve -- the release the reference
pins (requirements.txt:15) -- which truncates the curve at full recall; False gives the curve of releases >= 1.1.  Only the
``mP`` / ``mR`` fields of the report depend on it (oracle/map_oracle.py documents the difference).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.distributed as dist
from torch import nn

from . import _cabi
from . import functional as F_
from .loss_head import LOSS_HEADS_REGISTRY, LossHead, _cfg_get, _compact, _LayerNormF32

__all__ = ["BCELossHead", "multilabel_scores", "similarity_matrix"]


def similarity_matrix(q: torch.Tensor, k: torch.Tensor) -> torch.Tensor:
    """fp32 ``q @ k.T`` on the device by the library's similarity kernel (one sequential fp32 sum per entry: bit-stable)."""
    F_._require_cuda(q, k)
    q, k = F_._rows2d(q.float()), F_._rows2d(k.float())
    N, D = q.shape
    M = k.shape[0]
    if k.shape[1] != D:
        raise ValueError(f"feature dims differ: {D} vs {k.shape[1]}")
    lib = _cabi.lib()
    with torch.cuda.device(q.device):
        nbytes = lib.vpa_sim_workspace_bytes(N, M)
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=q.device)
        # g = 0, k = 0: the call leaves S (N x M fp32) in the workspace (include/vipant_b200.h)
        _cabi.check(lib.vpa_sim_rank_topk(q.data_ptr(), k.data_ptr(), N, M, D, q.stride(0), k.stride(0), None, 0, 0, None, None,
                                          None, ws.data_ptr(), nbytes, torch.cuda.current_stream().cuda_stream),
                    "vpa_sim_rank_topk")
    return ws[: N * M * 4].view(torch.float32).view(N, M)


def multilabel_scores(scores: torch.Tensor, labels: torch.Tensor, truncate_pr: bool = True, micro: bool = True):
    """Per-class AP / ROC-AUC / PR-curve middle point and the micro AP of an (N, C) score matrix against (N, C) labels.

    Returns a dict of host numpy arrays: ap, auc (NaN where undefined), p_mid, r_mid (C,), support (C,), flags (C,),
    and micro_ap (float or None).  scores: CUDA fp32; labels: CUDA, positive == 1.
    """
    F_._require_cuda(scores, labels)
    if scores.dim() != 2 or scores.shape != labels.shape:
        raise ValueError(f"scores {tuple(scores.shape)} and labels {tuple(labels.shape)} must be equal (N, C) matrices")
    s = scores.float()
    if s.stride(1) != 1:
        s = s.contiguous()
    y = labels
    if y.dtype == torch.bool:
        y = y.to(torch.uint8)
    if y.dtype not in (torch.uint8, torch.float32):
        y = y.float()
    if y.stride(1) != 1:
        y = y.contiguous()
    N, C = s.shape
    dev = s.device
    per = torch.empty((C, 4), dtype=torch.float64, device=dev)
    flags = torch.empty((C,), dtype=torch.int32, device=dev)
    support = torch.empty((C,), dtype=torch.int32, device=dev)
    mic = torch.empty((), dtype=torch.float64, device=dev) if micro else None
    lib = _cabi.lib()
    with torch.cuda.device(dev):
        nbytes = lib.vpa_multilabel_workspace_bytes(N, C)
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
        _cabi.check(lib.vpa_multilabel_scores(s.data_ptr(), s.stride(0), y.data_ptr(), _cabi.U8 if y.dtype == torch.uint8 else _cabi.F32,
                                              y.stride(0), N, C, int(truncate_pr), per.data_ptr(), flags.data_ptr(),
                                              support.data_ptr(), None if mic is None else mic.data_ptr(), ws.data_ptr(), nbytes,
                                              torch.cuda.current_stream().cuda_stream), "vpa_multilabel_scores")
    per = per.cpu().numpy()
    return dict(ap=per[:, 0], auc=per[:, 1], p_mid=per[:, 2], r_mid=per[:, 3], support=support.cpu().numpy(),
                flags=flags.cpu().numpy(), micro_ap=None if mic is None else float(mic.item()))


@LOSS_HEADS_REGISTRY.register()
class BCELossHead(LossHead):
    def __init__(self, cfg, **kwargs):
        super().__init__()
        self.normalized = False
        assert "output_dim" in kwargs, "`the label number` is not found in `kwargs`"
        nlabel = kwargs["output_dim"]
        layers = []
        embed_dim = _cfg_get(cfg, "embed_dim") or cfg.width
        sizes = [embed_dim] + list(cfg.layers) + [nlabel]
        for i in range(len(sizes) - 2):
            layers.extend([_LayerNormF32(sizes[i]), nn.Linear(sizes[i], sizes[i + 1])])
        layers.extend([_LayerNormF32(sizes[-2]), nn.Linear(sizes[-2], sizes[-1], bias=cfg.bias)])
        self.linear = nn.Sequential(*layers)
        self.logit_scale = (
            nn.Parameter(torch.ones([]) * np.log(1 / 0.07)) if cfg.scaling else
            torch.ones([], requires_grad=False) * np.log(1 / 1)
        )
        self.loss_fn = nn.BCEWithLogitsLoss()
        self.truncate_pr = bool(_cfg_get(cfg, "sklearn_pr_truncate", True))
        self.reduce = False

    def copy_state_dict(self, state_dict):
        key = "logit_scale"
        new_dict = self.state_dict()
        new_dict.update({key: state_dict[key]})
        self.load_state_dict(new_dict)

    def infer(self, x1, x2, *args, **kwargs):
        if not all(hasattr(self, k) for k in ("audios", "x1s", "x2s", "ids")):
            self.audios, self.x1s, self.x2s, self.ids = [], [], [], []
        self.audios.append(_compact(x1))
        logits_per_x1 = self.logit_scale.exp().to(x1.device) * self.linear(x1)
        loss_mean_x1 = self.loss_fn(logits_per_x1, x2.float())
        self.x1s.append(torch.sigmoid(logits_per_x1))
        self.x2s.append(x2)
        names = kwargs.get("names", None)
        if names is not None:
            self.ids.extend(names)
        return loss_mean_x1

    def zero_shot(self, text, gold_file):
        audios = torch.cat(self.audios)
        already = bool(self.normalized)
        a = F_.l2_normalize(audios, already_normalized=already)            # (the reference's `if True and not self.normalized`)
        t = F_.l2_normalize(text.to(audios.device), already_normalized=already)
        return self.report(gold_file=gold_file, text=None, x1s=similarity_matrix(a, t))

    def report(self, gold_file=None, x1s=None, x2s=None, **kwargs):
        text = kwargs.get("text", None)
        if text is not None:                                               # zero-shot classification
            return self.zero_shot(text, gold_file)
        x1s = torch.cat(self.x1s) if x1s is None else x1s                  # supervised classification: sigmoid outputs
        x2s = torch.cat(self.x2s) if x2s is None else x2s
        nsample = x1s.shape[0]
        m = multilabel_scores(x1s, x2s.to(x1s.device), truncate_pr=self.truncate_pr)
        ap = m["ap"]
        w = m["support"].astype(np.float64)
        ap_micro = m["micro_ap"]
        ap_macro = float(np.mean(ap))                                      # NaN when a class has no positive, as sklearn
        ap_weighted = float(np.sum(ap * w) / w.sum()) if w.sum() > 0 else float("nan")
        has_err = bool(np.isnan(ap).any() or np.isnan(m["auc"]).any())     # the reference's `except` / isnan branches
        mean_ap = np.mean(np.where(np.isnan(ap), 0.0, ap)) * 100.
        mean_auc = np.mean(np.where(np.isnan(m["auc"]), 0.0, m["auc"])) * 100.
        mean_p = np.mean(m["p_mid"]) * 100.
        mean_r = np.mean(m["r_mid"]) * 100.
        text = f"Err({has_err}) mAP = {mean_ap:2.2f} mAUC = {mean_auc:2.2f} mP = {mean_p:2.2f} mR = {mean_r:2.2f}"
        del self.audios, self.x1s, self.x2s, self.ids
        common = f"Mac-AP = {ap_macro:2.2f} Mic-AP = {ap_micro:2.2f} wAP = {ap_weighted:2.2f}"
        return f"{common} {text} @ {nsample}"

    def forward(self, x1, x2, *args, **kwargs):
        """x1: input features, x2: multi-hot labels."""
        if not self.training:
            if not dist.is_initialized() or dist.get_rank() == 0:
                return self.infer(x1, x2, *args, **kwargs)
            return None
        logits_per_x1 = self.logit_scale.exp().to(x1.device) * self.linear(x1)
        return self.loss_fn(logits_per_x1, x2.float())
