"""Drop-in loss heads: the reference's loss-head API on top of the B200 kernels.

Mirrors the public surface of ``/root/reference/cvap/module/decoder/loss_head.py`` for the InfoNCE path:
same class names, constructor ``(cfg, **kwargs)``, ``forward(x1, x2, *args, **kwargs)`` with the
``normalized`` / ``names`` kwargs, train/eval split, ``infer`` / ``report`` / ``copy_state_dict`` /
``stats``, the ``logit_scale`` parameter name (checkpoint key) and the exact report strings
(which ``cvap/monitor/esc50_clf.py:321`` regex-parses).  Differences, all by design:
  * arithmetic runs in libvipant_b200.so (no B x B logits, no argsort); CPU tensors raise;
  * ``cfg.precision`` ("bf16" default | "fp32"), ``cfg.gather`` (False default) and ``cfg.ddp_average`` (False default)
    are optional extra cfg fields; ``gather=True`` computes the GLOBAL-batch loss over the default process group, i.e. the
    semantics of the reference's `dp` mode (loss on the gathered batch) under one process per GPU;
    ``ddp_average=True`` is for a trainer that wraps the model in DistributedDataParallel (which AVERAGES gradients,
    cvap/monitor/cvap.py:37): the loss is multiplied by the world size and d logit_scale stays a per-rank partial like the
    feature gradients, so that after DDP's averaging every parameter gets exactly the `dp`-mode gradient;
  * BarlowLossHead / BarlowCELossHead and the BCE / LM heads of loss_more.py are out of scope.
"""
from __future__ import annotations

import itertools
import json
from collections import defaultdict

import numpy as np
import torch
import torch.distributed as dist
from torch import nn

from . import functional as F_

__all__ = ["LOSS_HEADS_REGISTRY", "build_loss_head", "LossHead", "CELossHead", "ClassificationHead",
           "VALCELossHead", "VACELossHead", "install_into_reference"]


class _Registry:
    """Minimal stand-in for fvcore's Registry: ``register()`` decorator / direct call, ``get(name)``."""

    def __init__(self, name):
        self._name = name
        self._map = {}

    def register(self, obj=None):
        if obj is None:
            return lambda cls: self.register(cls)
        self._map[obj.__name__] = obj
        return obj

    def get(self, name):
        if name not in self._map:
            raise KeyError(f"No object named '{name}' found in '{self._name}' registry!")
        return self._map[name]

    def __contains__(self, name):
        return name in self._map


LOSS_HEADS_REGISTRY = _Registry("LOSS_HEADS")


def build_loss_head(cfg, **kwargs):
    """Same contract as the reference factory (loss_head.py:22-23): keyed by ``cfg.name``."""
    return LOSS_HEADS_REGISTRY.get(cfg.name)(cfg, **kwargs)


_HEAD_IDS = itertools.count()      # construction order: identical on every rank, names the peer-memory segment of a head


def _is_primary():
    return not dist.is_initialized() or dist.get_rank() == 0


def _compact(x):
    """The batch as the stash keeps it: the tensor itself when it owns its storage (an encoder output), a copy when it is a
    view into something larger (`hidden[:, 0, :]` would keep the whole hidden state alive for the rest of the evaluation)."""
    x = x.detach()
    if x.is_contiguous() and x.untyped_storage().nbytes() <= 2 * x.numel() * x.element_size() + 512:
        return x
    return x.clone(memory_format=torch.contiguous_format)


def _cfg_get(cfg, key, default=None):
    try:
        v = getattr(cfg, key)
    except Exception:
        return default
    return default if v is None else v


class LossHead(nn.Module):
    """Eval-time scoring base: stash normalised features, then rank-based metrics (loss_head.py:25-244)."""

    def __init__(self):
        super().__init__()
        self.reduce = False
        self.normalized = True

    def copy_state_dict(self, state_dict):
        pass

    # -- accumulate ------------------------------------------------------------------------------
    def infer(self, x1, x2, *args, **kwargs):
        """Stash the batch (reference :34-46).  The L2 normalisation of :38-40 is applied when the stash is read (`_stash`):
        one fused kernel launch over the concatenated features instead of two launches per batch."""
        if not (hasattr(self, "x1s") and hasattr(self, "x2s") and hasattr(self, "ids")):
            self.x1s, self.x2s, self.ids = [], [], []
            self._stash_normalized = []
        F_._require_cuda(x1, x2)
        self.x1s.append(_compact(x1))
        self.x2s.append(_compact(x2))
        self._stash_normalized.append(bool(kwargs.get("normalized", False)))
        names = kwargs.get("names", None)
        if names is not None:
            self.ids.extend(names)
        return None

    def _stash(self):
        """(x1s, x2s): the stashed features as two fp32 matrices of unit-norm rows."""
        flags = getattr(self, "_stash_normalized", [True] * len(self.x1s))
        if len(set(flags)) <= 1:
            already = bool(flags[0]) if flags else True
            return (F_.l2_normalize(torch.cat(self.x1s), already_normalized=already),
                    F_.l2_normalize(torch.cat(self.x2s), already_normalized=already))
        return (torch.cat([F_.l2_normalize(x, already_normalized=f) for x, f in zip(self.x1s, flags)]),
                torch.cat([F_.l2_normalize(x, already_normalized=f) for x, f in zip(self.x2s, flags)]))

    # -- metrics ---------------------------------------------------------------------------------
    @staticmethod
    def retrieval_metrics(ranks, nsample=None, msg=""):
        """R@1/5/10/50, median and mean rank (+1) of a float32 rank vector (reference :67-77)."""
        nsample = nsample or ranks.shape[0]
        hits = {k: int((ranks < k).sum()) / nsample * 100. for k in (1, 5, 10, 50)}
        med = ranks.median() + 1
        avg = ranks.mean() + 1
        return (f"{msg}: R@1 {hits[1]:2.2f} R5 {hits[5]:2.2f} R10 {hits[10]:2.2f} R50 {hits[50]:2.2f} "
                f"MED {med:2.2f} AVG {avg:2.2f}")

    @staticmethod
    def _retrieval_eval_from_ranks(r12, r21):
        """`retrieval_eval` (:79-107) from the ranks the kernel already produced: A->T uses the best of the k
        captions per clip, T->A the rank of the clip; both as float32 vectors like the reference."""
        msg_12 = LossHead.retrieval_metrics(r12.min(-1)[0].float(), msg="A->T")
        msg_21 = LossHead.retrieval_metrics(r21.float(), msg="T->A")
        return f"{msg_12}\n{msg_21}"

    @staticmethod
    def retrieval_eval(x1s, x2s, k=5):
        n = x1s.shape[0]
        gt12 = torch.arange(n * k, device=x1s.device).view(n, k)
        gt21 = torch.arange(n * k, device=x1s.device) // k
        res = F_.sim_rank_fused(x1s, x2s, gt_q=gt12, gt_k=gt21)          # both directions, one pass over the similarity
        return LossHead._retrieval_eval_from_ranks(res["ranks_q"].cpu(), res["ranks_k"][:, 0].cpu())

    def _gold_cluster(self, gold_file, nsample, verbose=False):
        by_class, by_sample = defaultdict(list), defaultdict(str)
        with open(gold_file, "r") as fr:
            for iline, line in enumerate(fr):
                if iline + 1 > nsample:
                    break
                record = json.loads(line)
                key = " ".join(record["labels"])
                by_class[key].append(record["id"])
                by_sample[record["id"]] = key
        if verbose:
            for k, v in sorted(by_class.items(), key=lambda kv: -len(kv[1])):
                print(k, len(v))
            print(f"total {len(by_class)} groups")
        return by_class, by_sample

    def _class_stats(self, top1, by_class, by_sample, nsample, msg):
        """Per-class P@1 / R@1 / mAP / mAR from each row's nearest neighbour (reference :182-238, k = 1)."""
        k = 1
        per_class = defaultdict(dict)
        for idx, nb in enumerate(top1.flatten().tolist()):
            sample = self.ids[idx]
            cname = by_sample[sample]
            hit = 1 if self.ids[nb] in by_class[cname] else 0
            per_class[cname][sample] = hit
        p = r = p_cls = r_cls = 0.
        for cname, samples in per_class.items():
            nrel = len(by_class[cname])
            cp = cr = 0.
            for tp in samples.values():
                p += tp / k
                r += tp / nrel
                cp += tp / k
                cr += tp / nrel
            p_cls += cp / nrel
            r_cls += cr / nrel
        nclass = len(by_class)
        p, r = p / nsample * 100, r / nsample * 100
        p_cls, r_cls = p_cls / nclass * 100, r_cls / nclass * 100
        return f"{msg}: P@{k} {p:2.2f} R@{k} {r:2.2f} mAP {p_cls:2.2f} mAR {r_cls:2.2f}"

    def report(self, gold_file=None):
        """Reference :109-244.  ONE fused kernel pass over the similarity yields the ranks (and nearest neighbours) of both
        directions; they come to the host in one copy each and the reference's own expressions run on them there."""
        x1s, x2s = self._stash()
        n1, n2 = x1s.shape[0], x2s.shape[0]
        dev = x1s.device
        msg_12 = msg_21 = ""
        ref_metric = ""
        if n1 == n2:
            gt = torch.arange(n1, device=dev)
            want_top = gold_file is not None
            res = F_.sim_rank_fused(x1s, x2s, gt_q=gt, gt_k=gt, top1_q=want_top, top1_k=want_top)
            r12, r21 = res["ranks_q"][:, 0].cpu(), res["ranks_k"][:, 0].cpu()
            t12_1 = int((r12 < 1).sum()) / n1 * 100.
            t12_5 = int((r12 < 5).sum()) / n1 * 100.
            t21_1 = int((r21 < 1).sum()) / n1 * 100.
            t21_5 = int((r21 < 5).sum()) / n1 * 100.
            p_12 = f"I->A: t1 = {t12_1:2.2f} t5 = {t12_5:2.2f}"
            p_21 = f"A->I: t1 = {t21_1:2.2f} t5 = {t21_5:2.2f}"
            if gold_file is not None:
                by_class, by_sample = self._gold_cluster(gold_file, n1)
                msg_12 = self._class_stats(res["top1_q"][0].cpu(), by_class, by_sample, n1, "I->A")
                msg_21 = self._class_stats(res["top1_k"][0].cpu(), by_class, by_sample, n1, "A->I")
        elif n1 * 5 == n2:
            # 1 clip vs 5 captions: caption c belongs to clip c // 5 (collator order)
            gt12 = torch.arange(n2, device=dev).view(n1, 5)
            gt21 = torch.arange(n2, device=dev) // 5
            res = F_.sim_rank_fused(x1s, x2s, gt_q=gt12, gt_k=gt21)
            r12, r21 = res["ranks_q"].cpu(), res["ranks_k"][:, 0].cpu()        # (n1, 5), (n2,)
            t12_1 = (r12 < 1).sum(-1).sum() / (1 * r12.shape[0]) * 100.    # P@1   (0-d tensors, as in :143-146)
            t12_5 = (r12 < 5).sum(-1).sum() / (5 * r12.shape[0]) * 100.    # R@5
            mean12 = r12.min(-1)[0].float().mean() + 1
            p_12 = f"A->T: t1 = {t12_1:2.2f} t5 = {t12_5:2.2f} mR = {mean12:2.2f}"
            t21_1 = int((r21 < 1).sum()) / r21.shape[0] * 100.
            t21_5 = int((r21 < 5).sum()) / r21.shape[0] * 100.
            mean21 = r21.float().mean() + 1
            p_21 = f"T->A: t1 = {t21_1:2.2f} t5 = {t21_5:2.2f} mR = {mean21:2.2f}"
            ref_metric = self._retrieval_eval_from_ranks(r12, r21)      # no second similarity pass
        else:
            p_12, p_21 = f"{x1s.shape}x{x2s.shape}", "-"
        del self.x1s, self.x2s, self.ids
        if hasattr(self, "_stash_normalized"):
            del self._stash_normalized
        msg = "" if msg_12 == msg_21 == "" else f"\n{msg_12} {msg_21}\n"
        ref = "" if ref_metric == "" else f"\nREFERENCE\n{ref_metric}"
        return f"{msg}{p_12} {p_21} @ {n1}{ref}"


@LOSS_HEADS_REGISTRY.register()
class CELossHead(LossHead):
    """Contrastive (InfoNCE) head: reference loss_head.py:246-284 on the fused B200 path."""

    def __init__(self, cfg, **kwargs):
        super().__init__()
        self.logit_scale = (
            nn.Parameter(torch.ones([]) * np.log(1 / 0.07)) if cfg.scaling else
            torch.ones([], requires_grad=False) * np.log(1 / 1)
        )
        self.scale_max = _cfg_get(cfg, "scale_max") or float("inf")
        self.precision = _cfg_get(cfg, "precision", "bf16")
        self.gather = bool(_cfg_get(cfg, "gather", False))
        self.ddp_average = bool(_cfg_get(cfg, "ddp_average", False))
        self._segment_key = next(_HEAD_IDS)
        self.reduce = False

    def copy_state_dict(self, state_dict):
        key = "logit_scale"
        new_dict = self.state_dict()
        if key in new_dict and key in state_dict:
            new_dict.update({key: state_dict[key]})
        self.load_state_dict(new_dict)

    def forward(self, x1, x2, *args, **kwargs):
        if not self.training:
            if _is_primary():
                return self.infer(x1, x2, *args, **kwargs)
            return None
        group = dist.group.WORLD if (self.gather and dist.is_initialized() and dist.get_world_size() > 1) else None
        cap = None if self.scale_max == float("inf") else self.scale_max
        precision = self.precision
        if precision == "bf16" and not F_.tensor_core_supported(x1.shape[-1]):
            precision = "fp32"      # embed dims the tcgen05 tiling does not cover run on the exact kernel
        ddp = self.ddp_average and group is not None
        loss = F_.infonce_loss(x1, x2, self.logit_scale, scale_max=cap,
                               normalized=bool(kwargs.get("normalized", False)), precision=precision, group=group,
                               logit_scale_grad="local" if ddp else "sum", segment_key=self._segment_key)
        return loss * dist.get_world_size() if ddp else loss


class _LayerNormF32(nn.LayerNorm):
    """LayerNorm computed in fp32 regardless of the input dtype (what CLIP's LayerNorm does)."""

    def forward(self, x):
        return super().forward(x.float()).to(x.dtype)


@LOSS_HEADS_REGISTRY.register()
class ClassificationHead(LossHead):
    """Linear-probe head whose eval path doubles as the zero-shot scorer (reference :330-419)."""

    def __init__(self, cfg, **kwargs):
        super().__init__()
        self.normalized = False
        assert "output_dim" in kwargs, "`the label number` is not found in `kwargs`"
        nlabel = kwargs["output_dim"]
        self.linear = nn.Sequential(_LayerNormF32(cfg.embed_dim), nn.Linear(cfg.embed_dim, nlabel))
        self.logit_scale = nn.Parameter(torch.ones([]) * np.log(1 / 0.07))
        self.loss_fn = nn.CrossEntropyLoss()
        self.reduce = False

    def copy_state_dict(self, state_dict):
        new_dict = self.state_dict()
        new_dict.update({k: v for k, v in state_dict.items() if k in new_dict})
        self.load_state_dict(new_dict)

    def infer(self, x1, x2, *args, **kwargs):
        if not all(hasattr(self, k) for k in ("audios", "x1s", "x2s", "ids")):
            self.audios, self.x1s, self.x2s, self.ids = [], [], [], []
        self.audios.append(x1)
        self.x1s.append(self.linear(x1).argmax(-1))
        self.x2s.append(x2)
        names = kwargs.get("names", None)
        if names is not None:
            self.ids.extend(names)
        return None

    def report(self, gold_file=None, **kwargs):
        x1s = torch.cat(self.x1s)
        x2s = torch.cat(self.x2s)
        nsample = len(x1s)
        precision = (x1s == x2s).sum() / nsample * 100.
        text = kwargs.get("text", None)
        if text is not None:
            audios = torch.cat(self.audios)
            labels = x2s.unsqueeze(-1)
            # un-normalised similarity + argmax (the reference's normalisation is disabled: `if False and`)
            _, top1, _ = F_.sim_rank_topk(audios, text.to(audios.device), None, topk=1)
            predictions = top1
            label_map = kwargs.get("label_map", None)
            if isinstance(label_map, dict):
                mapped = [label_map[x] for x in predictions.flatten().tolist()]
                predictions = torch.tensor(mapped, device=predictions.device).view(predictions.shape)
            precision = (predictions == labels).sum() / x1s.shape[0] * 100.
        del self.audios, self.x1s, self.x2s, self.ids
        return f"A->T: p1 = {precision:2.2f} @ {nsample}"

    def forward(self, x1, x2, *args, **kwargs):
        """x1: features, x2: integer labels (supervised probe; not part of the InfoNCE hot path)."""
        if not self.training:
            if _is_primary():
                return self.infer(x1, x2, *args, **kwargs)
            return None
        logits = self.logit_scale.exp() * self.linear(x1)
        return self.loss_fn(logits, x2)


class _PairwiseCE(LossHead):
    """Shared machinery of the composite heads: a dict of named CELossHeads plus running loss sums.  In training the enabled
    pairs run as ONE fused multi-pair step (functional.infonce_multi_loss: every modality normalised once, one launch per
    sweep for all pairs) when the shapes allow it, and one after the other otherwise -- same values either way."""

    _pairs = ()          # ((key, attr_name), ...)

    def _build(self, cfg, **kwargs):
        self._total_loss = {}
        for key, attr in self._pairs:
            head = CELossHead(cfg, **kwargs) if getattr(cfg, key) else None
            setattr(self, attr, head)
            if head is not None:
                self._total_loss[key] = 0.

    def copy_state_dict(self, state_dict):
        pass

    def stats(self, nstep=1, **kwargs):
        return " ".join(f"{k} {v / nstep:.3f}" for k, v in self._total_loss.items())

    def _run(self, key, attr, xa, xb, train, *args, **kwargs):
        head = getattr(self, attr)
        if head is None or xa is None or xb is None:
            return 0.
        if not train:
            return head.infer(xa, xb, *args, **kwargs) or 0.
        loss = head(xa, xb, *args, **kwargs)
        self._total_loss[key] += loss.detach()
        return loss

    def _train_terms(self, spec, *args, **kwargs):
        """spec: [(key, attr, xa, xb)] in the reference's order.  Returns the per-pair losses (0. for a disabled pair)."""
        live = [(k, a, xa, xb) for k, a, xa, xb in spec if getattr(self, a) is not None and xa is not None and xb is not None]
        heads = [getattr(self, a) for _, a, _, _ in live]
        fused = (len(live) >= 2 and all(h.precision == "bf16" and not (h.gather and dist.is_initialized() and dist.get_world_size() > 1)
                                        for h in heads))
        if fused:
            feats, index = [], {}
            for _, _, xa, xb in live:
                for x in (xa, xb):
                    if id(x) not in index:
                        index[id(x)] = len(feats)
                        feats.append(x)
            fused = len(feats) <= 5 and len(live) <= 5 and F_.multi_pair_supported(feats)
        out = {}
        if fused:
            losses = F_.infonce_multi_loss(
                feats, [(index[id(xa)], index[id(xb)]) for _, _, xa, xb in live], [h.logit_scale for h in heads],
                [None if h.scale_max == float("inf") else h.scale_max for h in heads],
                normalized=bool(kwargs.get("normalized", False)))
            for i, (key, _, _, _) in enumerate(live):
                self._total_loss[key] += losses[i].detach()
                out[key] = losses[i]
        else:
            for key, attr, xa, xb in live:
                out[key] = self._run(key, attr, xa, xb, True, *args, **kwargs)
        return [out.get(k, 0.) for k, _, _, _ in spec]


@LOSS_HEADS_REGISTRY.register()
class VALCELossHead(_PairwiseCE):
    """vision/audio/language triple: va, lv, al InfoNCE pairs (reference :421-495)."""

    _pairs = (("va", "loss_head_va"), ("lv", "loss_head_lv"), ("al", "loss_head_al"))

    def __init__(self, cfg, **kwargs):
        super().__init__()
        self._build(cfg, **kwargs)

    def _all(self, x1, x2, x3, train, *args, **kwargs):
        if train:
            va, lv, al = self._train_terms([("va", "loss_head_va", x1, x2), ("lv", "loss_head_lv", x1, x3),
                                            ("al", "loss_head_al", x2, x3)], *args, **kwargs)
            return va + lv + al
        return (self._run("va", "loss_head_va", x1, x2, train, *args, **kwargs) +
                self._run("lv", "loss_head_lv", x1, x3, train, *args, **kwargs) +
                self._run("al", "loss_head_al", x2, x3, train, *args, **kwargs))

    def infer(self, x1, x2, x3, *args, **kwargs):
        return self._all(x1, x2, x3, False, *args, **kwargs)

    def report(self, gold_file=None):
        lines = [tag + getattr(self, attr).report(gold_file)
                 for tag, attr in (("VA: ", "loss_head_va"), ("LV: ", "loss_head_lv"), ("AL: ", "loss_head_al"))
                 if getattr(self, attr) is not None]
        return "\n" + "\n".join(lines).strip()

    def forward(self, x1, x2, x3, *args, **kwargs):
        """v: x1; a: x2; l: x3"""
        if not self.training:
            if _is_primary():
                return self.infer(x1, x2, x3, *args, **kwargs)
            return None
        return self._all(x1, x2, x3, True, *args, **kwargs)


@LOSS_HEADS_REGISTRY.register()
class VACELossHead(_PairwiseCE):
    """vp / ap / va / vv / aa weighted InfoNCE pairs (reference :497-598)."""

    _pairs = (("vp", "loss_head_vp"), ("ap", "loss_head_ap"), ("va", "loss_head_va"),
              ("vv", "loss_head_vv"), ("aa", "loss_head_aa"))

    def __init__(self, cfg, **kwargs):
        super().__init__()
        self._build(cfg, **kwargs)
        self.vp_w, self.ap_w, self.va_w, self.vv_w, self.aa_w = cfg.vp_w, cfg.ap_w, cfg.va_w, cfg.vv_w, cfg.aa_w

    def _terms(self, images, images_v1, audios_v1, images_v2, audios_v2, train, *args, **kwargs):
        if train:
            return tuple(self._train_terms([("vp", "loss_head_vp", images_v1, images), ("ap", "loss_head_ap", audios_v1, images),
                                            ("va", "loss_head_va", images_v1, audios_v1), ("vv", "loss_head_vv", images_v1, images_v2),
                                            ("aa", "loss_head_aa", audios_v1, audios_v2)], *args, **kwargs))
        return (self._run("vp", "loss_head_vp", images_v1, images, train, *args, **kwargs),
                self._run("ap", "loss_head_ap", audios_v1, images, train, *args, **kwargs),
                self._run("va", "loss_head_va", images_v1, audios_v1, train, *args, **kwargs),
                self._run("vv", "loss_head_vv", images_v1, images_v2, train, *args, **kwargs),
                self._run("aa", "loss_head_aa", audios_v1, audios_v2, train, *args, **kwargs))

    def infer(self, images, images_v1, audios_v1, images_v2=None, audios_v2=None, *args, **kwargs):
        return sum(self._terms(images, images_v1, audios_v1, images_v2, audios_v2, False, *args, **kwargs))

    def report(self, gold_file=None):
        lines = []
        for tag, attr in (("VP: ", "loss_head_vp"), ("AP: ", "loss_head_ap"), ("VA: ", "loss_head_va"),
                          ("VV: ", "loss_head_vv"), ("AA: ", "loss_head_aa")):
            head = getattr(self, attr)
            if head is not None and hasattr(head, "x1s"):
                lines.append(tag + head.report(gold_file))
        return "\n" + "\n".join(lines).strip()

    def forward(self, images, images_v1, audios_v1, images_v2=None, audios_v2=None, *args, **kwargs):
        if not self.training:
            if _is_primary():
                return self.infer(images, images_v1, audios_v1, images_v2, audios_v2, *args, **kwargs)
            return None
        vp, ap, va, vv, aa = self._terms(images, images_v1, audios_v1, images_v2, audios_v2, True, *args, **kwargs)
        return self.vp_w * vp + self.ap_w * ap + self.va_w * va + self.vv_w * vv + self.aa_w * aa


def install_into_reference(ref_module):
    """Swap the reference's heads for the B200 ones in an imported reference ``loss_head`` module.

    After this call ``ref_module.build_loss_head(cfg)`` (and the composite heads, which instantiate
    ``CELossHead`` from the module namespace) construct the classes of this file.  See INTEGRATION.md.
    """
    for cls in (CELossHead, ClassificationHead, VALCELossHead, VACELossHead):
        setattr(ref_module, cls.__name__, cls)
        reg = getattr(ref_module, "LOSS_HEADS_REGISTRY", None)
        store = getattr(reg, "_obj_map", None)
        if isinstance(store, dict):
            store[cls.__name__] = cls
    return ref_module
