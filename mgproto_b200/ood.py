"""OoD scoring of the reference's test loop (``train_and_test._testing_with_OoD``, train_and_test.py:163-242) on the device.

The reference pulls every batch's ``output[:, :, 0].exp()`` to the host, concatenates, and calls numpy for the
percentile; here the per-image statistics come from one kernel per batch (``mgp_ood_score`` on the level-0 log
evidences of ``MGProto.head_level0`` -- no top-T mining, no log p matrix), stay on the GPU, and the threshold / FPR95
/ AUROC are a sort and a scan at the end (one host read for the final numbers).
"""
from __future__ import annotations

import torch

from . import ops


class OoDScorer:
    """Accumulates in-distribution and out-of-distribution batches; mirrors the reference's quirk of thresholding on
    the SUM over classes of p(x|c) (5th percentile of the in-distribution set) while testing OoD images on the MEAN
    (train_and_test.py:199 vs :213)."""

    def __init__(self, model):
        self.model = model
        self.id_sum, self.id_mean, self.id_pred, self.id_label = [], [], [], []
        self.ood_sum, self.ood_mean = [], []

    @torch.no_grad()
    def _score(self, x_add):
        out0 = self.model.head_level0(x_add)                       # [B, C] level-0 log evidences
        return ops.ood_score(out0)

    @torch.no_grad()
    def add_in_distribution(self, x_add, labels=None):
        s, m, pred = self._score(x_add)
        self.id_sum.append(s)
        self.id_mean.append(m)
        self.id_pred.append(pred)
        if labels is not None:
            self.id_label.append(labels.to(pred.device))

    @torch.no_grad()
    def add_out_of_distribution(self, x_add):
        s, m, _ = self._score(x_add)
        self.ood_sum.append(s)
        self.ood_mean.append(m)

    @torch.no_grad()
    def results(self, percentile=5.0):
        """-> dict(threshold, FPR95 (reference definition), AUROC of the sum score (ID positive), accuracy | None)."""
        ids, oods, oodm = torch.cat(self.id_sum), torch.cat(self.ood_sum), torch.cat(self.ood_mean)
        thr = percentile_linear(ids, percentile)                    # np.percentile(prob_sum_over_c, 5), :199
        fpr95 = (oodm > thr).float().mean()                         # :213, :216
        auroc = auroc_rank(ids, oods)
        acc = None
        if self.id_label:
            acc = (torch.cat(self.id_pred) == torch.cat(self.id_label)).float().mean().item()
        return {"threshold": float(thr), "FPR95": float(fpr95), "AUROC": float(auroc), "accuracy": acc,
                "n_in": int(ids.numel()), "n_ood": int(oods.numel())}


def percentile_linear(x: torch.Tensor, q: float) -> torch.Tensor:
    """numpy.percentile(x, q) with its default linear interpolation, on the device."""
    v, _ = torch.sort(x.double())
    pos = (v.numel() - 1) * (q / 100.0)
    lo = int(pos)
    hi = min(lo + 1, v.numel() - 1)
    return v[lo] + (v[hi] - v[lo]) * (pos - lo)


def auroc_rank(pos: torch.Tensor, neg: torch.Tensor) -> torch.Tensor:
    """Area under the ROC curve of `score > t => positive` from average ranks (equals
    sklearn.metrics.roc_auc_score with ties counted half), on the device."""
    s = torch.cat([pos, neg]).double()
    n = s.numel()
    v, order = torch.sort(s)
    # average rank of each tie group
    first = torch.ones(n, dtype=torch.bool, device=s.device)
    first[1:] = v[1:] != v[:-1]
    gid = torch.cumsum(first.long(), 0) - 1
    cnt = torch.bincount(gid).double()
    start = torch.cumsum(cnt, 0) - cnt
    avg = start + (cnt + 1) / 2.0                                   # 1-based average rank of the group
    rank = torch.empty(n, dtype=torch.float64, device=s.device)
    rank[order] = avg[gid]
    np_, nn_ = pos.numel(), neg.numel()
    return (rank[:np_].sum() - np_ * (np_ + 1) / 2.0) / (np_ * nn_)
