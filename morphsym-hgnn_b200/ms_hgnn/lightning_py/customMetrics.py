"""Epoch-accumulated metrics with torchmetrics-style semantics, device-resident (no host sync).

Mirrors the reference's ``customMetrics.py`` (``CrossEntropyLossMetric`` L6-25, ``BinaryF1Score``
L27-54, ``CosineSimilarityMetric`` L56-91) and the three torchmetrics regressions it uses.  Calling a
metric (``metric(preds, target)``) returns the value of THIS batch and accumulates the epoch state;
``compute()`` returns the epoch value; ``reset()`` clears it.  torchmetrics / sklearn are not
needed: the confusion counts are four device-side sums.
"""
import torch
from torch import nn


class _Metric(nn.Module):
    def __init__(self):
        super().__init__()
        self._state = {}

    def _acc(self, name, value):
        v = value.detach().double()
        self._state[name] = v if name not in self._state else self._state[name] + v

    def reset(self):
        self._state = {}

    def forward(self, preds, target):
        batch = self.update(preds, target)
        return batch


class MeanSquaredError(_Metric):
    def __init__(self, squared: bool = True):
        super().__init__()
        self.squared = squared

    def update(self, preds, target):
        d = preds.reshape(-1) - target.reshape(-1)
        sse = (d * d).sum()
        self._acc("sse", sse); self._acc("n", torch.tensor(float(d.numel()), device=d.device))
        mse = sse / d.numel()
        return mse if self.squared else torch.sqrt(mse)

    def compute(self):
        mse = self._state["sse"] / self._state["n"]
        return mse if self.squared else torch.sqrt(mse)


class MeanAbsoluteError(_Metric):
    def update(self, preds, target):
        d = (preds.reshape(-1) - target.reshape(-1)).abs()
        self._acc("sae", d.sum()); self._acc("n", torch.tensor(float(d.numel()), device=d.device))
        return d.mean()

    def compute(self):
        return self._state["sae"] / self._state["n"]


class CrossEntropyLossMetric(_Metric):
    """sum-reduced CE over the rows, epoch value = summed_loss.float() / total_num (customMetrics.py:L25)."""

    def update(self, preds, target):
        if preds.size(0) != target.size(0):
            raise ValueError("Both tensors must have the same number of batches.")
        s = torch.nn.functional.cross_entropy(preds, target, reduction="sum")
        self._acc("sum", s); self._acc("n", torch.tensor(float(preds.shape[0]), device=preds.device))
        return s.float() / preds.shape[0]

    def accumulate_value(self, batch_mean: torch.Tensor, n_rows: int):
        """Accumulate a batch mean computed elsewhere (the native fused CE kernel)."""
        self._acc("sum", batch_mean.detach().double() * n_rows)
        self._acc("n", torch.tensor(float(n_rows), device=batch_mean.device))

    def compute(self):
        return self._state["sum"].float() / self._state["n"]


class MulticlassAccuracy(_Metric):
    def update(self, preds, target):
        ok = (preds == target).sum()
        self._acc("ok", ok); self._acc("n", torch.tensor(float(preds.numel()), device=preds.device))
        return ok.double() / preds.numel()

    def compute(self):
        return self._state["ok"] / self._state["n"]


class BinaryF1Score(_Metric):
    def update(self, preds, target):
        if preds.size(0) != target.size(0):
            raise ValueError("Both tensors must have the same number of batches.")
        p = preds.reshape(-1) != 0
        t = target.reshape(-1) != 0
        tp = (p & t).sum(); fp = (p & ~t).sum(); fn = (~p & t).sum()
        self._acc("tp", tp); self._acc("fp", fp); self._acc("fn", fn)
        return self._f1(tp.double(), fp.double(), fn.double())

    @staticmethod
    def _f1(tp, fp, fn):
        precision = tp / (tp + fp)
        recall = tp / (tp + fn)
        return torch.nan_to_num(2 * (precision * recall) / (precision + recall))

    def compute(self):
        return self._f1(self._state["tp"], self._state["fp"], self._state["fn"])


class CosineSimilarityMetric(_Metric):
    def update(self, preds, target):
        if preds.size() != target.size():
            raise ValueError("Prediction and target tensors must have the same shape")
        sim = torch.nn.functional.cosine_similarity(preds, target, dim=1)
        self._acc("sum", sim.sum()); self._acc("n", torch.tensor(float(preds.shape[0]), device=preds.device))
        return sim.mean()

    def compute(self):
        return self._state["sum"] / self._state["n"]
