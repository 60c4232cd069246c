"""Stand-in for ``lightning.LightningModule`` (lightning is not installed in this image).

Provides what the reference's modules use: ``log``, ``save_hyperparameters``, ``hparams``, ``freeze``,
``load_from_checkpoint``.  When ``lightning`` is importable the real class is the base, but ``log`` / ``logged`` and
``load_from_checkpoint`` below stay in force: ``trainer.py`` reads ``module.logged`` and logs outside a ``Trainer``, and the
checkpoint loader rebuilds ``dummy_batch`` (a torch_geometric object in reference checkpoints) without torch_geometric."""
import inspect

import torch
from torch import nn

try:  # pragma: no cover - not available offline
    import lightning as L
    _Base = L.LightningModule
    HAVE_LIGHTNING = True
except Exception:
    HAVE_LIGHTNING = False

    class _Base(nn.Module):
        def __init__(self):
            super().__init__()
            self.hparams = {}

        def save_hyperparameters(self, *args, **kwargs):
            frame = inspect.currentframe().f_back
            names = [p for p in inspect.signature(type(self).__init__).parameters if p != "self"]
            self.hparams = {n: frame.f_locals[n] for n in names if n in frame.f_locals}

        def freeze(self):
            for p in self.parameters():
                p.requires_grad = False
            self.eval()

        @property
        def device(self):
            try:
                return next(self.parameters()).device
            except StopIteration:
                return torch.device("cpu")


class LightningModule(_Base):
    def __init__(self):
        super().__init__()
        self.logged = {}

    def log(self, name, value, on_step=False, on_epoch=True, **kw):
        self.logged[name] = value
        if HAVE_LIGHTNING and getattr(self, "_trainer", None) is not None:      # pragma: no cover - attached to a real Trainer
            super().log(name, value, on_step=on_step, on_epoch=on_epoch, **kw)

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, map_location=None, strict=True, **overrides):
        from ..checkpoint import embedded_batch, load_checkpoint
        from ..synthetic import HeteroBatch
        ck = load_checkpoint(checkpoint_path, map_location or "cpu")
        hp = dict(ck.get("hyper_parameters", {}))
        eb = embedded_batch(ck)
        if eb is not None:
            x, ei, y = eb
            B = x["base"].shape[0] if "base" in x else 1
            hp["dummy_batch"] = HeteroBatch(x, ei, y, B)
        elif isinstance(hp.get("dummy_batch"), dict) and "x" in hp["dummy_batch"]:      # train_model's own checkpoints
            d = hp["dummy_batch"]
            hp["dummy_batch"] = HeteroBatch(d["x"], d["edge_index"], d["y"], d["batch_size"])
        hp.update(overrides)
        names = [p for p in inspect.signature(cls.__init__).parameters if p != "self"]
        obj = cls(**{k: v for k, v in hp.items() if k in names})
        obj.load_state_dict(ck["state_dict"], strict=strict)
        return obj
