"""Loads the REFERENCE's own model classes (staged by oracle/build_ref.py under oracle/_ref/refpy, or straight from
/root/reference where that exists) on top of the drop-in modules.  Test infrastructure only.

pytorch_lightning and hydra are not installed in this image: `LightningModule` is stubbed by an nn.Module that
provides the four members the reference's forward/_loss touch (`hparams.cfg`, `current_epoch`, `device`, `log`).
Nothing of the reference's model code is modified."""
import inspect
import os
import sys
import types

import torch
import torch.nn as nn
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Ns(dict):
    """Attribute access over nested dicts (what the reference does with its OmegaConf cfg)."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError:
            raise AttributeError(k)
        return Ns(v) if isinstance(v, dict) else v


class _LightningModule(nn.Module):
    current_epoch = 10_000  # > prepare_epochs: the clustering stage is active

    def save_hyperparameters(self, *a, **k):
        frame = inspect.currentframe().f_back
        cfg = frame.f_locals.get("cfg")
        if cfg is not None:
            object.__setattr__(self, "hparams", Ns({"cfg": cfg}))

    def log(self, *a, **k):
        pass

    def print(self, *a, **k):
        pass

    @property
    def device(self):
        return next(self.parameters()).device


def reference_root():
    staged = os.path.join(ROOT, "oracle", "_ref", "refpy")
    if os.path.exists(os.path.join(staged, "minsu3d", "model", "pointgroup.py")):
        return staged
    if os.path.exists("/root/reference/minsu3d/model/pointgroup.py"):
        return "/root/reference"
    return None


def load_reference_models():
    """-> (module minsu3d.model of the reference, cfg factory) or None when the reference python is not available."""
    ref = reference_root()
    if ref is None:
        return None
    pl = types.ModuleType("pytorch_lightning")
    pl.LightningModule = _LightningModule
    sys.modules["pytorch_lightning"] = pl
    sys.modules.setdefault("hydra", types.ModuleType("hydra"))
    import minsu3d_b200
    minsu3d_b200.install_as_reference_modules()
    for name in [m for m in sys.modules if m == "minsu3d" or m.startswith("minsu3d.")]:
        del sys.modules[name]
    if ref not in sys.path:
        sys.path.insert(0, ref)
    import minsu3d.model as ref_models

    def make_cfg(name):
        with open(os.path.join(ref, "config", "data", "scannetv2.yaml")) as f:
            data_cfg = yaml.safe_load(f)
        with open(os.path.join(ref, "config", "model", name + ".yaml")) as f:
            model_cfg = yaml.safe_load(f)
        return Ns({"model": model_cfg, "data": data_cfg})

    return ref_models, make_cfg
