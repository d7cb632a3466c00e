"""Generates tests/golden/reference_state_dicts.json: parameter / buffer names and shapes of the REFERENCE's own
PointGroup, HAIS and SoftGroup modules (checkpoint compatibility).

The reference classes are instantiated from /root/reference with its own YAML configs.  MinkowskiEngine resolves to
the drop-in (minsu3d_b200.install_as_reference_modules); pytorch_lightning and hydra are not installed and are
stubbed for construction only (LightningModule -> nn.Module: same registration of parameters and buffers, hence the
same state_dict keys as a Lightning `.ckpt` holds under "state_dict").

    python tests/golden/make_state_dict_golden.py        # CPU only, needs /root/reference
"""
import json
import os
import sys
import types

import torch.nn as nn
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("MINSU3D_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)


class _Ns(dict):
    """Attribute access over nested dicts (what the reference does with its OmegaConf cfg)."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError:
            raise AttributeError(k)
        return _Ns(v) if isinstance(v, dict) else v


class _LightningModule(nn.Module):
    def save_hyperparameters(self, *a, **k):
        pass

    def log(self, *a, **k):
        pass


def main():
    pl = types.ModuleType("pytorch_lightning")
    pl.LightningModule = _LightningModule
    sys.modules["pytorch_lightning"] = pl
    sys.modules["hydra"] = types.ModuleType("hydra")
    import minsu3d_b200
    minsu3d_b200.install_as_reference_modules()
    import minsu3d.model as ref_models

    with open(os.path.join(REF, "config", "data", "scannetv2.yaml")) as f:
        data_cfg = yaml.safe_load(f)
    out = {}
    for name, cls in (("pointgroup", ref_models.PointGroup), ("hais", ref_models.HAIS), ("softgroup", ref_models.SoftGroup)):
        with open(os.path.join(REF, "config", "model", name + ".yaml")) as f:
            model_cfg = yaml.safe_load(f)
        cfg = _Ns({"model": model_cfg, "data": data_cfg})
        model = cls(cfg)
        out[name] = {k: list(v.shape) for k, v in model.state_dict().items()}
        print(name, len(out[name]), "entries,", sum(p.numel() for p in model.parameters()), "parameters")
    with open(os.path.join(HERE, "reference_state_dicts.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)


if __name__ == "__main__":
    main()
