"""Checkpoint compatibility with the reference.

The reference trains with Lightning; its `.ckpt` files (README.md:146-151 publishes one per model) are
`torch.save`d dicts whose "state_dict" holds the LightningModule's parameters and buffers under the names the model
classes register (`backbone.unet.1.blocks.block0.conv_branch.2.kernel`, `...bn.running_mean`, `score_net...`).
The harness models register exactly the same names and shapes (tests/golden/reference_state_dicts.json is produced
from the reference's own classes), so loading is a strict `load_state_dict` plus validation with readable errors.
"""
import torch


def _load_file(path, trust_pickle):
    """torch.load with the safe unpickler first: a published `.ckpt` is a download from the internet and the
    plain unpickler executes arbitrary code.  Lightning checkpoints carry hyper-parameter containers (omegaconf)
    next to the tensors; those classes are allow-listed when importable.  Only `trust_pickle=True` -- an explicit
    statement by the caller -- falls back to the unrestricted unpickler."""
    try:
        return torch.load(path, map_location="cpu", weights_only=True)
    except Exception as first:  # noqa: BLE001 -- UnpicklingError and friends; retried below
        allow = []
        for mod, names in (("omegaconf.dictconfig", ["DictConfig"]), ("omegaconf.listconfig", ["ListConfig"]),
                           ("omegaconf.base", ["ContainerMetadata", "Metadata"]), ("omegaconf.nodes", ["AnyNode"])):
            try:
                m = __import__(mod, fromlist=names)
                allow += [getattr(m, n) for n in names if hasattr(m, n)]
            except ImportError:
                pass
        if allow:
            try:
                with torch.serialization.safe_globals(allow):
                    return torch.load(path, map_location="cpu", weights_only=True)
            except Exception:  # noqa: BLE001
                pass
        if trust_pickle:
            return torch.load(path, map_location="cpu", weights_only=False)
        raise ValueError("checkpoint %r needs the unrestricted unpickler (%s); pass trust_pickle=True only for "
                         "files you trust" % (path, first)) from first


def reference_state_dict(ckpt, trust_pickle=False):
    """The tensor dict inside a Lightning checkpoint (or the dict itself when it already is a state dict)."""
    if isinstance(ckpt, (str, bytes)) or hasattr(ckpt, "__fspath__"):
        ckpt = _load_file(ckpt, trust_pickle)
    if isinstance(ckpt, dict) and "state_dict" in ckpt and isinstance(ckpt["state_dict"], dict):
        ckpt = ckpt["state_dict"]
    if not isinstance(ckpt, dict) or not all(torch.is_tensor(v) for v in ckpt.values()):
        raise ValueError("not a checkpoint: expected a Lightning .ckpt dict with a 'state_dict' or a plain state dict")
    return ckpt


def load_reference_checkpoint(model, ckpt, strict=True, trust_pickle=False):
    """Load a reference `.ckpt` (path, Lightning dict or state dict) into a harness model.

    Raises ValueError naming every missing / unexpected key and every shape mismatch (strict=True), so that a
    checkpoint of a different model family or channel width fails before any tensor is copied.
    Returns (missing_keys, unexpected_keys) like `nn.Module.load_state_dict`.
    """
    state = reference_state_dict(ckpt, trust_pickle)
    own = model.state_dict()
    missing = [k for k in own if k not in state]
    unexpected = [k for k in state if k not in own]
    wrong = ["%s: checkpoint %s, model %s" % (k, tuple(state[k].shape), tuple(own[k].shape))
             for k in own if k in state and tuple(state[k].shape) != tuple(own[k].shape)]
    if wrong or (strict and (missing or unexpected)):
        raise ValueError("checkpoint does not fit the model: %d missing %s, %d unexpected %s, %d shape mismatches %s"
                         % (len(missing), missing[:4], len(unexpected), unexpected[:4], len(wrong), wrong[:4]))
    model.load_state_dict({k: v for k, v in state.items() if k in own}, strict=False)
    return missing, unexpected


def save_reference_checkpoint(model, path, epoch=0, global_step=0):
    """Write the model in the layout the reference's `test.py` expects from `ckpt_path` (Lightning dict)."""
    torch.save({"epoch": epoch, "global_step": global_step, "pytorch-lightning_version": "2.0.0",
                "state_dict": {k: v.detach().cpu() for k, v in model.state_dict().items()}}, path)
