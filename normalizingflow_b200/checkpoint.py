"""Checkpoint compatibility with the reference.

The reference's training loop writes ``{"model", "optim", "scheduler", "epoch", "loss"}`` with
``torch.save`` (applications/src/train.py:39-40) and ``load_model`` reads it back with
``map_location='cpu'`` and ``load_state_dict(..., strict=False)`` (applications/src/setup.py:102-109).
The layer classes here keep the reference's state-dict keys, so those files load unchanged."""
from __future__ import annotations

import torch

KEYS = ("model", "optim", "scheduler", "epoch", "loss")


def save_checkpoint(path, model, optimizer=None, scheduler=None, epoch=0, losses=()):
    """Write the dict of applications/src/train.py:39-40."""
    torch.save({"model": model.state_dict(),
                "optim": optimizer.state_dict() if optimizer is not None else {},
                "scheduler": scheduler.state_dict() if scheduler is not None else {},
                "epoch": int(epoch), "loss": list(losses)}, path)


def load_checkpoint(path, model, optimizer=None, scheduler=None, strict=False, device=None, trusted=False):
    """Load a reference-format checkpoint (applications/src/setup.py:102-109).  Returns
    (epoch, losses, load_state_dict result).  ``strict=False`` as in the reference; the returned
    result lists missing / unexpected keys so a caller can insist on an exact match.

    The reference's checkpoints hold only tensors, dicts, lists and numbers, so the file is read with
    ``weights_only=True`` (no arbitrary unpickling).  ``trusted=True`` allows the full unpickler for a
    file from a trusted source that fails the restricted load."""
    try:
        blob = torch.load(path, map_location="cpu", weights_only=True)
    except Exception as e:                              # noqa: BLE001 - pickle / torch raise several types
        if not trusted:
            raise ValueError(f"{path}: cannot be read with weights_only=True ({type(e).__name__}: {e}); pass "
                             "trusted=True to unpickle arbitrary objects from a file you trust") from e
        blob = torch.load(path, map_location="cpu", weights_only=False)
    if not isinstance(blob, dict) or "model" not in blob:
        raise ValueError(f"{path}: not a reference checkpoint (expected a dict with keys {KEYS})")
    result = model.load_state_dict(blob["model"], strict=strict)
    if device is not None:
        model.to(device)
    if optimizer is not None and blob.get("optim"):
        optimizer.load_state_dict(blob["optim"])
    if scheduler is not None and blob.get("scheduler"):
        scheduler.load_state_dict(blob["scheduler"])
    return int(blob.get("epoch", 0)), list(blob.get("loss", [])), result
