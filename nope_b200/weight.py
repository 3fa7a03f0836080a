"""`load_checkpoint` of the reference (src/utils/weight.py:6-37): load a checkpoint into a model,
keeping only the entries whose key exists in the model WITH THE SAME SHAPE (after removing `prefix`
from every key); everything else keeps the model's current value.  Works on the mirrors in this
package (`nope_b200.unet.UNet`, `nope_b200.model.PoseConditional`), whose `state_dict()` /
`load_state_dict()` speak the reference's key schema."""
import logging

import torch


def load_checkpoint(model, checkpoint_path, checkpoint_key=None, prefix=""):
    """checkpoint_path: a file for torch.load, or an already loaded dict.  Returns
    (loaded keys, keys the checkpoint has but the model cannot take, model keys not updated)."""
    if isinstance(checkpoint_path, dict):
        checkpoint = checkpoint_path
    else:
        # Lightning checkpoints carry hyper-parameters / optimizer state next to the tensors:
        # they need the full unpickler (only load files you trust)
        checkpoint = torch.load(checkpoint_path, map_location="cpu", weights_only=False)
    pretrained = checkpoint[checkpoint_key] if checkpoint_key is not None else checkpoint
    pretrained = {k.replace(prefix, ""): v for k, v in pretrained.items()}       # weight.py:13
    model_dict = model.state_dict()
    can_load = {k: v for k, v in pretrained.items()
                if k in model_dict and tuple(v.shape) == tuple(model_dict[k].shape)}
    cannot_load = [k for k in pretrained if k not in can_load]
    not_updated = [k for k in model_dict if k not in pretrained]
    logging.info("Cannot load: %s", sorted({k.split(".")[0] for k in cannot_load}))
    logging.info("Not update: %s", sorted({k.split(".")[0] for k in not_updated}))
    logging.info("Pretrained: %d/ Loaded: %d/ Cannot loaded: %d VS Current model: %d",
                 len(pretrained), len(can_load), len(cannot_load), len(model_dict))
    model_dict.update(can_load)
    model.load_state_dict(model_dict)
    return sorted(can_load), cannot_load, not_updated
