"""Checkpoint loading with the reference's rules (det3d/torchie/trainer/checkpoint.py:67-137,166-218), local files
only: `state_dict` wrapper and `module.` prefix are stripped, sparse-conv weights saved by spconv 1.x
((kH,kW,Cin,Cout)) are re-laid to the 2.x layout this package uses ((Cout,kH,kW,Cin)), unexpected keys / shape
mismatches are reported instead of raised unless strict, `num_batches_tracked` is never reported missing.  After
loading, the cached lowerings are dropped so the next forward re-packs the new weights."""
import os
from collections import OrderedDict

import torch

from . import layers


def _sparse_conv_weight_keys(module, prefix=""):
    """checkpoint.py:47-64: names of the weights of every sparse-conv child"""
    keys = set()
    for name, child in module.named_modules():
        if isinstance(child, layers._SparseConvBase):
            keys.add((name + "." if name else "") + "weight")
    return keys


def load_state_dict(module, state_dict, strict=False, logger=None):
    own = module.state_dict()
    sp_keys = _sparse_conv_weight_keys(module)
    unexpected, mismatched = [], []
    for name, param in state_dict.items():
        if name in sp_keys and name in own and own[name].shape != param.shape:
            native = param.transpose(-1, -2)                       # (.., Cin, Cout) -> (.., Cout, Cin)
            if native.shape == own[name].shape:
                param = native.contiguous()
            else:
                implicit = param.permute(param.dim() - 1, *range(param.dim() - 1))   # -> (Cout, k.., Cin)
                if implicit.shape == own[name].shape:
                    param = implicit.contiguous()
        if name not in own:
            unexpected.append(name)
            continue
        if isinstance(param, torch.nn.Parameter):
            param = param.data
        if param.size() != own[name].size():
            mismatched.append((name, tuple(own[name].size()), tuple(param.size())))
            continue
        own[name].copy_(param)
    missing = [k for k in set(own.keys()) - set(state_dict.keys()) if "num_batches_tracked" not in k]
    layers.invalidate(module)
    msgs = []
    if unexpected:
        msgs.append("unexpected key in source state_dict: {}\n".format(", ".join(unexpected)))
    if missing:
        msgs.append("missing keys in source state_dict: {}\n".format(", ".join(sorted(missing))))
    if mismatched:
        msgs.append("these keys have mismatched shape:\n" + "\n".join(
            f"{k}: expected {a}, loaded {b}" for k, a, b in mismatched))
    if msgs:
        text = "The model and loaded state dict do not match exactly\n" + "\n".join(msgs)
        if strict:
            raise RuntimeError(text)
        if logger is not None:
            logger.warning(text)
        else:
            print(text)
    return dict(unexpected=unexpected, missing=missing, mismatched=mismatched)


def load_checkpoint(model, filename, map_location=None, strict=False, logger=None):
    if "://" in filename:
        raise IOError("only local checkpoint files are supported (no network on the target machines)")
    if not os.path.isfile(filename):
        raise IOError("{} is not a checkpoint file".format(filename))
    checkpoint = torch.load(filename, map_location=map_location, weights_only=False)
    if isinstance(checkpoint, OrderedDict):
        state_dict = checkpoint
    elif isinstance(checkpoint, dict) and "state_dict" in checkpoint:
        state_dict = checkpoint["state_dict"]
    else:
        raise RuntimeError("No state_dict found in checkpoint file {}".format(filename))
    if list(state_dict.keys())[0].startswith("module."):
        state_dict = {k[7:]: v for k, v in state_dict.items()}
    load_state_dict(model.module if hasattr(model, "module") else model, state_dict, strict, logger)
    return checkpoint


def save_checkpoint(model, filename, optimizer=None, meta=None):
    """det3d/torchie/trainer/checkpoint.py:235-262: writes {"meta", "state_dict" (CPU tensors, reference key names and
    spconv-2.x layouts), "optimizer"} so a det3d trainer — or load_checkpoint above — reads the file back unchanged."""
    if meta is None:
        meta = {}
    elif not isinstance(meta, dict):
        raise TypeError("meta must be a dict or None, but got {}".format(type(meta)))
    d = os.path.dirname(filename)
    if d:
        os.makedirs(d, exist_ok=True)
    if hasattr(model, "module"):
        model = model.module
    checkpoint = {"meta": meta,
                  "state_dict": OrderedDict((k, v.detach().cpu()) for k, v in model.state_dict().items())}
    if optimizer is not None:
        checkpoint["optimizer"] = optimizer.state_dict()
    torch.save(checkpoint, filename)
    return filename
