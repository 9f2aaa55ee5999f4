"""The reference's checkpoint wire format (ref runners/runner_utils.py:736-831).

A checkpoint is ONE file ``ckpt_{epoch:06d}.pth`` written by ``torch.save`` holding the dict

    {"epoch": int,
     "optimizer_state_dict": torch.optim.Adam(...).state_dict()   # 44 parameters: coarse net, then fine net
     "scheduler_state_dict": ExponentialLR.state_dict(),          # only if a scheduler is used
     "scene_default": coarse NeRF.state_dict(),                   # keys fc_in.weight ... fc_out.bias
     "scene_fine":    fine NeRF.state_dict()}                     # only if a fine network is used

The hot path keeps both networks in one flat fp32 buffer stepped by one Adam "parameter" (engine.FlatParams), so the
functions below translate between that flat optimizer state and the reference's per-tensor one: a checkpoint written
here loads into the reference's ``_load_ckpt`` (44-parameter Adam) and vice versa.  Pure host code, no CUDA needed."""
from __future__ import annotations

import os
from pathlib import Path
from typing import Optional, Sequence

import torch

from .network import NeRF


def ckpt_path(ckpt_dir, epoch: int) -> str:
    """runner_utils.py:756: ckpt_{epoch zero-filled to 6}.pth"""
    return os.path.join(str(ckpt_dir), f"ckpt_{str(epoch).zfill(6)}.pth")


def _per_tensor_params(nets: Sequence[NeRF]):
    # runner_utils.py:684-686: list(default.parameters()) + list(fine.parameters()) = registration order per net
    return [p for net in nets for p in net.ordered_parameters()]


def reference_optimizer_state(optimizer: torch.optim.Optimizer, flat, nets: Sequence[NeRF]) -> dict:
    """state_dict() of the reference's Adam over the 44 parameter tensors, built from the flat optimizer's state."""
    sd = optimizer.state_dict()
    group = dict(sd["param_groups"][0])
    params = _per_tensor_params(nets)
    group["params"] = list(range(len(params)))
    group["fused"] = None  # the reference steps 44 separate tensors; "fused" describes this process, not the state
    state = {}
    flat_state = sd["state"].get(0)
    if flat_state is not None:
        off = 0
        for i, p in enumerate(params):
            n = p.numel()
            entry = {}
            for k, v in flat_state.items():
                if torch.is_tensor(v) and v.numel() == flat.flat.numel():
                    entry[k] = v.detach().reshape(-1)[off:off + n].reshape(p.shape).clone().cpu()
                elif torch.is_tensor(v):
                    entry[k] = v.detach().clone().cpu()  # step
                else:
                    entry[k] = v
            state[i] = entry
            off += n
    return {"state": state, "param_groups": [group]}


def load_reference_optimizer_state(optimizer: torch.optim.Optimizer, flat, nets: Sequence[NeRF], ref_sd: dict) -> None:
    """Loads a 44-parameter Adam state_dict (as the reference writes it) into the flat optimizer."""
    params = _per_tensor_params(nets)
    group = {k: v for k, v in ref_sd["param_groups"][0].items() if k not in ("params", "fused", "foreach")}
    cur = optimizer.state_dict()
    new_group = dict(cur["param_groups"][0])
    new_group.update(group)
    new_group["params"] = [0]
    state = {}
    if len(ref_sd["state"]) > 0:
        if sorted(ref_sd["state"].keys()) != list(range(len(params))):
            raise ValueError("optimizer state does not cover the 44 parameter tensors of (coarse, fine)")
        keys = [k for k, v in ref_sd["state"][0].items() if torch.is_tensor(v) and v.dim() > 0 and k != "step"]
        entry = {}
        for k in keys:
            pieces = []
            for i, p in enumerate(params):
                t = ref_sd["state"][i][k]
                if tuple(t.shape) != tuple(p.shape):
                    raise ValueError(f"optimizer state '{k}' of parameter {i} has shape {tuple(t.shape)}, expected {tuple(p.shape)}")
                pieces.append(t.reshape(-1).to(torch.float32))
            entry[k] = torch.cat(pieces).to(flat.flat.device)
        steps = {float(ref_sd["state"][i]["step"]) for i in range(len(params))}
        if len(steps) != 1:
            raise ValueError("the per-parameter step counters differ; cannot represent them with one flat parameter")
        # torch 1.11 (the reference's pin) stores an int, current torch a float32 scalar tensor
        entry["step"] = torch.tensor(steps.pop(), dtype=torch.float32)
        state[0] = entry
    optimizer.load_state_dict({"state": state, "param_groups": [new_group]})


def save_ckpt(ckpt_dir, epoch: int, coarse: NeRF, fine: Optional[NeRF], optimizer: torch.optim.Optimizer,
              scheduler=None, flat=None) -> str:
    """runner_utils.py:736-783.  `flat` = engine.FlatParams when the optimizer steps the flat buffer."""
    os.makedirs(str(ckpt_dir), exist_ok=True)
    nets = [coarse] + ([fine] if fine is not None else [])
    ckpt = {"epoch": epoch}
    ckpt["optimizer_state_dict"] = reference_optimizer_state(optimizer, flat, nets) if flat is not None else optimizer.state_dict()
    if scheduler is not None:
        ckpt["scheduler_state_dict"] = scheduler.state_dict()
    ckpt["scene_default"] = {k: v.detach().clone().cpu() for k, v in coarse.state_dict().items()}
    if fine is not None:
        ckpt["scene_fine"] = {k: v.detach().clone().cpu() for k, v in fine.state_dict().items()}
    path = ckpt_path(ckpt_dir, epoch)
    torch.save(ckpt, path)
    return path


def load_ckpt(ckpt_dir, coarse: NeRF, fine: Optional[NeRF], optimizer: Optional[torch.optim.Optimizer] = None,
              scheduler=None, flat=None) -> int:
    """runner_utils.py:786-831: loads the LATEST checkpoint of the directory (sorted file names) and returns its epoch;
    0 when the directory is missing or empty."""
    if ckpt_dir is None or not Path(ckpt_dir).exists():
        return 0
    files = sorted(Path(ckpt_dir).iterdir())
    if len(files) == 0:
        return 0
    ckpt = torch.load(files[-1], map_location="cpu")
    nets = [coarse] + ([fine] if fine is not None else [])
    with torch.no_grad():
        for net, key in zip(nets, ("scene_default", "scene_fine")):
            sd = ckpt[key]
            own = net.state_dict()
            if set(sd.keys()) != set(own.keys()):
                raise KeyError(f"{key}: state_dict keys differ from NeRF's ({sorted(set(sd) ^ set(own))})")
            for k, v in sd.items():  # copy in place: the parameters may alias the flat buffer
                own[k].copy_(v)
    if optimizer is not None:
        if flat is not None:
            load_reference_optimizer_state(optimizer, flat, nets, ckpt["optimizer_state_dict"])
        else:
            optimizer.load_state_dict(ckpt["optimizer_state_dict"])
        if scheduler is not None:
            scheduler.load_state_dict(ckpt["scheduler_state_dict"])
    return int(ckpt["epoch"])
