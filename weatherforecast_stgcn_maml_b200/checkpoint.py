"""Meta-training checkpoints with the reference's layout, and a real resume.

The reference saves ``meta_optimizer.state_dict()`` and ``scheduler.state_dict()`` next to the two model state_dicts
(train_hybrid_maml_v5.py:311-335, final :345-370) but never loads them back.  Here the fused AdamW keeps its moments in
flat device buffers (engine.AdamState) and the cosine schedule is a plain object (schedule.CosineWarmRestarts); this module
converts both to and from the dictionaries ``torch.optim.AdamW`` / ``CosineAnnealingWarmRestarts`` produce, so that a file
written here loads into the reference's objects and a file written by the reference resumes here.

Parameter indices follow the reference's optimiser construction (train_hybrid_maml_v5.py:245-249):
``list(hybrid_model.parameters()) + list(koppen_embed.parameters())`` -- the hybrid ``state_dict`` key order (the model has
no buffers) followed by ``embedding.weight``.  Parameters that never receive a gradient (the frozen GCN stack, the
Koppen table: SURVEY.md D5/D10) have no entry in ``state``, exactly as in torch.
"""
from __future__ import annotations

import torch

from .engine import V5Dims, flatten_trainable, trainable_layout, unflatten_trainable

MODEL_VERSION = "5.0"


def optimizer_state_dict(adam, hybrid_keys, dims: V5Dims, num_extra_params=1):
    """``torch.optim.AdamW.state_dict()`` of the reference's meta optimiser from the flat fused state."""
    names = {name for name, _, _ in trainable_layout(dims)}
    keys = list(hybrid_keys)
    state = {}
    if adam.step_count > 0:
        m = unflatten_trainable(adam.exp_avg.detach().cpu(), dims)
        v = unflatten_trainable(adam.exp_avg_sq.detach().cpu(), dims)
        for i, k in enumerate(keys):
            if k in names:
                state[i] = {"step": torch.tensor(float(adam.step_count)), "exp_avg": m[k].clone(), "exp_avg_sq": v[k].clone()}
    group = {"lr": adam.lr, "betas": tuple(adam.betas), "eps": adam.eps, "weight_decay": adam.weight_decay, "amsgrad": False,
             "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
             "decoupled_weight_decay": bool(adam.decoupled), "params": list(range(len(keys) + int(num_extra_params)))}
    return {"state": state, "param_groups": [group]}


def load_optimizer_state_dict(adam, osd, hybrid_keys, dims: V5Dims):
    """Inverse of ``optimizer_state_dict``: moments, step count and hyper-parameters into the fused optimiser."""
    keys = list(hybrid_keys)
    group = osd["param_groups"][0]
    adam.lr, adam.betas, adam.eps = float(group["lr"]), tuple(group["betas"]), float(group["eps"])
    adam.weight_decay = float(group["weight_decay"])
    m, v, steps = {}, {}, set()
    for name, shape, _ in trainable_layout(dims):
        entry = osd["state"].get(keys.index(name))
        if entry is None:
            m[name], v[name] = torch.zeros(shape), torch.zeros(shape)
            continue
        if tuple(entry["exp_avg"].shape) != tuple(shape):
            raise ValueError(f"optimizer state of {name}: shape {tuple(entry['exp_avg'].shape)} != {tuple(shape)}")
        m[name], v[name] = entry["exp_avg"].detach().cpu(), entry["exp_avg_sq"].detach().cpu()
        steps.add(int(float(entry["step"])))
    if len(steps) > 1:
        raise ValueError(f"parameters disagree on the step count: {sorted(steps)}")
    adam.step_count = steps.pop() if steps else 0
    adam.exp_avg.copy_(flatten_trainable(m, dims))
    adam.exp_avg_sq.copy_(flatten_trainable(v, dims))


def scheduler_state_dict(sched):
    """The fields of ``CosineAnnealingWarmRestarts.state_dict()`` that define the schedule."""
    return {"T_0": sched.T_0, "T_i": sched.T_i, "T_mult": sched.T_mult, "eta_min": sched.eta_min, "T_cur": sched.T_cur,
            "base_lrs": [sched.base_lr], "last_epoch": sched.last_epoch, "_last_lr": [sched.lr]}


def load_scheduler_state_dict(sched, ssd):
    sched.T_0, sched.T_i, sched.T_mult = int(ssd["T_0"]), int(ssd["T_i"]), int(ssd["T_mult"])
    sched.eta_min, sched.T_cur, sched.last_epoch = float(ssd["eta_min"]), int(ssd["T_cur"]), int(ssd["last_epoch"])
    sched.base_lr = float(ssd["base_lrs"][0])
    sched.lr = float(ssd["_last_lr"][0]) if "_last_lr" in ssd else sched.base_lr


def meta_checkpoint(trainer, koppen_state_dict, scheduler, epoch, best_loss, lstm_dropout=0.2, final_loss=None):
    """The dict train_hybrid_maml_v5.py:311-335 saves (``final_loss`` added for the final file, :345-370)."""
    d = trainer.dims
    sd = trainer.state_dict()
    total = sum(int(t.numel()) for t in sd.values())
    ckpt = {
        "hybrid_model_state_dict": sd,
        "koppen_embed_state_dict": {k: v.detach().cpu().clone() for k, v in koppen_state_dict.items()},
        "meta_optimizer_state_dict": optimizer_state_dict(trainer.adam, sd.keys(), d, num_extra_params=len(koppen_state_dict)),
        "scheduler_state_dict": scheduler_state_dict(scheduler),
        "epoch": int(epoch),
        "best_loss": float(best_loss),
        "model_version": MODEL_VERSION,
        "total_params": total,
        "config": {"input_channels": d.in_channels, "hidden_channels": d.hidden, "output_channels": d.out_channels,
                   "window_size": d.window, "forecast_horizon": d.horizon},
        "hybrid_config": {"lstm_hidden_size": d.lstm_hidden, "lstm_num_layers": d.lstm_layers, "lstm_dropout": lstm_dropout},
    }
    if final_loss is not None:
        ckpt["final_loss"] = float(final_loss)
    return ckpt


def resume(trainer, scheduler, ckpt):
    """Load a meta-training checkpoint (written here or by the reference) into a MetaTrainer and its schedule: weights,
    AdamW moments and step count, schedule position.  Returns ``(epoch, best_loss)``."""
    d = trainer.dims
    cfg, hyb = ckpt["config"], ckpt["hybrid_config"]
    want = (d.in_channels, d.hidden, d.out_channels, d.window, d.horizon, d.lstm_hidden, d.lstm_layers)
    got = (cfg["input_channels"], cfg["hidden_channels"], cfg["output_channels"], cfg["window_size"], cfg["forecast_horizon"],
           hyb["lstm_hidden_size"], hyb["lstm_num_layers"])
    if want != got:
        raise ValueError(f"checkpoint configuration {got} does not match the trainer's {want}")
    sd = ckpt["hybrid_model_state_dict"]
    trainer.load_state_dict(sd)
    load_optimizer_state_dict(trainer.adam, ckpt["meta_optimizer_state_dict"], sd.keys(), d)
    load_scheduler_state_dict(scheduler, ckpt["scheduler_state_dict"])
    trainer.set_lr(ckpt["meta_optimizer_state_dict"]["param_groups"][0]["lr"])
    return int(ckpt["epoch"]), float(ckpt["best_loss"])
