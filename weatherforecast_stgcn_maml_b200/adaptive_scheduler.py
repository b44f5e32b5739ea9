"""Drop-in for the reference's ``adaptive_scheduler`` module (adaptive_scheduler.py:7-95).

Pure host-side scalar logic (which Adam lr / weight decay a region name gets, and the 5-epoch
cosine cycle with a loss-based nudge).  It has no kernel content; it is restated so that the
fine-tune loop has the same lr trajectory.  ``create_climate_optimizer`` accepts either torch
parameters (returns ``torch.optim.Adam`` like the reference) or is bypassed by
``adapt_hybrid_v5.FineTuner``, which feeds the same (lr, weight_decay) to the fused Adam kernel.
"""
import math

import torch

TROPICAL_REGIONS = ("Indonesia", "Thailand", "QueensAustralia")
COLD_REGIONS = ("Moscow", "NorthSiberia", "Afghanistan")
CLIMATE_LR_MULT = {"tropical": 0.9, "temperate": 1.0, "cold": 1.1}
CLIMATE_WEIGHT_DECAY = {"tropical": 1e-5, "temperate": 1e-4, "cold": 5e-5}


def climate_zone(region_name):
    if region_name in TROPICAL_REGIONS:
        return "tropical"
    if region_name in COLD_REGIONS:
        return "cold"
    return "temperate"


def climate_hyperparameters(region_name, base_lr=0.0006):
    """(lr, weight_decay) of adaptive_scheduler.py:72-87."""
    zone = climate_zone(region_name)
    return base_lr * CLIMATE_LR_MULT[zone], CLIMATE_WEIGHT_DECAY[zone]


class ClimateAwareLRScheduler:
    def __init__(self, optimizer, region_name, base_lr=0.0006):
        self.optimizer = optimizer
        self.region_name = region_name
        self.base_lr = base_lr
        self.current_epoch = 0
        self.climate_multipliers = dict(CLIMATE_LR_MULT)
        self.climate_zone = self._get_climate_zone()
        self.lr_multiplier = self.climate_multipliers.get(self.climate_zone, 1.0)

    def _get_climate_zone(self):
        return climate_zone(self.region_name)

    def step(self, epoch_loss=None):
        self.current_epoch += 1
        cycle_length = 5
        cycle_progress = (self.current_epoch - 1) % cycle_length / cycle_length
        cosine_factor = 0.5 * (1 + math.cos(math.pi * cycle_progress))
        climate_lr = self.base_lr * self.lr_multiplier * cosine_factor
        if epoch_loss is not None and self.current_epoch > 3:
            if epoch_loss > 1.0:
                climate_lr *= 1.1
            elif epoch_loss < 0.2:
                climate_lr *= 0.95
        for group in self._groups():
            group["lr"] = climate_lr
        return climate_lr

    def _groups(self):
        if hasattr(self.optimizer, "param_groups"):
            return self.optimizer.param_groups
        return [self.optimizer.__dict__]  # engine.AdamState exposes ``lr`` as an attribute

    def get_last_lr(self):
        return [g["lr"] for g in self._groups()]


def create_climate_optimizer(model_params, region_name, base_lr=0.0006):
    lr, wd = climate_hyperparameters(region_name, base_lr)
    return torch.optim.Adam(model_params, lr=lr, weight_decay=wd), lr
