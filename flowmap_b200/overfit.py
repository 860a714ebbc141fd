"""The optimisation loop around the hot path: what flowmap/overfit.py:76-112 builds and
flowmap/model/model_wrapper_overfit.py:51-73,104-105 runs every step (Model.forward ->
sum of losses -> backward -> Adam), without the Lightning/Hydra shell.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Sequence

import torch
from torch import Tensor

from . import ops
from .loss import (LossFlowCfg, LossTrackingCfg, MappingHuberCfg, MappingL1Cfg, MappingL2Cfg,
                   get_losses)
from .model import (BackboneExplicitDepthCfg, ExtrinsicsProcrustesCfg, IntrinsicsRegressedCfg,
                    IntrinsicsSoftminCfg, Model, ModelCfg, RegressionCfg)
from .types import Batch, Flows


@dataclass
class OverfitCfg:
    """The values of config/overfit.yaml + config/{model,loss}/** that the step reads."""
    initial_depth: float = 0.1
    weight_sensitivity: float = 100.0
    use_correspondence_weights: bool = True
    procrustes_points: Optional[int] = None  # experiment/ablation_explicit_depth.yaml:11-12
    procrustes_randomize: bool = False
    intrinsics: str = "regressed"
    initial_focal: float = 0.85
    softmin_points: int = 8192
    softmin_min: float = 0.5
    softmin_max: float = 2.0
    softmin_candidates: int = 60
    regression_after: Optional[int] = 1000
    regression_window: int = 100
    flow_weight: float = 1000.0
    flow_enable_after: int = 0
    tracking_weight: float = 100.0
    tracking_enable_after: int = 50
    use_tracking: bool = False
    mapping: str = "huber"
    delta: float = 0.01
    lr: float = 3e-5


def _mapping_cfg(name: str, delta: float):
    return {"huber": MappingHuberCfg("huber", delta), "l1": MappingL1Cfg("l1"),
            "l2": MappingL2Cfg("l2")}[name]


def build_model_and_losses(cfg: OverfitCfg, num_frames: int, image_shape):
    if cfg.intrinsics == "regressed":
        icfg = IntrinsicsRegressedCfg("regressed", cfg.initial_focal)
    else:
        reg = None if cfg.regression_after is None else RegressionCfg(cfg.regression_after,
                                                                      cfg.regression_window)
        icfg = IntrinsicsSoftminCfg("softmin", cfg.softmin_points, cfg.softmin_min,
                                    cfg.softmin_max, cfg.softmin_candidates, reg)
    mcfg = ModelCfg(BackboneExplicitDepthCfg("explicit_depth", cfg.initial_depth,
                                             cfg.weight_sensitivity), icfg,
                    ExtrinsicsProcrustesCfg("procrustes", cfg.procrustes_points,
                                            cfg.procrustes_randomize),
                    cfg.use_correspondence_weights)
    model = Model(mcfg, num_frames, image_shape)
    lcfgs = [LossFlowCfg(cfg.flow_enable_after, cfg.flow_weight, "flow",
                         _mapping_cfg(cfg.mapping, cfg.delta))]
    if cfg.use_tracking:
        lcfgs.append(LossTrackingCfg(cfg.tracking_enable_after, cfg.tracking_weight, "tracking",
                                     _mapping_cfg(cfg.mapping, cfg.delta)))
    return model, get_losses(lcfgs)


class FusedAdam:
    """torch.optim.Adam(params, lr) semantics (model_wrapper_overfit.py:104-105) on the
    fm_adam_step kernel: one launch per parameter tensor, no foreach temporaries."""

    def __init__(self, params, lr: float, betas=(0.9, 0.999), eps: float = 1e-8):
        self.params = [p for p in params]
        self.lr, self.betas, self.eps = lr, betas, eps
        self.step_count = 0
        self.state = [(torch.zeros_like(p), torch.zeros_like(p)) for p in self.params]

    def zero_grad(self):
        for p in self.params:
            p.grad = None

    @torch.no_grad()
    def step(self):
        self.step_count += 1
        for p, (m, v) in zip(self.params, self.state):
            if p.grad is None:
                continue
            ops.adam_step(p.data, p.grad.contiguous(), m, v, self.step_count, self.lr, self.betas,
                          self.eps)


class Overfitter:
    """Holds the constant batch/flows/tracks and runs optimisation steps
    (model_wrapper_overfit.py:24-73)."""

    def __init__(self, cfg: OverfitCfg, batch: Batch, flows: Flows, tracks=None,
                 device="cuda"):
        self.cfg = cfg
        self.batch, self.flows = batch.to(device), flows.to(device)
        self.tracks = None if tracks is None else [t.to(device) for t in tracks]
        _, f, _, h, w = batch.videos.shape
        self.model, self.losses = build_model_and_losses(cfg, f, (h, w))
        self.model.to(device)
        self.optimizer = FusedAdam(self.model.parameters(), cfg.lr)
        self.global_step = 0

    def training_step(self):
        """Returns (total loss tensor (device), ModelOutput); no host sync."""
        self.optimizer.zero_grad()
        out = self.model(self.batch, self.flows, self.global_step)
        total = 0
        for loss_fn in self.losses:
            total = total + loss_fn.forward(self.batch, self.flows, self.tracks, out,
                                            self.global_step)
        total.backward()
        self.optimizer.step()
        self.global_step += 1
        return total.detach(), out


class ShardedOverfitter(Overfitter):
    """Pair-sharded optimisation (flowmap_b200.parallel): this rank holds the frames and
    pairs of its ShardPlan; one all-reduce per step carries the loss, the focal-length
    gradient and the boundary depth-gradient frames."""

    def __init__(self, cfg: OverfitCfg, batch: Batch, flows: Flows, plan, device="cuda",
                 group=None):
        from . import parallel
        if cfg.use_tracking or cfg.intrinsics != "regressed":
            raise NotImplementedError("pair sharding currently covers the flow loss with a "
                                      "regressed focal length (BASELINE config 4)")
        super().__init__(cfg, batch, flows, None, device)
        self.plan, self.group = plan, group
        _, _, _, h, w = batch.videos.shape
        self.reducer = parallel.StepReducer(plan, (h, w), self.flows.forward.device, 2, group)
        local = ops.mask_sum(self.flows.forward_mask, self.flows.backward_mask)
        self.losses[0].set_global_mask_sum(parallel.global_mask_sum(local, group))

    def training_step(self):
        self.optimizer.zero_grad()
        out = self.model(self.batch, self.flows, self.global_step)
        total = 0
        for loss_fn in self.losses:
            total = total + loss_fn.forward(self.batch, self.flows, None, out, self.global_step)
        total.backward()
        focal = self.model.intrinsics.focal_length
        scalars = torch.stack((total.detach(), focal.grad.reshape(())))
        red = self.reducer.reduce(scalars, self.model.backbone.depth.grad)
        focal.grad.copy_(red[1])
        self.optimizer.step()
        self.global_step += 1
        return red[0], out
