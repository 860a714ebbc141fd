"""Fused evaluation behind the reference-shaped API.

`Model.forward` -> `LossFlow.forward` / `LossTracking.forward` -> `total.backward()` is the surface
flowmap/model/model_wrapper_overfit.py:51-73 drives.  Evaluated op by op it costs one C-ABI call
and several ATen launches per module; here the same surface runs on the two halves of the fused
step (fm_overfit_step, FM_STEP_FORWARD / FM_STEP_BACKWARD):

  * `Model.forward` launches nothing and returns a `LazyModelOutput`;
  * the first `LossFlow.forward` of the step runs the forward half (candidate sweep in the softmin
    stage, Procrustes poses, flow loss with its direct gradients) and returns the loss value as the
    output of an autograd node; `LossTracking.forward` adds the tracking sweep;
  * all loss nodes hang off ONE root node whose inputs are the model's parameters; autograd calls
    its backward once, after every loss has reported its grad_output, and that call runs the
    backward half and hands the finished parameter gradients to autograd (no intermediate tensors,
    no per-op ATen glue).

Anything the fused step does not cover (a consumer that reads `model_output.extrinsics` under
autograd, another backbone, per-frame intrinsics, different mappings for the two losses) makes the
LazyModelOutput materialise itself through the per-op autograd Functions of flowmap_b200.ops: same
results, the former speed.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from . import ops
from .types import Batch, Flows, ModelOutput


class _StepRoot(torch.autograd.Function):
    """Root of one fused step: token = f(parameters).  Its backward runs the backward half."""

    @staticmethod
    def forward(ctx, step, *params):
        ctx.step = step
        return torch.empty((), dtype=torch.float32, device=params[0].device)

    @staticmethod
    def backward(ctx, _g_token):
        return (None, *ctx.step.run_backward())


class _LossNode(torch.autograd.Function):
    """One loss of the step: value computed by the forward half, gradient deferred to the root."""

    @staticmethod
    def forward(ctx, token, step, which, value):
        ctx.step, ctx.which = step, which
        return value.clone()  # the engine's buffer is overwritten by the next step

    @staticmethod
    def backward(ctx, g):
        ctx.step.scales[ctx.which] = g.reshape(()).to(torch.float32).contiguous()
        return g, None, None, None


class FusedStep:
    """State of one optimisation step evaluated through the fused halves."""

    def __init__(self, model, batch: Batch, flows: Flows, global_step: int):
        self.model, self.batch, self.flows, self.global_step = model, batch, flows, global_step
        self.engine = None
        self.token: Optional[Tensor] = None
        self.scales = {}
        self.flow_done = self.track_done = False
        self.dead = False  # the output was materialised through the per-op path instead

    # ---- forward half
    def flow_loss(self, loss_mod, tracks) -> Optional[Tensor]:
        if self.dead or self.flow_done:
            return None
        eng = self.model._fused_engine(self.batch, self.flows, tracks, loss_mod)
        if eng is None:
            return None
        self.engine = eng
        eng.cfg.flow_weight, eng.cfg.flow_enable_after = loss_mod.cfg.weight, loss_mod.cfg.enable_after
        eng._msum.copy_(loss_mod._mask_total(self.flows))
        params = self.model._fused_params(self.global_step)
        for p, buf in zip(params, self._grad_buffers(params)):
            if p.grad is not None and buf is not None and p.grad.data_ptr() == buf.data_ptr():
                p.grad = p.grad.clone()  # a kept gradient must not alias the buffer about to be rewritten
        self._params = params
        value = eng.forward_phase(self.global_step, training=self.model.training)
        self.token = _StepRoot.apply(self, *params)
        self.flow_done = True
        return _LossNode.apply(self.token, self, "flow", value)

    def track_loss(self, loss_mod, tracks) -> Optional[Tensor]:
        eng = self.engine
        if self.dead or not self.flow_done or self.track_done or eng is None or eng._packed is None:
            return None
        fm = eng.cfg
        lm = loss_mod.cfg.mapping
        if lm.name != fm.mapping or getattr(lm, "delta", fm.delta) != fm.delta:
            return None  # the fused step evaluates both losses with one mapping
        fm.tracking_weight = loss_mod.cfg.weight
        eng._args.track_weight = loss_mod.cfg.weight
        value = eng.tracking_forward_phase()
        self.track_done = True
        return _LossNode.apply(self.token, self, "tracking", value)

    # ---- backward half
    def _grad_buffers(self, params):
        eng = self.engine
        g = eng.gradients()
        bufs = [g["depth"]]
        if eng.cfg.use_correspondence_weights:
            bufs.append(g["weights"])
        if len(params) > len(bufs):
            bufs.append(g["focal"])
        return bufs

    def run_backward(self):
        eng = self.engine
        eng.backward_phase(self.scales.get("flow"), self.scales.get("tracking"), with_tracking=self.track_done)
        out = []
        for p, buf in zip(self._params, self._grad_buffers(self._params)):
            # a fresh alias of the persistent buffer: autograd adopts it as .grad without a copy
            out.append(buf.detach().view(p.shape))
        return tuple(out)

    def snapshot(self) -> ModelOutput:
        """Detached ModelOutput of the step the forward half evaluated (for logging / visualisers
        that read the output after the losses): poses and intrinsics come from the engine's buffers."""
        eng, model = self.engine, self.model
        _, f, _, h, w = self.batch.videos.shape
        with torch.no_grad():
            k4 = eng.intrinsics_k4()[None].clone()
            rt = eng.rt.clone()
            k = torch.zeros(1, f, 3, 3, device=k4.device)
            k[..., 0, 0], k[..., 1, 1], k[..., 0, 2], k[..., 1, 2], k[..., 2, 2] = \
                k4[..., 0], k4[..., 1], k4[..., 2], k4[..., 3], 1.0
            bo = model.backbone.forward(self.batch, self.flows)
            weights = bo.weights if model.cfg.use_correspondence_weights else torch.ones_like(bo.weights)
            return ModelOutput(bo.depths, None, k, ops.pose_chain(rt), weights, relative=rt, k4=k4,
                               k_mode="shared_focal")


class LazyModelOutput(ModelOutput):
    """ModelOutput of a fused step: `depths` is the parameter itself, everything else is computed
    when (and only if) somebody reads it: before the losses through the differentiable per-op path
    (which retires the fused step for this iteration), after them as detached values of the step
    the fused forward half evaluated."""

    def __init__(self, model, batch: Batch, flows: Flows, global_step: int):
        object.__setattr__(self, "_lazy", (model, batch, flows, global_step))
        object.__setattr__(self, "_full", None)
        object.__setattr__(self, "_fused", FusedStep(model, batch, flows, global_step))
        object.__setattr__(self, "depths", model.backbone.depth[None])

    def _materialize(self) -> ModelOutput:
        full = object.__getattribute__(self, "_full")
        if full is None:
            model, batch, flows, step = object.__getattribute__(self, "_lazy")
            fused = object.__getattribute__(self, "_fused")
            if fused.flow_done:  # the fused losses already consumed this output: values only
                full = fused.snapshot()
            else:
                fused.dead = True
                full = model._forward_materialized(batch, flows, step)
            object.__setattr__(self, "_full", full)
        return full

    @property
    def surfaces(self) -> Tensor:
        return self._materialize().surfaces

    def __getattr__(self, name):  # only reached for attributes not set in __init__
        if name.startswith("__"):
            raise AttributeError(name)
        return getattr(self._materialize(), name)
