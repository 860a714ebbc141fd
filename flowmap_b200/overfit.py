"""The optimisation loop around the hot path: what flowmap/overfit.py:76-112 builds and
flowmap/model/model_wrapper_overfit.py:51-73,104-105 runs every step (Model.forward ->
sum of losses -> backward -> Adam), without the Lightning/Hydra shell.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

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
        self.steps = [0 for _ in self.params]  # torch keeps one step counter per parameter
        self.state = [(torch.zeros_like(p), torch.zeros_like(p)) for p in self.params]

    def zero_grad(self):
        for p in self.params:
            p.grad = None

    @torch.no_grad()
    def step(self):
        for i, (p, (m, v)) in enumerate(zip(self.params, self.state)):
            if p.grad is None:  # torch.optim skips parameters that did not receive a gradient
                continue
            self.steps[i] += 1
            ops.adam_step(p.data, p.grad.contiguous(), m, v, self.steps[i], self.lr, self.betas,
                          self.eps)


class Overfitter:
    """Holds the constant batch/flows/tracks and runs optimisation steps
    (model_wrapper_overfit.py:24-73)."""

    def __init__(self, cfg: OverfitCfg, batch: Batch, flows: Flows, tracks=None,
                 device="cuda", model=None):
        self.cfg = cfg
        self.batch, self.flows = batch.to(device), flows.to(device)
        self.tracks = None if tracks is None else [t.to(device) for t in tracks]
        _, f, _, h, w = batch.videos.shape
        if model is None:
            self.model, self.losses = build_model_and_losses(cfg, f, (h, w))
            self.model.to(device)
            self.optimizer = FusedAdam(self.model.parameters(), cfg.lr)
        else:  # bound to a caller's Model (the autograd drop-in surface, flowmap_b200.fused)
            self.model, self.losses, self.optimizer = model, None, None
        self.global_step = 0
        # torch.optim.Adam counts the updates each parameter has received, not the trainer's
        # global_step (they differ when a run starts at global_step > 0 with a fresh optimiser, and
        # for the focal length, which sees its first gradient at the softmin -> regressed hand-over)
        self.optimizer_steps = 0
        self.focal_steps = 0

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


class FusedOverfitter(Overfitter):
    """Same optimisation as :class:`Overfitter` (explicit-depth backbone, Procrustes poses,
    regressed focal length, flow [+ tracking] loss, Adam) but each step is ONE C-ABI call,
    fm_overfit_step: no autograd graph, no intermediate tensors, the sigmoid of the weight
    logits and its chain rule evaluated inside the kernels, gradients accumulated into a
    single buffer per parameter.

    use_splat_plan=True selects the DETERMINISTIC backward (ops.SplatPlan, csrc/fm_tiled.cuh): the
    bilinear scatter of the Procrustes adjoint is transposed once per Flows into a static plan and
    evaluated as a gather with TMA-staged windows -- no atomics, bit-reproducible gradients.  It is
    parity-green but measured ~10 % slower per backward than the default global-RED kernel on
    B200 (profiles/README.md), hence opt-in."""

    def __init__(self, cfg: OverfitCfg, batch: Batch, flows: Flows, tracks=None, device="cuda",
                 use_splat_plan: bool = False, model=None):
        super().__init__(cfg, batch, flows, tracks, device, model=model)
        from ._lib import OverfitStepArgs, PackedTracksC, lib
        import ctypes
        if batch.videos.shape[0] != 1:
            raise ValueError("flowmap_b200: the fused step optimises one video (batch size 1), as "
                             "flowmap/overfit.py does")
        # the kernels read raw pointers: canonical (contiguous float32) copies, kept alive here
        # (FlowPredictor.rescale_flow returns a permuted view, flow_predictor.py:40-49)
        self.flows = Flows(*(ops._canon(getattr(self.flows, n), n)
                             for n in ("forward", "backward", "forward_mask", "backward_mask")))
        dev = self.flows.forward.device
        _, f, _, h, w = batch.videos.shape
        self._use_plan = use_splat_plan and cfg.procrustes_points is None and not cfg.procrustes_randomize
        self._plan = ops.SplatPlan(self.flows.backward) if self._use_plan else None
        bb = self.model.backbone
        self._depth, self._wlog = bb.depth.data, bb.weights.data
        self._softmin = cfg.intrinsics == "softmin"
        if self._softmin:
            intr = self.model.intrinsics
            self._focal = (intr.intrinsics_regressed.focal_length.data if cfg.regression_after
                           is not None else torch.zeros((), device=dev))
            n = cfg.softmin_candidates
            self._cand_f = intr.focal_length_candidates.float().contiguous()
            self._cand_k4 = ops.candidate_k4(self._cand_f, h, w, 1)
            self._sw_err = torch.empty(1, n, device=dev)
            self._sw_sm = torch.empty(1, n, device=dev)
            self._sw_gerr = torch.empty(1, n, device=dev)
            self._sw_rt = torch.empty(n, 3, 4, device=dev)
            self._sw_focal = torch.zeros(1, device=dev)
            self._sw_ws = torch.empty(lib().fm_softmin_workspace_bytes(1, n), dtype=torch.uint8,
                                      device=dev)
            # candidate 0's intrinsics for every frame: what the early moment pass of a sweep step uses
            self._k4_base = self._cand_k4.reshape(-1, 4)[0].expand(f, 4).contiguous()
            self.window = []
            self.injected_indices = None
        else:
            self._focal = self.model.intrinsics.focal_length.data
        z = lambda t: torch.zeros_like(t)  # noqa: E731
        self._state = [z(self._depth), z(self._depth), z(self._wlog), z(self._wlog),
                       z(self._focal), z(self._focal)]
        self._g_depth, self._g_w = torch.empty_like(self._depth), torch.empty_like(self._wlog)
        self._g_focal = torch.zeros_like(self._focal)
        self._k4 = torch.empty(f, 4, device=dev)
        self._g_k4 = torch.empty(f, 4, device=dev)
        self.rt = torch.empty(1, f - 1, 3, 4, device=dev)
        self._loss = torch.zeros((), device=dev)
        self._track_loss = torch.zeros((), device=dev)
        self._ws = ops.workspace(1, f, h, w, dev)
        self._msum = ops.mask_sum(self.flows.forward_mask, self.flows.backward_mask)
        self._indices = self.model.extrinsics.select_indices(h, w, dev) \
            if not cfg.procrustes_randomize else None
        a = OverfitStepArgs()
        P = lambda t: None if t is None else t.data_ptr()  # noqa: E731
        a.F, a.H, a.W = f, h, w
        a.depth = P(self._depth)
        a.weight_logits = P(self._wlog) if cfg.use_correspondence_weights else None
        a.weight_sensitivity = cfg.weight_sensitivity
        a.focal, a.k4 = P(self._focal), P(self._k4)
        a.fflow, a.bflow = P(self.flows.forward), P(self.flows.backward)
        a.fmask, a.bmask = P(self.flows.forward_mask), P(self.flows.backward_mask)
        a.mask_sum = P(self._msum)
        a.mapping, a.delta, a.flow_weight = ops.MAPPINGS[cfg.mapping], cfg.delta, cfg.flow_weight
        (a.m_depth, a.v_depth, a.m_weights, a.v_weights, a.m_focal, a.v_focal) = \
            [P(t) for t in self._state]
        a.lr, a.beta1, a.beta2, a.eps = cfg.lr, 0.9, 0.999, 1e-8
        a.g_depth, a.g_weights, a.g_focal, a.g_k4 = P(self._g_depth), P(self._g_w), \
            P(self._g_focal), P(self._g_k4)
        a.rt, a.loss, a.ws = P(self.rt), P(self._loss), P(self._ws)
        # step-dependent scalars live in device memory: every step is the same launch sequence
        self._clock = ops.StepClock(dev, cfg.lr)
        self._side_stream = torch.cuda.Stream(device=dev)  # parallel branch of the step (see _step_softmin)
        a.clock = self._clock.ptr
        self._total = torch.zeros((), device=dev)
        self._idx_buf = torch.empty(min(cfg.softmin_points, h * w), dtype=torch.int64, device=dev) \
            if self._softmin else None
        self.use_cuda_graph = False  # opt-in: replay the update step as ONE CUDA graph launch
        self._graphs, self._eager_runs = {}, {}
        self._set_plan_args(a)
        self._packed = None
        if cfg.use_tracking:
            assert self.tracks is not None
            pk = ops.PackedTracks(self.tracks, dev)
            self._packed = pk
            self._pk_c = PackedTracksC(P(pk.seg), P(pk.xy), P(pk.vis), pk.num_segments,
                                       pk.max_rows, pk.max_points, pk.total)
            self._ext = torch.empty(1, f, 4, 4, device=dev)
            self._g_ext = torch.empty(1, f, 4, 4, device=dev)
            self._g_rt = torch.empty(1, f - 1, 3, 4, device=dev)
            self._tg_k4 = torch.empty(f, 4, device=dev)
            self._tws = torch.empty(lib().fm_track_workspace_bytes(f, pk.total), dtype=torch.uint8,
                                    device=dev)
            a.track_weight = cfg.tracking_weight
            a.extrinsics, a.g_extrinsics, a.g_rt = P(self._ext), P(self._g_ext), P(self._g_rt)
            a.track_g_k4, a.track_loss, a.track_ws = P(self._tg_k4), P(self._track_loss), P(self._tws)
        self._args, self._ctypes = a, ctypes
        self._lib = lib()

    def _set_plan_args(self, a):
        pl = self._plan
        a.splat_plan = pl.ptr if pl is not None else None
        a.splat_overflow_max = pl.overflow_max if (pl is not None and pl.ok) else 0

    def set_flows(self, flows: Flows, mask_sum: Optional[Tensor] = None):
        """Point the step at another device-resident Flows of the same shape (the next batch of a
        prefetching loader) without rebuilding parameters or optimiser state.  `mask_sum` is the
        flow-loss normaliser (loss_flow.py:70) if the caller already has it."""
        old = self.flows
        canon = {}
        for name in ("forward", "backward", "forward_mask", "backward_mask"):
            t = ops._canon(getattr(flows, name), name)
            if t.shape != getattr(old, name).shape or t.device != getattr(old, name).device:
                raise ValueError(f"flowmap_b200: `{name}` does not match the optimiser's shapes / device")
            canon[name] = t
        flows = Flows(canon["forward"], canon["backward"], canon["forward_mask"], canon["backward_mask"])
        self.flows = flows  # the canonical tensors stay referenced while the kernels hold their pointers
        a = self._args
        a.fflow, a.bflow = flows.forward.data_ptr(), flows.backward.data_ptr()
        a.fmask, a.bmask = flows.forward_mask.data_ptr(), flows.backward_mask.data_ptr()
        self._msum.copy_(self._mask_sum(flows) if mask_sum is None else mask_sum)
        if self._plan is not None:  # new backward flows: new transpose
            self._plan.rebuild(flows.backward)
            self._set_plan_args(a)
        self._graphs.clear()  # captured launches hold the old pointers
        self._eager_runs.clear()

    def _mask_sum(self, flows: Flows) -> Tensor:
        return ops.mask_sum(flows.forward_mask, flows.backward_mask)

    def _softmin_stage(self) -> bool:
        c = self.cfg
        return self._softmin and not (c.regression_after is not None and
                                      self.global_step >= c.regression_after)

    def _step_softmin(self, update: bool):
        """Sweep stage (intrinsics_softmin.py:84-141): focal estimate from the candidate sweep,
        the step itself with that focal length, the sweep's backward, then Adam."""
        from ._lib import check
        c, a, L = self.cfg, self._args, self._lib
        _, f, _, h, w = self.batch.videos.shape
        dev = self.rt.device
        st = torch.cuda.current_stream().cuda_stream
        P = lambda t: None if t is None else t.data_ptr()  # noqa: E731
        n = c.softmin_candidates
        wl = P(self._wlog) if c.use_correspondence_weights else None
        sens = c.weight_sensitivity if c.use_correspondence_weights else 0.0
        # All-pixel Procrustes: the moment pass of the step does not have to wait for the focal length
        # the sweep is about to produce -- the sums for one K follow exactly from the sums for another
        # (fm_overfit_step_args.moments_k4) -- so it runs beside the sweep, on the candidate-0 intrinsics.
        early_moments = update and self._indices is None and self._plan is None
        cur = torch.cuda.current_stream()
        with torch.cuda.device(dev):
            if early_moments:
                self._side_stream.wait_stream(cur)
                check(L.fm_procrustes_moments(P(self._depth), P(self._k4_base), P(self.flows.backward), wl, sens,
                                              P(self._ws), f, h, w, st), "fm_procrustes_moments")
            with torch.cuda.stream(self._side_stream if early_moments else cur):
                sst = torch.cuda.current_stream().cuda_stream
                idx = self.injected_indices
                if idx is None:
                    if update:  # seeded by the step clock (replayable); intrinsics_softmin.py:90
                        idx = ops.random_subset_clock(self._clock, h * w, self._idx_buf)
                    else:
                        idx = ops.random_subset(h * w, min(c.softmin_points, h * w), dev)
                idx = idx.contiguous()
                check(L.fm_softmin_sweep_fwd(P(self._depth), wl, sens, P(self.flows.backward), P(idx),
                                             idx.numel(), P(self._cand_k4), n, P(self._sw_err),
                                             P(self._sw_rt), P(self._sw_ws), 1, f, h, w, sst),
                      "fm_softmin_sweep_fwd")
                check(L.fm_softmin_focal(P(self._sw_err), P(self._cand_f), n, 1, P(self._sw_sm),
                                         P(self._sw_focal), sst), "fm_softmin_focal")
            if early_moments:
                cur.wait_stream(self._side_stream)
            a.moments_k4 = P(self._k4_base) if early_moments else None
            # all-pixel dense path: the logits of pairs >= 1 are updated inside the step (their
            # gradient is final there); depth and pair 0 wait for the sweep's backward
            fuse = update and c.use_correspondence_weights and self._indices is None and w % 4 == 0
            a.focal = P(self._sw_focal)
            a.step = 1 if fuse else 0  # on / off: the bias corrections come from the step clock
            a.defer_adam = 1 if fuse else 0
            try:
                check(L.fm_overfit_step(self._ctypes.byref(a), st), "fm_overfit_step")
            finally:
                a.moments_k4 = None
            a.defer_adam = 0
            # The sweep's backward only touches the gradients of frames 0 / 1 (the candidate Procrustes
            # runs on the first pair): the depth update of every other frame runs beside it on a second
            # stream (a parallel branch of the captured step), frames 0 / 1 follow the sweep.
            side, cur = None, torch.cuda.current_stream()
            if update and f > 2:
                side = self._side_stream
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    ops.adam_step_clock(self._depth[2:], self._g_depth[2:], self._state[0][2:], self._state[1][2:],
                                        self._clock)
            check(L.fm_softmin_focal_bwd(P(self._sw_sm), P(self._cand_f), P(self._sw_focal),
                                         P(self._g_focal), n, 1, P(self._sw_gerr), st),
                  "fm_softmin_focal_bwd")
            check(L.fm_softmin_sweep_bwd(P(self._depth), wl, sens, P(self.flows.backward), P(idx),
                                         idx.numel(), P(self._cand_k4), n, P(self._sw_rt),
                                         P(self._sw_gerr), P(self._g_depth),
                                         P(self._g_w) if wl else None, P(self._sw_ws), 1, f, h, w, st),
                  "fm_softmin_sweep_bwd")
        if update:
            ops.adam_step_clock(self._depth[:2], self._g_depth[:2], self._state[0][:2], self._state[1][:2],
                                self._clock)
            if c.use_correspondence_weights:
                k = 1 if fuse else self._wlog.shape[0]  # pair 0 only when the rest was fused
                ops.adam_step_clock(self._wlog[:k], self._g_w[:k], self._state[2][:k], self._state[3][:k],
                                    self._clock)
            if side is not None:
                cur.wait_stream(side)

    # ---- split step: the two halves of one iteration WITHOUT the parameter update, for callers that
    # need the loss values before they decide on the backward (torch.autograd: flowmap_b200.fused)
    def _window(self):
        """The hand-over window of the softmin stage (intrinsics_softmin.py:133-139): the bound
        model's own list when there is one (drop-in surface), else this optimiser's."""
        intr = self.model.intrinsics
        return intr.window if hasattr(intr, "window") and self.optimizer is None else self.window

    def forward_phase(self, global_step: int, training: bool = True):
        """Poses + flow loss with its direct gradients (fm_overfit_step, FM_STEP_FORWARD; the
        candidate sweep first in the softmin stage).  Returns the weighted flow loss (device scalar,
        a buffer that the next call overwrites)."""
        from ._lib import check
        c, a, L = self.cfg, self._args, self._lib
        _, f, _, h, w = self.batch.videos.shape
        dev = self.rt.device
        st = torch.cuda.current_stream().cuda_stream
        P = lambda t: None if t is None else t.data_ptr()  # noqa: E731
        self.global_step = global_step
        if c.procrustes_randomize:
            self._indices = self.model.extrinsics.select_indices(h, w, dev)
        a.indices = None if self._indices is None else self._indices.data_ptr()
        a.num_indices = 0 if self._indices is None else self._indices.numel()
        a.flow_weight = c.flow_weight if global_step >= c.flow_enable_after else 0.0
        a.tracks, a.step, a.defer_adam = None, 0, 0
        self._sweep_idx = None
        with torch.cuda.device(dev):
            if self._softmin_stage():
                idx = self.injected_indices
                if idx is None:
                    idx = getattr(self.model.intrinsics, "injected_indices", None)
                if idx is None:
                    idx = ops.random_subset(h * w, min(c.softmin_points, h * w), dev)
                self._sweep_idx = idx = idx.contiguous()
                n = c.softmin_candidates
                wl = P(self._wlog) if c.use_correspondence_weights else None
                sens = c.weight_sensitivity if c.use_correspondence_weights else 0.0
                check(L.fm_softmin_sweep_fwd(P(self._depth), wl, sens, P(self.flows.backward), P(idx),
                                             idx.numel(), P(self._cand_k4), n, P(self._sw_err),
                                             P(self._sw_rt), P(self._sw_ws), 1, f, h, w, st),
                      "fm_softmin_sweep_fwd")
                check(L.fm_softmin_focal(P(self._sw_err), P(self._cand_f), n, 1, P(self._sw_sm),
                                         P(self._sw_focal), st), "fm_softmin_focal")
                a.focal = P(self._sw_focal)
                if training and c.regression_after is not None and \
                        global_step >= c.regression_after - c.regression_window:
                    self._window().append(self._sw_focal[0].clone())
            else:
                if self._softmin and global_step == c.regression_after and training:
                    self._focal.copy_(torch.stack(self._window()).mean())
                a.focal = self._focal.data_ptr()
            a.phase = 1  # FM_STEP_FORWARD
            try:
                check(L.fm_overfit_step(self._ctypes.byref(a), st), "fm_overfit_step (forward)")
            finally:
                a.phase = 0
        return self._loss

    def tracking_forward_phase(self):
        """Chained poses + tracking loss of the step begun by forward_phase (loss_tracking.py:28-61).
        Returns the weighted tracking loss (device scalar buffer)."""
        from ._lib import check
        c, a, L, pk = self.cfg, self._args, self._lib, self._packed
        _, f, _, h, w = self.batch.videos.shape
        P = lambda t: t.data_ptr()  # noqa: E731
        st = torch.cuda.current_stream().cuda_stream
        with torch.cuda.device(self.rt.device):
            check(L.fm_pose_chain(P(self.rt), P(self._ext), 1, f, st), "fm_pose_chain")
            check(L.fm_track_loss_fwd_sharded(
                P(self._depth), P(self._k4), P(self._ext), P(pk.seg), pk.num_segments, pk.max_rows,
                pk.max_points, P(pk.xy), P(pk.vis), pk.total, ops.MAPPINGS[c.mapping], c.delta,
                c.tracking_weight, P(self._track_loss), P(self._tws), f, h, w, 0, 0, f, 1, st),
                "fm_track_loss_fwd")
        return self._track_loss

    def backward_phase(self, flow_scale=None, track_scale=None, with_tracking: bool = False):
        """Second half: [tracking backward,] Procrustes backward, focal gradient[, the sweep's
        backward].  flow_scale / track_scale: device float scalars d total / d loss (None = 1).
        Leaves the gradients in gradients()."""
        from ._lib import check
        c, a, L = self.cfg, self._args, self._lib
        _, f, _, h, w = self.batch.videos.shape
        st = torch.cuda.current_stream().cuda_stream
        P = lambda t: None if t is None else t.data_ptr()  # noqa: E731
        a.tracks = self._ctypes.pointer(self._pk_c) if with_tracking else None
        a.flow_grad_scale, a.track_grad_scale = P(flow_scale), P(track_scale)
        a.phase, a.step, a.defer_adam = 2, 0, 0  # FM_STEP_BACKWARD
        with torch.cuda.device(self.rt.device):
            try:
                check(L.fm_overfit_step(self._ctypes.byref(a), st), "fm_overfit_step (backward)")
            finally:
                a.phase, a.tracks, a.flow_grad_scale, a.track_grad_scale = 0, None, None, None
            if self._sweep_idx is not None:
                idx, n = self._sweep_idx, c.softmin_candidates
                wl = P(self._wlog) if c.use_correspondence_weights else None
                sens = c.weight_sensitivity if c.use_correspondence_weights else 0.0
                check(L.fm_softmin_focal_bwd(P(self._sw_sm), P(self._cand_f), P(self._sw_focal),
                                             P(self._g_focal), n, 1, P(self._sw_gerr), st),
                      "fm_softmin_focal_bwd")
                check(L.fm_softmin_sweep_bwd(P(self._depth), wl, sens, P(self.flows.backward), P(idx),
                                             idx.numel(), P(self._cand_k4), n, P(self._sw_rt),
                                             P(self._sw_gerr), P(self._g_depth),
                                             P(self._g_w) if wl else None, P(self._sw_ws), 1, f, h, w, st),
                      "fm_softmin_sweep_bwd")

    def _step_body(self, update: bool, track_on: bool, sweep: bool):
        """One step as a fixed launch sequence (no host-side step numbers: see ops.StepClock)."""
        from ._lib import check
        c, a = self.cfg, self._args
        if update:
            self._clock.tick(tick_focal=not sweep)
        a.tracks = self._ctypes.pointer(self._pk_c) if track_on else None
        a.flow_weight = c.flow_weight if self.global_step >= c.flow_enable_after else 0.0
        if sweep:
            self._step_softmin(update)
        else:
            a.focal = self._focal.data_ptr()
            a.step = a.focal_step = 1 if update else 0  # on / off: the step clock carries the counts
            with torch.cuda.device(self.rt.device):
                check(self._lib.fm_overfit_step(self._ctypes.byref(a),
                                                torch.cuda.current_stream().cuda_stream),
                      "fm_overfit_step")
        if track_on:
            torch.add(self._loss, self._track_loss, out=self._total)
        else:
            self._total.copy_(self._loss)

    def _run_body(self, key, graphable: bool, body, ticks_focal: bool):
        """Run one step body: eagerly, or -- from its third run on -- as a replay of its CUDA graph
        (the body must tick the step clock first and be a fixed launch sequence)."""
        if graphable and self._eager_runs.get(key, 0) >= 2:
            g = self._graphs.get(key)
            if g is None:  # capture records the launches without running them: replayed right below
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    body()
                self._graphs[key] = g
                self._clock.steps -= 1  # the capture's host-side tick was not executed
                self._clock.focal_steps -= int(ticks_focal)
            g.replay()
            self._clock.steps += 1
            self._clock.focal_steps += int(ticks_focal)
        else:
            body()
            if graphable:
                self._eager_runs[key] = self._eager_runs.get(key, 0) + 1

    def training_step(self, update: bool = True):
        """Returns (total loss (device tensor), relative poses rt (1, F-1, 3, 4))."""
        c, a = self.cfg, self._args
        if c.procrustes_randomize:
            _, _, _, h, w = self.batch.videos.shape
            self._indices = self.model.extrinsics.select_indices(h, w, self.rt.device)
        a.indices = None if self._indices is None else self._indices.data_ptr()
        a.num_indices = 0 if self._indices is None else self._indices.numel()
        track_on = c.use_tracking and self.global_step >= c.tracking_enable_after
        sweep = self._softmin_stage()
        if update:
            self._clock.set(self.optimizer_steps, self.focal_steps)
            if self._softmin and not sweep and self.global_step == c.regression_after:
                self._focal.copy_(torch.stack(self.window).mean())  # hand-over: seed the regressed focal length once
        window_on = sweep and c.regression_after is not None and \
            self.global_step >= c.regression_after - c.regression_window
        key = (track_on, sweep, self.global_step >= c.flow_enable_after)
        graphable = (update and self.use_cuda_graph and not c.procrustes_randomize and not window_on and
                     getattr(self, "injected_indices", None) is None)
        self._run_body(key, graphable, lambda upd=update: self._step_body(upd, track_on, sweep), not sweep)
        if update:
            if window_on:
                self.window.append(self._sw_focal[0].clone())
            self.global_step += 1
            self.optimizer_steps += 1
            self.focal_steps += int(not sweep)
        return self._total.clone(), self.rt

    def extrinsics(self) -> Tensor:
        """Camera-to-world poses of the last step (projection.py:187-210)."""
        return ops.pose_chain(self.rt)

    def gradients(self):
        return {"depth": self._g_depth, "weights": self._g_w, "focal": self._g_focal}

    def intrinsics_k4(self) -> Tensor:
        """(F, 4) = (fx, fy, cx, cy) used by the last step."""
        return self._k4


class ShardedFusedOverfitter(FusedOverfitter):
    """Pair-sharded :class:`FusedOverfitter` (flowmap_b200.parallel).

    Every rank holds the frames / pairs of its ShardPlan and runs fm_overfit_step on them
    without the Adam part; ONE all-reduce carries the loss, the focal-length gradient and the
    boundary depth-gradient frames, then each rank applies Adam to its parameters (replicas of
    a boundary frame see identical gradients).

    Tracking loss (not pair-local: a track segment spans up to 41 frames): the step is split
    (FM_STEP_FORWARD / FM_STEP_BACKWARD).  In between, the relative poses are gathered (149 x 12
    floats), every rank chains them, evaluates the tracking loss for the SOURCE frames it owns
    against all target frames, the loss sum / valid count / per-frame pose and intrinsics sums
    (F x 10 doubles) are all-reduced, and every rank backpropagates the (now global) pose
    gradient through the chain to its own pairs.  `tracks` are the global segments.

    Softmin intrinsics: the candidate sweep lives on the rank that owns pair 0; its focal
    estimate is broadcast before the step and its backward runs after the step's all-reduce
    (the summed d loss / d focal), so the sweep adds two one-float collectives."""

    def __init__(self, cfg: OverfitCfg, batch: Batch, flows: Flows, plan, tracks=None, device="cuda",
                 group=None):
        from dataclasses import replace
        from . import parallel
        from ._lib import lib
        super().__init__(replace(cfg, use_tracking=False), batch, flows, None, device)
        self.cfg = cfg
        self.plan, self.group = plan, group
        self._args.clock = None  # this driver passes Adam's step numbers by value
        _, f_local, _, h, w = batch.videos.shape
        if f_local != plan.num_local_frames:
            raise ValueError("flowmap_b200: batch does not match the shard plan")
        dev = self.rt.device
        self.reducer = parallel.StepReducer(plan, (h, w), dev, 2, group)
        self._msum.copy_(parallel.global_mask_sum(self._msum, group))  # in place: args hold its address
        if self._softmin and plan.world > 1 and plan.pair_range[1] - plan.pair_range[0] < 2 and plan.rank == 0:
            raise ValueError("flowmap_b200: the rank owning pair 0 needs >= 2 pairs in the softmin stage")
        if cfg.use_tracking:
            if tracks is None:
                raise ValueError("flowmap_b200: use_tracking needs the (global) track segments")
            F = plan.num_pairs_total + 1
            pk = ops.PackedTracks([t.to(dev) for t in tracks], dev)
            self._packed, self._F = pk, F
            z = lambda *shape: torch.zeros(*shape, device=dev)  # noqa: E731
            self._rt_all, self._g_rt_all = z(1, F - 1, 3, 4), z(1, F - 1, 3, 4)
            self._ext_all, self._g_ext_all = z(1, F, 4, 4), z(1, F, 4, 4)
            self._tg_k4_all = z(F, 4)
            self._g_rt_local = z(1, f_local - 1, 3, 4)
            self._tws = torch.empty(lib().fm_track_workspace_bytes(F, pk.total), dtype=torch.uint8, device=dev)
            self._treduce = self._tws[:lib().fm_track_reduce_bytes(F)].view(torch.float64)
            self._src_range = parallel.source_frame_range(plan)

    def _mask_sum(self, flows: Flows) -> Tensor:
        from . import parallel
        return parallel.global_mask_sum(ops.mask_sum(flows.forward_mask, flows.backward_mask), self.group)

    def sync_boundary_depth(self):
        """Make the replicas of every shared boundary frame identical (owner = left rank)."""
        import torch.distributed as dist
        p = self.plan
        if p.world == 1:
            return
        buf = torch.zeros(p.world - 1, *self._depth.shape[1:], device=self._depth.device)
        if p.has_right:
            buf[p.rank].copy_(self._depth[-1])
        dist.all_reduce(buf, group=self.group)
        if p.has_left:
            self._depth[0].copy_(buf[p.rank - 1])

    def _first_rank(self) -> int:
        import torch.distributed as dist
        return 0 if self.group is None else dist.get_global_rank(self.group, 0)

    def _tracking_exchange(self, h: int, w: int):
        """Between the two halves of a split step; returns tracking's d loss / d focal (global)."""
        import torch.distributed as dist
        from . import parallel
        from ._lib import check
        c, L, pk, F = self.cfg, self._lib, self._packed, self._F
        a0, b0 = self.plan.pair_range
        P = lambda t: t.data_ptr()  # noqa: E731
        st = torch.cuda.current_stream().cuda_stream
        parallel.gather_pairs(self.plan, self.rt, self._rt_all, self.group)
        k4_all = self._k4[0].expand(F, 4).contiguous()  # one shared focal length
        args = (P(k4_all), P(self._ext_all), P(pk.seg), pk.num_segments, pk.max_rows, pk.max_points, P(pk.xy),
                P(pk.vis), pk.total, ops.MAPPINGS[c.mapping], c.delta, c.tracking_weight)
        tail = (F, h, w, a0, self._src_range[0], self._src_range[1])
        with torch.cuda.device(self.rt.device):
            check(L.fm_pose_chain(P(self._rt_all), P(self._ext_all), 1, F, st), "fm_pose_chain")
            check(L.fm_track_loss_fwd_sharded(P(self._depth), *args, None, P(self._tws), *tail, 1, st),
                  "fm_track_loss_fwd_sharded")  # shared focal: only the summed K gradient is used
            if self.plan.world > 1:
                dist.all_reduce(self._treduce, group=self.group)
            check(L.fm_track_loss_value(P(self._tws), c.tracking_weight, P(self._track_loss), st),
                  "fm_track_loss_value")
            check(L.fm_track_loss_bwd_sharded(P(self._depth), *args, None, P(self._g_depth), P(self._g_ext_all),
                                              P(self._tg_k4_all), P(self._tws), *tail, st),
                  "fm_track_loss_bwd_sharded")
            check(L.fm_pose_chain_bwd(P(self._rt_all), P(self._ext_all), P(self._g_ext_all),
                                      P(self._g_rt_all), 1, F, st), "fm_pose_chain_bwd")
        self._g_rt_local.copy_(self._g_rt_all[:, a0:b0])
        scale = (h * w) ** 0.5
        return (self._tg_k4_all[:, 0].double().sum() * (scale / w) +
                self._tg_k4_all[:, 1].double().sum() * (scale / h)).float()

    def _flow_only_body(self):
        """Flow loss with a regressed focal length, update step, as a fixed launch sequence: the step
        (weight-logit Adam fused in), the exchange started, Adam on the interior depth frames WHILE the
        boundary frames and the two scalars travel, then the boundary frames and the focal length."""
        from ._lib import check
        c, a, p, r = self.cfg, self._args, self.plan, self.reducer
        P = lambda t: None if t is None else t.data_ptr()  # noqa: E731
        clk = self._clock
        clk.tick(tick_focal=True)
        a.clock, a.tracks, a.step, a.focal_step, a.defer_adam, a.phase = clk.ptr, None, 1, 1, 2, 0
        a.focal = P(self._focal)
        # the step writes its two scalars (loss, d focal) straight into the all-reduce buffer
        loss_ptr, gf_ptr = a.loss, a.g_focal
        a.loss, a.g_focal = r.scal[0:1].data_ptr(), r.scal[1:2].data_ptr()
        try:
            with torch.cuda.device(self.rt.device):
                check(self._lib.fm_overfit_step(self._ctypes.byref(a), torch.cuda.current_stream().cuda_stream),
                      "fm_overfit_step")
        finally:
            a.clock, a.step, a.focal_step, a.defer_adam = None, 0, 0, 0
            a.loss, a.g_focal = loss_ptr, gf_ptr
        reqs = r.start(self._g_depth)
        stt, n = self._state, self._depth.shape[0]
        lo, hi = int(p.has_left and p.world > 1), n - int(p.has_right and p.world > 1)
        if hi > lo:
            ops.adam_step_clock(self._depth[lo:hi], self._g_depth[lo:hi], stt[0][lo:hi], stt[1][lo:hi], clk)
        red = r.finish(reqs, self._g_depth)
        if lo > 0:
            ops.adam_step_clock(self._depth[:1], self._g_depth[:1], stt[0][:1], stt[1][:1], clk)
        if hi < n:
            ops.adam_step_clock(self._depth[n - 1:], self._g_depth[n - 1:], stt[0][n - 1:], stt[1][n - 1:], clk)
        ops.adam_step_clock(self._focal.reshape(1), red[1:2], stt[4].reshape(1), stt[5].reshape(1), clk,
                            focal_clock=True)
        self._total.copy_(red[0])

    def training_step(self, update: bool = True):
        import torch.distributed as dist
        from ._lib import check
        c, a, p, L = self.cfg, self._args, self.plan, self._lib
        _, _, _, h_, w_ = self.batch.videos.shape
        if (update and not self._softmin and not c.procrustes_randomize and self._indices is None and
                c.use_correspondence_weights and w_ % 4 == 0 and
                not (c.use_tracking and self.global_step >= c.tracking_enable_after)):
            a.indices, a.num_indices = None, 0
            a.flow_weight = c.flow_weight if self.global_step >= c.flow_enable_after else 0.0
            self._clock.set(self.optimizer_steps, self.focal_steps)
            self._run_body(("flow", self.global_step >= c.flow_enable_after), self.use_cuda_graph,
                           self._flow_only_body, True)
            self.global_step += 1
            self.optimizer_steps += 1
            self.focal_steps += 1
            return self._total.clone(), self.rt
        _, _, _, h, w = self.batch.videos.shape
        P = lambda t: None if t is None else t.data_ptr()  # noqa: E731
        st = torch.cuda.current_stream().cuda_stream
        dev = self.rt.device
        if c.procrustes_randomize:
            self._indices = self.model.extrinsics.select_indices(h, w, dev)
        a.indices = None if self._indices is None else self._indices.data_ptr()
        a.num_indices = 0 if self._indices is None else self._indices.numel()
        a.flow_weight = c.flow_weight if self.global_step >= c.flow_enable_after else 0.0
        track_on = c.use_tracking and self.global_step >= c.tracking_enable_after
        sweep = self._softmin_stage()
        own_sweep = sweep and p.rank == 0
        # the weight logits' gradient is rank-local and final inside the step: update them there
        # (all pairs, or pairs >= 1 on the rank whose sweep still touches pair 0); depth and the
        # focal length wait for the exchange below
        fuse_w = update and c.use_correspondence_weights and self._indices is None and w % 4 == 0
        a.tracks = None
        a.step = self.optimizer_steps + 1 if fuse_w else 0
        a.defer_adam = (1 if own_sweep else 2) if fuse_w else 0
        if sweep:
            n = c.softmin_candidates
            wl = P(self._wlog) if c.use_correspondence_weights else None
            sens = c.weight_sensitivity if c.use_correspondence_weights else 0.0
            if own_sweep:
                idx = self.injected_indices
                if idx is None:
                    idx = ops.random_subset(h * w, min(c.softmin_points, h * w), dev)
                idx = idx.contiguous()
                f_local = self._depth.shape[0]
                with torch.cuda.device(dev):
                    check(L.fm_softmin_sweep_fwd(P(self._depth), wl, sens, P(self.flows.backward), P(idx),
                                                 idx.numel(), P(self._cand_k4), n, P(self._sw_err),
                                                 P(self._sw_rt), P(self._sw_ws), 1, f_local, h, w, st),
                          "fm_softmin_sweep_fwd")
                    check(L.fm_softmin_focal(P(self._sw_err), P(self._cand_f), n, 1, P(self._sw_sm),
                                             P(self._sw_focal), st), "fm_softmin_focal")
            if p.world > 1:
                dist.broadcast(self._sw_focal, src=self._first_rank(), group=self.group)
            a.focal = P(self._sw_focal)
        else:
            a.focal = P(self._focal)
            if self._softmin and self.global_step == c.regression_after and update:
                self._focal.copy_(torch.stack(self.window).mean())  # hand-over, identical on all ranks
        extra_focal = None
        with torch.cuda.device(dev):
            if track_on:
                a.phase = 1  # FM_STEP_FORWARD
                check(L.fm_overfit_step(self._ctypes.byref(a), st), "fm_overfit_step (forward)")
                extra_focal = self._tracking_exchange(h, w)
                a.phase, a.g_rt, a.track_g_k4 = 2, P(self._g_rt_local), None  # FM_STEP_BACKWARD
                check(L.fm_overfit_step(self._ctypes.byref(a), st), "fm_overfit_step (backward)")
                a.phase, a.g_rt = 0, None
            else:
                check(L.fm_overfit_step(self._ctypes.byref(a), st), "fm_overfit_step")
        a.step, a.defer_adam = 0, 0
        g_focal = self._g_focal.reshape(())
        if extra_focal is not None and p.rank == 0:  # global value, counted once
            g_focal = g_focal + extra_focal
        red = self.reducer.reduce(torch.stack((self._loss.reshape(()), g_focal)), self._g_depth)
        self._g_focal.copy_(red[1])
        if own_sweep:  # backward of the sweep with the summed focal gradient: frames 0/1, pair 0
            f_local = self._depth.shape[0]
            with torch.cuda.device(dev):
                check(L.fm_softmin_focal_bwd(P(self._sw_sm), P(self._cand_f), P(self._sw_focal),
                                             P(self._g_focal), n, 1, P(self._sw_gerr), st),
                      "fm_softmin_focal_bwd")
                check(L.fm_softmin_sweep_bwd(P(self._depth), wl, sens, P(self.flows.backward), P(idx),
                                             idx.numel(), P(self._cand_k4), n, P(self._sw_rt),
                                             P(self._sw_gerr), P(self._g_depth),
                                             P(self._g_w) if wl else None, P(self._sw_ws), 1, f_local, h, w, st),
                      "fm_softmin_sweep_bwd")
        if update:
            s_ = self.optimizer_steps + 1
            stt = self._state
            ops.adam_step(self._depth, self._g_depth, stt[0], stt[1], s_, c.lr)
            if c.use_correspondence_weights and (not fuse_w or own_sweep):
                k = 1 if fuse_w else self._wlog.shape[0]  # pair 0 only when the rest was fused
                ops.adam_step(self._wlog[:k], self._g_w[:k], stt[2][:k], stt[3][:k], s_, c.lr)
            if sweep:
                if c.regression_after is not None and self.global_step >= c.regression_after - c.regression_window:
                    self.window.append(self._sw_focal[0].clone())
            else:
                self.focal_steps += 1
                ops.adam_step(self._focal.reshape(1), self._g_focal.reshape(1), stt[4].reshape(1),
                              stt[5].reshape(1), self.focal_steps, c.lr)
            self.global_step += 1
            self.optimizer_steps += 1
        total = red[0] + self._track_loss if track_on else red[0]
        return total, self.rt


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
