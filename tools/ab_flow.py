"""A/B of the flow-loss kernel on a B200: k_flow_lean (default) vs the cp.async-staged variant
(k_flow_lean_staged, FM_FLOW_STAGED=1).  Same inputs through fm_procrustes_fwd ->
fm_flow_loss_fwd_bwd in both modes: loss, depth gradient, pose and intrinsics gradients must agree;
then the per-launch time of fm_flow_loss_fwd_bwd and of the fused step in both modes.

Usage: python tools/ab_flow.py    (writes gpurun_out/ab_flow.json)"""
import json
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from flowmap_b200 import ops  # noqa: E402
from flowmap_b200._lib import lib  # noqa: E402
from flowmap_b200.overfit import FusedOverfitter, OverfitCfg  # noqa: E402
from flowmap_b200.types import Batch, Flows, Tracks  # noqa: E402

dev = torch.device("cuda:0")
P = lambda x: x.data_ptr()  # noqa: E731
MODES = {"default": "0", "staged": "1"}


def set_mode(m):
    os.environ["FM_FLOW_STAGED"] = MODES[m]


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def setup(f, h, w, k_mode, seed=0):
    c = {k: v.to(dev).contiguous() for k, v in bench.synthetic_inputs(f, h, w, seed=seed).items()}
    L = lib()
    st = torch.cuda.current_stream().cuda_stream
    depths = c["depth"][None].contiguous()
    weights = torch.sigmoid(100.0 * c["wparam"])[None].contiguous()
    s_ = (h * w) ** 0.5
    k4 = torch.tensor([0.85 * s_ / w, 0.85 * s_ / h, 0.5, 0.5], device=dev).expand(1, f, 4).contiguous()
    msum = ops.mask_sum(c["fmask"], c["bmask"])
    ws = ops.workspace(1, f, h, w, dev)
    rt = torch.empty(1, f - 1, 3, 4, device=dev)
    g_depth = torch.zeros_like(depths)
    g_k4, g_rt = torch.zeros_like(k4), torch.zeros_like(rt)
    lossb = torch.zeros((), device=dev)
    assert L.fm_procrustes_fwd(P(depths), P(k4), P(c["bwd"]), P(weights), None, 0, P(rt), P(ws), 1, f, h, w, st) == 0

    def flow():
        assert L.fm_flow_loss_fwd_bwd(P(depths), P(k4), P(rt), P(c["fwd"]), P(c["bwd"]), P(c["fmask"]), P(c["bmask"]),
                                      P(msum), 0, 0.01, 1000.0, k_mode, P(lossb), P(g_depth), P(g_rt), P(g_k4), P(ws),
                                      1, f, h, w, st) == 0, L.fm_last_error()
    return flow, (lossb, g_depth, g_rt, g_k4)


def compare(f, h, w, k_mode):
    res = {}
    for m in MODES:
        set_mode(m)
        flow, outs = setup(f, h, w, k_mode)
        flow()
        torch.cuda.synchronize()
        res[m] = [o.clone() for o in outs]
    r = {"shape": [f, h, w], "k_mode": k_mode,
         "loss_rel": rel(res["staged"][0], res["default"][0]),
         "g_depth_equal": bool(torch.equal(res["staged"][1], res["default"][1])),
         "g_depth_rel": rel(res["staged"][1], res["default"][1]),
         "g_rt_rel": rel(res["staged"][2], res["default"][2]),
         "g_k4_rel": rel(res["staged"][3], res["default"][3])}
    # same per-pixel arithmetic, same thread -> pixel mapping: only the float64 atomics may reorder
    r["ok"] = r["g_depth_rel"] <= 1e-7 and r["loss_rel"] <= 1e-9 and r["g_rt_rel"] <= 1e-9 and r["g_k4_rel"] <= 1e-9
    print("compare", json.dumps(r), flush=True)
    return r


def time_flow(n=20):
    out = {}
    for m in MODES:
        set_mode(m)
        flow, _ = setup(bench.F_, bench.H_, bench.W_, 1)
        for _ in range(3):
            flow()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(n):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); flow(); b.record()
            torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        out[m] = tot / n
    print("fm_flow_loss_fwd_bwd 150x360x640  " + "  ".join(f"{m} {out[m]:.4f} ms" for m in MODES), flush=True)
    return out


def time_steps(steps=40):
    F, H, W = bench.F_, bench.H_, bench.W_
    inp = bench.synthetic_inputs(F, H, W, seed=0)
    batch = Batch(torch.zeros(1, 1, 1, 1, 1, device=dev).expand(1, F, 3, H, W), torch.arange(F, device=dev)[None], ["s"], ["d"])
    flows = Flows(*(inp[k].to(dev) for k in ("fwd", "bwd", "fmask", "bmask")))
    tracks = [Tracks(xy, vis, s) for xy, vis, s in bench.synthetic_track_arrays(F, seed=0)]

    def timed(o):
        with torch.no_grad():
            o.model.backbone.depth.copy_(inp["depth"])
            o.model.backbone.weights.copy_(inp["wparam"])
        o.global_step = bench.START_STEP
        for _ in range(5):
            o.training_step()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            out = o.training_step()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps, float(out[0])
    out = {}
    for m in MODES:
        set_mode(m)
        full, loss = timed(FusedOverfitter(OverfitCfg(intrinsics="softmin", use_tracking=True), batch, flows, tracks, device=dev))
        flow_only, _ = timed(FusedOverfitter(OverfitCfg(), batch, flows, device=dev))
        out[m] = {"full_ms": full, "flow_only_ms": flow_only, "loss": loss}
        print(f"fused step {m:8s} full {full:.4f} ms  flow-only {flow_only:.4f} ms  loss {loss:.6f}", flush=True)
    return out


def main():
    report = {"compare": []}
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)

    def save():
        report["all_ok"] = all(r["ok"] for r in report["compare"])
        (out / "ab_flow.json").write_text(json.dumps(report, indent=1))
    # the most informative results first (a short GPU visit may be cut off)
    report["compare"].append(compare(3, 360, 640, 1))
    report["flow_ms"] = time_flow()
    save()
    for f, h, w in [(2, 16, 24), (3, 24, 32), (5, 72, 96), (4, 128, 128), (3, 360, 640), (2, 720, 1280)]:
        for k_mode in (1, 2):  # shared focal / constant intrinsics: the two lean instantiations
            report["compare"].append(compare(f, h, w, k_mode))
    save()
    print("ALL_OK" if report["all_ok"] else "MISMATCH", flush=True)
    report["steps"] = time_steps()
    save()


if __name__ == "__main__":
    main()
