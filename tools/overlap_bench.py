"""Experiment: how much do the three pixel-parallel ops gain when chunks of pairs are pipelined
over several streams (different bottlenecks: L1 gather / FP32 issue / L2 RED; chunk working set
L2-resident)?  Timing only: every chunk is run as an independent sub-video through the per-op
C ABI, so boundary frames are written by two chunks (results are not used)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from flowmap_b200 import ops  # noqa: E402
from flowmap_b200._lib import lib  # noqa: E402

F, H, W = 150, 360, 640
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
depths = (0.1 + 0.05 * torch.rand(1, F, H, W, device=dev, generator=g))
weights = torch.sigmoid(torch.randn(1, F - 1, H, W, device=dev, generator=g))
fwd = 0.01 * torch.randn(1, F - 1, H, W, 2, device=dev, generator=g)
bwd = 0.01 * torch.randn(1, F - 1, H, W, 2, device=dev, generator=g)
fm = torch.rand(1, F - 1, H, W, device=dev, generator=g)
bm = torch.rand(1, F - 1, H, W, device=dev, generator=g)
s = (H * W) ** 0.5
k4 = torch.tensor([0.85 * s / W, 0.85 * s / H, 0.5, 0.5], device=dev).expand(1, F, 4).contiguous()
msum = ops.mask_sum(fm, bm)
rt = torch.empty(1, F - 1, 3, 4, device=dev)
g_depth, g_w, g_k4, g_rt = (torch.empty_like(depths), torch.empty_like(weights),
                            torch.empty_like(k4), torch.empty_like(rt))
P = lambda x: x.data_ptr()  # noqa: E731
L = lib()


def chunk_calls(a, b, ws, loss, st, stages=(0, 1, 2)):
    """pairs [a, b) = frames [a, b] as a sub-video."""
    f = b - a + 1
    d, k, r = depths[:, a:b + 1], k4[:, a:b + 1], rt[:, a:b]
    if 0 in stages:
        L.fm_procrustes_fwd(P(d), P(k), P(bwd[:, a:b]), P(weights[:, a:b]), None, 0, P(r), P(ws), 1, f, H, W, st)
    if 1 in stages:
        L.fm_flow_loss_fwd_bwd(P(d), P(k), P(r), P(fwd[:, a:b]), P(bwd[:, a:b]), P(fm[:, a:b]), P(bm[:, a:b]),
                               P(msum), 0, 0.01, 1000.0, 1, P(loss), P(g_depth[:, a:b + 1]), P(g_rt[:, a:b]),
                               P(g_k4[:, a:b + 1]), P(ws), 1, f, H, W, st)
    if 2 in stages:
        L.fm_procrustes_bwd(P(d), P(k), P(bwd[:, a:b]), P(weights[:, a:b]), None, 0, None, 1, None,
                            P(g_depth[:, a:b + 1]), P(g_w[:, a:b]), P(g_k4[:, a:b + 1]), P(ws), 1, f, H, W, st)


def run(chunk, nstreams, reps=20, order="chunk"):
    bounds = [(a, min(a + chunk, F - 1)) for a in range(0, F - 1, chunk)]
    wss = [ops.workspace(1, b - a + 1, H, W, dev) for a, b in bounds]
    losses = [torch.empty((), device=dev) for _ in bounds]
    streams = [torch.cuda.Stream() for _ in range(nstreams)] if nstreams > 0 else [torch.cuda.current_stream()]
    main = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def step():
        if nstreams > 0:
            for s_ in streams:
                s_.wait_stream(main)
        if order == "chunk":
            for i, (a, b) in enumerate(bounds):
                s_ = streams[i % len(streams)]
                chunk_calls(a, b, wss[i], losses[i], s_.cuda_stream)
        else:  # stage-skewed issue order: chunk i stage j issued at time i + j
            n = len(bounds)
            for t in range(n + 2):
                for j in (2, 1, 0):
                    i = t - j
                    if 0 <= i < n:
                        a, b = bounds[i]
                        chunk_calls(a, b, wss[i], losses[i], streams[i % len(streams)].cuda_stream, (j,))
        if nstreams > 0:
            for s_ in streams:
                main.wait_stream(s_)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


print("whole video, one stream: %.3f ms" % run(F - 1, 0))
for chunk in (4, 8, 16, 32):
    print("chunk %2d pairs, serial on one stream: %.3f ms" % (chunk, run(chunk, 0)))
    for ns in (2, 3, 4, 6):
        print("chunk %2d pairs, %d streams: %.3f ms   skewed issue: %.3f ms" %
              (chunk, ns, run(chunk, ns), run(chunk, ns, order="skew")))
