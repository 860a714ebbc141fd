"""Checks the analytic forward/backward formulas that the CUDA kernels instantiate
(flowmap_b200/csrc/fm_pixel.cuh, fm_procrustes.cuh) by compiling the same headers with g++
and driving them serially on the CPU (tests/host_emulation/emu.cpp, test-only), against the
golden vectors of the reference.  Runs in the build container (no GPU)."""
import ctypes
import subprocess
from pathlib import Path

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden, max_abs, rel_l2

EMU_DIR = ROOT / "tests" / "host_emulation"


@pytest.fixture(scope="module")
def emu():
    build = EMU_DIR / "_build"
    build.mkdir(exist_ok=True)
    so = build / "libemu.so"
    srcs = [EMU_DIR / "emu.cpp"] + sorted((ROOT / "flowmap_b200" / "csrc").glob("*.cuh"))
    if not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", str(so),
                               str(EMU_DIR / "emu.cpp")])
    lib = ctypes.CDLL(str(so))
    lib.emu_state_bytes.restype = ctypes.c_size_t
    return lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def flow_pose_grad(flowacc, rt, B, F):
    """numpy twin of flow_pose_grad() in fm_kernels.cu."""
    g = np.zeros((B * (F - 1), 12))
    for pair in range(B * (F - 1)):
        bi, i = divmod(pair, F - 1)
        a = bi * F + i
        fa, fb = flowacc[a], flowacc[a + 1]
        R = rt[pair].reshape(3, 4)[:, :3].astype(np.float64)
        G = np.zeros((3, 4))
        G[:, :3] = fa[1:10].reshape(3, 3) + fb[13:22].reshape(3, 3)
        G[:, 3] = fb[22:25] - R @ fa[10:13]
        g[pair] = G.reshape(-1)
    return g


def flow_k4_grad(flowacc, B, F):
    g = np.zeros((B * F, 4))
    for fr in range(B * F):
        i = fr % F
        g[fr] = flowacc[fr, 25:29]
        if i > 0:
            g[fr] += flowacc[fr - 1, 29:33]
        if i < F - 1:
            g[fr] += flowacc[fr + 1, 33:37]
    return g


def run_emulated_step(emu, g, mapping=0, focal=0.85, indices=None, delta=0.01, weight=1000.0,
                      lean=False):
    depth = np.ascontiguousarray(g["in_depth"], dtype=np.float32)
    F_, H, W = depth.shape
    B = 1
    wparam = torch.as_tensor(g["in_wparam"], dtype=torch.float32)
    w = torch.sigmoid(100.0 * wparam).numpy()
    ff = np.ascontiguousarray(g["in_fwd"], dtype=np.float32)
    fb = np.ascontiguousarray(g["in_bwd"], dtype=np.float32)
    mf = np.ascontiguousarray(g["in_fmask"], dtype=np.float32)
    mb = np.ascontiguousarray(g["in_bmask"], dtype=np.float32)
    s = (H * W) ** 0.5
    k4 = np.tile(np.array([focal * s / W, focal * s / H, 0.5, 0.5], dtype=np.float32), (F_, 1))
    BP = F_ - 1
    rt = np.zeros((BP, 12), dtype=np.float32)
    state = np.zeros(BP * emu.emu_state_bytes(), dtype=np.uint8)
    idx = None if indices is None else np.ascontiguousarray(indices, dtype=np.int64)
    n_idx = 0 if idx is None else len(idx)
    emu.emu_procrustes_fwd(_p(depth), _p(k4), _p(fb), _p(w), _p(idx), n_idx, _p(rt), _p(state),
                           B, F_, H, W)
    mask_sum = float(mf.astype(np.float64).sum() + mb.astype(np.float64).sum())
    g_depth = np.zeros_like(depth)
    flowacc = np.zeros((F_, 40), dtype=np.float64)
    if lean:  # shared-focal kernel: twist pose accumulators, one focal accumulator
        emu.emu_flow_lean(_p(depth), _p(k4), _p(rt), _p(ff), _p(fb), _p(mf), _p(mb),
                          ctypes.c_double(mask_sum), mapping, ctypes.c_float(delta),
                          ctypes.c_float(weight), 1, _p(g_depth), _p(flowacc), B, F_, H, W)
    else:
        emu.emu_flow(_p(depth), _p(k4), _p(rt), _p(ff), _p(fb), _p(mf), _p(mb),
                     ctypes.c_double(mask_sum), mapping, ctypes.c_float(delta),
                     ctypes.c_float(weight), _p(g_depth), _p(flowacc), B, F_, H, W)
    loss = flowacc[:, 0].sum()
    g_rt = flow_pose_grad(flowacc, rt, B, F_)
    g_w = np.zeros_like(w)
    k4acc = np.zeros((F_, 4), dtype=np.float64)
    direct = g_depth.copy()
    emu.emu_procrustes_bwd(_p(depth), _p(k4), _p(fb), _p(w), _p(idx), n_idx, _p(state), _p(g_rt),
                           _p(g_depth), _p(g_w), _p(k4acc), B, F_, H, W)
    g_k4 = flow_k4_grad(flowacc, B, F_) + k4acc
    g_focal = (g_k4[:, 0] * s / W + g_k4[:, 1] * s / H).sum()
    wt = torch.as_tensor(w, dtype=torch.float64)
    g_wparam = torch.as_tensor(g_w, dtype=torch.float64) * 100.0 * wt * (1 - wt)
    # chain the relative poses (float64 here; the float32 chain is checked on the GPU)
    P = [np.eye(4)]
    for i in range(BP):
        T = np.eye(4)
        T[:3] = rt[i].reshape(3, 4)
        P.append(P[-1] @ T)
    return dict(loss=loss, extrinsics=np.stack(P)[None], g_depth=g_depth, g_wparam=g_wparam.numpy(),
                g_focal=g_focal, direct=direct, rt=rt)


CASES = [("flow_huber", 0, 0.85, None), ("flow_l1", 1, 0.85, None), ("flow_l2", 2, 0.85, None),
         ("flow_rough", 0, 1.3, None), ("flow_pts1000", 0, 0.85, 1000)]


@pytest.mark.parametrize("lean", [False, True])
@pytest.mark.parametrize("name,mapping,focal,npts", CASES)
def test_emulated_step_matches_reference(emu, name, mapping, focal, npts, lean):
    g64 = load_golden(name, f64=True)
    g32 = load_golden(name, f64=False)
    _, H, W = g64["in_depth"].shape
    idx = None if npts is None else torch.linspace(0, H * W - 1, npts, dtype=torch.int64).numpy()
    r = run_emulated_step(emu, g64, mapping=mapping, focal=focal, indices=idx, lean=lean)
    # float32 noise floor of the reference itself (float32 run vs float64 run of the reference)
    ref_noise_d = rel_l2(g32["g_depth"], g64["g_depth"])
    ref_noise_w = rel_l2(g32["g_wparam"], g64["g_wparam"])
    assert abs(r["loss"] - float(g64["loss"])) <= 2e-5 * abs(float(g64["loss"]))
    assert max_abs(r["extrinsics"], g64["extrinsics"]) <= 5e-6
    assert rel_l2(r["g_depth"], g64["g_depth"]) <= max(1e-4, 3 * ref_noise_d)
    assert rel_l2(r["g_wparam"], g64["g_wparam"]) <= max(1e-4, 3 * ref_noise_w)
    assert abs(r["g_focal"] - float(g64["g_focal"])) <= 1e-4 * abs(float(g64["g_focal"]))


@pytest.mark.parametrize("total,rounds,grid", [
    (33525, 1, 444), (33525, 9, 444), (134100, 37, 444),   # 149 pairs of 360x640 / 720x1280, B200 grid
    (16762, 4, 444),                                        # one rank of an 8-way split at 720p
    (7, 3, 5), (5, 8, 444), (1000, 1, 1), (4096, 16, 3), (0, 2, 4),
])
def test_item_decomposition_covers_every_item_once_and_is_balanced(emu, total, rounds, grid):
    """block_item_range of the persistent dense kernels (item_span in fm_math.cuh): the (round, block)
    spans tile the item list exactly, and the rotation between rounds keeps the blocks' totals within
    a couple of items of each other (one odd item per round, moved around)."""
    cover = np.zeros(max(total, 1), dtype=np.int32)
    minmax = np.zeros(2, dtype=np.int64)
    emu.emu_item_cover(ctypes.c_longlong(total), rounds, grid, _p(cover), _p(minmax))
    assert (cover[:total] == 1).all()
    if total >= grid * rounds:
        assert minmax[1] - minmax[0] <= 2


@pytest.mark.parametrize("mapping", [0, 1, 2], ids=["huber", "l1", "l2"])
def test_packed_two_point_term_equals_the_scalar_term(emu, mapping):
    """lean_term2 (two points as one float32x2 computation: k_flow_lean's pixel pairs, k_track_src's
    point pairs) against lean_term on each point, incl. a point behind the camera plane
    (z + eps == 0: the nan_to_num branch of projection.py:56) next to an ordinary one."""
    rng = np.random.default_rng(7)
    for case in range(200):
        D = rng.uniform(0.5, 2.0, 2).astype(np.float32)
        dirs = rng.normal(0, 0.3, (2, 3)).astype(np.float32)
        dirs[:, 2] = rng.uniform(0.7, 1.3, 2)
        off = rng.normal(0, 0.05, 3).astype(np.float32)
        if case % 10 == 0:  # first point exactly on the plane z + eps = 0
            dirs[0, 2] = 0.0
            off[2] = np.float32(-1e-5)
        k4 = np.array([0.9, 1.2, 0.5, 0.5], dtype=np.float32)
        xy = rng.uniform(0, 1, (2, 2)).astype(np.float32)
        fl = rng.normal(0, 0.01, (2, 2)).astype(np.float32)
        out, out2 = np.zeros(14, np.float32), np.zeros(14, np.float32)
        emu.emu_lean_terms(_p(D), _p(dirs), _p(off), _p(k4), _p(xy), _p(fl), mapping, ctypes.c_float(0.01), 36, 48,
                           _p(out), _p(out2))
        ok = np.isfinite(out)
        assert (np.isfinite(out2) == ok).all()
        assert np.allclose(out2[ok], out[ok], rtol=2e-6, atol=1e-7), (case, out, out2)
