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


# ------------------------------------------------------------------------------------------
# Tiled Procrustes-adjoint scatter (k_distribute_tiled): fixed-point shared-memory window.
# ------------------------------------------------------------------------------------------
def test_fixed_point_cell_is_an_exact_integer_sum(emu):
    """(high, low) words of a window cell = the exact integer sum of the rounded contributions,
    whatever their sign pattern and however often the low word wraps (fm_pixel.cuh: fix_add)."""
    rng = np.random.default_rng(0)
    lim = 2.0 ** 30
    cases = [rng.uniform(-lim, lim, 4000), rng.uniform(0.9 * lim, lim, 3000) * 0.999,
             -rng.uniform(0.9 * lim, lim, 3000) * 0.999, rng.uniform(-3, 3, 1000), np.zeros(5),
             np.array([-1.0]), np.array([-0.4, 0.4, -0.5, 0.5, 1.5, 2.5])]
    for vals in cases:
        v = np.ascontiguousarray(vals, dtype=np.float32)
        v = v[np.abs(v) < lim]
        lo, hi, out = ctypes.c_uint(), ctypes.c_int(), ctypes.c_float()
        emu.emu_fix_accumulate(_p(v), len(v), ctypes.byref(lo), ctypes.byref(hi), ctypes.byref(out))
        exact = int(sum(int(np.rint(x)) for x in v.astype(np.float64)))  # rint: ties to even, like cvt.rni
        assert hi.value * 2 ** 32 + lo.value - 2 ** 31 == exact
        assert out.value == np.float32(exact)


def test_fixed_point_scale_maps_the_bound_below_2_pow_29(emu):
    for bound in [1e-25, 3e-7, 0.999, 1.0, 1.0001, 777.0, 6e12, 1e30]:
        s, i = ctypes.c_float(), ctypes.c_float()
        emu.emu_fix_scale(ctypes.c_float(bound), ctypes.byref(s), ctypes.byref(i))
        assert 2.0 ** 28 <= np.float32(bound) * np.float32(s.value) < 2.0 ** 29
        assert s.value * i.value == 1.0 and np.log2(s.value) == np.rint(np.log2(s.value))
    for bound in [0.0, float("nan"), float("inf"), 1e-45, 1e-30]:  # clamped, still an exact power of two
        s, i = ctypes.c_float(), ctypes.c_float()
        emu.emu_fix_scale(ctypes.c_float(bound), ctypes.byref(s), ctypes.byref(i))
        assert np.isfinite(s.value) and s.value > 0 and s.value * i.value == 1.0
        assert not bound * s.value >= 2.0 ** 29 or not np.isfinite(bound)


def _tiled_case(emu, F_, H, W, seed=0, sigma=0.01, outliers=0.0, wscale=1.0, shift=(0.0, 0.0),
                depth_scale=1.0, grt_scale=1.0, tall=False):
    rng = np.random.default_rng(seed)
    _, xx = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    depth = (depth_scale * (1.0 + 0.5 * rng.random((F_, H, W)) + 0.3 * np.sin(xx / 17.0)[None])).astype(np.float32)
    fb = (sigma * rng.standard_normal((F_ - 1, H, W, 2))).astype(np.float32)
    fb[..., 0] += shift[0]
    fb[..., 1] += shift[1]
    if outliers > 0:
        m = rng.random((F_ - 1, H, W)) < outliers
        fb[m] = (0.5 * rng.standard_normal((int(m.sum()), 2))).astype(np.float32)
    w = (wscale * rng.random((F_ - 1, H, W))).astype(np.float32)
    s = (H * W) ** 0.5
    k4 = np.tile(np.array([0.85 * s / W, 0.85 * s / H, 0.5, 0.5], dtype=np.float32), (F_, 1))
    BP = F_ - 1
    rt = np.zeros((BP, 12), dtype=np.float32)
    state = np.zeros(BP * emu.emu_state_bytes(), dtype=np.uint8)
    emu.emu_procrustes_fwd(_p(depth), _p(k4), _p(fb), _p(w), None, 0, _p(rt), _p(state), 1, F_, H, W)
    g_rt = (grt_scale * rng.standard_normal((BP, 12))).astype(np.float64)
    res = []
    for tiled in (False, True):
        gd, gw, k4acc = np.zeros_like(depth), np.zeros_like(w), np.zeros((F_, 4))
        stats = np.zeros(8, dtype=np.int64)
        if tiled:
            (emu.emu_procrustes_bwd_tiled64 if tall else emu.emu_procrustes_bwd_tiled)(_p(depth), _p(k4), _p(fb), _p(w), _p(state), _p(g_rt), _p(gd), _p(gw),
                                         _p(k4acc), _p(stats), 1, F_, H, W)
        else:
            emu.emu_procrustes_bwd(_p(depth), _p(k4), _p(fb), _p(w), None, 0, _p(state), _p(g_rt), _p(gd),
                                   _p(gw), _p(k4acc), 1, F_, H, W)
        res.append((gd, gw, k4acc, stats))
    return res


TILED_CASES = [
    dict(F_=3, H=72, W=96),                                  # iid +-6 px jitter (the bench's flows)
    dict(F_=3, H=40, W=64),                                  # partial last tile row
    dict(F_=3, H=24, W=32),                                  # image smaller than the window
    dict(F_=3, H=100, W=64),                                 # second half of the last 64-row tile partly outside
    dict(F_=2, H=136, W=96, shift=(0.0, 0.08)),              # vertical motion across several tile rows
    dict(F_=3, H=72, W=96, outliers=0.05),                   # far taps: float fallback
    dict(F_=3, H=72, W=96, wscale=5.0),                      # raw weights above 1
    dict(F_=3, H=72, W=96, shift=(0.3, -0.2)),               # large coherent motion: shifted window
    dict(F_=3, H=72, W=96, shift=(2.0, 0.0)),                # everything leaves the frame: border pile-up
    dict(F_=3, H=72, W=96, depth_scale=1e3),
    dict(F_=3, H=72, W=96, depth_scale=1e-3, grt_scale=1e-6),
    dict(F_=3, H=72, W=96, grt_scale=1e8),
]


@pytest.mark.parametrize("tall", [False, True], ids=["tile32x32", "tile32x64"])
@pytest.mark.parametrize("case", TILED_CASES, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()))
def test_tiled_scatter_matches_direct_scatter(emu, case, tall):
    """Serial twin of k_distribute_tiled / k_distribute_tiled64 vs the direct float scatter of
    k_distribute on the same inputs: identical aligned / weight / intrinsics parts, depth gradient
    equal up to the float32 rounding of the direct path's own cell sums."""
    (gd0, gw0, k0, _), (gd1, gw1, k1, st) = _tiled_case(emu, tall=tall, **case)
    assert np.array_equal(gw0, gw1)
    assert np.abs(k0 - k1).max() <= 1e-12 * max(1.0, np.abs(k0).max())
    assert rel_l2(gd1, gd0) <= 5e-7
    assert st[5] == 0                       # no touched cell outside the image
    assert st[2] == 0 or case.get("wscale", 1.0) > 1.0   # the bound holds: no range fallback
    taps = st[0] + st[1] + st[2]
    assert taps == 2 * (case["F_"] - 1) * case["H"] * case["W"]
    if not case.get("outliers"):
        assert st[1] <= 0.01 * taps          # windows follow the flow
    assert st[3] <= 0.01 * taps              # high-word adds are rare
    if case.get("wscale", 1.0) <= 1.0:
        # the bound of scatter_bound_consts holds (scaled values stay below 2^29) and is not
        # wastefully loose (the largest one uses at least 1/2000 of the range: >= 18 bits left)
        assert 500 <= st[6] <= 1_000_000
