"""The C-ABI library loads on a CPU-only box and exports every symbol the header declares."""
import ctypes
import re

from conftest import ROOT


def declared_symbols():
    text = (ROOT / "include" / "flowmap_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fm_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    from flowmap_b200.build import build
    from flowmap_b200 import _lib
    so = build()
    handle = ctypes.CDLL(str(so))
    names = declared_symbols()
    assert len(names) >= 10
    for name in names:
        assert hasattr(handle, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(names)
    assert _lib.lib().fm_version() >= 100
    # pure host-side query works without a GPU
    assert _lib.lib().fm_workspace_bytes(1, 150, 360, 640) > 0
    assert _lib.lib().fm_workspace_bytes(1, 1, 8, 8) == 0


def test_product_has_no_cpu_path():
    import pytest
    import torch
    from flowmap_b200 import ops
    d = torch.zeros(1, 2, 4, 4)
    with pytest.raises(ValueError, match="CUDA"):
        ops.procrustes_poses(d, None, torch.zeros(1, 2, 4), torch.zeros(1, 1, 4, 4, 2))


def test_package_does_not_import_oracle():
    import subprocess, sys
    code = ("import sys, flowmap_b200, flowmap_b200.model, flowmap_b200.loss, flowmap_b200.overfit;"
            "bad=[m for m in sys.modules if m.startswith('oracle')]; assert not bad, bad")
    subprocess.check_call([sys.executable, "-c", code], cwd=str(ROOT))


def test_install_patches_reference_registries():
    """flowmap_b200.install() against the reference checkout (build container only)."""
    import os, sys
    import pytest
    if not os.path.isdir("/root/reference/flowmap"):
        pytest.skip("reference checkout not present (GPU box)")
    code = (
        "import sys; sys.path.insert(0, '/root/reference'); sys.dont_write_bytecode = True\n"
        "import flowmap_b200, flowmap.model.model as rm, flowmap.loss as rl\n"
        "rep = flowmap_b200.install()\n"
        "from flowmap_b200.model import Model\nfrom flowmap_b200.loss import LossFlow, LossTracking\n"
        "assert rm.Model is Model and rl.LOSSES['flow'] is LossFlow and rl.LOSSES['tracking'] is LossTracking\n"
        "from flowmap.model.model import ModelCfg\n"
        "from flowmap.model.backbone.backbone_explicit_depth import BackboneExplicitDepthCfg\n"
        "from flowmap.model.intrinsics.intrinsics_softmin import IntrinsicsSoftminCfg, RegressionCfg\n"
        "from flowmap.model.extrinsics.extrinsics_procrustes import ExtrinsicsProcrustesCfg\n"
        "cfg = ModelCfg(BackboneExplicitDepthCfg('explicit_depth', 0.1, 100.0), IntrinsicsSoftminCfg('softmin', 8192, 0.5, 2.0, 60, RegressionCfg(1000, 100)), ExtrinsicsProcrustesCfg('procrustes', None, False), True)\n"
        "m = rm.Model(cfg, 4, (8, 12))\n"
        "names = sorted(n for n, _ in m.named_parameters())\n"
        "assert names == ['backbone.depth', 'backbone.weights', 'intrinsics.intrinsics_regressed.focal_length'], names\n"
        "from flowmap.loss import get_losses\nfrom flowmap.loss.loss_flow import LossFlowCfg\nfrom flowmap.loss.mapping.mapping_huber import MappingHuberCfg\n"
        "l = get_losses([LossFlowCfg(0, 1000.0, 'flow', MappingHuberCfg('huber', 0.01))])\n"
        "assert type(l[0]) is LossFlow\n"
        "from flowmap.flow.flow_predictor import FlowPredictor\nfrom flowmap_b200 import flow as fl\n"
        "assert FlowPredictor.rescale_flow is fl.rescale_flow and FlowPredictor.compute_consistency_mask is fl.compute_consistency_mask\n"
        "assert 'flowmap.flow.flow_predictor.FlowPredictor.rescale_mask' in rep\n")
    import subprocess
    from conftest import ROOT
    subprocess.check_call([sys.executable, "-c", code], cwd=str(ROOT))
