"""world_size-2 (and 3) gloo tests of the pair-sharding host logic (flowmap_b200.parallel):
shard plan, the single per-step all-reduce and the boundary-frame halo.  The per-shard
compute is the oracle (CPU); the CUDA kernels are covered by the -m gpu tests."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def test_shard_pairs():
    from flowmap_b200.parallel import shard_pairs
    assert shard_pairs(149, 1) == [(0, 149)]
    s = shard_pairs(149, 8)
    assert s[0][0] == 0 and s[-1][1] == 149
    assert all(a[1] == b[0] for a, b in zip(s, s[1:]))
    assert max(b - a for a, b in s) - min(b - a for a, b in s) <= 1
    with pytest.raises(ValueError):
        shard_pairs(3, 4)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, f, h, w, out_dir):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from flowmap_b200 import parallel
        from oracle import flowmap_oracle as O
        torch.set_num_threads(2)
        dt = torch.float64
        flows = O.synthetic_flows(f, h, w, seed=3, dtype=dt)
        gen = torch.Generator().manual_seed(4)
        depth = 1.0 + torch.rand(f, h, w, generator=gen, dtype=dt)
        wparam = 0.01 * torch.randn(f - 1, h, w, generator=gen, dtype=dt)
        plan = parallel.make_plan(f - 1)
        d_l, w_l, fl_l = parallel.shard_inputs(plan, depth, wparam, flows)
        st = O.OverfitOracle(O.OverfitConfig(intrinsics="regressed"), plan.num_local_frames, h, w, dtype=dt)
        with torch.no_grad():
            st.depth.copy_(d_l)
            st.weights.copy_(w_l)
        out = st.forward(fl_l, 0)
        # local loss normalised by the GLOBAL mask sum (what LossFlow does with set_global_mask_sum)
        local_den = fl_l.forward_mask.sum() + fl_l.backward_mask.sum()
        den = parallel.global_mask_sum(local_den.reshape(()).clone())
        loss = 1000.0 * O.flow_loss(out.surfaces, out.extrinsics, out.intrinsics, fl_l) * local_den / den
        loss.backward()
        red = parallel.StepReducer(plan, (h, w), "cpu", 2)
        g_depth = st.depth.grad.float().clone()
        scal = red.reduce(torch.stack((loss.detach().float(), st.focal.grad.float())), g_depth)
        torch.save({"range": plan.pair_range, "g_depth": g_depth, "g_w": st.weights.grad.float(),
                    "scalars": scal, "bytes": red.bytes_per_step()}, f"{out_dir}/r{rank}.pt")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_step_equals_unsharded(tmp_path, world):
    from oracle import flowmap_oracle as O
    f, h, w = 7, 12, 16
    mp.spawn(_worker, args=(world, _free_port(), f, h, w, str(tmp_path)), nprocs=world, join=True)
    dt = torch.float64
    flows = O.synthetic_flows(f, h, w, seed=3, dtype=dt)
    gen = torch.Generator().manual_seed(4)
    depth = 1.0 + torch.rand(f, h, w, generator=gen, dtype=dt)
    wparam = 0.01 * torch.randn(f - 1, h, w, generator=gen, dtype=dt)
    st = O.OverfitOracle(O.OverfitConfig(intrinsics="regressed"), f, h, w, dtype=dt)
    with torch.no_grad():
        st.depth.copy_(depth)
        st.weights.copy_(wparam)
    ref = st.training_step(flows)
    parts = [torch.load(f"{tmp_path}/r{r}.pt") for r in range(world)]
    for p in parts:
        a, b = p["range"]
        assert abs(float(p["scalars"][0]) - ref["loss"]) <= 1e-5 * abs(ref["loss"])
        assert abs(float(p["scalars"][1]) - float(ref["grads"]["focal"])) <= 1e-4 * abs(float(ref["grads"]["focal"]))
        gd = ref["grads"]["depth"][a:b + 1].float()
        assert float((p["g_depth"] - gd).norm() / gd.norm()) <= 1e-5
        gw = ref["grads"]["weights"][a:b].float()
        assert float((p["g_w"] - gw).norm() / gw.norm()) <= 1e-5
        assert p["bytes"] == 4 * (2 + (world - 1) * h * w)
    # replicas of a boundary frame hold bit-identical gradients after the reduce
    for left, right in zip(parts, parts[1:]):
        assert torch.equal(left["g_depth"][-1], right["g_depth"][0])
