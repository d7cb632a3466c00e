"""N > 1 host logic on CPU: world_size-2 gloo run of the bucketed gradient all-reduce (minsu3d_b200.dp).

Checks: (1) averaged gradients equal the mean of the per-rank gradients, (2) buckets are launched in the
same order on every rank even when one rank leaves parameters unused (no proposals -> ScoreNet unused),
(3) scene sharding covers every scene exactly once.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from minsu3d_b200 import dp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


class _Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Linear(8, 16)
        self.b = torch.nn.Linear(16, 16)
        self.unused_on_rank1 = torch.nn.Linear(16, 4)
        self.c = torch.nn.Linear(16, 2)

    def forward(self, x, use_extra):
        h = torch.relu(self.b(torch.relu(self.a(x))))
        out = self.c(h).sum()
        if use_extra:
            out = out + self.unused_on_rank1(h).sum()
        return out


def _worker(rank, world, port, result_dir, overlap=True):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = dp.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)
    net = _Net()
    bucketer = dp.GradBucketer(net.parameters(), bucket_mb=0.0005, overlap=overlap)  # tiny buckets -> several collectives
    assert len(bucketer.buckets) >= 3
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.randn(5, 8, generator=g)
    for step in range(2):
        bucketer.zero_grad()
        net(x, use_extra=(rank == 0)).backward()
        bucketer.finish()
    grads = {n: p.grad.clone() for n, p in net.named_parameters()}
    torch.save({"grads": grads, "launched_in_backward": bucketer.launched_in_backward, "x": x},
               os.path.join(result_dir, "rank%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("overlap", [True, False])
def test_bucketed_allreduce_world2_gloo(tmp_path, overlap):
    world, port = 2, _free_port()
    mp.start_processes(_worker, args=(world, port, str(tmp_path), overlap), nprocs=world, join=True,
                       start_method="spawn")
    res = [torch.load(os.path.join(tmp_path, "rank%d.pt" % r)) for r in range(world)]
    # reference: single-process gradients per rank, averaged
    torch.manual_seed(0)
    net = _Net()
    want = {n: torch.zeros_like(p) for n, p in net.named_parameters()}
    for rank in range(world):
        net.zero_grad()
        net(res[rank]["x"], use_extra=(rank == 0)).backward()
        for n, p in net.named_parameters():
            if p.grad is not None:
                want[n] += p.grad / world
    for rank in range(world):
        for n, g in res[rank]["grads"].items():
            assert torch.allclose(g, want[n], atol=1e-6), (rank, n)
    assert torch.equal(res[0]["grads"]["a.weight"], res[1]["grads"]["a.weight"])
    # rank 1 never produced a gradient for the extra head: its bucket had to be launched from finish()
    assert res[1]["launched_in_backward"] <= res[0]["launched_in_backward"]
    if not overlap:
        assert res[0]["launched_in_backward"] == 0  # pack-after-backward mode: collectives only in finish()


def _adam_run(net, bucketer, x, pattern):
    opt = torch.optim.Adam(net.parameters(), lr=1e-2)
    unused = []
    for use_extra in pattern:
        bucketer.zero_grad()
        net(x, use_extra=use_extra).backward()
        bucketer.finish()
        unused.append(sum(p.grad is None for p in net.parameters()))
        opt.step()
    return {n: p.detach().clone() for n, p in net.named_parameters()}, unused


def _worker_unused(rank, world, port, result_dir, overlap):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dp.init_from_env(backend="gloo")
    torch.manual_seed(0)
    net = _Net()
    bucketer = dp.GradBucketer(net.parameters(), bucket_mb=0.0005, overlap=overlap)
    x = torch.randn(5, 8, generator=torch.Generator().manual_seed(7))  # same batch on both ranks: mean grad = grad
    params, unused = _adam_run(net, bucketer, x, [True, False, False, True])
    torch.save({"params": params, "unused": unused}, os.path.join(result_dir, "rank%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("overlap", [True, False])
def test_globally_unused_parameters_are_skipped_like_world1(tmp_path, overlap):
    """A head that gets no gradient on ANY rank (no proposals anywhere) must keep p.grad = None so that Adam skips
    it -- no step-count advance, no stale-momentum update -- exactly as at world size 1 and as the reference's
    DDP(find_unused_parameters=True) does."""
    world, port = 2, _free_port()
    mp.start_processes(_worker_unused, args=(world, port, str(tmp_path), overlap), nprocs=world, join=True,
                       start_method="spawn")
    res = [torch.load(os.path.join(tmp_path, "rank%d.pt" % r)) for r in range(world)]
    torch.manual_seed(0)
    net = _Net()
    x = torch.randn(5, 8, generator=torch.Generator().manual_seed(7))
    want, want_unused = _adam_run(net, dp.GradBucketer(net.parameters()), x, [True, False, False, True])
    assert want_unused == [0, 2, 2, 0]
    for r in range(world):
        assert res[r]["unused"] == want_unused
        for n, v in res[r]["params"].items():
            assert torch.allclose(v, want[n], atol=1e-6), (r, n)


def test_single_process_bucketer_is_a_noop_wrapper():
    net = _Net()
    b = dp.GradBucketer(net.parameters())
    b.zero_grad()
    net(torch.randn(3, 8), True).backward()
    b.finish()
    assert all(p.grad is not None and p.grad.abs().sum() > 0 for p in net.parameters())
    assert b.grad_bytes() == sum(p.numel() for p in net.parameters()) * 4


def test_scene_sharding_partitions_the_dataset():
    for world in (1, 2, 4, 8):
        seen = sorted(i for r in range(world) for i in dp.shard_indices(312, r, world))  # 312 val scenes
        assert seen == list(range(312))
