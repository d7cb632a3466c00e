"""N > 1 on real GPUs: world_size-2 NCCL run of the scene-sharded train step (minsu3d_b200.dp + harness.train).

Averaged gradients after `GradBucketer.finish()` must equal the mean of the two ranks' single-process gradients, in
both bucket modes (launched from backward / after backward), and both ranks must end the step with identical
parameters.  Needs 2 GPUs (`gpurun --gpus 2`); skipped on a one-GPU box."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _make(rank_seed):
    from minsu3d_b200.harness import models, scenes
    cfg = models.Config.for_model("pointgroup", proposal_source="gt_noise", blocks=[1, 2, 3])
    batch = scenes.make_batch([40 + rank_seed], "cuda", n_points=15_000)
    return cfg, batch


def _worker(rank, world, port, out_dir, overlap):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from minsu3d_b200 import dp
    from minsu3d_b200.harness import train
    dp.init_from_env(backend="nccl")
    torch.cuda.set_device(rank)
    cfg, batch = _make(rank)
    tr = train.Trainer(cfg, torch.device("cuda", rank), bucket_mb=0.25, overlap_allreduce=overlap)
    assert len(tr.bucketer.buckets) >= 3
    tr.bucketer.zero_grad()
    total, _, _ = tr.model.training_loss(batch)
    total.backward()
    tr.bucketer.finish()
    grads = {n: (p.grad.detach().cpu().clone() if p.grad is not None else None) for n, p in tr.model.named_parameters()}
    tr.optimizer.step()
    torch.cuda.synchronize()
    params = {n: p.detach().cpu().clone() for n, p in tr.model.named_parameters()}
    torch.save({"grads": grads, "params": params, "launched": tr.bucketer.launched_in_backward},
               os.path.join(out_dir, "rank%d.pt" % rank))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize("overlap", [True, False])
def test_nccl_world2_gradients_equal_mean_of_ranks(tmp_path, overlap):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world, port = 2, _free_port()
    mp.start_processes(_worker, args=(world, port, str(tmp_path), overlap), nprocs=world, join=True, start_method="spawn")
    res = [torch.load(os.path.join(tmp_path, "rank%d.pt" % r)) for r in range(world)]
    # reference: the same two steps in one process, gradients averaged on the host
    from minsu3d_b200.harness import train
    want = None
    for rank in range(world):
        cfg, batch = _make(rank)
        tr = train.Trainer(cfg, torch.device("cuda", 0))
        tr.bucketer.zero_grad()
        total, _, _ = tr.model.training_loss(batch)
        total.backward()
        g = {n: (p.grad.detach().cpu() / world if p.grad is not None else None) for n, p in tr.model.named_parameters()}
        want = g if want is None else {n: (want[n] if g[n] is None else g[n] if want[n] is None else want[n] + g[n])
                                       for n in g}
    for rank in range(world):
        for n, g in res[rank]["grads"].items():
            if want[n] is None:
                assert g is None, n
                continue
            err = float((g - want[n]).norm() / want[n].norm().clamp_min(1e-12))
            assert err < 1e-4, "rank %d %s: rel l2 %.3e" % (rank, n, err)  # weight-gradient atomics: ~1e-6
    for n in res[0]["params"]:
        assert torch.equal(res[0]["params"][n], res[1]["params"][n]), n  # same reduced gradients -> same update
    if overlap:
        assert res[0]["launched"] >= 1  # at least one bucket went out from inside backward
    else:
        assert res[0]["launched"] == 0
