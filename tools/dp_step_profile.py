"""Where an N-GPU PointGroup train step spends its time (run under torchrun): per rank, CUDA-event time of
forward+loss / backward / GradBucketer.finish() (= exposed all-reduce tail) / optimizer, the rank's step time and the
skew between ranks.  Rank 0 prints one table.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 \
        tools/dp_step_profile.py [--no-overlap]
"""
import os
import sys

os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from minsu3d_b200 import dp, ops  # noqa: E402
from minsu3d_b200.harness import models, scenes, train  # noqa: E402


def main():
    overlap = "--no-overlap" not in sys.argv
    rank, world, local = dp.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cfg = models.Config.for_model("pointgroup", proposal_source="gt_noise")
    tr = train.Trainer(cfg, dev, reserve_gb=16.0, overlap_allreduce=overlap)
    pool = []
    for i in range(3):
        d = scenes.make_batch([(rank * 3 + i) * 4 + s for s in range(4)], dev, 100_000)
        d.pop("voxel_level_sizes", None)
        pool.append(d)
    for i in range(6):
        tr.step(pool[i % 3])
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    names = ("forward+loss", "backward", "finish (all-reduce tail)", "optimizer+repack", "step")
    acc = [0.0] * len(names)
    steps = 12
    launched = 0
    for i in range(steps):
        data = pool[i % 3]
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        ev[0].record()
        tr.bucketer.zero_grad()
        total, losses, _ = tr.model.training_loss(data)
        ev[1].record()
        total.backward()
        ev[2].record()
        tr.bucketer.finish()
        ev[3].record()
        tr.optimizer.found_inf = ops.deferred_failure_flag(dev)
        tr.optimizer.step()
        if tr.packed is not None:
            tr.packed.repack()
        ev[4].record()
        torch.cuda.synchronize()
        launched += tr.bucketer.launched_in_backward
        for j in range(4):
            acc[j] += ev[j].elapsed_time(ev[j + 1])
        acc[4] += ev[0].elapsed_time(ev[4])
    mine = torch.tensor([a / steps for a in acc] + [launched / steps], device=dev)
    allr = [torch.zeros_like(mine) for _ in range(world)] if world > 1 else [mine]
    if world > 1:
        dist.all_gather(allr, mine)
    if rank == 0:
        print("N = %d GPUs, overlap = %s, %d buckets of <= 8 MB, %.1f MB of gradients, ms per step (CUDA events), mean of %d steps"
              % (world, overlap, len(tr.bucketer.buckets), tr.bucketer.grad_bytes() / 1e6, steps))
        print("%-5s" % "rank" + "".join("%26s" % n for n in names) + "%22s" % "buckets sent in bwd")
        for r, t in enumerate(allr):
            t = t.tolist()
            print("%-5d" % r + "".join("%26.2f" % v for v in t[:5]) + "%22.1f" % t[5])
        st = torch.stack(allr)[:, 4]
        print("step time: min %.2f  max %.2f  (skew %.2f ms); the bench reports the max over ranks" % (
            float(st.min()), float(st.max()), float(st.max() - st.min())))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
