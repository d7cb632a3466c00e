"""Scene-sharded data parallelism: one process per GPU, bucketed gradient all-reduce over NCCL.

The reference's only parallelism is Lightning DDP by scene with `find_unused_parameters=True`
(config/model/base.yaml:12-20).  The hot path itself has no exchange step
(every op is independent across the batch index), so the only collective is the per-step gradient
all-reduce: 7.7 M (m=16) / 30.8 M (m=32) fp32 parameters.

Design: see GradBucketer.  The ~31 MB (m=16) all-reduce takes ~0.2 ms over NVLink/NVSwitch; what matters is where
it sits: launched bucket by bucket from backward it hides behind the remaining backward kernels, and the host cost
per parameter is one counter increment (no accumulate-into-view kernels, no per-parameter collectives).
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """torchrun-style rendezvous (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*). Returns (rank, world, local)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


HOOK_EVERY = 8


class GradBucketer:
    """Flat fp32 gradient buckets in reverse registration order (~ the order backward produces gradients).

    overlap=True (default): `p.grad` stays None, so autograd hands its gradient tensors over without an
    accumulate-into-view kernel per parameter; post-accumulate hooks on a subset of the parameters check for complete
    buckets.  When every gradient
    of a bucket has arrived (and every earlier bucket has been launched -- the collective order must be the same on
    all ranks) the bucket is packed with ONE multi-tensor copy and its all-reduce is launched asynchronously: NCCL
    runs it on its own stream while the rest of backward keeps the compute stream busy.  `finish()` launches what is
    left (buckets holding parameters that got no gradient on this rank), waits, averages and points `p.grad` at the
    reduced views.  overlap=False: everything in `finish()` (pack -> all-reduce -> scale).

    Parameters that received no gradient on ANY rank keep `p.grad = None` after `finish()` -- the optimizer then
    skips them exactly as it does at world size 1 and as the reference's DDP(find_unused_parameters=True) does
    (config/model/base.yaml:12-20): e.g. the ScoreNet in a step where no rank found a proposal.  The global OR of
    the per-parameter flags rides on the data group as one more (844-byte) all-reduce behind the buckets; only a rank
    that itself has an unused parameter reads it back -- the common step has no host-side collective and no device
    read (round 2 used a gloo all-reduce per step here: a host-level rendezvous of all ranks in every step)."""

    def __init__(self, params, bucket_mb=8.0, group=None, overlap=True):
        self.group = group
        self.overlap = overlap
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        params = [p for p in params if p.requires_grad]
        self.params = list(reversed(params))
        cap = int(bucket_mb * 1024 * 1024) // 4
        self.buckets = []
        cur, cur_n = [], 0
        for p in self.params:
            if cur and cur_n + p.numel() > cap:
                self.buckets.append(cur)
                cur, cur_n = [], 0
            cur.append(p)
            cur_n += p.numel()
        if cur:
            self.buckets.append(cur)
        self.flats, self.views, self.bucket_of = [], [], {}
        self._handles, self._hooks = [], []
        self._ready, self._next = [], 0
        self.launched_in_backward = 0
        self.last_unused = 0  # parameters left without a gradient by the last finish() (globally unused)
        for p in self.params:
            p.grad = None
        if self.world == 1:
            return  # nothing to exchange
        for bi, plist in enumerate(self.buckets):
            flat = torch.zeros(sum(p.numel() for p in plist), dtype=plist[0].dtype, device=plist[0].device)
            off = 0
            views = []
            for p in plist:
                views.append(flat[off:off + p.numel()].view_as(p))
                off += p.numel()
                self.bucket_of[p] = bi
            self.flats.append(flat)
            self.views.append(views)
        self._ready = [0] * len(self.buckets)
        self._index = {p: i for i, p in enumerate(self.params)}
        self.flags = torch.zeros(len(self.params), dtype=torch.float32, device=self.params[0].device)
        if overlap:
            # a hook on every 8th parameter and on the last one of every bucket: each firing checks whether the next
            # bucket in line is complete (a scan over its `p.grad`), which costs a few microseconds ~30 times per
            # backward instead of 211 Python hook dispatches; a bucket whose last gradient arrives at an un-hooked
            # parameter is launched at the next firing (or by finish())
            for plist in self.buckets:
                for j, p in enumerate(plist):
                    if j == len(plist) - 1 or j % HOOK_EVERY == HOOK_EVERY - 1:
                        self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    # -- per step -----------------------------------------------------------------------------
    def zero_grad(self):
        for p in self.params:
            p.grad = None
        self._handles = []
        if self.world > 1:
            self._ready = [0] * len(self.buckets)
            self._next = 0
            self.launched_in_backward = 0

    def _launch(self, bi):
        """Pack bucket bi (zeros where this rank has no gradient) and start its all-reduce."""
        flat, plist, views = self.flats[bi], self.buckets[bi], self.views[bi]
        dst = [v for p, v in zip(plist, views) if p.grad is not None]
        src = [p.grad for p in plist if p.grad is not None]
        if len(dst) < len(plist):
            flat.zero_()
        if dst:
            torch._foreach_copy_(dst, src)
        self._handles.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def _on_grad(self, p):
        while self._next < len(self.buckets) and all(q.grad is not None for q in self.buckets[self._next]):
            self._launch(self._next)
            self._next += 1
            self.launched_in_backward += 1

    def _launch_used_flags(self, used_local):
        """Start the SUM all-reduce of the per-parameter 'received a gradient' flags on the data group (same stream
        and ordering as the gradient buckets; 4 bytes per parameter).  The result is only READ by ranks that have a
        locally unused parameter: a rank whose parameters all received gradients already knows the global OR."""
        host = torch.tensor([1.0 if u else 0.0 for u in used_local], dtype=torch.float32,
                            pin_memory=self.flags.is_cuda)
        self.flags.copy_(host, non_blocking=True)
        self._handles.append(dist.all_reduce(self.flags, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self):
        """Launch the remaining buckets in order, wait for all of them, average, expose the reduced gradients."""
        if self.world == 1:
            return
        used_local = [p.grad is not None for p in self.params]
        while self._next < len(self.buckets):
            self._launch(self._next)
            self._next += 1
        self._launch_used_flags(used_local)
        for h in self._handles:
            h.wait()
        torch._foreach_mul_(self.flats, 1.0 / self.world)
        if all(used_local):
            used = used_local  # the global OR contains this rank's flags: no read, the host does not wait
        else:
            used = [f > 0.5 for f in self.flags.tolist()]  # rare (a rank without proposals): one small device read
        self.last_unused = used.count(False)
        for plist, views in zip(self.buckets, self.views):
            for p, v in zip(plist, views):
                p.grad = v if used[self._index[p]] else None

    def grad_bytes(self):
        return sum(p.numel() * p.element_size() for p in self.params)

    def remove_hooks(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []


def shard_indices(n_items, rank, world):
    """DistributedSampler-style interleaved shard: items rank, rank+world, ..."""
    return list(range(rank, n_items, world))
