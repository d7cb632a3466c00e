"""Scene-sharded data parallelism: one process per GPU, bucketed gradient all-reduce over NCCL.

The reference's only parallelism is Lightning DDP by scene with `find_unused_parameters=True`
(config/model/base.yaml:12-20).  The hot path itself has no exchange step
(every op is independent across the batch index), so the only collective is the per-step gradient
all-reduce: 7.7 M (m=16) / 30.8 M (m=32) fp32 parameters.

Design: parameters are packed once into flat fp32 buckets in reverse registration order (the
order gradients become ready in backward); `p.grad` are views into the bucket, so there is no
copy-in / copy-out.  A post-accumulate hook launches the bucket's asynchronous all-reduce as soon
as (a) all of its gradients are ready and (b) every earlier bucket has been launched, which keeps
the collective order identical on all ranks even when a rank has unused parameters (e.g. the
ScoreNet when a rank found no proposals).  NCCL runs the reduction on its own stream, overlapping
the rest of backward; `finish()` launches whatever is left, waits, and averages.

`overlap=False` ("pack after backward"): the ~30 MB all-reduce takes ~0.2 ms over NVLink/NVSwitch, far less than
the host cost of ~200 Python hook calls and ~200 accumulate-into-view kernels per step, so on NVLink machines the
trainer lets autograd hand over its gradient tensors untouched, packs them into the flat buckets with one
multi-tensor copy per bucket after backward, all-reduces the buckets and points `p.grad` at the reduced views.
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


def _host_group(group):
    """A process group that can reduce small CPU tensors: the data group itself when it is gloo, else a gloo side
    group over the same ranks (created once per process; every rank constructs its GradBucketer, so the collective
    `new_group` call is matched)."""
    if not dist.is_initialized():
        return None
    if dist.get_backend(group) == "gloo":
        return group
    key = id(group)
    g = _HOST_GROUPS.get(key)
    if g is None:
        ranks = dist.get_process_group_ranks(group) if group is not None else None
        g = _HOST_GROUPS[key] = dist.new_group(ranks=ranks, backend="gloo")
    return g


_HOST_GROUPS = {}


class GradBucketer:
    """Parameters that received no gradient on ANY rank keep `p.grad = None` after `finish()` -- the optimizer then
    skips them exactly as it does at world size 1 and as the reference's DDP(find_unused_parameters=True) does
    (config/model/base.yaml:12-20): e.g. the ScoreNet in a step where no rank found a proposal.  Which parameters
    are used is host knowledge on every rank (hooks fired / `p.grad is not None`), so the global OR is one tiny
    bit-mask all-reduce on the gloo side group: no device read, the host does not wait for the backward."""

    def __init__(self, params, bucket_mb=8.0, group=None, overlap=True):
        self.group = group
        self.overlap = overlap
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        params = [p for p in params if p.requires_grad]
        self.params = list(reversed(params))  # ~ order in which backward produces gradients
        cap = int(bucket_mb * 1024 * 1024) // 4
        self.buckets = []  # list of dict(flat, params, pending)
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
        if self.world == 1:
            # nothing to exchange: leave p.grad unset so that autograd hands its gradient tensors over without the
            # accumulate-into-view kernel per parameter that the flat buckets cost
            for p in self.params:
                p.grad = None
            return
        for bi, plist in enumerate(self.buckets):
            flat = torch.zeros(sum(p.numel() for p in plist), dtype=plist[0].dtype, device=plist[0].device)
            off = 0
            views = []
            for p in plist:
                views.append(flat[off:off + p.numel()].view_as(p))
                p.grad = views[-1] if overlap else None
                off += p.numel()
                self.bucket_of[p] = bi
            self.flats.append(flat)
            self.views.append(views)
        self._ready = [0] * len(self.buckets)
        self._next = 0
        self._handles = []
        self._hooks = []
        self._index = {p: i for i, p in enumerate(self.params)}
        self._used = [False] * len(self.params)  # overlap mode: set by the hooks
        self.host_group = _host_group(group)
        self.last_unused = 0  # parameters left without a gradient by the last finish() (globally unused)
        if self.world > 1 and overlap:
            for p in self.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
        self.launched_in_backward = 0

    # -- per step -----------------------------------------------------------------------------
    def zero_grad(self):
        """Keeps p.grad as views of the flat buckets (do not call optimizer.zero_grad(set_to_none=True))."""
        if self.world == 1 or not self.overlap:
            for p in self.params:
                p.grad = None
            self._handles = []
            return
        for flat in self.flats:
            flat.zero_()
        for plist, views in zip(self.buckets, self.views):
            for p, v in zip(plist, views):
                p.grad = v  # finish() may have set globally unused parameters to None
        self._ready = [0] * len(self.buckets)
        self._used = [False] * len(self.params)
        self._next = 0
        self._handles = []
        self.launched_in_backward = 0

    def _launch(self, bi):
        self._handles.append(dist.all_reduce(self.flats[bi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def _on_grad(self, p):
        bi = self.bucket_of[p]
        self._used[self._index[p]] = True
        self._ready[bi] += 1
        while self._next < len(self.buckets) and self._ready[self._next] == len(self.buckets[self._next]):
            self._launch(self._next)
            self._next += 1
            self.launched_in_backward += 1

    def _global_used(self, used_local):
        """OR over ranks of the per-parameter 'received a gradient' flags (host-side, 63 flags per int64 word)."""
        words = [0] * ((len(used_local) + 62) // 63)
        for i, u in enumerate(used_local):
            if u:
                words[i // 63] |= 1 << (i % 63)
        t = torch.tensor(words, dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.BOR, group=self.host_group)
        words = t.tolist()
        return [bool((words[i // 63] >> (i % 63)) & 1) for i in range(len(used_local))]

    def finish(self):
        """Launch the remaining buckets in order, wait for all of them, and average."""
        if self.world == 1:
            return
        if not self.overlap:
            used = self._global_used([p.grad is not None for p in self.params])
            self.last_unused = used.count(False)
            for flat, plist, views in zip(self.flats, self.buckets, self.views):
                dst = [v for p, v in zip(plist, views) if p.grad is not None]
                src = [p.grad for p in plist if p.grad is not None]
                if len(dst) < len(plist):
                    flat.zero_()  # parameters without a gradient on this rank contribute zeros
                if dst:
                    torch._foreach_copy_(dst, src)
                self._handles.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            for h in self._handles:
                h.wait()
            torch._foreach_mul_(self.flats, 1.0 / self.world)
            for plist, views in zip(self.buckets, self.views):
                for p, v in zip(plist, views):
                    p.grad = v if used[self._index[p]] else None
            return
        while self._next < len(self.buckets):
            self._launch(self._next)
            self._next += 1
        used = self._global_used(self._used)
        self.last_unused = used.count(False)
        for h in self._handles:
            h.wait()
        torch._foreach_mul_(self.flats, 1.0 / self.world)
        for p in self.params:
            if not used[self._index[p]]:
                p.grad = None  # restored to the bucket view by zero_grad()

    def grad_bytes(self):
        return sum(p.numel() * p.element_size() for p in self.params)

    def remove_hooks(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []


def shard_indices(n_items, rank, world):
    """DistributedSampler-style interleaved shard: items rank, rank+world, ..."""
    return list(range(rank, n_items, world))
