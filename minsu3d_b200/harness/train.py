"""One training step of PointGroup / HAIS / SoftGroup on the B200-native hot path.

Replaces the Lightning loop around `training_step` (minsu3d/model/general_model.py:52-66,
train.py:35-41) with a plain loop: forward (backbone + clustering + ScoreNet), losses, backward,
scene-sharded gradient all-reduce (minsu3d_b200.dp) and Adam (config/model/pointgroup.yaml:16-18).
"""
import torch

from .. import dp, ops
from . import models

HOST_KEYS = ("point_xyz", "vert_batch_ids", "sem_labels", "instance_ids", "instance_center_xyz",
             "instance_num_point", "instance_offsets", "instance_semantic_cls", "voxel_xyz", "voxel_features",
             "voxel_point_map")


def to_pinned_host(data):
    """Collated batch -> pinned host tensors (what a DataLoader with pin_memory=True hands over)."""
    out = {}
    for k, v in data.items():
        out[k] = v.detach().cpu().pin_memory() if torch.is_tensor(v) else v
    return out


def host_bytes(data):
    return sum(v.numel() * v.element_size() for v in data.values() if torch.is_tensor(v))


def reserve_device_pool(device, gigabytes):
    """Grow torch's caching allocator once, up front: the step's ~800 allocations per iteration are then carved from
    cached memory instead of occasionally reaching the driver (cuMemMap / cudaMalloc stalls of 10-100 ms when the
    voxel count of a batch exceeds everything seen so far)."""
    if device.type == "cuda" and gigabytes > 0:
        block = torch.empty(int(gigabytes * 2 ** 30), dtype=torch.uint8, device=device)
        # the allocator keeps a separate pool for requests under 1 MB (per-channel vectors, counters, small levels)
        small = [torch.empty(512 * 1024, dtype=torch.uint8, device=device) for _ in range(1024)]
        del block, small


class Trainer:
    def __init__(self, cfg: models.Config, device, bucket_mb=8.0, seed=123, reserve_gb=0.0, overlap_allreduce=True):
        torch.manual_seed(seed)  # config/config.yaml:17
        self.cfg = cfg
        self.device = torch.device(device)
        reserve_device_pool(self.device, reserve_gb)
        self.model = models.build_model(cfg).to(self.device)
        self.model.train()
        # buckets are packed and all-reduced from inside backward as they complete (dp.py)
        self.bucketer = dp.GradBucketer(self.model.parameters(), bucket_mb=bucket_mb, overlap=overlap_allreduce)
        # same update rule as the reference's torch.optim.Adam; the fused multi-tensor implementation keeps the
        # host cost of the ~200-parameter step at a few launches (the default foreach path costs ~5 ms of Python)
        self.optimizer = torch.optim.Adam(self.model.parameters(), lr=cfg.lr, fused=(self.device.type == "cuda"))
        self.last_losses = None
        self._copy_stream = None
        self._copy_event = None
        # tensor-core operand images of every convolution kernel: refreshed in one launch after each optimizer step
        # (the modules would otherwise re-pack lazily, one launch per layer per step)
        self.packed = None
        if self.device.type == "cuda":
            from ..MinkowskiEngine.modules import _ConvBase
            self.packed = ops.PackedSet([(m.kernel, m._packed) for m in self.model.modules() if isinstance(m, _ConvBase)])
            self.packed.repack()

    def step(self, data):
        """data already on the device.  Returns the total loss (device scalar)."""
        self.bucketer.zero_grad()
        total, losses, _ = self.model.training_loss(data)
        total.backward()
        self.bucketer.finish()
        # Size claims made in this forward (SparseTensor(coordinates_unique / level_sizes), ops.py) are validated on
        # the device BEFORE the weights change: the fused Adam skips the update when the flag is set (the GradScaler
        # `found_inf` mechanism), no host read; the ValueError is raised by the next ops.run_deferred_checks().
        if self.device.type == "cuda":
            self.optimizer.found_inf = ops.deferred_failure_flag(self.device)
        self.optimizer.step()
        if self.packed is not None:
            self.packed.repack()
        self.last_losses = losses
        return total.detach()

    def step_from_host(self, host_data):
        """The user-facing call: pinned host batch in, python float loss out (H2D + D2H inside)."""
        # the backbone needs only the voxel tensors: they are copied first on the compute stream; everything else
        # (per-point labels, instance targets, point coordinates: ~half of the bytes) follows on a copy stream while
        # the backbone forward is being enqueued and run; the model waits for it before its first use
        # (models.wait_late_inputs, right after the backbone)
        if self.device.type == "cuda":
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(device=self.device)
                self._copy_event = torch.cuda.Event()
            early = {k: v.to(self.device, non_blocking=True) for k, v in host_data.items()
                     if torch.is_tensor(v) and k in models.BACKBONE_INPUTS}
            main = torch.cuda.current_stream(self.device)
            self._copy_stream.wait_stream(main)  # (the previous step's readers of recycled buffers have been enqueued)
            with torch.cuda.stream(self._copy_stream):
                late = {k: v.to(self.device, non_blocking=True) for k, v in host_data.items()
                        if torch.is_tensor(v) and k not in models.BACKBONE_INPUTS}
                self._copy_event.record(self._copy_stream)
            for t in late.values():
                t.record_stream(main)
            data = {k: (early.get(k, late.get(k)) if torch.is_tensor(v) else v) for k, v in host_data.items()}
            data["_late_event"] = self._copy_event
        else:
            data = {k: (v.to(self.device, non_blocking=True) if torch.is_tensor(v) else v) for k, v in host_data.items()}
        loss = float(self.step(data).item())
        ops.run_deferred_checks()  # size claims of this step (loader-reported level sizes), stream already drained
        return loss
