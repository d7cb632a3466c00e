"""BASELINE.json configs[2] and [3] on ONE GPU: HAIS train step (m=32, hierarchical aggregation) and SoftGroup
inference (m=32, soft grouping + top-down refinement) on synthetic 100k-point scenes.  Scene-sharded multi-GPU
runs use the same code under torchrun (minsu3d_b200.dp.shard_indices); here the per-GPU rate is measured.

    python tools/bench_models.py > profiles/rNN_models.json
"""
import json
import os
import sys
import time

os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")
import torch  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from minsu3d_b200.harness import models, scenes, train  # noqa: E402


def timed(fn, n, warm):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(warm + i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3


def main():
    dev = torch.device("cuda", 0)
    out = {}
    # ---- config 2: HAIS train step, batch 8 scenes per GPU -------------------------------------------
    batch = 8
    pool = [scenes.make_batch([b * batch + s for s in range(batch)], dev, 100_000) for b in range(2)]
    tr = train.Trainer(models.Config.for_model("hais", proposal_source="gt_noise"), dev, reserve_gb=24.0)
    steps = 4
    t = timed(lambda i: tr.step(pool[i % 2]), steps, 4)
    out["hais_train"] = {"scenes_per_s": batch * steps / t, "ms_per_step": t / steps * 1e3, "batch_per_gpu": batch,
                         "losses": sorted(tr.last_losses), "m": 32}
    del tr
    torch.cuda.empty_cache()
    # ---- config 3: SoftGroup inference, scenes of 50k-250k points, batch 1 (test.py semantics) ---------
    model = models.build_model(models.Config.for_model("softgroup", proposal_source="gt_noise")).to(dev).eval()
    sizes = [50_000, 100_000, 150_000, 200_000, 250_000, 120_000, 80_000, 180_000]
    val = [scenes.collate([scenes.make_scene(500 + i, n)], dev) for i, n in enumerate(sizes)]

    def infer(i):
        with torch.no_grad():
            model(val[i % len(val)])

    n = 16
    t = timed(infer, n, 4)
    out["softgroup_inference"] = {"scenes_per_s": n / t, "ms_per_scene": t / n * 1e3, "m": 32,
                                  "points_per_scene": "50k-250k", "eta_312_scenes_1gpu_s": 312 * t / n}
    out["gpu"] = torch.cuda.get_device_name(0)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
