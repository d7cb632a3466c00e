#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native sparse-voxel hot path.

Metric (BASELINE.json): PointGroup train scenes/s, synthetic 100k-point scenes at 2 cm voxels,
configs[1]: full train step (backbone + ballquery_batch_p / bfs_cluster + ScoreNet), batch 4 per GPU.

    python bench.py --gpus N --steps K --warmup W            # own arm (one rank per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # CPU arm: the same train step restated on the host
    python bench.py --workload hais_train|softgroup_infer    # BASELINE.json configs[2] / [3] (same launch contract)

Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os

os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")  # no cudaFree/cudaMalloc stalls when batch sizes vary
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "pointgroup_train_scenes_per_s"
UNIT = "scenes/s"
SCENES_PER_GPU = 4
POINTS_PER_SCENE = 100_000


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), "measured"
    return 6650.0, 1590.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.samples[0][1]),
                "reasons": reasons, "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------
# CPU arm: the same PointGroup train step restated on the host (oracle/cpu_step.py), all host threads
# --------------------------------------------------------------------------------------------
CPU_SAMPLE = ("ONE synthetic 100k-point scene per step (a quarter of the GPU arm's 4-scene batch) through the full "
              "PointGroup train step on the host: restated MinkowskiEngine CPU-backend backbone + ScoreNet forward AND "
              "backward (oracle/, OpenMP; MinkowskiEngine itself is an un-vendored dependency and not installable), "
              "brute-force ball query x2 (oracle port of bfs_cluster.cu), the REFERENCE'S OWN pg_bfs_cluster (oracle/_ref, "
              "1 thread) x2, clusters_voxelization, losses, Adam")


def cpu_train_step_scenes_per_s(n_steps, warmup):
    """-> (scenes/s, seconds per step, per-stage seconds of the last step, used the reference's BFS binary?)."""
    import oracle
    from oracle import cpu_step
    from minsu3d_b200.harness import models, scenes
    oracle.build()
    torch.manual_seed(123)
    cfg = models.Config.for_model("pointgroup", proposal_source="gt_noise")
    stepper = cpu_step.CpuPointGroupStep(models.build_model(cfg), lr=cfg.lr)
    small = scenes.collate([scenes.make_scene(99, 30_000)], "cpu")
    for _ in range(max(0, min(warmup, 2))):  # warm-up on a 30k-point scene: libraries loaded, threads started
        stepper.step(small)
    batches = [scenes.collate([scenes.make_scene(s, POINTS_PER_SCENE)], "cpu") for s in range(min(n_steps, 3))]
    t0 = time.perf_counter()
    for i in range(n_steps):
        stepper.step(batches[i % len(batches)])
    dt = time.perf_counter() - t0
    return n_steps / dt, dt / n_steps, dict(stepper.timing), stepper.ref_ops is not None


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    steps = max(1, args.steps)
    sps, sec, stages, used_ref = cpu_train_step_scenes_per_s(steps, args.warmup)
    sample = "%d steps of: %s; BFS from the reference binary: %s" % (steps, CPU_SAMPLE, used_ref)
    line = {"impl": "reference", "metric": METRIC, "value": sps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: PointGroup full train step (backbone + 2x ballquery_batch_p/pg_bfs_cluster + "
                                   "ScoreNet + losses + Adam) on the host cores, bounded sample: 1 scene x 100k points per step",
                       "stage_seconds": {k: round(v, 3) for k, v in stages.items()}},
            "cpu_baseline": {"value": sps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": sps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# own arm: per-kernel roofline table, measured live (CUDA events on the launch stream, L2 flushed)
# --------------------------------------------------------------------------------------------
class _Timer:
    def __init__(self, device):
        self.flush = torch.empty(256 * 1024 * 1024 // 4, device=device)  # 256 MB > 126 MB of L2

    def __call__(self, fn, reps=5, warm=2):
        for _ in range(warm):
            fn()
        times = []
        for _ in range(reps):
            self.flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1) * 1e-3)
        return float(np.mean(times))


def _traffic(kernel_key):
    """DRAM bytes per launch of the same kernel from the committed `ncu` capture (profiles/r02_kernel_traffic.json)."""
    p = os.path.join(ROOT, "profiles", "r02_kernel_traffic.json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        return json.load(f).get(kernel_key)


def kernel_rooflines(data, device, cfg):
    """One entry per kernel class of SURVEY.md 8(d) on the tensors of the benchmark batch.  `achieved` = ALGORITHMIC
    bytes (the 8(d) formulas: every distinct byte once) / measured time; `peak` = measured HBM copy bandwidth."""
    from minsu3d_b200 import ops
    from minsu3d_b200.common_ops.functions import common_ops
    from minsu3d_b200.harness import models
    hbm, bf16, which = _peaks()
    timer = _Timer(device)
    rows = []

    def add(name, alg_bytes, t, **extra):
        e = {"kernel": name, "bound": "hbm", "achieved": alg_bytes / t / 1e9, "peak": hbm, "unit": "GB/s",
             "frac": alg_bytes / t / 1e9 / hbm, "us": t * 1e6, "algorithmic_bytes": int(alg_bytes),
             "traffic": _traffic(name.split(" ")[0])}
        e.update(extra)
        rows.append(e)
        return e

    coords = data["voxel_xyz"]
    m = coords.size(0)
    n_pts = data["point_xyz"].size(0)
    # T1 / V1: coordinate insert + unique (bytes: 16 N + 16 M read/written + 8 N + 8 M maps)
    t = timer(lambda: ops.coord_unique_async(coords, 1))
    add("coord_unique (T1/V1: hash insert + first-occurrence unique, %d rows)" % m, 16 * m + 16 * m + 8 * m + 8 * m, t)
    table, _, _, oc = ops.coord_unique(coords, 1)
    # T2: kernel map 3^3 (bytes: 16 M_in + 8 P)
    t = timer(lambda: ops.kernel_map(oc, table, 3, 1, with_tile_mask=True))
    nbr, tile_mask = ops.kernel_map(oc, table, 3, 1, with_tile_mask=True)
    pairs = int((nbr >= 0).sum().item())
    add("kernel_map (T2: 27 probes per row, %d rows, %d pairs)" % (m, pairs), 16 * m + 8 * pairs, t)
    perm, nbr_sorted, tms = ops.tile_order(nbr)
    # T3: convolution, level-0 16 -> 16 (dominant kernel) -- persistent tcgen05 kernel on mask-sorted tiles
    x16 = torch.randn(m, 16, device=device)
    w16 = torch.randn(27, 16, 16, device=device) * 0.05
    p16 = ops.conv_pack(w16)
    conv_bytes = lambda mm, pp, ci, co: 4 * (mm * ci + mm * co) + 4 * pp + 4 * 27 * ci * co
    t_conv = timer(lambda: ops.conv_table(x16, w16, nbr_sorted, m, 27, 16, 16, algo=ops.ALGO_TC_3XTF32, tile_mask=tms,
                                          out_rows=perm, packed=p16), reps=10, warm=3)
    t_fma = timer(lambda: ops.conv_table(x16, w16, nbr, m, 27, 16, 16, algo=ops.ALGO_SIMT))
    dominant = add("conv_tcp_kernel<3> (T3: 3^3 conv 16->16 on the level-0 map, tcgen05 3xTF32, persistent, mask-sorted "
                   "tiles, weights packed once per optimizer step)", conv_bytes(m, pairs, 16, 16), t_conv,
                   rows=m, pairs=pairs, useful_tflops=2.0 * pairs * 256 / t_conv / 1e12, fp32_fma_path_us=t_fma * 1e6)
    # T3 launch-weighted over the U-Net levels of this batch (one 3^3 c->c convolution per level, its real map)
    from minsu3d_b200.MinkowskiEngine.sparse_tensor import CoordinateManager
    mgr = CoordinateManager(D=3, device=device)
    key, _, _ = mgr.insert_and_map(coords, 1, assume_unique=True)
    tot_b, tot_t, per_level = 0.0, 0.0, []
    for lvl in range(7):
        c = cfg.m * (lvl + 1)
        km = mgr.kernel_map(key, key, 3)
        nb, tmk, orows = km.table_for(c, c)
        ml = km.n_out
        pl = int((km.nbr >= 0).sum().item())
        xl = torch.randn(ml, c, device=device)
        wl = torch.randn(27, c, c, device=device) * 0.05
        pk = ops.conv_pack(wl)
        tl = timer(lambda: ops.conv_table(xl, wl, nb, ml, 27, c, c, tile_mask=tmk, out_rows=orows, packed=pk))
        tot_b += conv_bytes(ml, pl, c, c)
        tot_t += tl
        per_level.append({"level": lvl, "rows": ml, "channels": c, "us": tl * 1e6,
                          "gbs": conv_bytes(ml, pl, c, c) / tl / 1e9})
        if lvl < 6:
            key = mgr.stride_key(key, 2)
    add("conv_tcp_kernel<3> (T3: launch-weighted over the 7 U-Net levels, one 3^3 c->c conv each)", tot_b, tot_t,
        per_level=per_level)
    # T3 weight gradient, level 0
    pin, pout, koff, _ = ops.pairs_from_nbr(nbr)
    g16 = torch.randn(m, 16, device=device)
    t = timer(lambda: ops.conv_wgrad(x16, g16, pin, pout, koff, 27, 16, 16, m * 27))
    add("conv_wgrad (T3: gW[k] = in[I_k]^T gout[O_k], 16x16, level-0 map)", 4 * (2 * m * 16) + 8 * pairs + 4 * 27 * 256, t,
        useful_tflops=2.0 * pairs * 256 / t / 1e12)
    # T5: BatchNorm + ReLU forward (stats + apply) and backward, C = 16
    gam, bet = torch.ones(16, device=device), torch.zeros(16, device=device)
    t = timer(lambda: ops.bn_forward(x16, 1e-5, 0.0, None, None, gam, bet, True))
    add("bn_forward (T5: batch statistics + BN-apply + ReLU, [%d, 16])" % m, 4 * 16 * m * 3, t)
    y16, mean, rstd = ops.bn_forward(x16, 1e-5, 0.0, None, None, gam, bet, True)
    t = timer(lambda: ops.bn_backward(x16, y16, g16, mean, rstd, gam, True, True))
    add("bn_backward (T5: dgamma/dbeta sums + dx, [%d, 16])" % m, 4 * 16 * m * 4, t)
    # V2: devoxelise gather + scatter-add
    v2p = data["voxel_point_map"]
    t = timer(lambda: ops.devoxelize(x16, v2p))
    add("gather_rows (V2: devoxelise [%d,16] -> [%d,16])" % (m, n_pts), 4 * 16 * (m + n_pts) + 8 * n_pts, t)
    gp = torch.randn(n_pts, 16, device=device)
    out_sc = torch.zeros(m, 16, device=device)
    t = timer(lambda: ops.check(ops.lib().b2s_scatter_add_rows(ops.ptr(gp), ops.ptr(v2p), n_pts, 16, ops.ptr(out_sc),
                                                                ops.stream()), "scatter"))
    add("scatter_add_rows (V2 backward)", 4 * 16 * (m + n_pts) + 8 * n_pts, t)
    # C1 / C2: ball query + BFS clustering on the batch's foreground points (raw and shifted coordinates)
    model = models.build_model(cfg).to(device)
    scores, offsets = model._cluster_inputs(data, {"semantic_scores": torch.zeros((n_pts, cfg.classes), device=device)})
    preds = scores.argmax(1).to(torch.int16)
    obj = model._object_points(preds)
    bidx = data["vert_batch_ids"][obj].contiguous()
    boffs = torch.cumsum(torch.bincount(bidx + 1), dim=0).int()
    lab = preds[obj].contiguous()
    n_fg = obj.numel()
    cluster = {"foreground_points": n_fg}
    wall = 0.0
    for tag, pts in (("raw", data["point_xyz"][obj].contiguous()),
                     ("shifted", (data["point_xyz"] + offsets)[obj].contiguous())):
        def bq():
            return ops.ballquery(pts, bidx, boffs, cfg.cluster_radius)
        idx, sl = bq()
        n_act = idx.numel()
        t_bq = timer(bq, reps=3, warm=1)
        add("ballquery (C1: count + fill, %s coords, %d points, %d pairs)" % (tag, n_fg, n_act),
            12 * n_fg + 8 * n_fg + n_fg + 4 * n_act, t_bq)

        def bfs():
            comp = ops.cluster_label(idx, sl, lab)
            return ops.cluster_extract(idx, sl, lab, comp, mode=0, thr_i=cfg.cluster_npoint_thre)
        ci, co = bfs()
        t_bfs = timer(bfs, reps=3, warm=1)
        add("bfs_cluster (C2: union-find labels + BFS order, %s coords, %d clusters)" % (tag, co.numel() - 1),
            4 * n_act + 8 * n_fg + 2 * n_fg + 8 * ci.size(0) + 4 * co.numel(), t_bfs)
        cluster["%s_ballquery_ms" % tag], cluster["%s_bfs_ms" % tag], cluster["%s_pairs" % tag] = t_bq * 1e3, t_bfs * 1e3, n_act
        wall += t_bq + t_bfs
    n_scenes = int(data["vert_batch_ids"].max().item()) + 1
    cluster["ms_per_scene"] = wall * 1e3 / n_scenes
    # S1 / S2: segmented reductions on proposal-shaped input
    offs = torch.arange(0, n_pts + 1, 1500, device=device, dtype=torch.int32)
    seg_x = torch.randn(int(offs[-1].item()), 16, device=device)
    t = timer(lambda: common_ops.roipool(seg_x, offs))
    add("roipool_fp (S2: segmented max + argmax, %d segments x 1500 rows x 16)" % (offs.numel() - 1),
        4 * 16 * seg_x.size(0) + 4 * offs.numel() + 8 * 16 * (offs.numel() - 1), t)
    xyz_seg = data["point_xyz"][:seg_x.size(0)].contiguous()
    t = timer(lambda: common_ops.sec_mean(xyz_seg, offs))
    add("sec_mean (S1: segmented mean, C = 3)", 4 * 3 * seg_x.size(0) + 4 * offs.numel() + 12 * (offs.numel() - 1), t)
    return dominant, rows, cluster, which


def reference_cluster_timings(data, device, cfg):
    """the reference's own brute-force CUDA ball query (same B200) and its
    single-threaded CPU BFS (oracle/_ref), on the foreground points of the benchmark batch."""
    from oracle import build_ref
    from minsu3d_b200.harness import models
    ref = build_ref.load()
    if ref is None:
        return {"unavailable": "oracle/_ref not built"}
    n_pts = data["point_xyz"].size(0)
    model = models.build_model(cfg).to(device)
    scores, offsets = model._cluster_inputs(data, {"semantic_scores": torch.zeros((n_pts, cfg.classes), device=device)})
    preds = scores.argmax(1).to(torch.int16)
    obj = model._object_points(preds)
    bidx = data["vert_batch_ids"][obj].contiguous()
    boffs = torch.cumsum(torch.bincount(bidx + 1), dim=0).int()
    lab = preds[obj].contiguous().cpu()
    n = obj.numel()
    out = {}
    wall = 0.0
    for tag, pts, mean_active in (("raw", data["point_xyz"][obj].contiguous(), cfg.cluster_meanActive),
                                  ("shifted", (data["point_xyz"] + offsets)[obj].contiguous(), 1000)):
        idx = torch.zeros(n * mean_active, dtype=torch.int32, device=device)
        sl = torch.zeros((n, 2), dtype=torch.int32, device=device)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n_act = ref.ballquery_batch_p(pts, bidx, boffs, idx, sl, n, mean_active, cfg.cluster_radius)
        torch.cuda.synchronize()
        t_bq = time.perf_counter() - t0
        idx_c, sl_c = idx[:n_act].cpu(), sl.cpu()
        ci, co = torch.zeros(0, dtype=torch.int32), torch.zeros(0, dtype=torch.int32)
        t0 = time.perf_counter()
        ref.pg_bfs_cluster(lab, idx_c, sl_c, ci, co, n, cfg.cluster_npoint_thre)
        t_bfs = time.perf_counter() - t0
        out["%s_ballquery_cuda_ms" % tag], out["%s_bfs_cpu_ms" % tag] = t_bq * 1e3, t_bfs * 1e3
        wall += t_bq + t_bfs
    n_scenes = int(data["vert_batch_ids"].max().item()) + 1
    out["ms_per_scene"] = wall * 1e3 / n_scenes
    out["note"] = ("reference COMMON_OPS compiled unmodified (oracle/build_ref.py): brute-force CUDA ball query on this B200 + "
                   "1-thread CPU BFS; excludes the reference's three D2H copies of the pair lists")
    return out


# --------------------------------------------------------------------------------------------
# own arm: timed loops
# --------------------------------------------------------------------------------------------
def _timed(fn, pool, steps, warmup, world, device):
    from minsu3d_b200 import _cabi
    for i in range(warmup):
        fn(pool[i % len(pool)])
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    _cabi.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(pool[(warmup + i) % len(pool)])
    e1.record()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=device)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())
    return ms, _cabi.launch_count()


def run_train(args, model_name):
    from minsu3d_b200 import _cabi, dp
    from minsu3d_b200.harness import models, scenes, train
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    rank, world, local = dp.init_from_env()
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    _cabi.lib()
    per_gpu = SCENES_PER_GPU if model_name == "pointgroup" else 8  # configs[2]: HAIS batch 8 / GPU
    cfg = models.Config.for_model(model_name, proposal_source="gt_noise")
    trainer = train.Trainer(cfg, device, reserve_gb=16.0 if model_name == "pointgroup" else 32.0,
                            overlap_allreduce=args.overlap)
    n_pool = 3
    pool_host, pool_dev = [], []
    for i in range(n_pool):
        seeds = [(rank * n_pool + i) * per_gpu + s for s in range(per_gpu)]
        d = scenes.make_batch(seeds, device, POINTS_PER_SCENE)
        if not args.size_hints:
            # the reference's loader hands over no level sizes: by default the step reads the pyramid's row counts from
            # the device (one host read); --size-hints lets the loader report them (validated on the device)
            d.pop("voxel_level_sizes", None)
        pool_dev.append(d)
        pool_host.append(train.to_pinned_host(d))
    h2d = int(np.mean([train.host_bytes(h) for h in pool_host]))
    sampler = ClockSampler(local)
    sampler.start()
    ms, launches = _timed(trainer.step, pool_dev, args.steps, args.warmup, world, device)
    ms_e2e, _ = _timed(trainer.step_from_host, pool_host, args.steps, 1, world, device)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    scenes_total = per_gpu * world * args.steps
    value = scenes_total / (ms * 1e-3)
    e2e = scenes_total / (ms_e2e * 1e-3)
    n_prop = 0
    out = trainer.model(pool_dev[0])
    if out.get("proposal_scores") is not None:
        n_prop = int(out["proposal_scores"][2].numel() - 1)
    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return
    metric = METRIC if model_name == "pointgroup" else "%s_train_scenes_per_s" % model_name
    line = {
        "metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": ("configs[1]: PointGroup full train step (MinkUNet m=16 backbone + 2x ballquery_batch_p/"
                                "pg_bfs_cluster + ScoreNet + losses + Adam), batch %d synthetic 100k-point scenes per GPU, "
                                "2 cm voxels" % per_gpu) if model_name == "pointgroup" else
                               ("configs[2]: HAIS train step (MinkUNet m=32 + ballquery_batch_p + hierarchical_aggregation + "
                                "intra-instance refinement + losses + Adam), batch %d synthetic 100k-point scenes per GPU" % per_gpu),
                   "global_batch_scenes": per_gpu * world, "points_per_scene": POINTS_PER_SCENE,
                   "voxels_per_gpu": int(pool_dev[0]["voxel_xyz"].size(0)), "proposals_per_gpu": n_prop,
                   "proposal_source": "gt_noise (GT labels/offsets + noise drive the clustering stage so that "
                                      "random-init weights yield proposals; network outputs still get their losses)",
                   "parallelism": "dp%d (scene-sharded, bucketed NCCL gradient all-reduce%s)" % (
                       world, ", overlapped with backward" if args.overlap else ""),
                   "cache": "per-step working set (activations, ~GBs) >> 126 MB L2; %d batches rotated" % n_pool,
                   "sizes": ("loader-reported level sizes (--size-hints), validated on the device" if args.size_hints else
                             "row counts of the strided maps are read back from the device (as the reference does); "
                             "voxel uniqueness is guaranteed by sparse_quantize and validated on the device"),
                   "conv_algo": "tcgen05 3xTF32 implicit GEMM, persistent kernel (fp32-class accuracy); fp32 FMA for the "
                                "3/6-channel input conv; weight gradient: deterministic mma.sync 3xTF32 (wgrad_det.cu)"},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches,
        "clocks": sampler.summary(),
    }
    if model_name == "pointgroup":
        dominant, table, cluster, which = kernel_rooflines(pool_dev[0], device, cfg)
        roof = dict(dominant)
        roof["peak_source"] = which + " (burst copy, MEASURED_PEAKS.json)"
        roof["traffic_source"] = "profiles/r02_kernel_traffic.json (ncu --set full of tools/experiments/tcp_ncu_target.py; static)"
        roof["timing"] = "CUDA events on the launch stream, L2 flushed (256 MB write) between launches"
        line["roofline"] = roof
        line["roofline_table"] = table
        line["cluster"] = cluster
        if world == 1 and not args.no_cpu_baseline:
            cluster["reference"] = reference_cluster_timings(pool_dev[0], device, cfg)
            sps, sec, stages, used_ref = cpu_train_step_scenes_per_s(1, 1)
            line["cpu_baseline"] = {
                "value": sps, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                "sample": "1 step of: %s (%.1f s; stages %s; reference BFS binary: %s)" % (
                    CPU_SAMPLE, sec, {k: round(v, 2) for k, v in stages.items()}, used_ref)}
    print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def run_softgroup_infer(args):
    """configs[3]: SoftGroup inference over 312 synthetic val-size scenes, scene i -> rank i % world (test.py is
    single-GPU), GPU post-processing (postprocess.softgroup_pred_instances) inside the timed region,
    instance counts gathered on rank 0."""
    from minsu3d_b200 import dp, postprocess
    from minsu3d_b200.harness import models, scenes
    rank, world, local = dp.init_from_env()
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    cfg = models.Config.for_model("softgroup", proposal_source="gt_noise")
    torch.manual_seed(123)
    model = models.build_model(cfg).to(device).eval()
    sizes = [50_000, 100_000, 150_000, 200_000, 250_000, 120_000, 80_000, 180_000]
    distinct = [scenes.collate([scenes.make_scene(500 + i, n)], device) for i, n in enumerate(sizes)]
    n_val = 312
    which = np.random.default_rng(0).integers(0, len(distinct), n_val)  # val scene i has the size of distinct[which[i]]
    mine = dp.shard_indices(n_val, rank, world)
    inst_classes = cfg.classes - len(cfg.ignore_classes)

    def one(i):
        d = distinct[int(which[i])]
        with torch.no_grad():
            out = model(d)
            if out.get("proposals_idx") is None:
                return 0
            res = postprocess.softgroup_pred_instances(d["point_xyz"], out["proposals_idx"], d["point_xyz"].size(0),
                                                       out["cls_scores"], out["iou_scores"], out["mask_scores"],
                                                       inst_classes, -0.5, 0.001, 100)
        return int(res["label_id"].numel())

    for i in range(min(args.warmup, len(mine))):
        one(mine[i])
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    counts = [one(i) for i in mine]
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    total_instances = sum(counts)
    if world > 1:
        gathered = [None] * world if rank == 0 else None
        torch.distributed.gather_object(counts, gathered, dst=0)
        t = torch.tensor([ms], device=device)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())
        if rank == 0:
            total_instances = sum(sum(g) for g in gathered)
    if rank == 0:
        print(json.dumps({
            "metric": "softgroup_inference_scenes_per_s", "value": n_val / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": 1, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[3]: SoftGroup (m=32) inference over 312 synthetic val-size scenes (8 distinct "
                                   "50k-250k-point scenes, drawn per val index with a fixed seed), scene i -> rank i % world, GPU post-processing, instance "
                                   "counts gathered on rank 0", "predicted_instances": total_instances,
                       "ms_per_scene_per_gpu": ms / max(len(mine), 1)}}), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--workload", default="pointgroup_train", choices=["pointgroup_train", "hais_train", "softgroup_infer"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--size-hints", action="store_true",
                    help="the loader reports the strided maps' row counts with the batch (no host read in the backbone)")
    ap.add_argument("--no-overlap", dest="overlap", action="store_false",
                    help="all-reduce the gradient buckets after backward instead of from inside it")
    ap.add_argument("--roofline-only", action="store_true",
                    help="only the per-kernel roofline table (the command the ncu captures in profiles/ profile)")
    args = ap.parse_args()
    if args.roofline_only:
        from minsu3d_b200.harness import models, scenes
        device = torch.device("cuda", 0)
        batch = scenes.make_batch(list(range(SCENES_PER_GPU)), device, POINTS_PER_SCENE)
        dominant, table, cluster, which = kernel_rooflines(batch, device, models.Config.for_model("pointgroup", proposal_source="gt_noise"))
        print(json.dumps({"roofline": dominant, "roofline_table": table, "cluster": cluster}), flush=True)
        return
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "softgroup_infer":
        run_softgroup_infer(args)
    else:
        run_train(args, "hais" if args.workload == "hais_train" else "pointgroup")


if __name__ == "__main__":
    main()
