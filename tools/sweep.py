"""BASELINE.json configs[4]: sparse-conv and clustering microbench sweep (50k-1M voxels, C = 16-256, 3^3 kernels).

    python tools/sweep.py [--quick] > profiles/rNN_sweep.json

Per (voxels, channels): kernel-map build, conv forward with tcgen05 3xTF32 / TF32 and the fp32 FMA path (CUDA
events, 5 launches after 2 warm-ups, L2 flushed), useful TFLOP/s = 2*P*Cin*Cout / t, algorithmic GB/s
(SURVEY.md 8(d)); clustering: ball query + label + BFS order in ms per 100k-point scene; CPU oracle beside it at
the sizes it finishes in seconds.
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))

import oracle  # noqa: E402  (CPU baseline leg only)
from minsu3d_b200 import ops  # noqa: E402
from minsu3d_b200.harness import scenes  # noqa: E402


def ev_time(fn, reps=5, warm=2, flush=None):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return float(np.median(ts))


def voxels_of(n_vox):
    """ScanNet-like surface voxels: scenes from the generator, 2 cm voxels, concatenated until n_vox rows."""
    out, b, total = [], 0, 0
    while total < n_vox:
        sc = scenes.make_scene(100 + b, 100_000)
        v = np.unique(np.floor((sc["xyz"] - sc["xyz"].min(0)) / 0.02).astype(np.int32), axis=0)
        out.append(np.concatenate((np.full((v.shape[0], 1), b, np.int32), v), 1))
        total += v.shape[0]
        b += 1
    c = np.concatenate(out)[:n_vox]
    return np.ascontiguousarray(c)


def main():
    quick = "--quick" in sys.argv
    dev = torch.device("cuda", 0)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    sizes = [50_000, 200_000] if quick else [50_000, 100_000, 200_000, 500_000, 1_000_000]
    chans = [16, 64] if quick else [16, 32, 64, 128, 256]
    rows = []
    for n_vox in sizes:
        c = voxels_of(n_vox)
        dc = torch.from_numpy(c).to(dev)
        t_unique = ev_time(lambda: ops.coord_unique(dc, 1))
        table, _, _, oc = ops.coord_unique(dc, 1)
        t_map = ev_time(lambda: ops.kernel_map(oc, table, 3, 1))
        nbr = ops.kernel_map(oc, table, 3, 1)
        pairs = int((nbr >= 0).sum().item())
        m = oc.size(0)
        cpu_map = None
        if n_vox <= 200_000:
            t0 = time.perf_counter()
            oracle.kernel_map(c, c, 3, 1)
            cpu_map = time.perf_counter() - t0
        for ch in chans:
            x = torch.randn(m, ch, device=dev)
            w = torch.randn(27, ch, ch, device=dev) * 0.05
            rec = {"voxels": m, "pairs": pairs, "channels": ch, "coord_unique_us": t_unique * 1e6,
                   "kernel_map_us": t_map * 1e6, "cpu_kernel_map_s": cpu_map}
            flops = 2.0 * pairs * ch * ch
            alg_bytes = 4 * (2 * m * ch) + 4 * pairs + 4 * 27 * ch * ch
            for name, algo in (("tc3xtf32", 2), ("tctf32", 3), ("fma", 1)):
                if algo == 1 and ch > 64 and n_vox > 200_000:
                    continue
                t = ev_time(lambda: ops.conv_table(x, w, nbr, m, 27, ch, ch, algo=algo), flush=flush)
                rec[name + "_us"] = t * 1e6
                rec[name + "_useful_tflops"] = flops / t / 1e12
                rec[name + "_alg_gbs"] = alg_bytes / t / 1e9
            if n_vox <= 100_000 and ch <= 64:
                xn, wn = x.cpu().numpy(), w.cpu().numpy()
                nb = nbr.cpu().numpy()
                t0 = time.perf_counter()
                oracle.conv_fwd(xn, wn, nb, m)
                rec["cpu_conv_s"] = time.perf_counter() - t0
                rec["cpu_cores"] = os.cpu_count()
            rows.append(rec)
            print(json.dumps(rec), file=sys.stderr, flush=True)
    # clustering: one 100k-point scene, raw and shifted coordinates
    clus = []
    for n_scene in ([1] if quick else [1, 4]):
        data = scenes.make_batch(list(range(n_scene)), dev, 100_000)
        fg = torch.nonzero(data["sem_labels"] > 1).view(-1)
        xyz = data["point_xyz"][fg].contiguous()
        shift = (xyz + (data["instance_center_xyz"][fg] - xyz) * 0.85).contiguous()
        bidx = data["vert_batch_ids"][fg].contiguous()
        offs = torch.cumsum(torch.bincount(bidx.long() + 1), 0).int()
        lab = data["sem_labels"][fg].contiguous()
        for tag, pts in (("raw", xyz), ("shifted", shift)):
            t_bq = ev_time(lambda: ops.ballquery(pts, bidx, offs, 0.03), reps=3, warm=1)
            idx, sl = ops.ballquery(pts, bidx, offs, 0.03)
            t_lab = ev_time(lambda: ops.cluster_label(idx, sl, lab), reps=3, warm=1)
            comp = ops.cluster_label(idx, sl, lab)
            t_ext = ev_time(lambda: ops.cluster_extract(idx, sl, lab, comp, 0, thr_i=50), reps=3, warm=1)
            rec = {"scenes": n_scene, "coords": tag, "fg_points": int(fg.numel()), "n_active": int(idx.numel()),
                   "ballquery_ms": t_bq * 1e3, "label_ms": t_lab * 1e3, "select_order_ms": t_ext * 1e3,
                   "cluster_ms_per_scene": (t_bq + t_lab + t_ext) * 1e3 / n_scene,
                   "ballquery_alg_gbs": (21 * fg.numel() + 4 * idx.numel()) / t_bq / 1e9}
            if n_scene == 1:
                xn, bn, on = pts.cpu().numpy(), bidx.cpu().numpy(), offs.cpu().numpy()
                t0 = time.perf_counter()
                ci, csl = oracle.ballquery(xn, bn, on, 0.03)
                rec["cpu_ballquery_s"] = time.perf_counter() - t0
                t0 = time.perf_counter()
                oracle.pg_bfs_cluster(lab.cpu().numpy(), ci, csl, 50)
                rec["cpu_bfs_s_1thread"] = time.perf_counter() - t0
            clus.append(rec)
            print(json.dumps(rec), file=sys.stderr, flush=True)
    print(json.dumps({"conv": rows, "clustering": clus, "gpu": torch.cuda.get_device_name(0)}, indent=1))


if __name__ == "__main__":
    main()
