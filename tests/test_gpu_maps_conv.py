"""GPU parity: coordinate hash / kernel maps (bit-exact) and sparse convolution (fp32 tolerance).

CUDA path (libb2s through the C ABI) vs the CPU oracle (oracle/oracle.c) on identical seeded inputs.
Tolerance for fp32 features/gradients: 1e-4 relative to the output's max-abs (BASELINE.json
north_star's example tolerance); integer/index results must be bit-exact.
"""
import numpy as np
import pytest
import torch

import oracle
from helpers import random_voxels, surface_voxels

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _close(got, want, tol=RTOL):
    got, want = got.detach().cpu().double().numpy(), np.asarray(want, np.float64)
    assert got.shape == want.shape
    scale = max(np.abs(want).max(), 1e-6)
    err = np.abs(got - want).max() / scale
    assert err < tol, "max rel err %.3e" % err


@pytest.mark.parametrize("n,dup", [(0, 0.0), (1, 0.0), (5000, 0.3), (200_000, 0.2)])
def test_coord_unique_matches_oracle(n, dup):
    from minsu3d_b200 import ops
    rng = np.random.default_rng(n + 1)
    c = random_voxels(rng, n, dup=dup) if n else np.zeros((0, 4), np.int32)
    ui, inv, oc = oracle.coord_unique(c, 1)
    table, g_ui, g_inv, g_oc = ops.coord_unique(_dev(c), 1)
    assert np.array_equal(g_ui.cpu().numpy(), ui)
    assert np.array_equal(g_inv.cpu().numpy(), inv)
    assert np.array_equal(g_oc.cpu().numpy(), oc)


@pytest.mark.parametrize("quant", [2, 4, 16])
def test_stride_map_matches_oracle(quant):
    from minsu3d_b200 import ops
    rng = np.random.default_rng(quant)
    c = random_voxels(rng, 50_000)
    c[:, 1:] = (c[:, 1:] // (quant // 2)) * (quant // 2)  # input lives on stride quant/2
    c = oracle.coord_unique(c, 1)[2]
    ui, inv, oc = oracle.coord_unique(c, quant)
    _, g_ui, g_inv, g_oc = ops.coord_unique(_dev(c), quant)
    assert np.array_equal(g_oc.cpu().numpy(), oc)
    assert np.array_equal(g_inv.cpu().numpy(), inv)


def test_coord_range_error():
    from minsu3d_b200 import ops
    c = np.array([[0, 0, 0, 0], [0, 20000, 0, 0]], np.int32)
    with pytest.raises(ValueError):
        ops.coord_unique(_dev(c), 1)


@pytest.mark.parametrize("ksize,n", [(3, 3000), (3, 120_000), (1, 1000), (5, 2000)])
def test_kernel_map_and_pairs_bit_exact(ksize, n):
    from minsu3d_b200 import ops
    rng = np.random.default_rng(ksize * 7 + n)
    c = surface_voxels(rng, n)
    nbr = oracle.kernel_map(c, c, ksize, 1)
    table, _, _, oc = ops.coord_unique(_dev(c), 1)
    g_nbr = ops.kernel_map(oc, table, ksize, 1)
    assert np.array_equal(g_nbr.cpu().numpy(), nbr)
    pin, pout, koff = oracle.pairs_from_nbr(nbr)
    g_in, g_out, g_koff, p = ops.pairs_from_nbr(g_nbr, exact=True)
    assert p == pin.size
    assert np.array_equal(g_koff.cpu().numpy(), koff)
    assert np.array_equal(g_in.cpu().numpy(), pin)
    assert np.array_equal(g_out.cpu().numpy(), pout)
    if ksize % 2 == 1:  # symmetry used by the data-gradient kernel: nbr[i, K-1-k] = o <=> nbr[o, k] = i
        K = ksize ** 3
        o, k = np.nonzero(nbr >= 0)
        assert np.array_equal(nbr[nbr[o, k], K - 1 - k], o)


def test_strided_kernel_map_bit_exact():
    from minsu3d_b200 import ops
    rng = np.random.default_rng(3)
    c = surface_voxels(rng, 60_000)
    _, inv, oc = oracle.coord_unique(c, 2)
    nbr = oracle.kernel_map(c, oc, 2, 1)
    assert (nbr >= 0).sum() == c.shape[0]  # every fine row has exactly one (coarse row, offset)
    t_in, _, _, c_in = ops.coord_unique(_dev(c), 1)
    _, _, g_inv, g_oc = ops.coord_unique(c_in, 2)
    g_nbr = ops.kernel_map(g_oc, t_in, 2, 1)
    assert np.array_equal(g_oc.cpu().numpy(), oc)
    assert np.array_equal(g_nbr.cpu().numpy(), nbr)
    # the parent found through the table equals the inverse map of the stride insert
    o, k = np.nonzero(nbr >= 0)
    assert np.array_equal(inv[nbr[o, k]], o)


def test_coordinate_pyramid_single_sync_matches_level_by_level_oracle():
    """ops.coord_pyramid builds every strided map on the device (upper-bound buffers + device-side counts)."""
    from minsu3d_b200 import ops
    rng = np.random.default_rng(17)
    c = surface_voxels(rng, 90_000, batch=3)
    levels = ops.coord_pyramid(_dev(c), 1, 6)
    cur = c
    for (stride, table, oc), want_stride in zip(levels, (2, 4, 8, 16, 32, 64)):
        assert stride == want_stride
        _, _, want = oracle.coord_unique(cur, stride)
        assert np.array_equal(oc.cpu().numpy(), want)
        # the table answers lookups for the level's own coordinates: 1x1x1 kernel map = identity
        nbr = ops.kernel_map(oc, table, 1, stride)
        assert np.array_equal(nbr.cpu().numpy()[:, 0], np.arange(want.shape[0]))
        cur = want


CONV_SHAPES = [(6, 16), (16, 16), (32, 16), (32, 32), (48, 48), (64, 32), (64, 64), (96, 112), (224, 224), (20, 24)]


ALGOS = {"fma": (1, RTOL), "tc3xtf32": (2, RTOL), "tctf32": (3, 5e-3)}


@pytest.mark.parametrize("algo_name", list(ALGOS))
@pytest.mark.parametrize("cin,cout", CONV_SHAPES)
def test_conv3_forward_backward(cin, cout, algo_name):
    """algo fma = fp32 FMA path; tc3xtf32 = tcgen05 3xTF32 (same 1e-4 tolerance as fp32); tctf32 = plain TF32,
    whose 10-bit mantissa gives ~1e-3 and is therefore opt-in only."""
    from minsu3d_b200 import ops
    algo, tol = ALGOS[algo_name]
    if algo != 1 and (cin % 16 or cout % 16):
        pytest.skip("tcgen05 path needs channel counts that are multiples of 16")
    rng = np.random.default_rng(cin * 1000 + cout)
    n = 6000 if cin * cout > 4096 else 20_000
    c = surface_voxels(rng, n)
    n = c.shape[0]
    nbr = oracle.kernel_map(c, c, 3, 1)
    x = rng.standard_normal((n, cin)).astype(np.float32)
    w = (rng.standard_normal((27, cin, cout)) / np.sqrt(27 * cin)).astype(np.float32)
    g = rng.standard_normal((n, cout)).astype(np.float32)
    want = oracle.conv_fwd(x, w, nbr, n)
    want_gin, want_gw = oracle.conv_bwd(x, w, g, nbr)
    d_nbr = _dev(nbr)
    got = ops.conv_table(_dev(x), _dev(w), d_nbr, n, 27, cin, cout, algo=algo)
    _close(got, want, tol)
    gin = ops.conv_table(_dev(g), _dev(w), d_nbr, n, 27, cout, cin, w_transposed=True, k_reversed=True, algo=algo)
    _close(gin, want_gin, tol)
    pin, pout, koff, maxp = ops.pairs_from_nbr(d_nbr)[:3] + (n * 27,)
    gw = ops.conv_wgrad(_dev(x), _dev(g), pin, pout, koff, 27, cin, cout, maxp, algo=ops.ALGO_SIMT)
    _close(gw, want_gw)


@pytest.mark.parametrize("n,c", [(1, 16), (300, 16), (70_000, 16), (325_422, 16), (90_000, 32), (5000, 64), (300, 224)])
def test_bn_one_launch_equals_two_kernel_path(n, c):
    """b2s_bn_forward (one cooperative launch: statistics -> device-wide barrier -> apply) gives the same bits as
    b2s_bn_stats followed by b2s_bn_apply (two kernels), running statistics included."""
    from minsu3d_b200 import ops
    torch.manual_seed(n + c)
    x = torch.randn(n, c, device="cuda") * 3 + 1
    gamma, beta = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda")
    rm1, rv1 = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
    rm2, rv2 = rm1.clone(), rv1.clone()
    y, mean, rstd = ops.bn_forward(x, 1e-4, 0.1, rm1, rv1, gamma, beta, True)
    mean2, rstd2 = ops.bn_stats(x, 1e-4, 0.1, rm2, rv2)
    y2 = ops.bn_apply(x, mean2, rstd2, gamma, beta, True)
    for a, b in ((y, y2), (mean, mean2), (rstd, rstd2), (rm1, rm2), (rv1, rv2)):
        assert torch.equal(a, b)


@pytest.mark.parametrize("cin,cout", [(16, 16), (32, 16), (16, 32), (32, 32)])
@pytest.mark.parametrize("n", [1, 17, 5000, 60_000])
def test_conv3_warp_stream_kernel(cin, cout, n):
    """algo 4 (csrc/conv_ws.cu, mma.sync 3xTF32 straight from the gathered rows): forward with the residual add and the
    data gradient (transposed weights, reversed offsets) vs the oracle at 1e-4, row-order and mask-sorted tables."""
    from minsu3d_b200 import ops
    rng = np.random.default_rng(cin * 100 + cout + n)
    c = surface_voxels(rng, n)
    n = c.shape[0]
    nbr = oracle.kernel_map(c, c, 3, 1)
    x = rng.standard_normal((n, cin)).astype(np.float32)
    w = (rng.standard_normal((27, cin, cout)) / np.sqrt(27 * cin)).astype(np.float32)
    g = rng.standard_normal((n, cout)).astype(np.float32)
    sc = rng.standard_normal((n, cout)).astype(np.float32)
    want = oracle.conv_fwd(x, w, nbr, n) + sc
    want_gin, _ = oracle.conv_bwd(x, w, g, nbr)
    d_nbr = _dev(nbr)
    tables = [(d_nbr, {})]
    if n >= 32768:
        perm, nbs, tms = ops.tile_order(d_nbr)
        tables.append((nbs, dict(out_rows=perm, tile_mask=tms)))
    for tb, kw in tables:
        got = ops.conv_table(_dev(x), _dev(w), tb, n, 27, cin, cout, algo=ops.ALGO_WARP_STREAM, add_src=_dev(sc), **kw)
        _close(got, want)
        gin = ops.conv_table(_dev(g), _dev(w), tb, n, 27, cout, cin, w_transposed=True, k_reversed=True,
                             algo=ops.ALGO_WARP_STREAM, **kw)
        _close(gin, want_gin)


WGRAD_SHAPES = [(16, 16), (32, 16), (16, 32), (32, 32), (48, 32), (48, 48), (64, 64), (80, 80), (96, 112), (224, 112),
                (128, 256), (3, 16), (6, 16), (16, 20), (16, 3), (20, 24), (40, 1)]


@pytest.mark.parametrize("ca,cg", WGRAD_SHAPES)
def test_wgrad_deterministic_kernel(ca, cg):
    """T3 weight gradient, default path for channel counts that are multiples of 16 (csrc/wgrad_det.cu, mma.sync
    3xTF32 straight from the gathered rows): 1e-4 of max-abs vs the oracle for every register tiling incl. ragged
    channel tiles, and bit-identical results run to run (per-warp partials summed in a fixed order, no atomics)."""
    from minsu3d_b200 import ops
    rng = np.random.default_rng(ca * 1000 + cg)
    n = 3000 if ca * cg > 4096 else 20_000
    c = surface_voxels(rng, n)
    n = c.shape[0]
    nbr = oracle.kernel_map(c, c, 3, 1)
    x = rng.standard_normal((n, ca)).astype(np.float32)
    w = np.zeros((27, ca, cg), np.float32)
    g = rng.standard_normal((n, cg)).astype(np.float32)
    _, want_gw = oracle.conv_bwd(x, w, g, nbr)
    pin, pout, koff, _ = ops.pairs_from_nbr(_dev(nbr))
    gw = ops.conv_wgrad(_dev(x), _dev(g), pin, pout, koff, 27, ca, cg, n * 27, algo=ops.ALGO_TC_3XTF32)
    _close(gw, want_gw)
    again = ops.conv_wgrad(_dev(x), _dev(g), pin, pout, koff, 27, ca, cg, n * 27, algo=ops.ALGO_TC_3XTF32)
    assert torch.equal(gw, again)
    # exact pair count instead of the n*K bound: a different warp split, same sums within tolerance
    exact = int(koff[-1].item())
    _close(ops.conv_wgrad(_dev(x), _dev(g), pin, pout, koff, 27, ca, cg, exact, algo=ops.ALGO_TC_3XTF32), want_gw)


@pytest.mark.parametrize("n", [0, 1, 7, 129, 1000])
def test_wgrad_deterministic_small_and_identity(n):
    """Edge cases: empty / one-row maps (most offsets have no pair) and the K = 1 identity map of a 1x1 convolution."""
    from minsu3d_b200 import ops
    rng = np.random.default_rng(n)
    ca, cg = 32, 48
    if n:
        c = surface_voxels(rng, n)
        n = c.shape[0]
        nbr = oracle.kernel_map(c, c, 3, 1)
    else:
        nbr = np.zeros((0, 27), np.int32)
    x = rng.standard_normal((n, ca)).astype(np.float32)
    g = rng.standard_normal((n, cg)).astype(np.float32)
    pin, pout, koff, _ = ops.pairs_from_nbr(_dev(nbr))
    gw = ops.conv_wgrad(_dev(x), _dev(g), pin, pout, koff, 27, ca, cg, n * 27)
    if n:
        _, want = oracle.conv_bwd(x, np.zeros((27, ca, cg), np.float32), g, nbr)
        _close(gw, want)
    else:
        assert float(gw.abs().max()) == 0.0
    ident = torch.arange(n, dtype=torch.int32, device="cuda")
    koff1 = torch.tensor([0, n], dtype=torch.int32, device="cuda")
    gw1 = ops.conv_wgrad(_dev(x), _dev(g), ident, ident, koff1, 1, ca, cg, n)
    want1 = x.astype(np.float64).T @ g.astype(np.float64)
    if n:
        _close(gw1[0], want1)
    else:
        assert float(gw1.abs().max()) == 0.0


@pytest.mark.parametrize("cin,cout", [(16, 16), (16, 20), (16, 3), (16, 1)])
def test_point_linear_matches_torch(cin, cout):
    """ops.linear (harness heads): forward = F.linear bit for bit, gradients vs torch autograd in float64."""
    from minsu3d_b200 import ops
    torch.manual_seed(cin * 100 + cout)
    n = 50_000
    x = torch.randn(n, cin, device="cuda", requires_grad=True)
    w = (torch.randn(cout, cin, device="cuda") / cin ** 0.5).requires_grad_()
    b = torch.randn(cout, device="cuda", requires_grad=True)
    g = torch.randn(n, cout, device="cuda")
    y = ops.linear(x, w, b)
    assert torch.equal(y, torch.nn.functional.linear(x, w, b))
    y.backward(g)
    xd, wd, bd = (t.detach().double().requires_grad_() for t in (x, w, b))
    torch.nn.functional.linear(xd, wd, bd).backward(g.double())
    _close(x.grad, xd.grad.cpu().numpy())
    _close(w.grad, wd.grad.cpu().numpy())
    _close(b.grad, bd.grad.cpu().numpy())


@pytest.mark.parametrize("n", [1, 100, 5000, 150_000])
def test_tile_order_bit_exact(n):
    """Mask-sorted tile schedule: permutation, permuted table and tile masks equal the numpy restatement."""
    from minsu3d_b200 import ops
    rng = np.random.default_rng(n)
    c = surface_voxels(rng, n, batch=3)
    n = c.shape[0]
    nbr = oracle.kernel_map(c, c, 3, 1)
    perm, nbr_sorted, tile_mask = oracle.tile_order(nbr)
    g_perm, g_sorted, g_mask = ops.tile_order(_dev(nbr))
    assert np.array_equal(g_perm.cpu().numpy(), perm)
    assert np.array_equal(g_sorted.cpu().numpy(), nbr_sorted)
    assert np.array_equal(g_mask.cpu().numpy().view(np.uint32), tile_mask)
    assert np.array_equal(np.sort(perm), np.arange(n))
    if n >= 5000:  # the point of the schedule: far fewer active offsets per tile than the 27 of a shuffled order
        active = np.array([bin(int(m)).count("1") for m in tile_mask]).mean()
        assert active < 20, active


@pytest.mark.parametrize("algo_name", ["tc3xtf32", "tctf32"])
@pytest.mark.parametrize("cin,cout", [(16, 16), (32, 16), (64, 64), (96, 112)])
def test_conv3_sorted_tiles_equal_row_order(cin, cout, algo_name):
    """b2s_conv_table_rows over the mask-sorted table = b2s_conv_table over the first-occurrence table
    (same products in the same k order; skipped slabs are exact zeros), forward and data gradient; and both
    match the oracle."""
    from minsu3d_b200 import ops
    algo, tol = ALGOS[algo_name]
    rng = np.random.default_rng(cin * 77 + cout)
    c = surface_voxels(rng, 20_000 if cin * cout <= 4096 else 6000, batch=2)
    n = c.shape[0]
    nbr = oracle.kernel_map(c, c, 3, 1)
    x = rng.standard_normal((n, cin)).astype(np.float32)
    w = (rng.standard_normal((27, cin, cout)) / np.sqrt(27 * cin)).astype(np.float32)
    g = rng.standard_normal((n, cout)).astype(np.float32)
    d_nbr = _dev(nbr)
    perm, nbr_sorted, tile_mask = ops.tile_order(d_nbr)
    plain = ops.conv_table(_dev(x), _dev(w), d_nbr, n, 27, cin, cout, algo=algo)
    got = ops.conv_table(_dev(x), _dev(w), nbr_sorted, n, 27, cin, cout, algo=algo, tile_mask=tile_mask, out_rows=perm)
    assert torch.equal(got, plain)
    _close(got, oracle.conv_fwd(x, w, nbr, n), tol)
    plain_g = ops.conv_table(_dev(g), _dev(w), d_nbr, n, 27, cout, cin, w_transposed=True, k_reversed=True, algo=algo)
    got_g = ops.conv_table(_dev(g), _dev(w), nbr_sorted, n, 27, cout, cin, w_transposed=True, k_reversed=True,
                           algo=algo, tile_mask=tile_mask, out_rows=perm)
    assert torch.equal(got_g, plain_g)
    _close(got_g, oracle.conv_bwd(x, w, g, nbr)[0], tol)
    with pytest.raises(RuntimeError):  # the fp32 FMA path takes no permutation: refused, never silently ignored
        ops.conv_table(_dev(x), _dev(w), nbr_sorted, n, 27, cin, cout, algo=ops.ALGO_SIMT, tile_mask=tile_mask,
                       out_rows=perm)


def test_conv_module_uses_sorted_tiles_on_large_maps(monkeypatch):
    """MinkowskiConvolution picks the sorted schedule by map size; both schedules give the same features/gradients."""
    from minsu3d_b200 import MinkowskiEngine as ME
    from minsu3d_b200.MinkowskiEngine.sparse_tensor import KernelMap
    rng = np.random.default_rng(21)
    c = surface_voxels(rng, 30_000)
    n = c.shape[0]
    x = rng.standard_normal((n, 16)).astype(np.float32)
    g = _dev(rng.standard_normal((n, 32)).astype(np.float32))
    conv = ME.MinkowskiConvolution(16, 32, kernel_size=3, dimension=3).cuda()
    res = []
    for min_rows in (1 << 30, 0):
        monkeypatch.setattr(KernelMap, "SORTED_MIN_ROWS", min_rows)
        xt = _dev(x).requires_grad_(True)
        st = ME.SparseTensor(features=xt, coordinates=_dev(c))
        y = conv(st)
        km = next(iter(st.coordinate_manager._kmaps.values()))
        assert (km._sorted is not None) == (min_rows == 0)
        conv.kernel.grad = None
        y.F.backward(g)
        res.append((y.F.detach(), xt.grad, conv.kernel.grad.clone()))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
    _close(res[1][2], res[0][2].cpu().numpy(), 1e-5)  # weight gradient: atomic flush order varies run to run


@pytest.mark.parametrize("cin,cout", [(16, 32), (32, 48), (112, 96), (64, 32)])
def test_strided_and_transposed_conv_modules(cin, cout):
    """MinkowskiConvolution(k=2,s=2) and MinkowskiConvolutionTranspose(k=2,s=2) incl. autograd."""
    from minsu3d_b200 import MinkowskiEngine as ME
    rng = np.random.default_rng(cin + cout)
    c = surface_voxels(rng, 15_000)
    n = c.shape[0]
    x = rng.standard_normal((n, cin)).astype(np.float32)
    _, _, oc = oracle.coord_unique(c, 2)
    m = oc.shape[0]
    nbr_down = oracle.kernel_map(c, oc, 2, 1)
    down = ME.MinkowskiConvolution(cin, cout, kernel_size=2, stride=2, dimension=3).cuda()
    up = ME.MinkowskiConvolutionTranspose(cout, cin, kernel_size=2, stride=2, dimension=3).cuda()
    xt = _dev(x).requires_grad_(True)
    st = ME.SparseTensor(features=xt, coordinates=_dev(c))
    y = down(st)
    z = up(y)
    assert np.array_equal(y.C.cpu().numpy(), oc)
    assert z.coordinate_map_key == st.coordinate_map_key
    wd = down.kernel.detach().cpu().numpy()
    wu = up.kernel.detach().cpu().numpy()
    want_y = oracle.conv_fwd(x, wd, nbr_down, m)
    want_z = oracle.convT_fwd(want_y, wu, nbr_down, n)
    _close(y.F, want_y)
    _close(z.F, want_z)
    gz = rng.standard_normal((n, cin)).astype(np.float32)
    z.F.backward(_dev(gz))
    # oracle gradients by the definition: convT is the adjoint structure of the strided conv
    gy = np.zeros((m, cout), np.float64)
    gwu = np.zeros(wu.shape, np.float64)
    gwd = np.zeros(wd.shape, np.float64)
    gx = np.zeros((n, cin), np.float64)
    for k in range(8):
        o = np.nonzero(nbr_down[:, k] >= 0)[0]
        i = nbr_down[o, k]
        gy[o] += gz[i].astype(np.float64) @ wu[k].astype(np.float64).T
        gwu[k] = want_y[o].astype(np.float64).T @ gz[i].astype(np.float64)
    for k in range(8):
        o = np.nonzero(nbr_down[:, k] >= 0)[0]
        i = nbr_down[o, k]
        gx[i] = gy[o] @ wd[k].astype(np.float64).T
        gwd[k] = x[i].astype(np.float64).T @ gy[o]
    _close(up.kernel.grad, gwu)
    _close(down.kernel.grad, gwd)
    _close(xt.grad, gx)


def test_conv1x1_and_empty():
    from minsu3d_b200 import ops
    rng = np.random.default_rng(5)
    x = rng.standard_normal((5000, 32)).astype(np.float32)
    w = rng.standard_normal((32, 16)).astype(np.float32)
    got = ops.conv_table(_dev(x), _dev(w), None, 5000, 1, 32, 16, algo=ops.ALGO_SIMT)
    _close(got, x.astype(np.float64) @ w.astype(np.float64))
    empty = ops.conv_table(_dev(x[:0]), _dev(w), None, 0, 1, 32, 16, algo=ops.ALGO_SIMT)
    assert empty.shape == (0, 16)


def test_sparse_conv_equals_dense_conv3d():
    """Independent of the oracle: scatter to a dense grid and run torch conv3d (appendix A cross-check)."""
    from minsu3d_b200 import MinkowskiEngine as ME
    rng = np.random.default_rng(11)
    c = random_voxels(rng, 4000, extent=16, batch=2)
    c[:, 1:] = np.abs(c[:, 1:]) % 12
    c = oracle.coord_unique(c, 1)[2]
    n, cin, cout = c.shape[0], 8, 12
    x = rng.standard_normal((n, cin)).astype(np.float32)
    conv = ME.MinkowskiConvolution(cin, cout, kernel_size=3, dimension=3).cuda()
    y = conv(ME.SparseTensor(features=_dev(x), coordinates=_dev(c))).F.detach().cpu()
    dense = torch.zeros(2, cin, 12, 12, 12, dtype=torch.float64)
    ct = torch.from_numpy(c).long()
    dense[ct[:, 0], :, ct[:, 3], ct[:, 2], ct[:, 1]] = torch.from_numpy(x).double()
    k = conv.kernel.detach().cpu().double()  # [27, cin, cout], kidx = ix + 3 iy + 9 iz
    wd = k.view(3, 3, 3, cin, cout).permute(4, 3, 0, 1, 2)  # [cout, cin, iz, iy, ix]
    out = torch.nn.functional.conv3d(dense, wd, padding=1)
    want = out[ct[:, 0], :, ct[:, 3], ct[:, 2], ct[:, 1]]
    _close(y, want.numpy())


@pytest.mark.parametrize("c,relu", [(16, True), (32, False), (112, True), (224, True)])
def test_batchnorm_relu_matches_torch(c, relu):
    from minsu3d_b200 import MinkowskiEngine as ME
    rng = np.random.default_rng(c)
    n = 30_000
    x = (rng.standard_normal((n, c)) * 2 + 0.5).astype(np.float32)
    coords = oracle.coord_unique(random_voxels(rng, n, extent=80), 1)[2]
    n = coords.shape[0]
    x = x[:n]
    mod = ME.MinkowskiBatchNorm(c).cuda()
    ref = torch.nn.BatchNorm1d(c).cuda()
    with torch.no_grad():
        mod.bn.weight.uniform_(0.5, 1.5)
        mod.bn.bias.uniform_(-0.5, 0.5)
        ref.load_state_dict(mod.bn.state_dict())
    xa = _dev(x).requires_grad_(True)
    xb = _dev(x).requires_grad_(True)
    st = mod(ME.SparseTensor(features=xa, coordinates=_dev(coords)))
    ya = (ME.MinkowskiReLU(inplace=True)(st) if relu else st).F
    yb = ref(xb)
    if relu:
        # share the ReLU mask: pre-activations within 1e-6 of zero may legitimately land on either side
        _close(ya, torch.relu(yb).detach().cpu().numpy(), 1e-5)
        yb = yb * (ya.detach() > 0).float()
    _close(ya, yb.detach().cpu().numpy(), 1e-5)
    g = _dev(rng.standard_normal((n, c)).astype(np.float32))
    ya.backward(g)
    yb.backward(g)
    _close(xa.grad, xb.grad.cpu().numpy(), 1e-4)
    _close(mod.bn.weight.grad, ref.weight.grad.cpu().numpy(), 1e-4)
    _close(mod.bn.bias.grad, ref.bias.grad.cpu().numpy(), 1e-4)
    _close(mod.bn.running_mean, ref.running_mean.cpu().numpy(), 1e-5)
    _close(mod.bn.running_var, ref.running_var.cpu().numpy(), 1e-5)


def test_devoxelize_gather_scatter():
    from minsu3d_b200 import ops
    rng = np.random.default_rng(9)
    m, n, c = 7000, 20_000, 16
    feat = rng.standard_normal((m, c)).astype(np.float32)
    idx = rng.integers(0, m, n)
    f = _dev(feat).requires_grad_(True)
    out = ops.devoxelize(f, _dev(idx))
    assert np.array_equal(out.detach().cpu().numpy(), feat[idx])
    g = rng.standard_normal((n, c)).astype(np.float32)
    out.backward(_dev(g))
    want = np.zeros((m, c), np.float64)
    np.add.at(want, idx, g.astype(np.float64))
    _close(f.grad, want, 1e-5)
