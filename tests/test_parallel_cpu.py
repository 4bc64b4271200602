"""CPU, world_size 2, gloo: the host-side multi-GPU logic of cmf_b200.parallel.

The CUDA kernels cannot run here, so compute is stood in for by PyTorch CPU ops IN THE TEST ONLY; what is under
test is the sharding arithmetic: gradient all-reduce + global masked-mean loss == single-process result, and
row-band halo exchange + GroupNorm-statistic all-reduce == un-sharded conv+GroupNorm for the three halo patterns
of the aggregation network (stride-1 conv, stride-2 conv, transposed conv)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _gn_from_sums(y, sums, groups, eps=1e-5):
    B, C = y.shape[:2]
    cpg = C // groups
    n = cpg * y[0, 0].numel()
    s = sums.view(B, groups, cpg, 2).sum(2)
    mean = s[..., 0] / n
    var = s[..., 1] / n - mean * mean
    rstd = torch.rsqrt(var + eps)
    shape = (B, groups, 1) + (1,) * (y.dim() - 2)
    yn = (y.double().view(B, groups, -1) - mean.view(B, groups, 1)) * rstd.view(B, groups, 1)
    return yn.view_as(y)


def _worker(rank, world, port, h_total):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "explicit-context-mapping-for-stereo-matching_b200"))
    from cmf_b200 import parallel as par

    torch.manual_seed(0)  # identical "full" tensors on every rank
    # ---------------- 1. gradient all-reduce (bucketed) ------------------------------------------------
    params = [torch.nn.Parameter(torch.zeros(7, 3)), torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(2, 2))]
    for i, p in enumerate(params):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    ncoll = par.allreduce_gradients(params, bucket_bytes=64)  # forces several buckets
    assert ncoll >= 2
    for i, p in enumerate(params):
        assert torch.equal(p.grad, torch.full_like(p, 3.0 * (i + 1)))

    # ---------------- 2. DP equivalence of the global masked-mean loss ---------------------------------
    net = torch.nn.Conv2d(3, 3, 3, padding=1)
    par.broadcast_parameters(net)
    x = torch.randn(4, 3, 8, 10)
    target = torch.rand(4, 8, 10) * 250 - 20  # some pixels fall outside (0, 192): uneven masks per shard
    mask = (target < 192) & (target > 0)

    def outputs(inp):
        y = net(inp)
        return [y[:, i:i + 1] * 50 for i in range(3)]

    net.zero_grad()
    par_loss = None
    full = sum(w * F.smooth_l1_loss(o.squeeze(1)[mask], target[mask]) for w, o in zip((0.5, 0.7, 1.0), outputs(x)))
    full.backward()
    want = [p.grad.clone() for p in net.parameters()]
    net.zero_grad()
    sl = slice(rank * 2, rank * 2 + 2)
    par_loss = par.masked_smooth_l1_dp(outputs(x[sl]), target[sl], 192)
    par_loss.backward()
    par.allreduce_gradients(list(net.parameters()), average=False)
    for g, p in zip(want, net.parameters()):
        torch.testing.assert_close(p.grad, g, rtol=1e-5, atol=1e-6)
    tot = par_loss.detach().clone()
    dist.all_reduce(tot)
    torch.testing.assert_close(tot, full.detach(), rtol=1e-5, atol=1e-6)

    # ---------------- 3. row bands: stride-1 conv + GroupNorm ------------------------------------------
    B, C, D, W = 1, 8, 4, 6
    xfull = torch.randn(B, C, D, h_total, W, dtype=torch.float64)
    wgt = torch.randn(16, C, 3, 3, 3, dtype=torch.float64) * 0.2
    r0, r1 = par.band_rows(h_total, world, rank)
    xb = xfull[:, :, :, r0:r1].contiguous()
    ext = par.exchange_row_halo(xb, 1, 1, dim=3)
    lo = xfull[:, :, :, r0 - 1:r0] if r0 > 0 else torch.zeros_like(xfull[:, :, :, :1])
    hi = xfull[:, :, :, r1:r1 + 1] if r1 < h_total else torch.zeros_like(xfull[:, :, :, :1])
    assert torch.equal(ext, torch.cat([lo, xb, hi], 3))
    yb = F.conv3d(ext, wgt, None, 1, 1)[:, :, :, 1:-1].contiguous()  # outer rows saw zero padding: dropped
    yfull = F.conv3d(xfull, wgt, None, 1, 1)
    torch.testing.assert_close(yb, yfull[:, :, :, r0:r1])
    sums = torch.stack([yb.sum((2, 3, 4)), (yb * yb).sum((2, 3, 4))], -1)
    par.allreduce_gn_sums(sums)
    full_sums = torch.stack([yfull.sum((2, 3, 4)), (yfull * yfull).sum((2, 3, 4))], -1)
    torch.testing.assert_close(sums, full_sums)
    # the scaled form the band forward uses (sums * band_rows / full_rows, so that gn_apply's own element count fits)
    part = torch.stack([yb.sum((2, 3, 4)), (yb * yb).sum((2, 3, 4))], -1)
    scaled = par.allreduce_gn_sums(part, scale=float(r1 - r0) / h_total)
    torch.testing.assert_close(scaled, full_sums * (float(r1 - r0) / h_total))
    # normalise with the GLOBAL statistics (n = full volume), then gather the bands
    B_, C_ = yb.shape[:2]
    n_full = (C_ // 8) * yfull[0, 0].numel()
    s = sums.view(B_, 8, C_ // 8, 2).sum(2)
    mean = s[..., 0] / n_full
    rstd = torch.rsqrt(s[..., 1] / n_full - mean * mean + 1e-5)
    yn = ((yb.view(B_, 8, C_ // 8, -1) - mean.view(B_, 8, 1, 1)) * rstd.view(B_, 8, 1, 1)).view_as(yb)
    got = par.gather_bands(yn, dim=3)
    torch.testing.assert_close(got, F.group_norm(yfull, 8, eps=1e-5), rtol=1e-9, atol=1e-9)

    # ---------------- 3b. in-place halo fill of an activation allocated WITH spare rows (no re-copy of the band)
    P = 2
    padded = torch.full((B, C, D, (r1 - r0) + 2 * P, W), float("nan"), dtype=torch.float64)
    padded[:, :, :, P:P + (r1 - r0)] = xb
    par.fill_row_halo_(padded, P, 1, 1, dim=3)
    assert torch.equal(padded[:, :, :, P - 1:P + (r1 - r0) + 1], torch.cat([lo, xb, hi], 3))
    assert torch.isnan(padded[:, :, :, 0]).all() and torch.isnan(padded[:, :, :, -1]).all()  # untouched spare rows
    par.fill_row_halo_(padded, P, 2, 2, dim=3)
    lo2 = xfull[:, :, :, r0 - 2:r0] if r0 > 0 else torch.zeros_like(xfull[:, :, :, :2])
    hi2 = xfull[:, :, :, r1:r1 + 2] if r1 < h_total else torch.zeros_like(xfull[:, :, :, :2])
    assert torch.equal(padded, torch.cat([lo2, xb, hi2], 3))

    # ---------------- 4. stride-2 conv: only a TOP halo row is needed (band edges are multiples of 4) ---
    ext = par.exchange_row_halo(xb, 1, 0, dim=3)
    w2 = torch.randn(8, C, 3, 3, 3, dtype=torch.float64) * 0.2
    y2_full = F.conv3d(xfull, w2, None, 2, 1)
    # with the explicit top row in place the band conv pads only d/w (and the bottom h edge, unused for even heights)
    y2 = F.conv3d(F.pad(ext, (1, 1, 0, 1, 1, 1)), w2, None, 2, 0)
    torch.testing.assert_close(y2, y2_full[:, :, :, r0 // 2:r1 // 2])

    # ---------------- 5. transposed conv (k3 s2 p1 op1): only a BOTTOM halo row is needed ---------------
    wt = torch.randn(C, 4, 3, 3, 3, dtype=torch.float64) * 0.2
    yt_full = F.conv_transpose3d(xfull, wt, None, stride=2, padding=1, output_padding=1)
    ext = par.exchange_row_halo(xb, 0, 1, dim=3)
    yt = F.conv_transpose3d(ext, wt, None, stride=2, padding=1, output_padding=1)[:, :, :, :2 * (r1 - r0)]
    torch.testing.assert_close(yt, yt_full[:, :, :, 2 * r0:2 * r1])
    dist.barrier()
    dist.destroy_process_group()


def test_band_rows_partition():
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "explicit-context-mapping-for-stereo-matching_b200"))
    from cmf_b200 import parallel as par

    for h, n in ((512, 8), (144, 2), (36, 4), (20, 3)):
        edges = [par.band_rows(h, n, r) for r in range(n)]
        assert edges[0][0] == 0 and edges[-1][1] == h
        for (a0, a1), (b0, b1) in zip(edges, edges[1:]):
            assert a1 == b0 and a0 % 4 == 0 and a1 > a0
    with pytest.raises(ValueError):
        par.band_rows(8, 4, 0)
    with pytest.raises(ValueError):
        par.band_rows(18, 2, 0)


@pytest.mark.timeout(300)
def test_world_size_2_gloo():
    mp.spawn(_worker, args=(2, _free_port(), 24), nprocs=2, join=True)


def _rows_conv_reference(ext, w, stride, dilation, h_offset, out_rows):
    """CPU statement of the row-window contract of cmfb200_conv*_rows_fwd (include/cmfb200.h): output row m reads input
    rows m*stride - pad + h_offset + kh*dilation of `ext`, rows outside `ext` read zero; W is padded as usual."""
    import torch.nn.functional as F

    pad = dilation
    grow = out_rows * stride + h_offset + 2 * pad  # enough zero rows below so that every window exists
    z = F.conv2d(F.pad(ext, (0, 0, 0, grow)), w, None, 1, pad, dilation)  # stride-1 responses, centre row r <-> z[r]
    rows = [m * stride + h_offset for m in range(out_rows)]
    return z[:, :, rows][:, :, :, ::stride]


@pytest.mark.parametrize("stride,dilation", [(1, 1), (1, 2), (2, 1)])
def test_row_window_contract_reproduces_the_unsharded_conv(stride, dilation):
    """The halo sizes / h_offset values used by cmfsm._cg_band and _conv2_band: a band extended by its neighbours' rows
    (zeros at the image border) and convolved under the row-window contract equals the band's rows of the whole-image
    convolution, for every band."""
    import torch.nn.functional as F

    g = torch.Generator().manual_seed(40)
    x = torch.randn(1, 3, 32, 11, generator=g)
    w = torch.randn(4, 3, 3, 3, generator=g)
    full = F.conv2d(x, w, None, stride, dilation, dilation)
    n, band = 4, 8
    top, bottom = (2, 0) if stride == 2 else (dilation, dilation)
    for r in range(n):
        lo, hi = r * band, (r + 1) * band
        t = x[:, :, lo - top:lo] if r > 0 else x.new_zeros(1, 3, top, 11)
        b = x[:, :, hi:hi + bottom] if r < n - 1 else x.new_zeros(1, 3, bottom, 11)
        ext = torch.cat([t, x[:, :, lo:hi], b], 2)
        got = _rows_conv_reference(ext, w, stride, dilation, top, band // stride)
        torch.testing.assert_close(got, full[:, :, lo // stride:hi // stride], rtol=1e-5, atol=1e-5)


def test_row_window_contract_transposed_conv():
    """Transposed conv (k3 s2 p1 op1) on a band + ONE bottom halo row, producing output rows only for the band's own
    input rows (H_compute = band): equals the band's 2x rows of the whole-volume result."""
    import torch.nn.functional as F

    g = torch.Generator().manual_seed(41)
    x = torch.randn(1, 3, 16, 7, generator=g)
    w = torch.randn(3, 2, 3, 3, generator=g)
    full = F.conv_transpose2d(x, w, None, 2, 1, 1)
    n, band = 4, 4
    for r in range(n):
        lo, hi = r * band, (r + 1) * band
        halo = x[:, :, hi:hi + 1] if r < n - 1 else x.new_zeros(1, 3, 1, 7)
        ext = torch.cat([x[:, :, lo:hi], halo], 2)
        got = F.conv_transpose2d(ext, w, None, 2, 1, 1)[:, :, :2 * band]  # rows of the band's own inputs only
        torch.testing.assert_close(got, full[:, :, 2 * lo:2 * hi], rtol=1e-5, atol=1e-5)
