"""Torch-facing wrappers of the libcmfb200 kernels (device memory, streams: PyTorch; compute: ours).

Every function takes/returns CUDA fp32 tensors, enqueues on the current stream of the tensor's device and
never synchronises.  Nothing here has a CPU or eager-PyTorch fallback: a CPU tensor or a missing
library raises.
"""
import ctypes

import threading

import torch

from . import lib as _lib

GN_GROUPS = 32  # group_norm_group_num, cmf/models/cmfsm.py:33
GN_EPS = 1e-5


# ---- optional per-kernel CUDA-event timing (bench.py / profiling only; off by default) ------------
_TIMING = None  # None, or a list receiving (kernel name, start event, end event)


def enable_event_timing(on=True):
    """When on, every kernel wrapper records a CUDA event pair on the launching stream."""
    global _TIMING
    _TIMING = [] if on else None


def drain_event_timing():
    """Synchronise and return {name: (launches, total ms)}; clears the record."""
    global _TIMING
    out = {}
    if _TIMING:
        torch.cuda.synchronize()
        for name, a, b in _TIMING:
            n, ms = out.get(name, (0, 0.0))
            out[name] = (n + 1, ms + a.elapsed_time(b))
        _TIMING = []
    return out


class _timed:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        self.on = _TIMING is not None and not torch.cuda.is_current_stream_capturing()
        if self.on:
            self.a = torch.cuda.Event(enable_timing=True)
            self.a.record()

    def __exit__(self, *exc):
        if self.on:
            b = torch.cuda.Event(enable_timing=True)
            b.record()
            _TIMING.append((self.name, self.a, b))
        return False


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _req(*tensors, dtype=torch.float32):
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise _lib.CmfB200Error("cmf_b200 kernels need CUDA tensors (no CPU fallback); got device %s" % t.device)
        if t.dtype != dtype:
            raise _lib.CmfB200Error("cmf_b200 kernel expected dtype %s, got %s" % (dtype, t.dtype))
        if not t.is_contiguous():
            raise _lib.CmfB200Error("cmf_b200 kernels need contiguous tensors")


# ------------------------------------------------------------------------------------------ K1
def cost_volume_concat(L, R, D):
    """[B,C,h,w] x2 -> [B,2C,D,h,w]; cmf/models/cmfsm.py:667-682."""
    _req(L, R)
    if L.shape != R.shape:
        raise ValueError("left/right feature shapes differ: %s vs %s" % (tuple(L.shape), tuple(R.shape)))
    B, C, h, w = L.shape
    cost = torch.empty((B, 2 * C, D, h, w), device=L.device, dtype=torch.float32)
    with torch.cuda.device(L.device), _timed("cost_volume_concat_fwd"):
        _lib.check(_lib.load().cmfb200_cost_volume_concat_fwd(_p(L), _p(R), _p(cost), B, C, h, w, D, _stream()),
                   "cost_volume_concat_fwd")
    return cost


def pitched_2d(t):
    """(height, width, pitch) in elements when `t` is `height` contiguous blocks of `width` elements `pitch` apart (a dense
    tensor, or rows narrowed out of a contiguous one); None for any other striding."""
    dims = [(n, st) for n, st in zip(t.shape, t.stride()) if n != 1]
    width, i = 1, len(dims) - 1
    while i >= 0 and dims[i][1] == width:
        width *= dims[i][0]
        i -= 1
    if i < 0:
        return 1, width, width
    pitch, height = dims[i][1], dims[i][0]
    expect = pitch * height
    for n, st in reversed(dims[:i]):
        if st != expect:
            return None
        height *= n
        expect *= n
    return height, width, pitch


def copy_rows(dst, src):
    """dst.copy_(src) for the boundary rows of band activations (row-band halo exchange): both sides are blocks of
    contiguous rows at a fixed pitch (or dense), moved 16 bytes per thread by `cmfb200_copy_2d`; either side may live in a
    peer GPU's mailbox.  Any other layout, dtype pair or alignment goes through ATen's copy_."""
    if dst.is_cuda and src.is_cuda and dst.dtype == src.dtype and dst.shape == src.shape and dst.numel() > 0:
        a, b = pitched_2d(dst), pitched_2d(src)
        if a is not None and b is not None:
            es = dst.element_size()
            if a[0] == 1 and b[0] > 1:      # dense side: cut it into the other side's blocks
                a = (b[0], b[1], b[1])
            elif b[0] == 1 and a[0] > 1:
                b = (a[0], a[1], a[1])
            ok = a[0] == b[0] and a[1] == b[1]
            vals = (dst.data_ptr(), src.data_ptr(), a[2] * es, b[2] * es, a[1] * es)
            if ok and all(v % 16 == 0 for v in vals):
                with torch.cuda.device(dst.device), _timed("copy_2d"):
                    _lib.check(_lib.load().cmfb200_copy_2d(_p(dst), a[2] * es, _p(src), b[2] * es, a[1] * es, a[0], _stream()),
                               "copy_2d")
                return dst
    return dst.copy_(src)


def sum_peers(dst, peers, scale=1.0):
    """dst = scale * (peers[0] + peers[1] + ...) over double tensors of dst's shape (the ranks' mailbox views of the band
    GroupNorm sums), summed in list order by ONE kernel."""
    n = dst.numel()
    for t in peers:
        if t.dtype != torch.float64 or t.numel() != n or not t.is_contiguous() or not t.is_cuda:
            raise _lib.CmfB200Error("sum_peers: every peer buffer must be a contiguous CUDA double tensor of dst's size")
    if dst.dtype != torch.float64 or not dst.is_contiguous():
        raise _lib.CmfB200Error("sum_peers: dst must be a contiguous double tensor")
    ptrs = (ctypes.c_void_p * len(peers))(*[t.data_ptr() for t in peers])
    with torch.cuda.device(dst.device), _timed("sum_peers"):
        _lib.check(_lib.load().cmfb200_sum_peers_f64(_p(dst), ptrs, len(peers), n, float(scale), _stream()), "sum_peers_f64")
    return dst


def cost_volume_corr(L, R, D, normalize=False):
    """Correlation cost volume [B,D,h,w]: mean over channels of L[x] * R[x-d] (cosine similarity with `normalize`)."""
    _req(L, R)
    if L.shape != R.shape:
        raise ValueError("left/right feature shapes differ: %s vs %s" % (tuple(L.shape), tuple(R.shape)))
    B, C, h, w = L.shape
    out = torch.empty((B, D, h, w), device=L.device, dtype=torch.float32)
    with torch.cuda.device(L.device), _timed("cost_volume_corr_fwd"):
        _lib.check(_lib.load().cmfb200_cost_volume_corr_fwd(_p(L), _p(R), _p(out), B, C, h, w, D, int(normalize), _stream()),
                   "cost_volume_corr_fwd")
    return out


def masked_smooth_l1_sums(o1, o2, o3, disp, maxdisp):
    """One pass over the three outputs and the target (train.py:162-174): returns a [4] double tensor
    [sum smooth_l1(o1-d), sum smooth_l1(o2-d), sum smooth_l1(o3-d), count] over valid pixels 0 < d < maxdisp."""
    _req(o1, o2, o3, disp)
    n = disp.numel()
    if not (o1.numel() == o2.numel() == o3.numel() == n):
        raise ValueError("masked_smooth_l1: outputs and target differ in size")
    sums = torch.zeros(4, device=disp.device, dtype=torch.float64)
    with torch.cuda.device(disp.device), _timed("masked_smooth_l1_fwd"):
        _lib.check(_lib.load().cmfb200_masked_smooth_l1_fwd(_p(o1), _p(o2), _p(o3), _p(disp), _p(sums), n, float(maxdisp),
                                                            _stream()), "masked_smooth_l1_fwd")
    return sums


def masked_smooth_l1_grads(o1, o2, o3, disp, scale3, maxdisp):
    """Gradients of the three outputs: scale3[i] * clamp(o_i - d, -1, 1) on valid pixels (scale3: 3 floats on device)."""
    _req(o1, o2, o3, disp, scale3)
    gs = [torch.empty_like(o) for o in (o1, o2, o3)]
    with torch.cuda.device(disp.device), _timed("masked_smooth_l1_bwd"):
        _lib.check(_lib.load().cmfb200_masked_smooth_l1_bwd(_p(o1), _p(o2), _p(o3), _p(disp), _p(scale3), _p(gs[0]),
                                                            _p(gs[1]), _p(gs[2]), disp.numel(), float(maxdisp), _stream()),
                   "masked_smooth_l1_bwd")
    return gs


def cost_volume_concat_bwd(g, C):
    _req(g)
    B, C2, D, h, w = g.shape
    dL = torch.empty((B, C, h, w), device=g.device, dtype=torch.float32)
    dR = torch.empty_like(dL)
    with torch.cuda.device(g.device):
        _lib.check(_lib.load().cmfb200_cost_volume_concat_bwd(_p(g), _p(dL), _p(dR), B, C, h, w, D, _stream()),
                   "cost_volume_concat_bwd")
    return dL, dR


# ------------------------------------------------------------------------------------------ K2 / K3
class _SumsPool:
    """One zeroed fp64 buffer per forward pass from which the GroupNorm-statistics buffers [B,C,2] are sliced, instead
    of one `torch.zeros` (= one fill launch) per conv layer (~190 launches per forward)."""

    def __init__(self):
        self.buf, self.used = None, 0

    def take(self, B, C, device):
        n = B * C * 2
        if self.buf is None or self.buf.device != device or self.used + n > self.buf.numel():
            self.buf = torch.zeros(max(1 << 16, 4 * n), device=device, dtype=torch.float64)
            self.used = 0
        out = self.buf[self.used:self.used + n].view(B, C, 2)
        self.used += n
        return out


_POOL = threading.local()


class sums_pool:
    """Context manager: inside it `_new_sums` slices from one pooled zero buffer (re-created on every entry, so the
    statistics of different forward passes never alias).  Without it every call allocates its own zeros."""

    def __enter__(self):
        self.prev = getattr(_POOL, "pool", None)
        _POOL.pool = _SumsPool()
        return self

    def __exit__(self, *exc):
        _POOL.pool = self.prev
        return False


def _new_sums(B, C, device):
    pool = getattr(_POOL, "pool", None)
    if pool is not None:
        return pool.take(B, C, device)
    return torch.zeros((B, C, 2), device=device, dtype=torch.float64)


def pack_conv3d_weight(weight, transposed=False):
    """nn.Conv3d [Cout,Cin,3,3,3] (or nn.ConvTranspose3d [Cin,Cout,3,3,3]) -> packed [Cin,27,Cout]."""
    weight = weight.detach()
    _req(weight)
    if transposed:
        Cin, Cout = weight.shape[:2]
    else:
        Cout, Cin = weight.shape[:2]
    if tuple(weight.shape[2:]) != (3, 3, 3):
        raise ValueError("expected a 3x3x3 kernel, got %s" % (tuple(weight.shape),))
    packed = torch.empty((Cin, 27, Cout), device=weight.device, dtype=torch.float32)
    with torch.cuda.device(weight.device):
        _lib.check(_lib.load().cmfb200_pack_conv3d_weight(_p(weight), _p(packed), Cout, Cin, int(transposed), _stream()),
                   "pack_conv3d_weight")
    return packed


def conv3d_k3(x, packed, stride=1, transposed=False, want_stats=False):
    """3x3x3 conv (pad 1) or transposed conv (s2,p1,op1).  Returns (y, gn_sums or None)."""
    _req(x, packed)
    B, Cin, D, H, W = x.shape
    if packed.shape[0] != Cin:
        raise ValueError("weight Cin %d != input channels %d" % (packed.shape[0], Cin))
    Cout = packed.shape[2]
    sums = _new_sums(B, Cout, x.device) if want_stats else None
    L = _lib.load()
    if transposed:
        y = torch.empty((B, Cout, 2 * D, 2 * H, 2 * W), device=x.device, dtype=torch.float32)
    else:
        Do, Ho, Wo = (D - 1) // stride + 1, (H - 1) // stride + 1, (W - 1) // stride + 1
        y = torch.empty((B, Cout, Do, Ho, Wo), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device), _timed("deconv3d_k3s2_fwd" if transposed else "conv3d_k3_fwd"):
        if transposed:
            _lib.check(L.cmfb200_deconv3d_k3s2_fwd(_p(x), _p(packed), _p(y), _p(sums), B, Cin, Cout, D, H, W, _stream()),
                       "deconv3d_k3s2_fwd")
        else:
            _lib.check(L.cmfb200_conv3d_k3_fwd(_p(x), _p(packed), _p(y), _p(sums), B, Cin, Cout, D, H, W, stride,
                                               _stream()), "conv3d_k3_fwd")
    return y, sums


def conv3d_k3_rows(x, packed, out_rows, stride=1, transposed=False, row_offset=0, want_stats=True):
    """Row-window conv for row-band sharding: `x` [B,Cin,D,H_in,W] carries halo rows; only the band's rows are produced.
    conv: output row m reads input rows m*stride - 1 + row_offset + kh, `out_rows` output rows;
    transposed: `out_rows` = number of INPUT rows that produce output (the last row of x may be a halo row)."""
    _req(x, packed)
    B, Cin, D, H, W = x.shape
    Cout = packed.shape[2]
    sums = _new_sums(B, Cout, x.device) if want_stats else None
    L = _lib.load()
    if transposed:
        y = torch.empty((B, Cout, 2 * D, 2 * out_rows, 2 * W), device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device), _timed("deconv3d_k3s2_fwd"):
            _lib.check(L.cmfb200_deconv3d_k3s2_rows_fwd(_p(x), _p(packed), _p(y), _p(sums), B, Cin, Cout, D, H, out_rows, W,
                                                        _stream()), "deconv3d_k3s2_rows_fwd")
    else:
        Do, Wo = (D - 1) // stride + 1, (W - 1) // stride + 1
        y = torch.empty((B, Cout, Do, out_rows, Wo), device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device), _timed("conv3d_k3_fwd"):
            _lib.check(L.cmfb200_conv3d_k3_rows_fwd(_p(x), _p(packed), _p(y), _p(sums), B, Cin, Cout, D, H, W, stride,
                                                    row_offset, out_rows, _stream()), "conv3d_k3_rows_fwd")
    return y, sums


def pack_conv2d_weight(weight):
    """nn.Conv2d [Cout,Cin,k,k] -> packed [Cin,k*k,Cout]."""
    weight = weight.detach()
    _req(weight)
    Cout, Cin, k, k2 = weight.shape
    if k != k2 or k not in (1, 3):
        raise ValueError("expected a 1x1 or 3x3 kernel, got %s" % (tuple(weight.shape),))
    packed = torch.empty((Cin, k * k, Cout), device=weight.device, dtype=torch.float32)
    with torch.cuda.device(weight.device):
        _lib.check(_lib.load().cmfb200_pack_conv2d_weight(_p(weight), _p(packed), Cout, Cin, k, _stream()),
                   "pack_conv2d_weight")
    return packed


def conv2d(x, packed, ksize, stride=1, dilation=1, want_stats=False):
    """2-D conv, padding (k//2)*dilation (reference convbn(), cmf/models/cmfsm.py:36-46).  Returns (y, gn_sums)."""
    _req(x, packed)
    B, Cin, H, W = x.shape
    if packed.shape[0] != Cin or packed.shape[1] != ksize * ksize:
        raise ValueError("packed weight %s does not match Cin=%d k=%d" % (tuple(packed.shape), Cin, ksize))
    Cout = packed.shape[2]
    pad = (ksize // 2) * dilation
    Ho = (H + 2 * pad - (ksize - 1) * dilation - 1) // stride + 1
    Wo = (W + 2 * pad - (ksize - 1) * dilation - 1) // stride + 1
    sums = _new_sums(B, Cout, x.device) if want_stats else None
    y = torch.empty((B, Cout, Ho, Wo), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device), _timed("conv2d_fwd"):
        _lib.check(_lib.load().cmfb200_conv2d_fwd(_p(x), _p(packed), _p(y), _p(sums), B, Cin, Cout, H, W, ksize, stride,
                                                  dilation, _stream()), "conv2d_fwd")
    return y, sums


def conv2d_rows(x, packed, ksize, out_rows, stride=1, dilation=1, row_offset=0, want_stats=True):
    """Row-window 2-D conv for row-band sharding: `x` carries halo rows, output row m reads input rows
    m*stride - pad + row_offset + kh*dilation; `out_rows` rows are produced (and only they enter the statistics)."""
    _req(x, packed)
    B, Cin, H, W = x.shape
    Cout = packed.shape[2]
    pad = (ksize // 2) * dilation
    Wo = (W + 2 * pad - (ksize - 1) * dilation - 1) // stride + 1
    sums = _new_sums(B, Cout, x.device) if want_stats else None
    y = torch.empty((B, Cout, out_rows, Wo), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device), _timed("conv2d_fwd"):
        _lib.check(_lib.load().cmfb200_conv2d_rows_fwd(_p(x), _p(packed), _p(y), _p(sums), B, Cin, Cout, H, W, ksize, stride,
                                                       dilation, row_offset, out_rows, _stream()), "conv2d_rows_fwd")
    return y, sums


_TC3_SHAPES = {(3, 1): (32, 64, 128), (3, 2): (128,), (1, 1): (32, 128)}  # (k, dilation) -> Cout of conv_tc3


def tc3_supported(cin, cout, ksize, dilation):
    return cin % 16 == 0 and cout in _TC3_SHAPES.get((ksize, dilation if ksize == 3 else 1), ())


def conv_s1_nchw(x, weight, dilation=1, want_stats=False, engine="tc3"):
    """Stride-1 'same' conv of an NCHW / NCDHW fp32 tensor with an nn.Conv2d / nn.Conv3d weight, result NCHW fp32
    (+ GroupNorm sums): on the tensor cores (three-term split, conv_tc3) when the shape is one it has, else FFMA.
    The training path uses this for forward and dgrad; activations stay plain fp32 tensors in the autograd graph."""
    weight = weight.detach()
    cout, cin, k = weight.shape[0], weight.shape[1], weight.shape[-1]
    # the tiny SPP-branch maps (1x2 .. 8x16 pooled values per channel) stay on the FFMA kernel: their GroupNorm
    # statistics need double-from-the-first-element sums (conv_common.cuh), and there is no work to speed up
    big = x[0, 0].numel() >= 4096
    if engine == "tc3" and big and tc3_supported(cin, cout, k, dilation):
        return conv_tc3(f32_to_c8s3(x.contiguous()), pack_tc3_weight(weight.contiguous()), dilation, want_stats, out_nchw=True)
    if x.dim() == 5:
        return conv3d_k3(x.contiguous(), pack_conv3d_weight(weight.contiguous()), 1, want_stats=want_stats)
    return conv2d(x.contiguous(), pack_conv2d_weight(weight.contiguous()), k, 1, dilation, want_stats)


def conv_wgrad(x, dy, ksize, stride=1, dilation=1):
    """Weight gradient of a 2-D ([B,C,H,W]) or 3-D ([B,C,D,H,W], 3x3x3) convolution with 'same' padding:
    returns dW [Cout, Cin, (3,) k, k].  For a transposed conv pass (x := grad of its output, dy := its input, stride 2)."""
    x, dy = x.contiguous(), dy.contiguous()
    _req(x, dy)
    three_d = x.dim() == 5
    B, Cin = x.shape[:2]
    Cout = dy.shape[1]
    D, H, W = (tuple(x.shape[2:]) if three_d else (1,) + tuple(x.shape[2:]))
    KD = 3 if three_d else 1
    want = tuple((v - 1) // stride + 1 for v in x.shape[2:])
    if tuple(dy.shape[2:]) != want or dy.shape[0] != B:
        raise ValueError("conv_wgrad: dy %s does not match x %s at stride %d" % (tuple(dy.shape), tuple(x.shape), stride))
    dw = torch.zeros((Cout, Cin) + ((3,) if three_d else ()) + (ksize, ksize), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device), _timed("conv_wgrad"):
        _lib.check(_lib.load().cmfb200_conv_wgrad(_p(x), _p(dy), _p(dw), B, Cin, Cout, D, H, W, KD, ksize, stride, dilation,
                                                  _stream()), "conv_wgrad")
    return dw


def conv3d_dgrad(dy, weight, stride=1, transposed=False, engine="tc3"):
    """Input gradient of the 3x3x3 conv / transposed conv layers, on the forward kernels with re-arranged weights:
    stride 1 -> the same conv with flipped taps and swapped channel roles (tensor cores when `engine` is tc3);
    stride 2 -> the transposed conv kernel; transposed -> the stride-2 conv kernel."""
    weight = weight.detach()
    if transposed:  # dx = conv3d(dy, Wt, stride 2, pad 1): Wt [Cin_t, Cout_t] is already [out, in] of that conv
        return conv3d_k3(dy.contiguous(), pack_conv3d_weight(weight.contiguous()), 2)[0]
    if stride == 2:  # dx = conv_transpose3d(dy, W, s2, p1, op1): W [Cout, Cin] is its [in, out] layout
        return conv3d_k3(dy.contiguous(), pack_conv3d_weight(weight.contiguous(), transposed=True), transposed=True)[0]
    wd = weight.transpose(0, 1).flip(2, 3, 4).contiguous()
    return conv_s1_nchw(dy, wd, 1, False, engine)[0]


def conv2d_dgrad(dy, weight, x_shape, stride=1, dilation=1, engine="tc3"):
    """Input gradient of the 2-D conv layers (3x3 / 1x1, stride 1 with any dilation, stride 2) on the forward kernels."""
    weight = weight.detach()
    k = weight.shape[-1]
    dy = dy.contiguous()
    if stride == 1:
        wd = weight.transpose(0, 1).flip(2, 3).contiguous()
        return conv_s1_nchw(dy, wd, dilation, False, engine)[0]
    if k == 1:  # 1x1 stride 2: the gradient lands on the even pixels
        t = conv2d(dy, pack_conv2d_weight(weight.transpose(0, 1).contiguous()), 1, 1, 1)[0]
        dx = torch.zeros(x_shape, device=dy.device, dtype=torch.float32)
        dx[:, :, ::2, ::2] = t
        return dx
    # 3x3 stride 2: a transposed conv = the 3-D transposed-conv kernel on a depth-1 volume, weights in the kd = 1 slice
    w3 = torch.zeros(tuple(weight.shape[:2]) + (3, 3, 3), device=dy.device, dtype=torch.float32)
    w3[:, :, 1] = weight
    dx = conv3d_k3(dy.unsqueeze(2).contiguous(), pack_conv3d_weight(w3, transposed=True), transposed=True)[0]
    return dx[:, :, 0].contiguous()


def spp_pool(x):
    """AvgPool2d 64/32/16/8 (stride = kernel) of [B,C,H,W]; returns (p64, p32, p16, p8) = branch1..branch4 inputs."""
    _req(x)
    B, C, H, W = x.shape
    ps = [torch.empty((B, C, H // k, W // k), device=x.device, dtype=torch.float32) for k in (8, 16, 32, 64)]
    with torch.cuda.device(x.device), _timed("spp_pool_fwd"):
        _lib.check(_lib.load().cmfb200_spp_pool_fwd(_p(x), _p(ps[0]), _p(ps[1]), _p(ps[2]), _p(ps[3]), B, C, H, W,
                                                    _stream()), "spp_pool_fwd")
    return ps[3], ps[2], ps[1], ps[0]


def spp_upsample_concat(raw, skip, b4, b3, b2, b1, full_rows=None, row_offset=0):
    """cat([raw, skip, up(b4), up(b3), up(b2), up(b1)], 1) with bilinear (align_corners=False) upsampling.
    Row bands: raw/skip hold rows [row_offset, row_offset+H) of an image with `full_rows` rows; b* cover all of it."""
    _req(raw, skip, b4, b3, b2, b1)
    B, _, H, W = skip.shape
    if raw.shape[1] != 64 or skip.shape[1] != 128 or any(t.shape[1] != 32 for t in (b4, b3, b2, b1)):
        raise ValueError("spp_upsample_concat expects 64 + 128 + 4 x 32 channels")
    cat = torch.empty((B, 320, H, W), device=skip.device, dtype=torch.float32)
    with torch.cuda.device(skip.device), _timed("spp_upsample_concat_fwd"):
        _lib.check(_lib.load().cmfb200_spp_upsample_concat_fwd(_p(raw), _p(skip), _p(b4), _p(b3), _p(b2), _p(b1), _p(cat),
                                                               B, H, W, H if full_rows is None else full_rows,
                                                               row_offset, _stream()), "spp_upsample_concat_fwd")
    return cat


def gn_stats(x):
    _req(x)
    B, C = x.shape[:2]
    spatial = x[0, 0].numel()
    sums = _new_sums(B, C, x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().cmfb200_gn_stats(_p(x), _p(sums), B, C, spatial, _stream()), "gn_stats")
    return sums


def gn_apply(x, sums, gamma, beta, residual=None, relu=False, out=None, groups=GN_GROUPS, eps=GN_EPS):
    """y = GroupNorm(x) (+residual) (ReLU).  `out` may be x itself (in place)."""
    gamma, beta = gamma.detach(), beta.detach()
    _req(x, gamma, beta, residual)
    B, C = x.shape[:2]
    spatial = x[0, 0].numel()
    if residual is not None and residual.shape != x.shape:
        raise ValueError("residual shape %s != %s" % (tuple(residual.shape), tuple(x.shape)))
    y = torch.empty_like(x) if out is None else out
    with torch.cuda.device(x.device), _timed("gn_apply"):
        _lib.check(_lib.load().cmfb200_gn_apply(_p(x), _p(sums), _p(gamma), _p(beta), _p(residual), _p(y), B, C, groups,
                                                spatial, eps, int(relu), _stream()), "gn_apply")
    return y


def conv3d_gn(x, packed, gamma, beta, stride=1, transposed=False, residual=None, relu=False):
    """convbn_3d (+residual) (+ReLU): conv with fused statistics, then one normalise pass in place."""
    y, sums = conv3d_k3(x, packed, stride, transposed, want_stats=True)
    return gn_apply(y, sums, gamma, beta, residual, relu, out=y)


# ------------------------------------------------------------------------------------------ K5 / K4
def conv3d_cout1_backward(x, weight, grad_y):
    """Backward of nn.Conv3d(32, 1, 3, padding=1, bias=False): returns (dx, dw)."""
    weight = weight.detach().contiguous()
    grad_y = grad_y.contiguous()
    _req(x, weight, grad_y)
    B, Cin, D, H, W = x.shape
    if tuple(weight.shape) != (1, Cin, 3, 3, 3) or grad_y.numel() != B * D * H * W:
        raise ValueError("conv3d_cout1_backward: weight %s / grad %s do not match x %s"
                         % (tuple(weight.shape), tuple(grad_y.shape), tuple(x.shape)))
    dx, dw = torch.empty_like(x), torch.empty_like(weight)
    with torch.cuda.device(x.device), _timed("conv3d_cout1_bwd"):
        _lib.check(_lib.load().cmfb200_conv3d_cout1_bwd(_p(x), _p(weight), _p(grad_y), _p(dx), _p(dw), B, Cin, D, H, W,
                                                        _stream()), "conv3d_cout1_bwd")
    return dx, dw


def gn_backward(grad_out, x, sums, gamma, out=None, want_dres=False, groups=GN_GROUPS, eps=GN_EPS):
    """Backward of `gn_apply`: (dx, d_gamma, d_beta, d_residual or None).  `out` = the forward output when the forward
    applied a ReLU (mask), else None."""
    gamma = gamma.detach()
    grad_out = grad_out.contiguous()
    _req(grad_out, x, gamma, out)
    B, C = x.shape[:2]
    spatial = x[0, 0].numel()
    dx = torch.empty_like(x)
    dres = torch.empty_like(x) if want_dres else None
    bsums = torch.empty((B, C, 2), device=x.device, dtype=torch.float64)
    with torch.cuda.device(x.device), _timed("gn_bwd"):
        _lib.check(_lib.load().cmfb200_gn_bwd(_p(grad_out), _p(x), _p(out), _p(sums), _p(gamma), _p(bsums), _p(dx),
                                              _p(dres), B, C, groups, spatial, eps, _stream()), "gn_bwd")
    tot = bsums.sum(0)
    return dx, tot[:, 1].float(), tot[:, 0].float(), dres


def ctxmap_weights(lr, hr, w0, w1, w2, w3, valid_rows=None):
    """eight_related_context_mapping: [B,32,h,w],[B,32,H,W] -> [B,9,H,W] (cmf/models/cmfsm.py:443-593).
    `valid_rows` = (y0, y1): low-res rows of `lr` that lie inside the image (row bands pass halo rows)."""
    ws = [w.detach().reshape(w.shape[0], w.shape[1]).contiguous() for w in (w0, w1, w2, w3)]
    _req(lr, hr, *ws)
    B, C, h, w = lr.shape
    H, W = hr.shape[2:]
    if C != 32 or hr.shape[1] != 32 or tuple(ws[0].shape) != (32, 66):
        raise ValueError("context mapping expects 32-channel features and a 66-input similarity MLP")
    scale = W // w
    if H != h * scale or W != w * scale:
        raise ValueError("hr %dx%d is not an integer multiple of lr %dx%d" % (H, W, h, w))
    out = torch.empty((B, 9, H, W), device=lr.device, dtype=torch.float32)
    vy0, vy1 = (0, h) if valid_rows is None else valid_rows
    with torch.cuda.device(lr.device), _timed("ctxmap_weights_fwd"):
        _lib.check(_lib.load().cmfb200_ctxmap_weights_fwd(_p(lr), _p(hr), _p(ws[0]), _p(ws[1]), _p(ws[2]), _p(ws[3]),
                                                          _p(out), B, h, w, scale, vy0, vy1, _stream()), "ctxmap_weights_fwd")
    return out


def ctxmap_weights_bwd(lr, hr, w0, w1, w2, w3, weights9, grad):
    """Backward of `ctxmap_weights`: returns (d_lr, d_hr, d_w0, d_w1, d_w2, d_w3) with the weights' own shapes.
    The kernel back-propagates the MLP per (pixel, neighbour); the two linear 1x1 maps of layer 0 are finished here
    as plain fp32 GEMMs (torch.matmul)."""
    shapes = [tuple(t.shape) for t in (w0, w1, w2, w3)]
    ws = [t.detach().reshape(t.shape[0], t.shape[1]).contiguous() for t in (w0, w1, w2, w3)]
    grad = grad.contiguous()
    _req(lr, hr, weights9, grad, *ws)
    B, _, h, w = lr.shape
    H, W = hr.shape[2:]
    d_ahr = torch.empty((B, 32, H, W), device=lr.device, dtype=torch.float32)
    d_alr = torch.empty((B, 32, h, w), device=lr.device, dtype=torch.float32)
    wbuf = torch.empty(712, device=lr.device, dtype=torch.float32)
    with torch.cuda.device(lr.device), _timed("ctxmap_weights_bwd"):
        _lib.check(_lib.load().cmfb200_ctxmap_weights_bwd(_p(lr), _p(hr), _p(ws[0]), _p(ws[1]), _p(ws[2]), _p(ws[3]),
                                                          _p(weights9), _p(grad), _p(d_ahr), _p(d_alr), _p(wbuf), B, h, w,
                                                          W // w, _stream()), "ctxmap_weights_bwd")
    w0m = ws[0]
    ga, gl = d_ahr.view(B, 32, H * W), d_alr.view(B, 32, h * w)
    d_hr = torch.matmul(w0m[:, 32:64].t(), ga).view(B, 32, H, W)
    d_lr = torch.matmul(w0m[:, :32].t(), gl).view(B, 32, h, w)
    d_w0 = torch.empty((32, 66), device=lr.device, dtype=torch.float32)
    d_w0[:, :32] = torch.bmm(gl, lr.reshape(B, 32, h * w).transpose(1, 2)).sum(0)
    d_w0[:, 32:64] = torch.bmm(ga, hr.reshape(B, 32, H * W).transpose(1, 2)).sum(0)
    d_w0[:, 64:] = wbuf[648:712].view(32, 2)
    return (d_lr, d_hr, d_w0.view(shapes[0]), wbuf[:512].view(shapes[1]), wbuf[512:640].view(shapes[2]),
            wbuf[640:648].view(shapes[3]))


def softargmin_ctxmap(c1, c2, c3, weights9, scale, want_lowres=False):
    """cmf/models/cmfsm.py:703-769.  c_i [B,D,h,w], weights9 [B,9,H,W] -> 3 x [B,1,H,W] (+ [3,B,h,w])."""
    _req(c1, c2, c3, weights9)
    B, D, h, w = c1.shape
    H, W = h * scale, w * scale
    if tuple(weights9.shape) != (B, 9, H, W):
        raise ValueError("weights9 shape %s != %s" % (tuple(weights9.shape), (B, 9, H, W)))
    outs = [torch.empty((B, 1, H, W), device=c1.device, dtype=torch.float32) for _ in range(3)]
    low = torch.empty((3, B, h, w), device=c1.device, dtype=torch.float32) if want_lowres else None
    with torch.cuda.device(c1.device), _timed("softargmin_ctxmap_fwd"):
        _lib.check(_lib.load().cmfb200_softargmin_ctxmap_fwd(_p(c1), _p(c2), _p(c3), _p(weights9), _p(outs[0]),
                                                             _p(outs[1]), _p(outs[2]), _p(low), B, D, h, w, scale,
                                                             _stream()), "softargmin_ctxmap_fwd")
    return (tuple(outs), low) if want_lowres else tuple(outs)


# ------------------------------------------------------------------------------------------ bf16 / C8 path
# C8 layout: bf16 [B, C/8, D, H, W, 8] (include/cmfb200.h); tensors below carry that 6-D shape.
BF16 = torch.bfloat16


def pack_igemm_weight(weight, transposed=False):
    """nn.Conv3d weight (fp32) -> bf16 [27, Cin/8, Cout, 8] for the tcgen05 implicit GEMM."""
    weight = weight.detach()
    _req(weight)
    Cin, Cout = (weight.shape[0], weight.shape[1]) if transposed else (weight.shape[1], weight.shape[0])
    packed = torch.empty((27, Cin // 8, Cout, 8), device=weight.device, dtype=BF16)
    with torch.cuda.device(weight.device):
        _lib.check(_lib.load().cmfb200_pack_igemm_weight_bf16(_p(weight), _p(packed), Cout, Cin, int(transposed),
                                                              _stream()), "pack_igemm_weight_bf16")
    return packed


def f32_to_c8(x):
    """[B,C,D,H,W] fp32 -> C8 bf16."""
    _req(x)
    B, C = x.shape[:2]
    sp = tuple(x.shape[2:])
    y = torch.empty((B, C // 8) + sp + (8,), device=x.device, dtype=BF16)
    with torch.cuda.device(x.device), _timed("f32_to_c8_bf16"):
        _lib.check(_lib.load().cmfb200_f32_to_c8_bf16(_p(x), _p(y), B, C, x[0, 0].numel(), _stream()), "f32_to_c8_bf16")
    return y


def c8_to_f32(x):
    """C8 bf16 -> [B,C,D,H,W] fp32."""
    _req(x, dtype=BF16)
    B, NC = x.shape[:2]
    sp = tuple(x.shape[2:-1])
    y = torch.empty((B, NC * 8) + sp, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device), _timed("c8_bf16_to_f32"):
        _lib.check(_lib.load().cmfb200_c8_bf16_to_f32(_p(x), _p(y), B, NC * 8, y[0, 0].numel(), _stream()),
                   "c8_bf16_to_f32")
    return y


def cost_volume_concat_c8(L, R, D):
    """K1 in C8/bf16: [B,C,h,w] fp32 x2 -> [B, 2C/8, D, h, w, 8] bf16."""
    _req(L, R)
    B, C, h, w = L.shape
    cost = torch.empty((B, 2 * C // 8, D, h, w, 8), device=L.device, dtype=BF16)
    with torch.cuda.device(L.device), _timed("cost_volume_concat_c8_bf16"):
        _lib.check(_lib.load().cmfb200_cost_volume_concat_c8_bf16(_p(L), _p(R), _p(cost), B, C, h, w, D, _stream()),
                   "cost_volume_concat_c8_bf16")
    return cost


def conv3d_igemm(x, packed, want_stats=True):
    """tcgen05 implicit-GEMM 3x3x3 s1 conv on C8/bf16.  Returns (raw y C8 bf16, gn_sums or None)."""
    _req(x, packed, dtype=BF16)
    B, NC, D, H, W, _ = x.shape
    Cin, Cout = NC * 8, packed.shape[2]
    if packed.shape[1] != NC:
        raise ValueError("weight Cin %d != input channels %d" % (packed.shape[1] * 8, Cin))
    y = torch.empty((B, Cout // 8, D, H, W, 8), device=x.device, dtype=BF16)
    sums = _new_sums(B, Cout, x.device) if want_stats else None
    with torch.cuda.device(x.device), _timed("conv3d_igemm_bf16_fwd"):
        _lib.check(_lib.load().cmfb200_conv3d_igemm_bf16_fwd(_p(x), _p(packed), _p(y), _p(sums), B, Cin, Cout, D, H, W,
                                                             _stream()), "conv3d_igemm_bf16_fwd")
    return y, sums


def conv3d_igemm_cout1(x, packed32):
    """classifN.2 on tensor cores: C8/bf16 [B,4,D,H,W,8] x pack_igemm_weight(weight zero-padded to 32 couts)
    -> fp32 [B,D,H,W]."""
    _req(x, packed32, dtype=BF16)
    B, NC, D, H, W, _ = x.shape
    if tuple(packed32.shape) != (27, NC, 32, 8) or NC != 4:
        raise ValueError("expected packed weights [27,4,32,8] and a 32-channel input, got %s / %d channels"
                         % (tuple(packed32.shape), NC * 8))
    y = torch.empty((B, D, H, W), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device), _timed("conv3d_igemm_cout1_bf16_fwd"):
        _lib.check(_lib.load().cmfb200_conv3d_igemm_cout1_bf16_fwd(_p(x), _p(packed32), _p(y), B, NC * 8, D, H, W, _stream()),
                   "conv3d_igemm_cout1_bf16_fwd")
    return y


def deconv3d_igemm(x, packed, want_stats=True):
    """tcgen05 transposed conv (k3 s2 p1 op1) on C8/bf16: [B,Cin/8,D,H,W,8] -> [B,Cout/8,2D,2H,2W,8]."""
    _req(x, packed, dtype=BF16)
    B, NC, D, H, W, _ = x.shape
    Cin, Cout = NC * 8, packed.shape[2]
    y = torch.empty((B, Cout // 8, 2 * D, 2 * H, 2 * W, 8), device=x.device, dtype=BF16)
    sums = _new_sums(B, Cout, x.device) if want_stats else None
    with torch.cuda.device(x.device), _timed("deconv3d_igemm_bf16_fwd"):
        _lib.check(_lib.load().cmfb200_deconv3d_igemm_bf16_fwd(_p(x), _p(packed), _p(y), _p(sums), B, Cin, Cout, D, H, W,
                                                               _stream()), "deconv3d_igemm_bf16_fwd")
    return y, sums


def c8_parity_split(x):
    """C8 [B,C/8,D,H,W,8] -> parity-split [B,8,C/8,D/2,H/2,W/2,8] (input format of the stride-2 igemm)."""
    _req(x, dtype=BF16)
    B, NC, D, H, W, _ = x.shape
    y = torch.empty((B, 8, NC, D // 2, H // 2, W // 2, 8), device=x.device, dtype=BF16)
    with torch.cuda.device(x.device), _timed("c8_parity_split"):
        _lib.check(_lib.load().cmfb200_c8_parity_split(_p(x), _p(y), B, NC * 8, D, H, W, _stream()), "c8_parity_split")
    return y


def conv3d_s2_igemm(x_split, packed, want_stats=True):
    """tcgen05 stride-2 conv on the parity-split input: -> [B,Cout/8,D/2,H/2,W/2,8]."""
    _req(x_split, packed, dtype=BF16)
    B, _, NC, Do, Ho, Wo, _ = x_split.shape
    Cin, Cout = NC * 8, packed.shape[2]
    y = torch.empty((B, Cout // 8, Do, Ho, Wo, 8), device=x_split.device, dtype=BF16)
    sums = _new_sums(B, Cout, x_split.device) if want_stats else None
    with torch.cuda.device(x_split.device), _timed("conv3d_s2_igemm_bf16_fwd"):
        _lib.check(_lib.load().cmfb200_conv3d_s2_igemm_bf16_fwd(_p(x_split), _p(packed), _p(y), _p(sums), B, Cin, Cout, Do,
                                                                Ho, Wo, _stream()), "conv3d_s2_igemm_bf16_fwd")
    return y, sums


def gn_apply_c8(x, sums, gamma, beta, residual=None, relu=False, out=None, want_split=False, f32_out=False,
                groups=GN_GROUPS, eps=GN_EPS):
    """GroupNorm(+residual)(+ReLU) on C8/bf16.  With `want_split` also returns the parity-split copy that the
    stride-2 implicit GEMM consumes: (y, y_split).  With `f32_out` the result is returned ONLY as un-rounded fp32
    [B,C,D,H,W] (the input of the fp32 classifier tail)."""
    gamma, beta = gamma.detach(), beta.detach()
    _req(gamma, beta)
    _req(x, residual, dtype=BF16)
    B, NC, D, H, W, _ = x.shape
    y32 = torch.empty((B, NC * 8, D, H, W), device=x.device, dtype=torch.float32) if f32_out else None
    y = None if f32_out else (torch.empty_like(x) if out is None else out)
    split = torch.empty((B, 8, NC, D // 2, H // 2, W // 2, 8), device=x.device, dtype=BF16) if want_split else None
    with torch.cuda.device(x.device), _timed("gn_apply_c8_bf16"):
        _lib.check(_lib.load().cmfb200_gn_apply_c8_bf16(_p(x), _p(sums), _p(gamma), _p(beta), _p(residual), _p(y),
                                                        _p(split), _p(y32), B, NC * 8, groups, D, H, W, eps, int(relu),
                                                        _stream()), "gn_apply_c8_bf16")
    if f32_out:
        return y32
    return (y, split) if want_split else y


# ------------------------------------------------------------------------------------------ tc3 (fp32-accurate tcgen05)
# C8S3: bf16 [B, C/8, 3, (D,) H, W, 8] -- three bf16 terms whose sum is the fp32 value; C8F: fp32 [B, C/8, (D,) H, W, 8].
def pack_tc3_weight(weight):
    """nn.Conv2d [Cout,Cin,k,k] / nn.Conv3d [Cout,Cin,3,3,3] weight (fp32) -> split-bf16 packing of conv_tc3."""
    weight = weight.detach()
    _req(weight)
    Cout, Cin = weight.shape[:2]
    KD = 3 if weight.dim() == 5 else 1
    k = weight.shape[-1]
    packed = torch.empty((KD * (Cin // 16), k, k, 2, 3, Cout, 8), device=weight.device, dtype=BF16)
    with torch.cuda.device(weight.device):
        _lib.check(_lib.load().cmfb200_pack_tc3_weight(_p(weight), _p(packed), Cout, Cin, KD, k, _stream()),
                   "pack_tc3_weight")
    return packed


def conv_tc3(x_s3, packed, dilation=1, want_stats=True, out_nchw=False, row_off=0, out_rows=None):
    """Stride-1 'same' conv (2-D or 3-D by the rank of x_s3) on the split-bf16 tensor-core path.
    Returns (raw fp32 y in C8F -- or NCHW/NCDHW when out_nchw --, gn_sums or None).
    Row bands: `x_s3` carries halo rows; `out_rows` rows are produced, output row h centred on input row h + row_off."""
    _req(x_s3, packed, dtype=BF16)
    three_d = x_s3.dim() == 7
    B, NC = x_s3.shape[:2]
    sp = tuple(x_s3.shape[3:-1])
    D, H, W = sp if three_d else (1,) + sp
    KS, k, _, _, _, Cout, _ = packed.shape
    Cin = NC * 8
    KD = KS // (Cin // 16)
    if x_s3.shape[2] != 3 or KS != KD * (Cin // 16) or KD != (3 if three_d else 1):
        raise ValueError("conv_tc3: input %s does not match packed weight %s" % (tuple(x_s3.shape), tuple(packed.shape)))
    Ho = H if out_rows is None else out_rows
    osp = sp[:-2] + (Ho, W)
    y = torch.empty(((B, Cout) + osp) if out_nchw else ((B, Cout // 8) + osp + (8,)), device=x_s3.device, dtype=torch.float32)
    sums = _new_sums(B, Cout, x_s3.device) if want_stats else None
    with torch.cuda.device(x_s3.device), _timed("conv_tc3_fwd"):
        _lib.check(_lib.load().cmfb200_conv_tc3_rows_fwd(_p(x_s3), _p(packed), _p(y), _p(sums), B, Cin, Cout, D, H, W, KD, k,
                                                         dilation, int(out_nchw), row_off, Ho, _stream()), "conv_tc3_fwd")
    return y, sums


def _split3_torch(w):
    """Exact three-term bf16 split of an fp32 tensor (the same round-to-nearest steps as the kernels)."""
    t0 = w.to(BF16)
    r1 = w - t0.float()
    t1 = r1.to(BF16)
    return t0, t1, (r1 - t1.float()).to(BF16)


def _pack_tc3_taps(w_oi, steps):
    """w_oi: fp32 [Cout, Cin, 3, 3, 3] (conv orientation).  steps: list of (kc, [(kd,kh,kw), ...]) in kernel K-step
    order.  Returns bf16 [sum taps, 2 chunks, 3 terms, Cout, 8] flattened = the B operand stream of conv_tc3g."""
    Cout = w_oi.shape[0]
    terms = _split3_torch(w_oi)
    out = []
    for kc, taps in steps:
        for kd, kh, kw in taps:
            blk = torch.stack([t[:, kc * 16:(kc + 1) * 16, kd, kh, kw] for t in terms])  # [3, Cout, 16]
            out.append(blk.view(3, Cout, 2, 8).permute(2, 0, 1, 3))  # [2 chunks, 3 terms, Cout, 8]
    return torch.stack(out).contiguous()


def _axis_taps(bit, transposed):
    """Kernel index k of the taps a parity bit contributes, in the order of csrc/conv_tc3_s2.cu."""
    if not bit:
        return [1]
    return [2, 0] if transposed else [0, 2]


def pack_tc3_s2_weight(weight):
    """nn.Conv3d(k3, s2, p1) weight [Cout,Cin,3,3,3] -> packed operand of `conv_tc3_s2` (K steps: input parity class q,
    depth tap, 16-channel chunk; 1..4 in-plane taps each)."""
    w = weight.detach().float()
    KC = w.shape[1] // 16
    steps = []
    for q in range(8):
        qd, qh, qw = q >> 2, (q >> 1) & 1, q & 1
        for kd in _axis_taps(qd, False):
            for kc in range(KC):
                steps.append((kc, [(kd, kh, kw) for kh in _axis_taps(qh, False) for kw in _axis_taps(qw, False)]))
    return _pack_tc3_taps(w, steps)


def pack_tc3_deconv_weight(weight):
    """nn.ConvTranspose3d(k3, s2, p1, op1) weight [Cin,Cout,3,3,3] -> packed operand of `deconv_tc3` (eight output parity
    classes back to back; per class: depth tap, 16-channel chunk; 1..4 in-plane taps each)."""
    w = weight.detach().float().transpose(0, 1).contiguous()  # [Cout, Cin, 3,3,3]: out[2i-1+k] += W[ci][co][k] in[i]
    KC = w.shape[1] // 16
    steps = []
    for p in range(8):
        pd, ph, pw = p >> 2, (p >> 1) & 1, p & 1
        for kd in _axis_taps(pd, True):
            for kc in range(KC):
                steps.append((kc, [(kd, kh, kw) for kh in _axis_taps(ph, True) for kw in _axis_taps(pw, True)]))
    return _pack_tc3_taps(w, steps)


def conv_tc3_s2(x_split, packed, Cout=64, want_stats=True, pad=0):
    """Stride-2 3x3x3 conv on the tensor cores: parity-split C8S3 [B,8,Cin/8,3,D/2,H/2 + 2 pad,W/2,8] -> raw C8F
    [B,Cout/8,D/2,H/2,W/2,8] (`pad` spare cell rows of the input for row bands)."""
    _req(x_split, packed, dtype=BF16)
    B, _, NC, _, Do, Hp, Wo, _ = x_split.shape
    Ho = Hp - 2 * pad
    y = torch.empty((B, Cout // 8, Do, Ho, Wo, 8), device=x_split.device, dtype=torch.float32)
    sums = _new_sums(B, Cout, x_split.device) if want_stats else None
    with torch.cuda.device(x_split.device), _timed("conv_tc3_s2_fwd"):
        _lib.check(_lib.load().cmfb200_conv_tc3_s2_rows_fwd(_p(x_split), _p(packed), _p(y), _p(sums), B, NC * 8, Cout, Do, Ho,
                                                            Wo, pad, _stream()), "conv_tc3_s2_fwd")
    return y, sums


def deconv_tc3(x_s3, packed, Cout, want_stats=True, pad=0):
    """Transposed 3x3x3 conv (s2, p1, op1) on the tensor cores: C8S3 [B,Cin/8,3,D,H + 2 pad,W,8] -> raw C8F
    [B,Cout/8,2D,2H,2W,8] (`pad` spare rows of the input for row bands)."""
    _req(x_s3, packed, dtype=BF16)
    B, NC, _, D, Hp, W, _ = x_s3.shape
    H = Hp - 2 * pad
    y = torch.empty((B, Cout // 8, 2 * D, 2 * H, 2 * W, 8), device=x_s3.device, dtype=torch.float32)
    sums = _new_sums(B, Cout, x_s3.device) if want_stats else None
    with torch.cuda.device(x_s3.device), _timed("deconv_tc3_fwd"):
        _lib.check(_lib.load().cmfb200_deconv_tc3_rows_fwd(_p(x_s3), _p(packed), _p(y), _p(sums), B, NC * 8, Cout, D, H, W, pad,
                                                           _stream()), "deconv_tc3_fwd")
    return y, sums


def gn_apply_tc3(raw, sums, gamma, beta, raw_c8f, res_s3=None, res_nchw=None, relu=False, want_s3=True, want_nchw=False,
                 groups=GN_GROUPS, eps=GN_EPS, pad=0, want_split=False, push=None, nchw_pad=False):
    """GroupNorm (+residual) (+ReLU) of the tc3 pipeline.  raw: C8F (raw_c8f) or NCHW/NCDHW fp32; the result is returned
    as (C8S3 or None, NCHW fp32 or None).  sums=None: layout conversion / three-term split only.
    Row bands: with `pad` the C8S3 result (and `res_s3`) carry `pad` halo rows above and below (left un-written);
    `push` = (up, dn, rows): bf16 landing buffers IN THE NEIGHBOUR RANKS' memory (or None) that receive the first / last
    `rows` rows of the C8S3 result -- the next conv's halo exchange fused into this kernel.  `nchw_pad`: the NCHW result
    carries the same `pad` rows (un-written)."""
    _req(raw, res_nchw)
    _req(res_s3, dtype=BF16)
    if sums is not None:
        gamma, beta = gamma.detach(), beta.detach()
        _req(gamma, beta)
    if raw_c8f:
        B, NC = raw.shape[:2]
        C, sp = NC * 8, tuple(raw.shape[2:-1])
    else:
        B, C = raw.shape[:2]
        sp = tuple(raw.shape[2:])
    spatial = 1
    for v in sp:
        spatial *= v
    psp = sp[:-2] + (sp[-2] + 2 * pad, sp[-1])
    if res_s3 is not None and tuple(res_s3.shape) != (B, C // 8, 3) + psp + (8,):
        raise ValueError("gn_apply_tc3: residual %s does not match %s (pad %d)" % (tuple(res_s3.shape), (B, C // 8, 3) + psp, pad))
    y_s3 = torch.empty((B, C // 8, 3) + psp + (8,), device=raw.device, dtype=BF16) if want_s3 else None
    y_nchw = (torch.empty((B, C) + (psp if nchw_pad else sp), device=raw.device, dtype=torch.float32)
              if want_nchw else None)
    csp = tuple(v // 2 for v in sp)
    y_split = (torch.empty((B, 8, C // 8, 3) + csp[:-2] + (csp[-2] + 2 * pad, csp[-1]) + (8,), device=raw.device, dtype=BF16)
               if want_split else None)
    with torch.cuda.device(raw.device), _timed("gn_apply_tc3"):
        _lib.check(_lib.load().cmfb200_gn_apply_tc3_padded(_p(raw), int(raw_c8f), _p(sums),
                                                           _p(gamma) if sums is not None else None,
                                                           _p(beta) if sums is not None else None, _p(res_s3), _p(res_nchw),
                                                           _p(y_s3), _p(y_nchw), B, C, groups, spatial, eps, int(relu), pad,
                                                           sp[-2], sp[-1], _p(y_split), _p(push[0]) if push else None,
                                                           _p(push[1]) if push else None, push[2] if push else 0,
                                                           int(bool(nchw_pad and pad)), _stream()), "gn_apply_tc3")
    if want_split:
        return y_s3, y_nchw, y_split
    return y_s3, y_nchw


def cost_volume_concat_c8s3(L, R, D, pad=0):
    """K1 in C8S3: [B,C,h,w] fp32 x2 -> [B, 2C/8, 3, D, h + 2 pad, w, 8] bf16 (terms sum to the fp32 volume exactly;
    `pad` un-written halo rows above and below for row-band sharding)."""
    _req(L, R)
    B, C, h, w = L.shape
    cost = torch.empty((B, 2 * C // 8, 3, D, h + 2 * pad, w, 8), device=L.device, dtype=BF16)
    with torch.cuda.device(L.device), _timed("cost_volume_concat_c8s3"):
        _lib.check(_lib.load().cmfb200_cost_volume_concat_c8s3_padded(_p(L), _p(R), _p(cost), B, C, h, w, D, pad, _stream()),
                   "cost_volume_concat_c8s3")
    return cost


def f32_to_c8s3(x, pad=0):
    """[B,C,...] fp32 -> C8S3 (exact three-term bf16 split)."""
    return gn_apply_tc3(x, None, None, None, raw_c8f=False, pad=pad)[0]


def c8s3_to_f32(x_s3):
    """C8S3 -> [B,C,...] fp32 (sum of the three terms; exact)."""
    t = x_s3.float().sum(2)  # [B, C/8, ..., 8]
    nd = t.dim()
    perm = (0, 1, nd - 1) + tuple(range(2, nd - 1))
    t = t.permute(*perm).contiguous()
    return t.view((t.shape[0], t.shape[1] * 8) + tuple(t.shape[3:]))


def c8f_to_f32(y):
    """C8F fp32 [B,C/8,...,8] -> [B,C,...] fp32 (tests / tools)."""
    nd = y.dim()
    perm = (0, 1, nd - 1) + tuple(range(2, nd - 1))
    t = y.permute(*perm).contiguous()
    return t.view((t.shape[0], t.shape[1] * 8) + tuple(t.shape[3:]))


def softargmin_ctxmap_bwd(c1, c2, c3, weights9, g1, g2, g3, scale):
    """Backward of `softargmin_ctxmap`: returns (dc1, dc2, dc3, dweights9)."""
    gs = [g.contiguous() for g in (g1, g2, g3)]
    _req(c1, c2, c3, weights9, *gs)
    B, D, h, w = c1.shape
    dcs = [torch.empty_like(c1) for _ in range(3)]
    dw = torch.empty_like(weights9)
    with torch.cuda.device(c1.device), _timed("softargmin_ctxmap_bwd"):
        _lib.check(_lib.load().cmfb200_softargmin_ctxmap_bwd(_p(c1), _p(c2), _p(c3), _p(weights9), _p(gs[0]), _p(gs[1]),
                                                             _p(gs[2]), _p(dcs[0]), _p(dcs[1]), _p(dcs[2]), _p(dw), B, D, h,
                                                             w, scale, _stream()), "softargmin_ctxmap_bwd")
    return dcs[0], dcs[1], dcs[2], dw


# ---- cmfsm_sub_8 variant ---------------------------------------------------------------------------
def ctxmap_weights5(lr, hr, w0, w1, w2, w3):
    """six_related_context_mapping (reference-image half, cmf/models/cmfsm_sub_8.py:440-572):
    [B,32,h,w],[B,32,H,W] -> [B,5,H,W] = softmax(logits)*logits, neighbours c,r,l,t,b."""
    ws = [w.detach().reshape(w.shape[0], w.shape[1]).contiguous() for w in (w0, w1, w2, w3)]
    _req(lr, hr, *ws)
    B, C, h, w = lr.shape
    H, W = hr.shape[2:]
    if C != 32 or hr.shape[1] != 32 or tuple(ws[0].shape) != (32, 66):
        raise ValueError("context mapping expects 32-channel features and a 66-input similarity MLP")
    scale = W // w
    if H != h * scale or W != w * scale:
        raise ValueError("hr %dx%d is not an integer multiple of lr %dx%d" % (H, W, h, w))
    out = torch.empty((B, 5, H, W), device=lr.device, dtype=torch.float32)
    with torch.cuda.device(lr.device), _timed("ctxmap_weights5_fwd"):
        _lib.check(_lib.load().cmfb200_ctxmap_weights5_fwd(_p(lr), _p(hr), _p(ws[0]), _p(ws[1]), _p(ws[2]), _p(ws[3]),
                                                           _p(out), B, h, w, scale, _stream()), "ctxmap_weights5_fwd")
    return out


def softargmin_ctxmap5(c1, c2, c3, weights5, scale):
    """cmf/models/cmfsm_sub_8.py:757-802.  c_i [B,D,h,w] (independent), weights5 [B,5,H,W] -> 3 x [B,1,H,W]."""
    _req(c1, c2, c3, weights5)
    B, D, h, w = c1.shape
    H, W = h * scale, w * scale
    if tuple(weights5.shape) != (B, 5, H, W):
        raise ValueError("weights5 shape %s != %s" % (tuple(weights5.shape), (B, 5, H, W)))
    outs = [torch.empty((B, 1, H, W), device=c1.device, dtype=torch.float32) for _ in range(3)]
    with torch.cuda.device(c1.device), _timed("softargmin_ctxmap5_fwd"):
        _lib.check(_lib.load().cmfb200_softargmin_ctxmap5_fwd(_p(c1), _p(c2), _p(c3), _p(weights5), _p(outs[0]),
                                                              _p(outs[1]), _p(outs[2]), None, B, D, h, w, scale,
                                                              _stream()), "softargmin_ctxmap5_fwd")
    return tuple(outs)


def spp_upsample_concat_sized(raw, skip, branches):
    """cat([raw, skip, up(b) for b in branches], 1) for four branch maps of arbitrary size (cmfsm_sub_8)."""
    _req(raw, skip, *branches)
    B, _, H, W = skip.shape
    raw_c = raw.shape[1]
    if raw_c not in (64, 128) or skip.shape[1] != 128 or len(branches) != 4 or any(t.shape[1] != 32 for t in branches):
        raise ValueError("spp_upsample_concat_sized expects (64|128) + 128 + 4 x 32 channels")
    cat = torch.empty((B, raw_c + 256, H, W), device=skip.device, dtype=torch.float32)
    sizes = [int(v) for t in branches for v in t.shape[2:]]
    with torch.cuda.device(skip.device), _timed("spp_upsample_concat_fwd"):
        _lib.check(_lib.load().cmfb200_spp_upsample_concat_sized_fwd(_p(raw), _p(skip), *[_p(t) for t in branches], _p(cat),
                                                                     B, raw_c, H, W, *sizes, _stream()),
                   "spp_upsample_concat_sized_fwd")
    return cat


# ---- cmfsm_sub_16 variant --------------------------------------------------------------------------
def ctxmap_weights3(lr, hr, w0, w1, w2, w3):
    """Target-image half of six_related_context_mapping (cmf/models/cmfsm_sub_16.py:488-573): the right image's
    [B,32,h,w],[B,32,H,W] -> [B,3,H,W] = softmax(logits)*logits, neighbours centre, right, left."""
    ws = [w.detach().reshape(w.shape[0], w.shape[1]).contiguous() for w in (w0, w1, w2, w3)]
    _req(lr, hr, *ws)
    B, C, h, w = lr.shape
    H, W = hr.shape[2:]
    scale = W // w
    if C != 32 or hr.shape[1] != 32 or H != h * scale or W != w * scale:
        raise ValueError("ctxmap_weights3: bad shapes %s / %s" % (tuple(lr.shape), tuple(hr.shape)))
    out = torch.empty((B, 3, H, W), device=lr.device, dtype=torch.float32)
    with torch.cuda.device(lr.device), _timed("ctxmap_weights3_fwd"):
        _lib.check(_lib.load().cmfb200_ctxmap_weights3_fwd(_p(lr), _p(hr), _p(ws[0]), _p(ws[1]), _p(ws[2]), _p(ws[3]),
                                                           _p(out), B, h, w, scale, _stream()), "ctxmap_weights3_fwd")
    return out


def volume_mapping(c1, c2, c3, weights5, weights3, scale):
    """cmf/models/cmfsm_sub_16.py:760-850: raw classifier volumes [B,D',h,w] + spatial / target weights -> 3 x [B,H,W]."""
    _req(c1, c2, c3, weights5, weights3)
    B, Dl, h, w = c1.shape
    H, W = h * scale, w * scale
    if tuple(weights5.shape) != (B, 5, H, W) or tuple(weights3.shape) != (B, 3, H, W):
        raise ValueError("volume_mapping: weights %s / %s do not match %s at scale %d"
                         % (tuple(weights5.shape), tuple(weights3.shape), tuple(c1.shape), scale))
    outs = [torch.empty((B, H, W), device=c1.device, dtype=torch.float32) for _ in range(3)]
    with torch.cuda.device(c1.device), _timed("volume_mapping_fwd"):
        _lib.check(_lib.load().cmfb200_volume_mapping_fwd(_p(c1), _p(c2), _p(c3), _p(weights5), _p(weights3), _p(outs[0]),
                                                          _p(outs[1]), _p(outs[2]), B, Dl, h, w, scale, _stream()),
                   "volume_mapping_fwd")
    return tuple(outs)


# ---- bilinear_cmf* baselines -----------------------------------------------------------------------
def trilinear_softargmin(c1, c2, c3, maxdisp, H, W):
    """cmf/models/bilinear_cmf.py:418-452: raw classifier volumes [B,D',h,w] -> cumulative sums -> trilinear upsampling
    to [B,maxdisp,H,W] -> softmax over maxdisp -> regression.  Returns 3 x [B,H,W]."""
    _req(c1, c2, c3)
    B, Dl, h, w = c1.shape
    outs = [torch.empty((B, H, W), device=c1.device, dtype=torch.float32) for _ in range(3)]
    with torch.cuda.device(c1.device), _timed("trilinear_softargmin_fwd"):
        _lib.check(_lib.load().cmfb200_trilinear_softargmin_fwd(_p(c1), _p(c2), _p(c3), _p(outs[0]), _p(outs[1]), _p(outs[2]),
                                                                B, Dl, h, w, maxdisp, H, W, _stream()),
                   "trilinear_softargmin_fwd")
    return tuple(outs)
