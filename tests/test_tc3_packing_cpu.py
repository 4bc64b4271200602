"""CPU: the host-side weight packing of the stride-2 / transposed tensor-core convs (cmf_b200.ops.pack_tc3_s2_weight,
pack_tc3_deconv_weight) walked in the K-step order of csrc/conv_tc3_s2.cu reproduces the convolution it decomposes.
(PyTorch CPU ops stand in for the kernel IN THE TEST ONLY; the kernel itself is tested on the GPU.)"""
import torch
import torch.nn.functional as F

from cmf_b200 import ops


def _unpack(packed):
    """[taps, 2 chunks, 3 terms, Cout, 8] bf16 -> fp32 [taps, Cout, 16] (the three terms sum to the fp32 weight)."""
    w = packed.float().sum(2)  # [taps, 2, Cout, 8]
    return w.permute(0, 2, 1, 3).reshape(w.shape[0], w.shape[2], 16)


def test_stride2_decomposition_over_parity_classes():
    g = torch.Generator().manual_seed(0)
    Cin, Cout, D, H, W = 32, 64, 4, 6, 8
    x = torch.randn(1, Cin, D, H, W, generator=g)
    wgt = torch.randn(Cout, Cin, 3, 3, 3, generator=g) * 0.1
    taps = _unpack(ops.pack_tc3_s2_weight(wgt))
    Do, Ho, Wo = D // 2, H // 2, W // 2
    out = torch.zeros(1, Cout, Do, Ho, Wo)
    xp = F.pad(x, (2, 0, 2, 0, 2, 0))  # two zero voxels on the low side of every axis = one zero CELL
    ti = 0
    for q in range(8):
        qd, qh, qw = q >> 2, (q >> 1) & 1, q & 1
        sub = xp[:, :, qd::2, qh::2, qw::2]  # parity class q with one zero cell in front (cell index c <-> sub[c + 1])
        for kdi in range(1 + qd):
            for kc in range(Cin // 16):
                for khi in range(1 + qh):
                    for kwi in range(1 + qw):
                        od = (kdi - 1) if qd else 0
                        oh = (khi - 1) if qh else 0
                        ow = (kwi - 1) if qw else 0
                        win = sub[:, kc * 16:(kc + 1) * 16, 1 + od:1 + od + Do, 1 + oh:1 + oh + Ho, 1 + ow:1 + ow + Wo]
                        out += torch.einsum("oc,bcdhw->bodhw", taps[ti], win)
                        ti += 1
    assert ti == taps.shape[0] == 27 * (Cin // 16)
    torch.testing.assert_close(out, F.conv3d(x, wgt, None, 2, 1), rtol=1e-5, atol=1e-5)


def test_transposed_decomposition_over_output_parity_classes():
    g = torch.Generator().manual_seed(1)
    Cin, Cout, D, H, W = 64, 32, 2, 3, 4
    x = torch.randn(1, Cin, D, H, W, generator=g)
    wgt = torch.randn(Cin, Cout, 3, 3, 3, generator=g) * 0.1
    taps = _unpack(ops.pack_tc3_deconv_weight(wgt))
    out = torch.zeros(1, Cout, 2 * D, 2 * H, 2 * W)
    xp = F.pad(x, (0, 1, 0, 1, 0, 1))  # one zero cell behind every axis (the TMA box reads cell i + 1)
    ti = 0
    for p in range(8):
        pd, ph, pw = p >> 2, (p >> 1) & 1, p & 1
        for kdi in range(1 + pd):
            for kc in range(Cin // 16):
                for khi in range(1 + ph):
                    for kwi in range(1 + pw):
                        win = xp[:, kc * 16:(kc + 1) * 16, kdi:kdi + D, khi:khi + H, kwi:kwi + W]
                        out[:, :, pd::2, ph::2, pw::2] += torch.einsum("oc,bcdhw->bodhw", taps[ti], win)
                        ti += 1
    assert ti == taps.shape[0] == 27 * (Cin // 16)
    torch.testing.assert_close(out, F.conv_transpose3d(x, wgt, None, 2, 1, 1), rtol=1e-5, atol=1e-5)


def test_pitched_2d_views():
    """ops.pitched_2d: the (height, width, pitch) description of band boundary rows that cmfb200_copy_2d moves."""
    import torch
    from cmf_b200 import ops

    x = torch.zeros(1, 4, 3, 6, 10, 12, 8)
    assert ops.pitched_2d(x) == (1, x.numel(), x.numel())
    rows = x.narrow(4, 2, 3)                      # 3 of the 10 rows: blocks of 3*12*8 elements, 10*12*8 apart
    assert ops.pitched_2d(rows) == (4 * 3 * 6, 3 * 12 * 8, 10 * 12 * 8)
    assert ops.pitched_2d(x.narrow(4, 0, 10)) == (1, x.numel(), x.numel())
    assert ops.pitched_2d(x[..., ::2, :]) == (4 * 3 * 6 * 10 * 6, 8, 16)  # every other column: 8-element blocks
    assert ops.pitched_2d(x.transpose(3, 4)) is None                      # permuted dims: not a pitched block copy
    assert ops.pitched_2d(x.narrow(4, 2, 3).narrow(3, 1, 2)) is None   # two narrowed dims do not collapse
    y = torch.zeros(2, 32, 5, 7)
    assert ops.pitched_2d(y.narrow(2, 1, 2)) == (64, 14, 35)
