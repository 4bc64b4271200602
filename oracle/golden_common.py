"""Seeded inputs / weights shared by oracle/gen_golden.py (which runs the real reference) and tests/.

TEST INFRASTRUCTURE.  Everything is derived from `torch.Generator().manual_seed(n)` CPU streams, which
are stable across machines, so fixtures only need to store OUTPUTS.
"""
import torch

GOLDEN_THREADS = 8
WEIGHT_SEED = 0  # torch.manual_seed(0) before get_model('cmfsm') -- SURVEY.md appendix B
INPUT_SEED = 1
SEED_K5_W, SEED_K5_X = 11, 12
SEED_HG_W, SEED_HG_X = 21, 22
SEED_HEAD_W, SEED_HEAD_X = 31, 32
SEED_CLS_W, SEED_CLS_X = 41, 42
K1_CROP_CH, K1_CROP_ROWS = 2, 2
K4_CROP = (20, 30, 40, 56)  # low-res cell window y0,y1,x0,x1 of the C1 run


def seeded_pair(B, H, W, seed=INPUT_SEED):
    """Uniform [0,1) synthetic stereo pair."""
    g = torch.Generator().manual_seed(seed)
    return torch.rand(B, 3, H, W, generator=g), torch.rand(B, 3, H, W, generator=g)


def structured_pair(H, W, delta=20, seed=INPUT_SEED):
    """Smooth texture shifted by `delta` pixels: true disparity = delta."""
    g = torch.Generator().manual_seed(seed)
    base = torch.nn.functional.avg_pool2d(torch.rand(1, 3, H, W + 64, generator=g), 5, 1, 2)
    return base[..., 64:].contiguous(), base[..., 64 - delta:W + 64 - delta].contiguous()


def seeded_weights(state_dict, seed, gn_affine=False):
    """Deterministic replacement weights for a module's state_dict (same keys/shapes).

    conv weights ~ N(0, 2/fan) (keeps activations O(1)); GroupNorm affine ~ 1 +- 0.25 / +- 0.25 when
    `gn_affine`, else left at 1/0.
    """
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, v in state_dict.items():
        if v.dim() >= 3:
            fan = v[0].numel()
            out[k] = torch.randn(v.shape, generator=g) * (2.0 / fan) ** 0.5
        elif k.endswith(".weight"):
            out[k] = 1.0 + (0.5 * torch.rand(v.shape, generator=g) - 0.25 if gn_affine else 0.0) + torch.zeros_like(v)
        else:
            out[k] = (0.5 * torch.rand(v.shape, generator=g) - 0.25) if gn_affine else torch.zeros_like(v)
    return out


def k5_inputs():
    g = torch.Generator().manual_seed(SEED_K5_X)
    return torch.randn(2, 32, 6, 10, generator=g), torch.randn(2, 32, 24, 40, generator=g)


def hourglass_inputs():
    g = torch.Generator().manual_seed(SEED_HG_X)
    x = torch.randn(1, 32, 8, 8, 16, generator=g)
    presqu = torch.randn(1, 64, 4, 4, 8, generator=g)
    postsqu = torch.randn(1, 64, 4, 4, 8, generator=g)
    return x, presqu, postsqu


def head_input():
    g = torch.Generator().manual_seed(SEED_HEAD_X)
    return torch.randn(1, 64, 6, 8, 20, generator=g)


def classif_input():
    g = torch.Generator().manual_seed(SEED_CLS_X)
    return torch.randn(2, 32, 6, 8, 12, generator=g)


def flying_padded_pair(seed=INPUT_SEED, structured=False):
    """BASELINE config 2 input: a 540x960 pair fed as 576x960 the way the reference test loader pads it
    (cmf/loader/Flying3d.py:67-72: the last 36 rows are appended once more).  Same stream as bench.py."""
    if structured:
        left, right = structured_pair(540, 960, delta=20, seed=seed)
    else:
        g = torch.Generator().manual_seed(seed)
        left, right = torch.rand(1, 3, 540, 960, generator=g), torch.rand(1, 3, 540, 960, generator=g)
    return tuple(torch.cat([x, x[:, :, -36:]], 2).contiguous() for x in (left, right))


def kitti_padded_pair(seed=INPUT_SEED, delta=20):
    """BASELINE config 4 input: a 375x1242 structured pair (true disparity `delta`) padded to 384x1248 the way
    the reference KITTI loader does (cmf/loader/KITTI.py:100-108: the first 9 rows are prepended, then the first
    6 columns)."""
    left, right = structured_pair(375, 1242, delta=delta, seed=seed)
    out = []
    for x in (left, right):
        x = torch.cat([x[:, :, :384 - 375], x], 2)
        out.append(torch.cat([x[:, :, :, :1248 - 1242], x], 3).contiguous())
    return tuple(out)
