"""Multi-GPU plumbing for the cmfsm hot path: one process per GPU, `torch.distributed` (NCCL over NVLink on the
GPU box, gloo in the CPU tests).  SURVEY.md section 8e.

Two sharding modes, both without any hidden collective:

* **independent samples** (training, batched inference): every rank owns whole stereo pairs.  GroupNorm is
  per-sample, so the forward needs no communication; training needs ONE gradient all-reduce per step
  (`allreduce_gradients`, bucketed flat buffers, the B200-native replacement of the reference's
  `nn.DataParallel` reduce, train.py:78-81) and, for exact equivalence with the reference's global masked mean
  loss (train.py:162-174), an all-reduce of the valid-pixel counts (`masked_smooth_l1_dp`).
* **row bands** of a single high-resolution pair (BASELINE config 5): a rank owns `h/N` rows of the 1/4-resolution
  volume.  Every 3-D conv needs a 1-row halo from each neighbour at its own resolution (`exchange_row_halo`) and
  every GroupNorm needs the per-(sample,channel) sum / sum-of-squares summed over bands (`allreduce_gn_sums`),
  64-128 doubles per layer.
"""
import torch
import torch.distributed as dist


def world(group=None):
    """Number of ranks of `group` (default: the whole job); 1 when torch.distributed is not initialised."""
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def rank(group=None):
    """This process's rank INSIDE `group` (default group: the global rank)."""
    return dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0


def _peer(group, r):
    """Global rank of rank `r` of `group` (point-to-point ops address peers by global rank)."""
    return r if group is None else dist.get_global_rank(group, r)


# ------------------------------------------------------------------------------------------ data parallel
def allreduce_gradients(params, bucket_bytes=32 << 20, average=False, group=None):
    """Sum (or average) `.grad` of `params` over all ranks through flat buckets (one collective per bucket).

    5,255,368 fp32 gradients = 21 MB -> a single 32 MB bucket by default: on NVSwitch the all-reduce cost is
    latency- not link-bound, so fewer, larger collectives win.  Returns the number of collectives.
    """
    n = world(group)
    params = list(params)
    if n == 1 or not params:
        return 0
    # bucket over ALL parameters (a missing gradient is a zero contribution), so that every rank issues the same
    # collectives even when a parameter received no gradient on some ranks only
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    grads = [p.grad for p in params]
    buckets, cur, cur_bytes = [], [], 0
    for g in grads:
        nbytes = g.numel() * g.element_size()
        if cur and (cur_bytes + nbytes > bucket_bytes or g.dtype != cur[0].dtype):
            buckets.append(cur)
            cur, cur_bytes = [], 0
        cur.append(g)
        cur_bytes += nbytes
    if cur:
        buckets.append(cur)
    for bucket in buckets:
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            flat.div_(n)
        off = 0
        for g in bucket:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
    return len(buckets)


def broadcast_parameters(module, src=0, group=None):
    """Make every rank start from rank `src`'s weights (DataParallel replicates from device 0 every step)."""
    if world(group) == 1:
        return
    src = _peer(group, src)
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)


def masked_smooth_l1_dp(outputs, target, maxdisp=192, weights=(0.5, 0.7, 1.0), group=None):
    """The reference loss (train.py:162-174): sum_i w_i * smooth_l1(out_i[mask], target[mask]) with the MEAN taken
    over the valid pixels (0 < target < maxdisp) of the GLOBAL batch.  Each rank contributes sum/global_count, so that
    summing gradients over ranks (allreduce_gradients(average=False)) reproduces the single-process DataParallel
    gradient exactly.  Two fused kernels (cmfb200_masked_smooth_l1_fwd/_bwd) instead of the reference's mask, gathers
    and reductions; CPU tensors (the gloo tests of the sharding arithmetic) take the plain PyTorch statement."""
    if outputs[0].is_cuda:
        from . import autograd_ops as aops

        return aops.masked_smooth_l1(outputs, target, maxdisp, weights, distributed=world(group) > 1, group=group)
    mask = ((target < maxdisp) & (target > 0)).detach()
    count = mask.sum().to(torch.float64)
    if world(group) > 1:
        dist.all_reduce(count, op=dist.ReduceOp.SUM, group=group)
    count = count.clamp_min(1.0).to(target.dtype)
    loss = target.new_zeros(())
    for wgt, out in zip(weights, outputs):
        out = out.squeeze(1) if out.dim() == 4 else out
        loss = loss + wgt * torch.nn.functional.smooth_l1_loss(out[mask], target[mask], reduction="sum") / count
    return loss


def dp_train_step(model, optimizer, left, right, disparity, maxdisp=192):
    """One data-parallel training step on this rank's shard of the batch.  Returns the (global) loss value."""
    optimizer.zero_grad(set_to_none=True)
    outputs = model(left, right)
    loss = masked_smooth_l1_dp(outputs, disparity, maxdisp)  # mask = (disparity < maxdisp) & (disparity > 0), train.py:162
    loss.backward()
    allreduce_gradients(list(model.parameters()), average=False)
    optimizer.step()
    total = loss.detach().clone()
    if world() > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM)
    return float(total)


# ------------------------------------------------------------------------------------------ row bands
def band_rows(h, n, r, multiple=4):
    """Rows [r0, r1) of a height-`h` volume owned by rank `r` of `n`; boundaries are multiples of `multiple`
    (=4 keeps the two stride-2 levels of the hourglass aligned with band edges)."""
    if h % multiple:
        raise ValueError("height %d is not a multiple of %d" % (h, multiple))
    units = h // multiple
    if units < n:
        raise ValueError("cannot cut %d rows into %d bands of multiples of %d" % (h, n, multiple))
    base, extra = divmod(units, n)
    u0 = r * base + min(r, extra)
    u1 = u0 + base + (1 if r < extra else 0)
    return u0 * multiple, u1 * multiple


def exchange_row_halo(x, top=1, bottom=1, dim=-2, group=None):
    """Return `x` extended along the row axis `dim` with `top` rows from the previous rank's bottom edge and
    `bottom` rows from the next rank's top edge (zeros at the image border = the conv's zero padding).

    Grouped point-to-point send/recv: each rank posts at most 2 sends + 2 recvs (`batch_isend_irecv`), which
    NCCL fuses into one launch."""
    n, r = world(group), rank(group)
    dim = dim % x.dim()
    shape_t = list(x.shape)
    shape_t[dim] = top
    shape_b = list(x.shape)
    shape_b[dim] = bottom
    halo_t, halo_b = x.new_zeros(shape_t), x.new_zeros(shape_b)
    if n > 1:
        ops = []
        rows = x.shape[dim]
        if r > 0:
            if bottom:
                ops.append(dist.P2POp(dist.isend, x.narrow(dim, 0, bottom).contiguous(), _peer(group, r - 1), group))
            if top:
                ops.append(dist.P2POp(dist.irecv, halo_t, _peer(group, r - 1), group))
        if r < n - 1:
            if top:
                ops.append(dist.P2POp(dist.isend, x.narrow(dim, rows - top, top).contiguous(), _peer(group, r + 1), group))
            if bottom:
                ops.append(dist.P2POp(dist.irecv, halo_b, _peer(group, r + 1), group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
    parts = ([halo_t] if top else []) + [x] + ([halo_b] if bottom else [])
    return torch.cat(parts, dim)


def fill_row_halo_(x, pad, top, bottom, dim=-3, group=None):
    """In-place halo exchange for activations that were ALLOCATED with `pad` spare rows above and below the band's rows
    along `dim` (ops.gn_apply_tc3(pad=...), ops.cost_volume_concat_c8s3(pad=...)): fills the `top` rows just above the
    band with the previous rank's bottom edge and the `bottom` rows just below with the next rank's top edge (zeros at
    the image border = the conv's zero padding).  Only the boundary rows move; the activation itself is never copied."""
    n, r = world(group), rank(group)
    dim = dim % x.dim()
    rows = x.shape[dim] - 2 * pad
    if top > pad or bottom > pad:
        raise ValueError("halo (%d, %d) exceeds the %d spare rows" % (top, bottom, pad))
    recv_t = recv_b = None
    ops = []
    if n > 1:
        if r > 0:
            if bottom:
                ops.append(dist.P2POp(dist.isend, x.narrow(dim, pad, bottom).contiguous(), _peer(group, r - 1), group))
            if top:
                recv_t = torch.empty_like(x.narrow(dim, pad - top, top), memory_format=torch.contiguous_format)
                ops.append(dist.P2POp(dist.irecv, recv_t, _peer(group, r - 1), group))
        if r < n - 1:
            if top:
                ops.append(dist.P2POp(dist.isend, x.narrow(dim, pad + rows - top, top).contiguous(), _peer(group, r + 1), group))
            if bottom:
                recv_b = torch.empty_like(x.narrow(dim, pad + rows, bottom), memory_format=torch.contiguous_format)
                ops.append(dist.P2POp(dist.irecv, recv_b, _peer(group, r + 1), group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
    if top:
        dst = x.narrow(dim, pad - top, top)
        dst.copy_(recv_t) if recv_t is not None else dst.zero_()
    if bottom:
        dst = x.narrow(dim, pad + rows, bottom)
        dst.copy_(recv_b) if recv_b is not None else dst.zero_()
    return x


def allreduce_gn_sums(sums, group=None):
    """Sum per-band GroupNorm statistics ([B,C,2] double: sum, sum of squares) over all bands, in place."""
    if world(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return sums


def gather_bands(x, dim=-2, group=None):
    """All-gather equally sized bands along `dim` (every rank gets the full tensor)."""
    n = world(group)
    if n == 1:
        return x
    parts = [torch.empty_like(x) for _ in range(n)]
    dist.all_gather(parts, x.contiguous(), group=group)
    return torch.cat(parts, dim % x.dim())
