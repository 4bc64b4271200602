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
import os

import torch
import torch.distributed as dist


def _copy_rows(dst, src):
    """Boundary-row copy of the halo exchange: the pitched 16-byte kernel on CUDA tensors, copy_ on CPU (gloo tests)."""
    if dst.is_cuda:
        from . import ops
        return ops.copy_rows(dst, src)
    return dst.copy_(src)


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


# ------------------------------------------------------------------------------------------ peer-memory mailbox
class _Mailbox:
    """Symmetric-memory mailbox of the row-band forward (torch.distributed._symmetric_memory: CUDA VMM peer mappings over
    NVLink + device-side signal barriers).  Every rank owns one buffer that all peers can read directly; an exchange is
    "write my boundary rows into my buffer -> barrier -> copy the neighbours' rows out of THEIR buffers" -- no NCCL call, no
    host synchronisation, and (unlike NCCL point-to-point) capturable in a CUDA graph.  Two slots alternate so that one
    barrier per exchange suffices: the barrier of exchange k+1 orders every read of exchange k before the next write of
    that slot."""

    def __init__(self, group, nbytes, device):
        import torch.distributed._symmetric_memory as sm

        self.group = group if group is not None else dist.group.WORLD
        self.cap = int(nbytes)
        self.buf = sm.empty(2 * self.cap, dtype=torch.uint8, device=device)
        self.hdl = sm.rendezvous(self.buf, self.group)
        self.n, self.r = self.hdl.world_size, self.hdl.rank
        self.k = 0

    def slot(self):
        self.k += 1
        return (self.k & 1) * self.cap, self.k & 1

    @staticmethod
    def _nbytes(shape, dtype):
        n = torch.empty(0, dtype=dtype).element_size()
        for v in shape:
            n *= v
        return n

    def mine(self, off, like=None, shape=None, dtype=None):
        shape, dtype = (like.shape, like.dtype) if like is not None else (shape, dtype)
        return self.buf[off:off + self._nbytes(shape, dtype)].view(dtype).view(shape)

    def peer(self, rank, off, like=None, shape=None, dtype=None):
        shape, dtype = (like.shape, like.dtype) if like is not None else (shape, dtype)
        return self.hdl.get_buffer(rank, (self._nbytes(shape, dtype),), torch.uint8, off).view(dtype).view(shape)

    def barrier(self, channel):
        self.hdl.barrier(channel=channel)


_MAILBOX = {}


def _align(nb):
    return (nb + 255) & ~255


def _mailbox(group, device, need):
    """The mailbox of (group, device), grown (collectively: every rank sees the same shapes) when a message needs more;
    None when peer memory is switched off (CMF_B200_BANDS_COMM=nccl) or unavailable."""
    if os.environ.get("CMF_B200_BANDS_COMM", "symm") != "symm":
        return None
    key = (id(group), device.index)
    mb = _MAILBOX.get(key)
    if mb is False:
        return None
    if mb is None or mb.cap < need:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("the peer-memory mailbox must be sized by an eager warm-up pass before graph capture")
        if mb is not None:  # growing: nobody may still be reading the old buffer
            torch.cuda.synchronize(device)
            dist.barrier(group=group)
        try:
            mb = _Mailbox(group, max(need, 1 << 22) if mb is None else max(need, 2 * mb.cap), device)
        except Exception:  # no symmetric-memory support on this stack: NCCL send/recv + all-reduce instead
            _MAILBOX[key] = False
            return None
        _MAILBOX[key] = mb
    return mb


class HaloPush:
    """A reserved mailbox slot into which the PRODUCING kernel (gn_apply_tc3) pushes a band's boundary rows: `up` / `dn`
    are bf16 views of the landing regions in the previous / next rank's mailbox (peer memory), `rows` rows each.  The
    consumer side (`fill_row_halo_`) only waits on the slot's barrier and copies out of ITS OWN mailbox."""

    def __init__(self, mb, off, ch, nb, rows_shape, dim, group):
        self.mb, self.off, self.ch, self.nb, self.shape, self.dim, self.done = mb, off, ch, nb, rows_shape, dim, False
        self.k = mb.k  # exchange counter at reservation: the slot is overwritten two exchanges later
        n, r = mb.n, mb.r
        # my TOP rows land in rank r-1's "from next" region (off + nb); my BOTTOM rows in rank r+1's "from previous" (off)
        self.up = mb.peer(r - 1, off + nb, shape=rows_shape, dtype=torch.bfloat16) if r > 0 else None
        self.dn = mb.peer(r + 1, off, shape=rows_shape, dtype=torch.bfloat16) if r < n - 1 else None
        self.rows = rows_shape[dim]

    def args(self):
        return (self.up, self.dn, self.rows)


def reserve_halo_push(full_shape, pad, dim, device, group=None):
    """Reserve the next mailbox slot for a C8S3 band activation of (padded) shape `full_shape` about to be produced by
    gn_apply_tc3; returns a HaloPush (its `.args()` go to the kernel) or None when peer memory is not in use."""
    n = world(group)
    if n == 1:
        return None
    dim = dim % len(full_shape)
    rows_shape = list(full_shape)
    rows_shape[dim] = pad
    nelem = 1
    for v in rows_shape:
        nelem *= v
    nb = _align(nelem * 2)
    mb = _mailbox(group, device, 2 * nb)
    if mb is None:
        return None
    off, ch = mb.slot()
    return HaloPush(mb, off, ch, nb, tuple(rows_shape), dim, group)


def band_fence(device, group=None):
    """End-of-forward fence of the peer-memory transport: one more barrier, so that every read of this forward's last
    exchange has completed on every rank before the next forward (or the next replay of a captured graph, whose slot
    sequence is baked in) writes that slot again.  Returns True when the mailbox transport is active."""
    mb = _MAILBOX.get((id(group), device.index))
    if not mb:
        return False
    _off, ch = mb.slot()
    mb.barrier(ch)
    return True


def mailbox_capacity(group=None, device=None):
    """Bytes per slot of the current mailbox (0 = none): lets a caller see whether a warm-up pass re-allocated it."""
    mb = _MAILBOX.get((id(group), device.index if device is not None else torch.cuda.current_device()))
    return mb.cap if mb else 0


def _exchange(x, dim, pad_lo, n_up, n_dn, top, bottom, group):
    """Core of the two halo functions.  Sends x[pad_lo : pad_lo+bottom] (my first rows) to the previous rank and my last
    `top` rows to the next one; returns (rows received from the previous rank or None, from the next rank or None)."""
    n, r = world(group), rank(group)
    rows_hi = n_up  # index one past my last row along dim
    up = x.narrow(dim, pad_lo, bottom) if bottom else None        # -> previous rank's bottom halo
    dn = x.narrow(dim, rows_hi - top, top) if top else None      # -> next rank's top halo
    mb = None
    if x.is_cuda and n > 1:
        nb_up = _align(up.numel() * up.element_size()) if up is not None else 0
        nb_dn = _align(dn.numel() * dn.element_size()) if dn is not None else 0
        mb = _mailbox(group, x.device, nb_up + nb_dn)
    got_t = got_b = None
    if n == 1:
        return None, None
    if mb is not None:
        off, ch = mb.slot()
        if up is not None and r > 0:
            _copy_rows(mb.mine(off, up), up)
        if dn is not None and r < n - 1:
            _copy_rows(mb.mine(off + nb_up, dn), dn)
        mb.barrier(ch)
        if top and r > 0:
            got_t = mb.peer(r - 1, off + nb_up, dn)   # the previous rank's "down" region (same shape as my own)
        if bottom and r < n - 1:
            got_b = mb.peer(r + 1, off, up)           # the next rank's "up" region
        return got_t, got_b
    ops = []
    if r > 0:
        if bottom:
            ops.append(dist.P2POp(dist.isend, up.contiguous(), _peer(group, r - 1), group))
        if top:
            got_t = torch.empty_like(dn, memory_format=torch.contiguous_format)
            ops.append(dist.P2POp(dist.irecv, got_t, _peer(group, r - 1), group))
    if r < n - 1:
        if top:
            ops.append(dist.P2POp(dist.isend, dn.contiguous(), _peer(group, r + 1), group))
        if bottom:
            got_b = torch.empty_like(up, memory_format=torch.contiguous_format)
            ops.append(dist.P2POp(dist.irecv, got_b, _peer(group, r + 1), group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return got_t, got_b


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
    Transport: the peer-memory mailbox (`_Mailbox`) on CUDA tensors, NCCL/gloo grouped send/recv otherwise."""
    dim = dim % x.dim()
    got_t, got_b = _exchange(x, dim, 0, x.shape[dim], 0, top, bottom, group)
    parts = []
    if top:
        parts.append(got_t if got_t is not None else x.new_zeros(x.narrow(dim, 0, top).shape))
    parts.append(x)
    if bottom:
        parts.append(got_b if got_b is not None else x.new_zeros(x.narrow(dim, 0, bottom).shape))
    return torch.cat(parts, dim)


def fill_row_halo_(x, pad, top, bottom, dim=-3, group=None):
    """In-place halo exchange for activations that were ALLOCATED with `pad` spare rows above and below the band's rows
    along `dim` (ops.gn_apply_tc3(pad=...), ops.cost_volume_concat_c8s3(pad=...)): fills the `top` rows just above the
    band with the previous rank's bottom edge and the `bottom` rows just below with the next rank's top edge (zeros at
    the image border = the conv's zero padding).  Only the boundary rows move; the activation itself is never copied."""
    dim = dim % x.dim()
    rows = x.shape[dim] - 2 * pad
    if top > pad or bottom > pad:
        raise ValueError("halo (%d, %d) exceeds the %d spare rows" % (top, bottom, pad))
    push = getattr(x, "_halo_push", None)
    if push is not None:  # the producing kernel already pushed `pad` boundary rows to the neighbours: wait + local copy
        if not push.done:
            mb, n, r = push.mb, push.mb.n, push.mb.r
            if mb.k - push.k >= 2:
                raise RuntimeError("halo push consumed too late: its mailbox slot has been reused (%d exchanges ago)"
                                   % (mb.k - push.k))
            mb.barrier(push.ch)
            like = x.narrow(dim, 0, pad)
            lo, hi = x.narrow(dim, 0, pad), x.narrow(dim, pad + rows, pad)
            _copy_rows(lo, mb.mine(push.off, like)) if r > 0 else lo.zero_()             # the previous rank's bottom rows
            _copy_rows(hi, mb.mine(push.off + push.nb, like)) if r < n - 1 else hi.zero_()  # the next rank's top rows
            push.done = True
        return x
    got_t, got_b = _exchange(x, dim, pad, pad + rows, 0, top, bottom, group)
    if top:
        dst = x.narrow(dim, pad - top, top)
        _copy_rows(dst, got_t) if got_t is not None else dst.zero_()
    if bottom:
        dst = x.narrow(dim, pad + rows, bottom)
        _copy_rows(dst, got_b) if got_b is not None else dst.zero_()
    return x


def allreduce_gn_sums(sums, group=None, scale=1.0):
    """Sum per-band GroupNorm statistics ([B,C,2] double: sum, sum of squares) over all bands, in place, times `scale`.
    Peer-memory path: every rank publishes its sums, one barrier, ONE kernel adds the n buffers in rank order (identical
    result on every rank); otherwise one all-reduce."""
    n = world(group)
    if n == 1:
        return sums if scale == 1.0 else sums.mul_(scale)
    mb = _mailbox(group, sums.device, _align(sums.numel() * sums.element_size())) if sums.is_cuda else None
    if mb is None:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
        return sums if scale == 1.0 else sums.mul_(scale)
    off, ch = mb.slot()
    mb.mine(off, sums).copy_(sums)
    mb.barrier(ch)
    if n <= 16 and sums.is_contiguous():
        from . import ops
        return ops.sum_peers(sums, [mb.peer(r, off, sums) for r in range(n)], scale)
    sums.copy_(mb.peer(0, off, sums))
    for r in range(1, n):
        sums.add_(mb.peer(r, off, sums))
    return sums if scale == 1.0 else sums.mul_(scale)


def gather_bands(x, dim=-2, group=None):
    """All-gather equally sized bands along `dim` (every rank gets the full tensor)."""
    n = world(group)
    if n == 1:
        return x
    x = x.contiguous()
    mb = _mailbox(group, x.device, _align(x.numel() * x.element_size())) if x.is_cuda else None
    if mb is None:
        parts = [torch.empty_like(x) for _ in range(n)]
        dist.all_gather(parts, x, group=group)
        return torch.cat(parts, dim % x.dim())
    off, ch = mb.slot()
    mb.mine(off, x).copy_(x)
    mb.barrier(ch)
    return torch.cat([mb.peer(r, off, x) for r in range(n)], dim % x.dim())
