"""Global-negative InfoNCE and sharded OT across the GPUs of one box.

The reference never gathers embeddings: its ``utils.gather_tensors`` (utils.py:192-206) has no call
site and under DDP every rank scores local negatives only (train.py:222-225).  This module is the
"every rank scores against the global negative set" variant BASELINE.json asks for; its oracle is
the reference loss evaluated single-process on the rank-order concatenation of all ranks' inputs.

Partitioning (column shard by text, one exchange step):
  rank r owns images [r*b, (r+1)*b) and their b*T descriptions.
  1. all_gather(image_features)              B_g*D elements       (+ labels_per_image)
  2. local fused GEMM  L_r = s * I^_all T^_r^t  -> per-row (max, sum, positive logit) over the local
     columns; the text-side cross-entropy of the local positive columns is complete locally
  3. all_gather(row statistics [B_g, 4]) + [4] text-side sums     -> loss_i, loss_t on every rank
  4. backward: G_r local; dtxt_r = s G_r^t I^_all needs no exchange;
     reduce_scatter(s G_r T^_r  [B_g, D] fp32) -> d I^ rows of this rank -> normalisation backward
     all_reduce(dlogit_scale)
OT shards by sample with no data-path exchange; only the scalar loss is all-reduced.

The collectives are torch.distributed calls on the current stream (NCCL over NVLink on B200, gloo
in the CPU tests).  ``compute`` is the per-rank kernel backend: :class:`CudaBackend` in production;
the CPU tests inject an oracle-backed stand-in to check the choreography with world_size 2.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from . import _lib as L
from . import functional as F_


class CudaBackend:
    """Per-rank compute through the C ABI phases (include/clip_event_b200.h)."""

    def __init__(self):
        self.lib = L.load()

    def fwd_partial(self, img_all, txt, ls, labels_i_all, labels_t, index_pos, col_offset):
        R, D = img_all.shape
        C, P = txt.shape[0], index_pos.numel()
        dt = L.dtype_code(img_all.dtype)
        nbytes = self.lib.ce_contrastive_workspace_bytes(R, C, P, D, dt)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=img_all.device)
        stats = torch.empty(R * 4 + 4, dtype=torch.float32, device=img_all.device)   # [row_part | sums]
        row_part, sums = stats[: R * 4], stats[R * 4:]
        L.check(self.lib.ce_contrastive_fwd_partial(
            img_all.data_ptr(), txt.data_ptr(), ls.data_ptr(), labels_i_all.data_ptr(), labels_t.data_ptr(),
            index_pos.data_ptr(), R, C, P, D, int(col_offset), L.CE_IMG_CE_OVERBATCH, 1, 0, dt,
            row_part.data_ptr(), sums.data_ptr(),
            ws.data_ptr(), nbytes, L.stream_ptr()), "contrastive fwd_partial")
        return stats, (ws, R, C, P, D, dt)

    def fwd_finish(self, stats_all, world, state):
        """stats_all: [world, R*4 + 4] -- the all-gathered per-rank [row_part | sums] records."""
        ws, R, C, P, D, dt = state
        out = torch.empty(2, dtype=torch.float32, device=ws.device)
        L.check(self.lib.ce_contrastive_fwd_finish(
            stats_all.data_ptr(), stats_all.data_ptr() + R * 16, R * 4 + 4, world, R, C, P, D, dt, out.data_ptr(),
            out.data_ptr() + 4, ws.data_ptr(), ws.numel(), L.stream_ptr()), "contrastive fwd_finish")
        return out[0], out[1]

    def bwd_partial(self, img_all, txt, ls, labels_i_all, labels_t, index_pos, col_offset, g_i, g_t,
                    R_total, P_total, state, dimg_out=None, dls_out=None):
        ws, R, C, P, D, dt = state
        dtxt = torch.empty_like(txt)
        dimg_hat = dimg_out if dimg_out is not None else torch.empty(R, D, dtype=torch.float32, device=txt.device)
        dls = dls_out if dls_out is not None else torch.empty(1, dtype=torch.float32, device=txt.device)
        L.check(self.lib.ce_contrastive_bwd_partial(
            img_all.data_ptr(), txt.data_ptr(), ls.data_ptr(), labels_i_all.data_ptr(), labels_t.data_ptr(),
            index_pos.data_ptr(), R, C, P, D, int(col_offset), L.CE_IMG_CE_OVERBATCH, 1, 0, dt,
            g_i.data_ptr(), g_t.data_ptr(),
            int(R_total), int(P_total), dtxt.data_ptr(), dimg_hat.data_ptr(), dls.data_ptr(),
            ws.data_ptr(), ws.numel(), L.stream_ptr()), "contrastive bwd_partial")
        return dtxt, dimg_hat, dls

    def bwd_finish(self, img_rows, dimg_hat_rows):
        rows, D = img_rows.shape
        out = torch.empty_like(img_rows)
        L.check(self.lib.ce_contrastive_bwd_finish(
            img_rows.data_ptr(), dimg_hat_rows.data_ptr(), rows, D, L.dtype_code(img_rows.dtype),
            out.data_ptr(), L.stream_ptr()), "contrastive bwd_finish")
        return out


def shard_bounds(n_global: int, world: int, rank: int):
    """Contiguous rank-order shard [lo, hi) of n_global items (n_global must divide evenly)."""
    if n_global % world != 0:
        raise RuntimeError("global batch %d is not divisible by world size %d" % (n_global, world))
    per = n_global // world
    return rank * per, (rank + 1) * per


def global_labels_for_rank(b_local: int, T: int, world: int, rank: int, device=None):
    """The reference's label contract (dataset_voa.py:617-663) restated for rank-order shards:
    labels_per_image are GLOBAL column indices, labels_per_text GLOBAL row indices, index_pos
    LOCAL column indices."""
    rows = torch.arange(rank * b_local, (rank + 1) * b_local, dtype=torch.int64, device=device)
    return rows * T, rows.repeat_interleave(T), torch.arange(b_local, dtype=torch.int64, device=device) * T


def _reduce_scatter_sum(out, inp, rank, group):
    """reduce_scatter (NCCL); gloo has none, so the CPU tests all-reduce and slice."""
    if dist.get_backend(group) == "gloo":
        tmp = inp.clone()
        dist.all_reduce(tmp, op=dist.ReduceOp.SUM, group=group)
        out.copy_(tmp.view(-1, *out.shape)[rank])
    else:
        dist.reduce_scatter_tensor(out, inp, op=dist.ReduceOp.SUM, group=group)


def _reduce_scatter_sums(pairs, rank, group):
    """Several reduce-scatters as ONE NCCL launch (c10d's coalescing fast path: one ncclGroup)."""
    if dist.get_backend(group) == "gloo":
        for out, inp in pairs:
            _reduce_scatter_sum(out, inp, rank, group)
        return
    with dist._coalescing_manager(group=group):
        for out, inp in pairs:
            dist.reduce_scatter_tensor(out, inp, op=dist.ReduceOp.SUM, group=group)


# --------------------------------------------------------------------------------------------
# exchange steps: NCCL collectives, or loads from the peers' symmetric buffers over NVLink
# --------------------------------------------------------------------------------------------
class NcclExchange:
    """The three exchange steps of the sharded loss head as collectives of the process group (NCCL; gloo in
    the CPU tests)."""

    def __init__(self, group):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    def gather_rows(self, x):
        out = torch.empty(self.world * x.numel(), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(out, x.reshape(-1), group=self.group)
        return out

    def gather_stats(self, stats):
        out = torch.empty(self.world * stats.numel(), dtype=torch.float32, device=stats.device)
        dist.all_gather_into_tensor(out, stats, group=self.group)
        return out.view(self.world, stats.numel())

    def grad_buffers(self, R, D, device):
        return None, None

    def reduce_scatter(self, dimg_hat, dls, b):
        mine = torch.empty(b, dimg_hat.shape[1], dtype=torch.float32, device=dimg_hat.device)
        dls_tot = torch.empty(1, dtype=torch.float32, device=dimg_hat.device)
        _reduce_scatter_sums([(mine, dimg_hat), (dls_tot, dls.expand(self.world).contiguous())], self.rank, self.group)
        return mine, dls_tot


class SymmExchange:
    """The same three steps without a collective library in the data path: every rank keeps its contribution in a
    symmetric-memory buffer (torch.distributed._symmetric_memory: one allocation mapped into every peer), a
    cross-rank barrier orders the step, and a copy / reduce kernel of this library reads the peers' buffers
    over NVLink (``ce_p2p_gather`` / ``ce_p2p_reduce_f32``).  The gradient GEMMs write their partial d I^
    straight into the symmetric buffer; the reduce-scatter is one kernel that sums this rank's rows over the
    peers in rank order.  Buffers are created (a collective rendezvous) on first use per shape and reused.

    Ordering: a buffer written in step n+1 was last read by the peers before their next barrier arrival, which
    precedes this rank passing that barrier -- three barriers per step keep writers behind readers."""

    _CACHE = {}

    def __init__(self, group):
        import torch.distributed._symmetric_memory as symm
        self.symm = symm
        self.group = dist.group.WORLD if group is None else group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.lib = L.load()

    def _buf(self, tag, numel, dtype, device):
        key = (id(self.group), tag, int(numel), dtype, str(device))
        hit = SymmExchange._CACHE.get(key)
        if hit is None:
            t = self.symm.empty(int(numel), dtype=dtype, device=device)
            h = self.symm.rendezvous(t, self.group)
            import ctypes
            ptrs = (ctypes.c_int64 * self.world)(*[int(p) for p in h.buffer_ptrs])
            hit = SymmExchange._CACHE[key] = (t, h, ptrs)
        return hit

    def _gather(self, tag, x):
        flat = x.reshape(-1)
        n = flat.numel()
        pad = (-n * flat.element_size()) % 16 // flat.element_size()
        t, h, ptrs = self._buf(tag, n + pad, flat.dtype, flat.device)
        t[:n].copy_(flat)
        h.barrier(channel=0)
        out = torch.empty(self.world * (n + pad), dtype=flat.dtype, device=flat.device)
        L.check(self.lib.ce_p2p_gather(ptrs, self.world, (n + pad) * flat.element_size(), out.data_ptr(), L.stream_ptr()),
                "p2p gather")
        return out.view(self.world, n + pad)[:, :n]

    def gather_rows(self, x):
        return self._gather("img", x).reshape(-1)

    def gather_stats(self, stats):
        return self._gather("stats", stats)

    def grad_buffers(self, R, D, device):
        """[R, D] fp32 partial gradient + one slot for dlogit_scale, both inside this rank's symmetric buffer."""
        t, _, _ = self._buf("dimg", R * D + 16, torch.float32, device)
        return t[: R * D].view(R, D), t[R * D: R * D + 1]

    def reduce_scatter(self, dimg_hat, dls, b):
        R, D = dimg_hat.shape
        t, h, ptrs = self._buf("dimg", R * D + 16, torch.float32, dimg_hat.device)
        if dimg_hat.data_ptr() != t.data_ptr():      # the caller did not use grad_buffers()
            t[: R * D].copy_(dimg_hat.reshape(-1))
            t[R * D: R * D + 1].copy_(dls.reshape(1))
        h.barrier(channel=0)
        mine = torch.empty(b, D, dtype=torch.float32, device=dimg_hat.device)
        dls_tot = torch.empty(1, dtype=torch.float32, device=dimg_hat.device)
        L.check(self.lib.ce_p2p_reduce_f32(ptrs, self.world, self.rank * b * D, b * D, mine.data_ptr(), R * D, 1,
                                           dls_tot.data_ptr(), L.stream_ptr()), "p2p reduce")
        return mine, dls_tot


_EXCHANGES = {}


def make_exchange(group, device):
    """Symmetric-memory exchange on CUDA when torch offers it (``CE_DIST_EXCHANGE=nccl`` forces the collectives),
    NCCL / gloo collectives otherwise."""
    import os
    key = (id(group), str(device))
    ex = _EXCHANGES.get(key)
    if ex is None:
        want = os.environ.get("CE_DIST_EXCHANGE", "symm")
        ex = None
        if want != "nccl" and torch.device(device).type == "cuda" and dist.get_backend(group) == "nccl" and dist.get_world_size(group) <= 16:
            try:
                ex = SymmExchange(group)
                ex._buf("probe", 4, torch.float32, device)        # a collective: every rank takes the same branch or none
            except Exception:
                ex = None
        if ex is None:
            ex = NcclExchange(group)
        _EXCHANGES[key] = ex
    return ex


_CANONICAL = {}


def _canonical_labels(n: int, T: int, device):
    """arange(n) * T, built once per (n, T, device): two launches less in every step."""
    key = (n, T, str(device))
    t = _CANONICAL.get(key)
    if t is None:
        t = _CANONICAL[key] = torch.arange(n, dtype=torch.int64, device=device) * T
    return t


class _GlobalContrastive(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, txt, logit_scale, labels_i, labels_t, index_pos, group, compute, ddp_average=False):
        world = dist.get_world_size(group)
        rank = dist.get_rank(group)
        dev = img.device
        img_c, txt_c = img.detach().contiguous(), txt.detach().contiguous()
        ls = logit_scale.detach().to(device=dev, dtype=torch.float32).reshape(1).contiguous()
        labels_i = None if labels_i is None else labels_i.to(device=dev, dtype=torch.int64).contiguous()
        labels_t = labels_t.to(device=dev, dtype=torch.int64).contiguous()
        index_pos = index_pos.to(device=dev, dtype=torch.int64).contiguous()
        b, D = img_c.shape
        C = txt_c.shape[0]
        # 1. gather images and their labels (texts stay local)
        ex = make_exchange(group, dev)
        img_all = ex.gather_rows(img_c).view(world * b, D)
        if labels_i is None:
            # canonical contract (dataset_voa.py:617-621): image r's positive is column r*T -- no exchange
            lab_all = _canonical_labels(world * b, C // b, dev)
        else:
            lab_all = torch.empty(world * b, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(lab_all, labels_i, group=group)
        # 2. local GEMM + statistics
        col_offset = rank * C
        stats, state = compute.fwd_partial(img_all, txt_c, ls, lab_all, labels_t, index_pos, col_offset)
        # 3. exchange the statistics: one record [row_part (R x 4) | sums (4)] per rank
        stats_all = ex.gather_stats(stats).contiguous()
        loss_i, loss_t = compute.fwd_finish(stats_all, world, state)
        ctx.saved = (img_all, txt_c, ls, lab_all, labels_t, index_pos, col_offset, state)
        ctx.ex = ex
        ctx.meta = (group, compute, world, rank, b, logit_scale.dtype, logit_scale.shape, bool(ddp_average))
        return loss_i, loss_t

    @staticmethod
    def backward(ctx, g_i, g_t):
        img_all, txt_c, ls, lab_all, labels_t, index_pos, col_offset, state = ctx.saved
        group, compute, world, rank, b, ls_dtype, ls_shape, ddp_average = ctx.meta
        dev = txt_c.device
        def scalar(g):   # a missing upstream gradient is a zero; no launch when both are present
            if g is None:
                return torch.zeros(1, dtype=torch.float32, device=dev)
            return g.detach().to(torch.float32).reshape(1).contiguous()
        gi, gt = scalar(g_i), scalar(g_t)
        if ddp_average:
            # DDP (train.py:222-225) AVERAGES parameter gradients over ranks.  The feature gradients
            # below are d(global-mean loss)/d(local features): their per-rank contributions to an
            # encoder parameter must be SUMMED, so they are handed to DDP multiplied by the world
            # size; logit_scale is replicated and already carries the full (all-reduced) gradient,
            # whose average over ranks is itself.
            gi, gt = gi * world, gt * world
        R_total = img_all.shape[0]
        P_total = index_pos.numel() * world   # ranks hold equal shards
        ex = ctx.ex
        dimg_buf, dls_buf = ex.grad_buffers(R_total, img_all.shape[1], dev)
        extra = () if dimg_buf is None else (dimg_buf, dls_buf)
        dtxt, dimg_hat, dls = compute.bwd_partial(img_all, txt_c, ls, lab_all, labels_t, index_pos, col_offset,
                                                  gi, gt, R_total, P_total, state, *extra)
        # gradient return: reduce-scatter of s G T^_local with the dlogit_scale partials riding along (one NCCL
        # launch, or one kernel over the peers' symmetric buffers)
        mine, dls_tot = ex.reduce_scatter(dimg_hat, dls, b)
        if ddp_average:
            dls_tot = dls_tot / world
        dimg = compute.bwd_finish(img_all[rank * b:(rank + 1) * b], mine)
        return dimg, dtxt, dls_tot.reshape(ls_shape).to(ls_dtype), None, None, None, None, None, None


def global_contrastive(image_features, text_features, logit_scale, labels_per_image, labels_per_text,
                       index_pos, group=None, compute=None, ddp_average=False):
    """loss_i, loss_t over the GLOBAL batch; gradients for this rank's images and descriptions.

    ``labels_per_image`` are global column indices (None = the canonical ``row * T``, which needs
    no exchange), ``labels_per_text`` global row indices and ``index_pos`` local column indices
    (see :func:`global_labels_for_rank`).

    Gradient convention: by default every rank receives exactly its slice of the gradient of the
    GLOBAL-mean loss (the oracle is the single-process reference on the concatenated batch).  With
    ``ddp_average=True`` the feature gradients are multiplied by the world size, which is what an
    encoder wrapped in DistributedDataParallel (gradient AVERAGING, train.py:222-225) needs to end up
    with the true parameter gradient; ``logit_scale`` gets the full gradient in both modes.
    """
    if compute is None:
        compute = CudaBackend()
    return _GlobalContrastive.apply(image_features, text_features, logit_scale, labels_per_image,
                                    labels_per_text, index_pos, group, compute, ddp_average)


def sharded_alignment(entitytxt_vec, object_vec, entitytxt_num, object_num, group=None, ot_fn=None):
    """loss_ot = 0.01 * sum over the GLOBAL batch of the OT distance: local kernel + one scalar
    all-reduce.  The gradient of the summed loss w.r.t. local nodes needs no exchange."""
    fn = ot_fn if ot_fn is not None else F_.ot_alignment
    loss, _ = fn(entitytxt_vec, object_vec, entitytxt_num, object_num)
    total = loss.detach().clone()
    dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    # value = global sum, gradient = local term (d total / d local nodes == d local loss / d local nodes)
    return loss + (total - loss.detach())


# --------------------------------------------------------------------------------------------
# the whole sharded loss head in one call: three collectives, two streams, eager gradients
# --------------------------------------------------------------------------------------------
def _cuda_ot_eager(etxt, obj, tnum, onum, need_grad, stream_ptr):
    tm, kind = F_._mask_args(F_.num_mask(tnum).contiguous())
    om, _ = F_._mask_args(F_.num_mask(onum).contiguous())
    loss, dist_b, detxt, dobj, gbuf, ws = F_._ot_launch(etxt, obj, tm, om, kind, True, F_.IPOT_BETA, F_.IPOT_ITERS,
                                                       F_.IPOT_K, F_.OT_LOSS_WEIGHT, need_grad, stream_ptr)
    return loss, detxt, dobj, (gbuf, ws, dist_b)


class _GlobalLossHeadStep(torch.autograd.Function):
    """Sharded version of ``functional._LossHeadStep``: (loss_i, loss_t, loss_ot) over the GLOBAL batch
    and the gradients of their sum for this rank's inputs, formed in the forward call.

    Exchange steps (NCCL launches) per step: all-gather of the image embeddings, all-gather of the
    per-rank statistics record -- the local OT loss rides in its spare slot -- and ONE coalesced
    reduce-scatter carrying both the image-gradient return and the dlogit_scale partials.  The OT
    chain runs on the library's side stream next to the GEMM chain and needs no exchange of its own.
    """

    @staticmethod
    def forward(ctx, img, txt, logit_scale, etxt, obj, labels_i, labels_t, index_pos, tnum, onum, group, compute,
                ot_eager, ddp_average):
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        dev = img.device
        cuda = img.is_cuda
        img_c, txt_c = img.detach().contiguous(), txt.detach().contiguous()
        etxt_c, obj_c = etxt.detach().contiguous(), obj.detach().contiguous()
        ls = logit_scale.detach().to(device=dev, dtype=torch.float32).reshape(1).contiguous()
        labels_i = None if labels_i is None else labels_i.to(device=dev, dtype=torch.int64).contiguous()
        labels_t = labels_t.to(device=dev, dtype=torch.int64).contiguous()
        index_pos = index_pos.to(device=dev, dtype=torch.int64).contiguous()
        b, D = img_c.shape
        C = txt_c.shape[0]
        need_c, need_o = any(ctx.needs_input_grad[:3]), any(ctx.needs_input_grad[3:5])
        # OT chain first, on the side stream
        if cuda:
            cur, side = torch.cuda.current_stream(), F_.side_stream(dev)
            side.wait_stream(cur)
            loss_ot_local, detxt, dobj, keep = ot_eager(etxt_c, obj_c, tnum, onum, need_o, side.cuda_stream)
            ot_done = torch.cuda.Event()
            ot_done.record(side)
        else:
            loss_ot_local, detxt, dobj, keep = ot_eager(etxt_c, obj_c, tnum, onum, need_o, 0)
        # 1. gather the images
        ex = make_exchange(group, dev)
        img_all = ex.gather_rows(img_c).view(world * b, D)
        if labels_i is None:
            lab_all = _canonical_labels(world * b, C // b, dev)
        else:
            lab_all = torch.empty(world * b, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(lab_all, labels_i, group=group)
        # 2. local GEMM + statistics
        col_offset = rank * C
        stats, state = compute.fwd_partial(img_all, txt_c, ls, lab_all, labels_t, index_pos, col_offset)
        # 3. one record per rank: [row_part (R x 4) | text-side sum, P, local OT loss, 0]
        R = world * b
        if cuda:
            cur.wait_event(ot_done)
        stats[R * 4 + 2:R * 4 + 3].copy_(loss_ot_local.reshape(1).to(torch.float32))
        stats_all = ex.gather_stats(stats).contiguous()
        loss_i, loss_t = compute.fwd_finish(stats_all, world, state)
        loss_ot = stats_all[:, R * 4 + 2].sum()
        # 4. gradients of (loss_i + loss_t) for unit upstream gradients
        dimg = dtxt = dls_tot = None
        if need_c:
            one = F_.unit_gradient(dev)
            g = one * world if ddp_average else one
            dimg_buf, dls_buf = ex.grad_buffers(R, D, dev)
            extra = () if dimg_buf is None else (dimg_buf, dls_buf)
            dtxt, dimg_hat, dls = compute.bwd_partial(img_all, txt_c, ls, lab_all, labels_t, index_pos, col_offset,
                                                      g, g, R, index_pos.numel() * world, state, *extra)
            mine, dls_tot = ex.reduce_scatter(dimg_hat, dls, b)
            if ddp_average:
                dls_tot = dls_tot / world
            dimg = compute.bwd_finish(img_all[rank * b:(rank + 1) * b], mine)
        if cuda:
            cur.wait_stream(side)
        ctx.stash = (dimg, dtxt, dls_tot, detxt, dobj, keep)
        ctx.ls_meta = (logit_scale.dtype, logit_scale.shape)
        ctx.set_materialize_grads(False)
        return loss_i, loss_t, loss_ot.to(loss_i.dtype)

    @staticmethod
    def backward(ctx, g_i, g_t, g_ot):
        if getattr(ctx, "consumed", False):
            raise RuntimeError("clip_event_b200 sharded loss head step: backward a second time; run the step again")
        ctx.consumed = True
        dimg, dtxt, dls, detxt, dobj, _ = ctx.stash
        ctx.stash = None
        out = [None] * 14
        use_c = dimg is not None and (g_i is not None or g_t is not None)
        use_o = detxt is not None and g_ot is not None
        if use_c and (g_i is None or g_t is None):
            raise RuntimeError("loss_i and loss_t must be back-propagated together on the fused step")
        if (use_c and dimg.is_cuda) or (use_o and detxt.is_cuda):
            # ONE scale launch over every gradient buffer; it returns on the device when the upstream gradients are 1
            F_.head_step_scale((dimg, dtxt, dls) if use_c else (), (detxt, dobj) if use_o else (),
                               g_i if use_c else None, g_t if use_c else None, g_ot if use_o else None,
                               "sharded loss head step backward")
        else:
            if use_c:
                gf, gtf = g_i.detach().float(), g_t.detach().float()
                g = torch.where(gf == gtf, gf, torch.full_like(gf, float("nan")))
                dimg, dtxt, dls = (dimg.float() * g).to(dimg.dtype), (dtxt.float() * g).to(dtxt.dtype), dls * g
            if use_o:
                g = g_ot.detach().float()
                detxt, dobj = (detxt.float() * g).to(detxt.dtype), (dobj.float() * g).to(dobj.dtype)
        if use_c:
            ls_dtype, ls_shape = ctx.ls_meta
            out[0], out[1], out[2] = dimg, dtxt, dls.reshape(ls_shape).to(ls_dtype)
        if use_o:
            out[3], out[4] = detxt, dobj
        return tuple(out)


def global_loss_head_step(image_features, text_features, logit_scale, labels_per_image, labels_per_text, index_pos,
                          entitytxt_vec, object_vec, entitytxt_num, object_num, group=None, compute=None,
                          ot_eager=None, ddp_average=False):
    """(loss_i, loss_t, loss_ot) over the global batch in one call (labels as in :func:`global_contrastive`).
    loss_ot is the GLOBAL sum (replicated); its gradient w.r.t. this rank's nodes is the local term, which is
    what the reference's per-rank loss under DDP averaging corresponds to (no scaling in either mode)."""
    if compute is None:
        compute = CudaBackend()
    if ot_eager is None:
        ot_eager = _cuda_ot_eager
    return _GlobalLossHeadStep.apply(image_features, text_features, logit_scale, entitytxt_vec, object_vec,
                                     labels_per_image, labels_per_text, index_pos, entitytxt_num, object_num,
                                     group, compute, ot_eager, ddp_average)
