"""Particles sharded over the GPUs of one node, one process per GPU (torch.distributed, NCCL over
NVLink / NVSwitch); the map, likelihood table, scan and parameters are replicated.

The reference is single-process; every per-particle kernel shards trivially, and the
only exchange steps are the ones SURVEY 8(e) lists:

  softmax   all-reduce MAX of the score maxima, all-reduce SUM of sum exp(s - max)   (2 x 2 doubles)
  estimate  all-reduce SUM of the 6 raw and the 9 central weighted sums              (15 doubles)
  resample  all-reduce MAX of the weight maximum (fixed-point scale), all-gather of the per-rank
            fixed-point weight totals -> exclusive prefix offsets; every rank then emits the
            offspring whose thresholds fall into its own cumulative-weight interval and an
            all-to-all over NVLink returns an equal share of outputs to every rank, so per-rank
            counts stay fixed and output slot m holds the same particle for 1, 2, 4 or 8 ranks
            (the fixed-point cumulative sums are exact integers, hence decomposition-independent).

Random draws are keyed by the GLOBAL particle index (first_index + i), so a sharded run consumes
the same Philox streams as a single-GPU run of the same total size.
"""
import ctypes as C
import math

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .localizer import Localizer, _dbl3, _ptr, assemble_estimate, compute_motion


# ------------------------------------------------------------------------------------ host planning
def resample_threshold(m, r, n_out, grand_total):
    """T_m = ceil((r + m * (1/n_out)) * float(total)) -- the same IEEE operations as the kernel
    k_search_fixed (f64 multiply, add, multiply, ceil; Python floats never fuse)."""
    step = 1.0 / n_out
    t = math.ceil((r + m * step) * float(grand_total))
    return min(max(t, 0), (1 << 64) - 1)


def count_thresholds_le(x, r, n_out, grand_total):
    """#{m in [0, n_out) : T_m <= x}; T_m is non-decreasing in m -> binary search."""
    lo, hi = 0, n_out
    while lo < hi:
        mid = (lo + hi) // 2
        if resample_threshold(mid, r, n_out, grand_total) <= x:
            lo = mid + 1
        else:
            hi = mid
    return lo


def plan_resample(totals, r, n_out, world):
    """Given every rank's fixed-point weight total, decide which global output slots each rank emits.

    Returns (offsets, grand_total, m_lo, m_hi, send) with send[k][d] = number of outputs rank k sends
    to rank d (outputs are dealt out in equal contiguous shares of n_out / world)."""
    totals = [int(t) for t in totals]
    offsets, acc = [], 0
    for t in totals:
        offsets.append(acc)
        acc += t
    grand = acc
    m_lo, m_hi = [], []
    for k in range(world):
        lo = 0 if k == 0 else count_thresholds_le(offsets[k], r, n_out, grand)
        hi = n_out if k == world - 1 else count_thresholds_le(offsets[k] + totals[k], r, n_out, grand)
        m_lo.append(lo)
        m_hi.append(max(hi, lo))
    # ranks without weight emit nothing; make the ranges a partition of [0, n_out)
    for k in range(1, world):
        m_lo[k] = max(m_lo[k], m_hi[k - 1])
        m_hi[k] = max(m_hi[k], m_lo[k])
    m_hi[world - 1] = n_out
    share = n_out // world
    send = [[0] * world for _ in range(world)]
    for k in range(world):
        for d in range(world):
            a, b = d * share, (d + 1) * share if d < world - 1 else n_out
            send[k][d] = max(0, min(m_hi[k], b) - max(m_lo[k], a))
    return offsets, grand, m_lo, m_hi, send


def exchange(send_buf, send_counts, recv_counts, group=None):
    """all_to_all_single with split sizes along dim 0 (NCCL on GPU tensors, gloo on CPU tensors)."""
    out = send_buf.new_empty((int(sum(recv_counts)),) + tuple(send_buf.shape[1:]))
    dist.all_to_all_single(out, send_buf, output_split_sizes=[int(c) for c in recv_counts],
                           input_split_sizes=[int(c) for c in send_counts], group=group)
    return out


# ------------------------------------------------------------------------------------ the filter
class ShardedLocalizer(Localizer):
    """Localizer whose particle set is the union over ranks of equally sized shards."""

    def __init__(self, device=0, params=None, mode=None, seed=0, resample_mode="fixed", max_attempts=1000,
                 group=None, peer_push=True, native_comm=True, allow_fallback=False):
        """peer_push / native_comm select the exchange implementation explicitly.  If the selected one cannot be set
        up (no symmetric memory, ...) construction RAISES unless allow_fallback=True, in which case the next
        simpler one is used and the reason is kept in `symm_error`."""
        self.allow_fallback = allow_fallback
        # "fixed": 64-bit fixed-point sums (the default: independent of the decomposition, cheapest exchanges);
        # "reference": pu:416-446's sequential float32 sums, continued from rank to rank inside the persistent step
        # kernel -- bit-identical to the single-GPU reference arithmetic; step() / step_staged() only, native exchanges
        if resample_mode not in ("fixed", "reference"):
            raise ValueError(resample_mode)
        if resample_mode == "reference" and not (peer_push and native_comm):
            raise ValueError("sharded reference-arithmetic resampling needs the native peer-memory exchanges")
        self.group = group
        self.use_peer_push = peer_push
        self.use_native_comm = native_comm and peer_push
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        super().__init__(device=device, params=params, mode=mode, seed=seed, resample_mode=resample_mode,
                         max_attempts=max_attempts)

    # -- buffers -----------------------------------------------------------------------------
    def _make_sets(self, n):
        """Pose buffers in symmetric memory (every rank can store into every other rank's buffers over
        NVLink): lets resampling push offspring straight to their destination.  Falls back to private
        buffers + NCCL all-to-all when symmetric memory is unavailable."""
        self.symm = None
        self.peer_tab = None
        if self.use_peer_push:
            try:
                import torch.distributed._symmetric_memory as symm_mem
                buf = symm_mem.empty(9 * n, dtype=torch.float64, device=self.device)
                hdl = symm_mem.rendezvous(buf, self.group if self.group is not None else dist.group.WORLD)
                ptrs = [int(p) for p in hdl.buffer_ptrs]
                tab = torch.tensor([[[ptrs[d] + (3 * j + k) * n * 8 for d in range(self.world)] for k in range(3)]
                                    for j in range(3)], dtype=torch.int64)
                self.peer_tab = tab.to(self.device)                      # [set][component][rank]
                self.symm, self.symm_buf = hdl, buf
                return [[buf[(3 * j + k) * n:(3 * j + k + 1) * n] for k in range(3)] for j in range(3)]
            except Exception as e:      # noqa: BLE001
                if not self.allow_fallback:
                    raise RuntimeError("ShardedLocalizer: symmetric (peer-mapped) pose buffers unavailable: %r; pass "
                                       "peer_push=False or allow_fallback=True for the NCCL all-to-all path" % (e,)) from e
                self.symm_error = repr(e)
                self.symm = None
        return super()._make_sets(n)

    def _alloc(self, n):
        self.first_index = self.rank * n
        super()._alloc(n)
        self.n_global = n * self.world
        self.native = False
        if self.symm is not None and self.use_native_comm:
            try:
                self._init_native_comm(n)
            except Exception as e:      # noqa: BLE001
                if not self.allow_fallback:
                    raise RuntimeError("ShardedLocalizer: peer-memory mailboxes unavailable: %r; pass native_comm=False "
                                       "or allow_fallback=True for NCCL scalar exchanges" % (e,)) from e
                self.symm_error = repr(e)
                self.native = False
        d = self.device
        self.st_post = torch.zeros(4, dtype=torch.float64, device=d)   # {max, sum, sum as 2^-40 fixed point, -}
        self.st_pre = torch.zeros(4, dtype=torch.float64, device=d)
        self.m9 = torch.zeros(9, dtype=torch.float64, device=d)
        self.c9 = torch.zeros(9, dtype=torch.float64, device=d)
        self.wmax = torch.zeros(1, dtype=torch.float32, device=d)
        self.total = torch.zeros(1, dtype=torch.int64, device=d)
        self.totals_all = torch.zeros(self.world, dtype=torch.int64, device=d)
        self.send_cap = 0
        self.send = None

    def _init_native_comm(self, n):
        """Hand the peer-mapped mailboxes and pose buffers to the library: from here on every exchange of the
        step runs inside libmcl's own kernels over NVLink peer memory (mcl_comm_init), and the whole sharded
        step is sequenced by the same C calls as the single-GPU one."""
        import torch.distributed._symmetric_memory as symm_mem
        grp = self.group if self.group is not None else dist.group.WORLD
        self.mail = symm_mem.empty(2048, dtype=torch.int64, device=self.device)
        self.mail.zero_()
        self.mail_hdl = symm_mem.rendezvous(self.mail, grp)
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)                                   # every mailbox is zeroed before first use
        W = self.world
        peers = (C.c_uint64 * W)(*[int(p) for p in self.mail_hdl.buffer_ptrs])
        ptrs = [int(p) for p in self.symm.buffer_ptrs]
        pose = (C.c_uint64 * (9 * W))(*[ptrs[d] + (3 * j + k) * n * 8 for j in range(3) for k in range(3) for d in range(W)])
        self.h.call("mcl_comm_init", self.rank, W, C.c_void_p(self.mail.data_ptr()), peers, pose)
        self.native = True

    def comm_error(self):
        e = C.c_int(0)
        self.h.call("mcl_comm_status", C.byref(e))
        return e.value

    def _set_roles(self, cur, prev, spare, ws, tick):
        self.h.call("mcl_filter_set_roles", (C.c_int * 4)(cur, prev, spare, ws), int(tick))

    # -- update ------------------------------------------------------------------------------
    def _update_core(self, uniforms=None):
        if self.use_adaptive:
            raise NotImplementedError("the KLD-adaptive (AMCL) modes are single-GPU: %r" % self.params["localization_mode"])
        if self.native:
            return Localizer._update_core(self, uniforms)
        if self.assym:
            raise NotImplementedError("asymmetric MH on sharded particles needs the native exchange path")
        cur, prev, spare, ws, tick = self._roles()
        S, n, h = self.sets, self.n, self.h
        h.call("mcl_likelihood", *[_ptr(t) for t in S[cur]], n, _ptr(self.score_post))
        if self.use_mh:
            h.call("mcl_likelihood", *[_ptr(t) for t in S[prev]], n, _ptr(self.score_pre))
        # staged softmax: max -> all-reduce MAX -> sum exp -> all-reduce SUM -> weights
        h.call("mcl_softmax_max", _ptr(self.score_post), n, _ptr(self.st_post))
        if self.use_mh:
            h.call("mcl_softmax_max", _ptr(self.score_pre), n, _ptr(self.st_pre))
        mx = torch.stack((self.st_post[0], self.st_pre[0]))
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=self.group)
        self.st_post[0], self.st_pre[0] = mx[0], mx[1]
        h.call("mcl_softmax_sumexp", _ptr(self.score_post), n, _ptr(self.st_post))
        if self.use_mh:
            h.call("mcl_softmax_sumexp", _ptr(self.score_pre), n, _ptr(self.st_pre))
        # the sums travel as exact 2^-40 fixed-point integers: identical for any number of ranks
        q = torch.stack((self.st_post.view(torch.int64)[2], self.st_pre.view(torch.int64)[2]))
        dist.all_reduce(q, op=dist.ReduceOp.SUM, group=self.group)
        sm = q.to(torch.float64) * (2.0 ** -40)
        self.st_post[1], self.st_pre[1] = sm[0], sm[1]
        if not self.use_mh:
            h.call("mcl_softmax_weights", _ptr(self.score_post), n, _ptr(self.st_post), _ptr(self.wbuf[ws]))
            return
        h.call("mcl_softmax_weights", _ptr(self.score_post), n, _ptr(self.st_post), _ptr(self.w_post))
        h.call("mcl_softmax_weights", _ptr(self.score_pre), n, _ptr(self.st_pre), _ptr(self.w_pre))
        up = None
        if uniforms is not None:
            u = uniforms if torch.is_tensor(uniforms) else torch.from_numpy(
                np.ascontiguousarray(uniforms, dtype=np.float64))
            u = u.to(self.device)
            up = _ptr(u)
        tick += 1
        h.call("mcl_mh_accept", *[_ptr(t) for t in S[prev]], *[_ptr(t) for t in S[cur]], _ptr(self.w_post),
               _ptr(self.w_pre), n, up, self.seed, tick, self.first_index, *[_ptr(t) for t in S[spare]],
               _ptr(self.wbuf[ws]), None)
        self._set_roles(spare, prev, cur, ws, tick)

    # -- estimate ----------------------------------------------------------------------------
    def _estimate_sums(self):
        cur, _, _, ws, _ = self._roles()
        S, n, h = self.sets, self.n, self.h
        h.call("mcl_estimate_moments_async", *[_ptr(t) for t in S[cur]], _ptr(self.wbuf[ws]), n, _ptr(self.m9))
        dist.all_reduce(self.m9[:6], op=dist.ReduceOp.SUM, group=self.group)
        h.call("mcl_estimate_means_async", _ptr(self.m9))
        h.call("mcl_estimate_central_async", *[_ptr(t) for t in S[cur]], _ptr(self.wbuf[ws]), n,
               C.c_void_p(self.m9.data_ptr() + 6 * 8), _ptr(self.c9))
        dist.all_reduce(self.c9, op=dist.ReduceOp.SUM, group=self.group)

    def estimate(self):
        if self.native:
            return Localizer.estimate(self)
        with self._lock:
            self._bind_stream()
            self._estimate_sums()
            m = self.m9.cpu().numpy()
            c = self.c9.cpu().numpy()
        return assemble_estimate([m[0], m[1], m[6], m[7], m[8]] + list(c) + [0, 0])

    def estimate_async(self, out18):
        if self.native:
            return Localizer.estimate_async(self, out18)
        with self._lock:
            self._bind_stream()
            self._estimate_sums()
            out18[:9].copy_(self.m9)
            out18[9:].copy_(self.c9)

    def finish(self):
        if self.native:
            return Localizer.finish(self)        # the library sequences estimate + resample with its own exchanges
        est = self.estimate()
        self.resample()
        return est

    def finish_async(self, out18):
        if self.native:
            return Localizer.finish_async(self, out18)
        self.estimate_async(out18)
        self.resample()

    # -- resample ----------------------------------------------------------------------------
    def resample(self, r=None):
        if self.native:
            return Localizer.resample(self, r)
        with self._lock:
            self._bind_stream()
            cur, prev, spare, ws, tick = self._roles()
            S, n, h, W = self.sets, self.n, self.h, self.world
            tick += 1
            if r is None:
                r = h.lib.mcl_resample_offset(self.seed, tick, self.n_global)     # same on every rank
            w = self.wbuf[ws]
            h.call("mcl_weights_max", _ptr(w), n, _ptr(self.wmax))
            dist.all_reduce(self.wmax, op=dist.ReduceOp.MAX, group=self.group)
            h.call("mcl_resample_scan", _ptr(w), n, _ptr(self.wmax), self.n_global, _ptr(self.total))
            dist.all_gather_into_tensor(self.totals_all, self.total, group=self.group)
            if self.symm is not None:
                # fused gather + exchange: offspring are stored straight into the destination ranks' spare set
                h.call("mcl_resample_push", n, _ptr(self.totals_all), self.rank, W, float(r), self.n_global, n,
                       *[_ptr(t) for t in S[cur]], C.c_void_p(self.peer_tab[spare].data_ptr()))
                self.symm.barrier()
                self._set_roles(spare, prev, cur, ws, tick)
                return
            totals = [int(t) for t in self.totals_all.cpu().tolist()]              # sync: the plan is host logic
            offsets, grand, m_lo, m_hi, send = plan_resample(totals, float(r), self.n_global, W)
            k = self.rank
            cnt = m_hi[k] - m_lo[k]
            if cnt > self.send_cap:
                cap = max(cnt, 2 * n)
                self.send = [torch.empty(cap, dtype=torch.float64, device=self.device) for _ in range(3)]
                self.send_idx = torch.empty(cap, dtype=torch.int32, device=self.device)
                self.send_cap = cap
            if cnt > 0:
                h.call("mcl_resample_search", n, offsets[k], grand, m_lo[k], cnt, float(r), self.n_global,
                       _ptr(self.send_idx))
                h.call("mcl_gather", *[_ptr(t) for t in S[cur]], _ptr(self.send_idx), cnt,
                       *[_ptr(t) for t in self.send])
            recv_counts = [send[j][k] for j in range(W)]
            for c3 in range(3):
                out = exchange(self.send[c3][:cnt] if cnt > 0 else S[spare][c3][:0], send[k], recv_counts,
                               self.group)
                S[spare][c3].copy_(out)
            self._set_roles(spare, prev, cur, ws, tick)

    # -- whole step ---------------------------------------------------------------------------
    def step(self, odom, ranges, angle_min=None, angle_max=None, angles=None):
        if self.native:
            return Localizer.step(self, odom, ranges, angle_min, angle_max, angles)
        self.predict(odom)
        self.update(ranges, angle_min, angle_max, angles)
        est = self.estimate()
        self.resample()
        return est

    def step_staged(self, odom, k, out18=None):
        if self.native:
            return Localizer.step_staged(self, odom, k, out18)
        self.predict(odom)
        self.update_staged(k)
        self.estimate_async(out18 if out18 is not None else self.est18)
        self.resample()

    def update_chain(self, ranges, angle_min=None, angle_max=None, angles=None, iters=32):
        if not self.native:
            raise NotImplementedError("the MH chain on sharded particles needs the native exchange path")
        return Localizer.update_chain(self, ranges, angle_min, angle_max, angles, iters)

    def gather_particles(self):
        """(n_global, 3) on every rank (tests / visualisation)."""
        local = torch.from_numpy(self.particles()).to(self.device)
        out = torch.empty((self.n_global, 3), dtype=torch.float64, device=self.device)
        dist.all_gather_into_tensor(out, local, group=self.group)
        return out.cpu().numpy()
