"""Multi-GPU evaluation: one process per GPU, contiguous support-block sharding.

Every generator is a map over its supports and its rows / COO slots are affine in the support
index, so rank r of W evaluates ``k ∈ [K·r/W, K·(r+1)/W)`` of every generator and OWNS the matching
slices of c, Jacobian values and Hessian values: ``cons!``, ``jac_coord!`` and ``hess_coord!`` need no
communication at all.  x is replicated.  ``obj`` needs an all-reduce of one double and ``grad!`` an
all-reduce of the slice of g that several ranks contribute to (finite / first-stage variables that
every support's objective term references) — NCCL over NVLink through ``torch.distributed`` on
GPUs, gloo in the CPU tests.  The KKT factorisation stays on one GPU (non-target cost).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import lib as _lib
from . import model as _m
from .core import ExaCore


class ShardedExaModel:
    """NLPModels callbacks over a model sharded across the ranks of a torch.distributed group."""

    def __init__(self, core: ExaCore, device: Optional[int] = None, group=None, flags: int = _lib.IEXA_F_DEFAULT,
                 library=None, evaluator=None):
        import torch
        import torch.distributed as dist
        self.dist, self.torch, self.group = dist, torch, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.model = _m.ExaModel(core, device=0 if device is None else device, rank=self.rank, world=self.world,
                                 flags=flags, library=library)
        self.meta = self.model.meta
        self._init_shared()
        self._ev = evaluator  # tests inject a host evaluator; default: the CUDA engine
        self._shared_t = None

    @classmethod
    def wrap(cls, model: "_m.ExaModel", group=None) -> "ShardedExaModel":
        """the sharded view of an ExaModel that was already built with (rank, world)"""
        import torch
        import torch.distributed as dist
        self = cls.__new__(cls)
        self.dist, self.torch, self.group = dist, torch, group
        self.rank, self.world = model.rank, model.world
        self.model, self.meta = model, model.meta
        self._init_shared()
        self._ev, self._shared_t = None, None
        return self

    def _init_shared(self):
        """0-based variable indices whose gradient entries are partial sums on several ranks (iexa_shared_ranges: shared /
        finite variables, shard-boundary entries of shifted references, variables indexed through product or restricted
        iterators).  ``shared_all``: the set covers most of g — all-reduce the whole vector instead of a gathered slice."""
        m = self.model
        n = m.L.iexa_shared_ranges(m.h, None, 0)
        segs = (_lib.Segment * max(n, 1))()
        m.L.iexa_shared_ranges(m.h, segs, n)
        self.shared_ranges = [(s.global_start, s.global_start + s.length) for s in segs[:n]]
        total = sum(hi - lo for lo, hi in self.shared_ranges)
        self.shared_all = total > 0.25 * max(self.meta.nvar, 1)
        self.shared_idx = (np.concatenate([np.arange(lo, hi, dtype=np.int64) for lo, hi in self.shared_ranges])
                           if (self.shared_ranges and not self.shared_all) else np.zeros(0, dtype=np.int64))
        self._peer = None

    # ---- layout ---------------------------------------------------------------------------------
    def segments(self, which: int):
        """[(global_start, local_start, length)] of rows (0), Jacobian slots (1), Hessian slots (2)."""
        m = self.model
        n = m.L.iexa_segments(m.h, which, None, 0)
        segs = (_lib.Segment * max(n, 1))()
        m.L.iexa_segments(m.h, which, segs, n)
        return [(s.global_start, s.local_start, s.length) for s in segs[:n]]

    def x_ranges(self):
        """[(start, length)] (0-based) of the parts of x this rank's callbacks read: its own supports of every
        variable block, the replicated finite / shared variables and the shard-boundary halos.  A distributed
        solver keeps only these current on this rank (halo exchange instead of a broadcast of the iterate)."""
        m = self.model
        n = m.L.iexa_x_ranges(m.h, None, 0)
        segs = (_lib.Segment * max(n, 1))()
        m.L.iexa_x_ranges(m.h, segs, n)
        return [(s.global_start, s.length) for s in segs[:n]]

    # ---- distributed iterate: who owns which part of x, and the halo exchange ----------------------
    @staticmethod
    def _intersect(a, b):
        """intersection of two sorted lists of disjoint half-open intervals [(lo, hi)]"""
        out, i, j = [], 0, 0
        while i < len(a) and j < len(b):
            lo, hi = max(a[i][0], b[j][0]), min(a[i][1], b[j][1])
            if lo < hi:
                out.append((lo, hi))
            if a[i][1] < b[j][1]:
                i += 1
            else:
                j += 1
        return out

    @staticmethod
    def _subtract(a, b):
        """a \\ b for sorted lists of disjoint half-open intervals"""
        out, j = [], 0
        for lo, hi in a:
            cur = lo
            while j < len(b) and b[j][1] <= cur:
                j += 1
            k = j
            while k < len(b) and b[k][0] < hi:
                if b[k][0] > cur:
                    out.append((cur, b[k][0]))
                cur = max(cur, b[k][1])
                k += 1
            if cur < hi:
                out.append((cur, hi))
        return out

    def x_partition(self):
        """(owned, recv, send): every entry of x that some rank reads is OWNED by the lowest rank that reads it (shared
        variables by rank 0; a shard-boundary halo by the lower neighbour).  ``owned``: this rank's half-open intervals;
        ``recv[q]`` / ``send[r]``: the intervals this rank needs from owner q / owes to reader r.  Computed once."""
        if getattr(self, "_partition", None) is not None:
            return self._partition
        mine = [(s0, s0 + ln) for s0, ln in self.x_ranges()]
        reads = [mine]
        if self.world > 1:
            reads = [None] * self.world
            self.dist.all_gather_object(reads, mine, group=self.group)
        owned_all, seen = [], []
        for r in range(self.world):
            owned_all.append(self._subtract(reads[r], seen))
            seen = sorted(seen + owned_all[r])
            merged = []
            for lo, hi in seen:                      # keep `seen` a list of disjoint, merged intervals
                if merged and lo <= merged[-1][1]:
                    merged[-1] = (merged[-1][0], max(merged[-1][1], hi))
                else:
                    merged.append((lo, hi))
            seen = merged
        recv = {q: self._intersect(reads[self.rank], owned_all[q]) for q in range(self.world) if q != self.rank}
        send = {r: self._intersect(reads[r], owned_all[self.rank]) for r in range(self.world) if r != self.rank}
        self._partition = (owned_all[self.rank], {q: v for q, v in recv.items() if v}, {r: v for r, v in send.items() if v})
        return self._partition

    # ---- NVLink peer-memory path (csrc/halo.cu): one small kernel per rank and exchange ---------------------------
    def enable_peer_halo(self, x) -> bool:
        """map the peers' x buffers and flag blocks (CUDA IPC) so that ``exchange_x(x)`` and the small all-reduces run as
        single kernels over NVLink peer memory.  ``x`` (a CUDA tensor) must be THE iterate buffer of this rank from now on.
        Collective.  Returns False (and keeps the NCCL path) when IPC is not available for this buffer."""
        torch, dist = self.torch, self.dist
        if self.world == 1 or not (isinstance(x, torch.Tensor) and x.is_cuda):
            return False
        m = self.model
        L = m.L
        h = C.c_void_p()
        ok = L.iexa_halo_create(C.byref(h), x.device.index, self.rank, self.world) == 0
        hx, hf, off = (C.c_ubyte * 64)(), (C.c_ubyte * 64)(), C.c_int64()
        ok = ok and L.iexa_halo_export(h, C.c_void_p(x.data_ptr()), hx, C.byref(off), hf) == 0
        mine = (bytes(hx), int(off.value), bytes(hf), bool(ok))
        allh = [None] * self.world
        dist.all_gather_object(allh, mine, group=self.group)
        if not all(a[3] for a in allh):
            if ok:
                L.iexa_halo_destroy(h)
            return False
        for q, (bx, o, bf, _) in enumerate(allh):
            if q == self.rank:
                continue
            ax, af = (C.c_ubyte * 64).from_buffer_copy(bx), (C.c_ubyte * 64).from_buffer_copy(bf)
            if L.iexa_halo_connect(h, q, ax, o, af) != 0:
                ok = False
        _, recv, send = self.x_partition()
        for r, ivs in send.items():
            a = np.ascontiguousarray(np.array(ivs, dtype=np.int64).reshape(-1))
            ok = ok and L.iexa_halo_set_sends(h, r, len(ivs), a.ctypes.data) == 0
        rp = np.ascontiguousarray(np.array(sorted(recv), dtype=np.int32))
        ok = ok and L.iexa_halo_set_recvs(h, len(rp), rp.ctypes.data if len(rp) else None) == 0
        flag = torch.tensor([1.0 if ok else 0.0], device=x.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if flag.item() < 1.0:
            L.iexa_halo_destroy(h)
            return False
        self._peer = (h, x.data_ptr())
        self._red = torch.zeros(1024, dtype=torch.float64, device=x.device)
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        return True

    def peer_status(self) -> int:
        return int(self.model.L.iexa_halo_status(self._peer[0])) if self._peer else 0

    def halo_entries(self) -> int:
        _, recv, _ = self.x_partition()
        return sum(hi - lo for v in recv.values() for lo, hi in v)

    def close_peer_halo(self):
        if self._peer:
            self.model.L.iexa_halo_destroy(self._peer[0])
            self._peer = None

    def allreduce_obj_grad_(self, f_dev, g):
        """the collective part of obj + grad! in ONE reduction: [f, g[shared]] summed over the ranks, in place (f_dev: 1-element
        CUDA tensor holding this rank's objective partial, g: this rank's dense gradient partial).  Small payloads go
        through the peer-memory all-reduce (deterministic, one kernel), larger ones through NCCL."""
        if self.world == 1:
            return
        torch, dist = self.torch, self.dist
        if self.shared_all:
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)
            if self._peer:
                _lib.check(self.model.L, self.model.L.iexa_halo_allreduce_small(self._peer[0], C.c_void_p(f_dev.data_ptr()), 1,
                                                                                 C.c_void_p(torch.cuda.current_stream(g.device).cuda_stream)))
            else:
                dist.all_reduce(f_dev, op=dist.ReduceOp.SUM, group=self.group)
            return
        ns = len(self.shared_idx)
        if self._shared_t is None or self._shared_t.device != g.device:
            self._shared_t = torch.from_numpy(self.shared_idx).to(g.device)
        if self._peer and ns + 1 <= 1024:
            buf = self._red[:ns + 1]
            buf[0:1] = f_dev
            if ns:
                buf[1:] = g[self._shared_t]
            _lib.check(self.model.L, self.model.L.iexa_halo_allreduce_small(self._peer[0], C.c_void_p(buf.data_ptr()), ns + 1,
                                                                             C.c_void_p(torch.cuda.current_stream(g.device).cuda_stream)))
        else:
            buf = torch.cat([f_dev, g[self._shared_t]]) if ns else f_dev
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
        f_dev.copy_(buf[0:1])
        if ns:
            g[self._shared_t] = buf[1:]

    def exchange_x(self, x):
        """make every range of x this rank READS current, given that each rank holds current values on the ranges it
        OWNS: the shared slice and the shard-boundary halos cross ranks — pushed straight into the readers' x over NVLink
        peer memory by one small kernel (after ``enable_peer_halo``), else NCCL point-to-point sends — the distributed
        solver's alternative to broadcasting the whole iterate."""
        if self.world == 1:
            return x
        torch, dist = self.torch, self.dist
        if self._peer and isinstance(x, torch.Tensor) and x.data_ptr() == self._peer[1]:
            _lib.check(self.model.L, self.model.L.iexa_halo_exchange(self._peer[0], C.c_void_p(x.data_ptr()),
                                                                      C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)))
            return x
        _, recv, send = self.x_partition()
        xt = x if isinstance(x, torch.Tensor) else torch.from_numpy(x)
        key = str(xt.device)
        if getattr(self, "_xidx", {}).get("dev") != key:   # index tensors of the packed ranges, built once per device
            mk = lambda ivs: torch.cat([torch.arange(lo, hi, dtype=torch.int64) for lo, hi in ivs]).to(xt.device)
            self._xidx = {"dev": key, "send": {r: mk(v) for r, v in sorted(send.items())},
                          "recv": {q: mk(v) for q, v in sorted(recv.items())}}
        ops, inbox = [], {}
        for r, idx in self._xidx["send"].items():           # one gather kernel per reader
            ops.append(dist.P2POp(dist.isend, xt[idx], r, group=self.group))
        for q, idx in self._xidx["recv"].items():
            inbox[q] = torch.empty(idx.numel(), dtype=xt.dtype, device=xt.device)
            ops.append(dist.P2POp(dist.irecv, inbox[q], q, group=self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for q, idx in self._xidx["recv"].items():           # one scatter kernel per owner
            xt[idx] = inbox[q]
        return x

    def scatter_local(self, which: int, global_vec):
        """this rank's slice of a GLOBAL vector (e.g. the multipliers y) in local layout"""
        n = (self.model.loc_ncon, self.model.loc_nnzj, self.model.loc_nnzh)[which]
        out = global_vec.new_zeros(max(n, 1)) if isinstance(global_vec, self.torch.Tensor) else np.zeros(max(n, 1))
        for gs, ls, ln in self.segments(which):
            out[ls:ls + ln] = global_vec[gs:gs + ln]
        return out

    def gather_global(self, which: int, local_vec):
        """assemble the GLOBAL vector from every rank's local slice (tests / single-GPU consumers:
        this gather is part of the non-target KKT cost)"""
        torch, dist = self.torch, self.dist
        total = (self.meta.ncon, self.meta.nnzj, self.meta.nnzh)[which]
        t = torch.as_tensor(local_vec)
        out = torch.zeros(total, dtype=t.dtype, device=t.device)
        for gs, ls, ln in self.segments(which):
            out[gs:gs + ln] = t[ls:ls + ln]
        if self.world > 1:
            dist.all_reduce(out, op=dist.ReduceOp.SUM, group=self.group)  # slices are disjoint
        return out

    def gather_rows(self, local_vec, offset: int, size: int):
        """rows ``offset .. offset+size`` (0-based, global numbering) of a row-sharded vector — the multipliers or the
        constraint values of ONE constraint (``map_dual``, infiniteopt_backend.jl:490-508, on sharded buffers): every
        rank contributes the rows it owns, an all-reduce of ``size`` doubles assembles them on all ranks."""
        torch = self.torch
        t = torch.as_tensor(local_vec)
        out = torch.zeros(size, dtype=t.dtype, device=t.device)
        for gs, ls, ln in self.segments(0):
            lo, hi = max(gs, offset), min(gs + ln, offset + size)
            if lo < hi:
                out[lo - offset:hi - offset] = t[ls + (lo - gs):ls + (hi - gs)]
        if self.world > 1:
            self.dist.all_reduce(out, op=self.dist.ReduceOp.SUM, group=self.group)
        return out

    # ---- callbacks ------------------------------------------------------------------------------
    def _reduce_scalar(self, v: float, like=None) -> float:
        if self.world == 1:
            return v
        torch, dist = self.torch, self.dist
        dev = like.device if isinstance(like, torch.Tensor) else "cpu"
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return float(t.item())

    def obj(self, x) -> float:
        torch = self.torch
        if self._ev is None and isinstance(x, torch.Tensor) and x.is_cuda:
            # device path: the rank's partial stays on the GPU (iexa_obj_device, no host synchronisation), NCCL
            # all-reduces the one double, and the only synchronisation is the final read
            m = self.model
            if getattr(self, "_fdev", None) is None or self._fdev.device != x.device:
                self._fdev = torch.zeros(1, dtype=torch.float64, device=x.device)
            st = torch.cuda.current_stream(x.device).cuda_stream
            _lib.check(m.L, m.L.iexa_obj_device(m.h, C.c_void_p(x.data_ptr()), C.c_void_p(self._fdev.data_ptr()), C.c_void_p(st)))
            if self.world > 1:
                self.dist.all_reduce(self._fdev, op=self.dist.ReduceOp.SUM, group=self.group)
            return float(self._fdev.item())
        part = self._ev.obj(x) if self._ev else _m.obj(self.model, x)
        return self._reduce_scalar(part, x)

    def grad_(self, x, g):
        """dense g; entries of shared variables are complete on every rank after the call, all other
        entries hold this rank's (exclusive) contributions — zero where another rank owns the support."""
        if self._ev:
            self._ev.grad_(x, g)
        else:
            if self.world > 1:   # the engine writes only inside this rank's read ranges (iexa_grad): zero the rest here
                g.zero_() if isinstance(g, self.torch.Tensor) else g.fill(0.0)
            _m.grad_(self.model, x, g)
        if self.world > 1 and self.shared_all:
            gt = g if isinstance(g, self.torch.Tensor) else self.torch.from_numpy(g)
            self.dist.all_reduce(gt, op=self.dist.ReduceOp.SUM, group=self.group)
        elif self.world > 1 and len(self.shared_idx):
            torch, dist = self.torch, self.dist
            gt = g if isinstance(g, torch.Tensor) else torch.from_numpy(g)
            if self._shared_t is None or self._shared_t.device != gt.device:
                self._shared_t = torch.from_numpy(self.shared_idx).to(gt.device)
            buf = gt[self._shared_t].contiguous()
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
            gt[self._shared_t] = buf
        return g

    def grad_full_(self, x, g):
        """dense g complete on every rank (all-reduce of the whole vector; for single-GPU consumers)"""
        if self._ev:
            self._ev.grad_(x, g)
        else:
            if self.world > 1:
                g.zero_() if isinstance(g, self.torch.Tensor) else g.fill(0.0)
            _m.grad_(self.model, x, g)
        if self.world > 1:
            gt = g if isinstance(g, self.torch.Tensor) else self.torch.from_numpy(g)
            self.dist.all_reduce(gt, op=self.dist.ReduceOp.SUM, group=self.group)
        return g

    def cons_(self, x, c):
        return self._ev.cons_(x, c) if self._ev else _m.cons_(self.model, x, c)

    def jac_coord_(self, x, vals):
        return self._ev.jac_coord_(x, vals) if self._ev else _m.jac_coord_(self.model, x, vals)

    def hess_coord_(self, x, y_local, vals, obj_weight: float = 1.0):
        if self._ev:
            return self._ev.hess_coord_(x, y_local, vals, obj_weight)
        return _m.hess_coord_(self.model, x, y_local, vals, obj_weight)
