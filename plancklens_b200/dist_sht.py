"""m-partitioned transforms over the GPUs of one NVSwitch box (BASELINE.json configs[4]).

One process per GPU (torchrun); `torch.distributed` only provides the plumbing: the CUDA-IPC handle exchange at
set-up, the two barriers per transform and the final sum of the per-rank alm rows.  The data exchange itself is
fused into the libplk_b200 kernels (peer-memory stores over NVLink, see include/plk.h "distributed transforms").

`DistPlan` has the call signatures of `sht.Plan` (alm2map, alm2map_spin, map2alm, map2alm_spin), so
`qest.qe_device` runs unchanged on it.  Conventions:

* alm arguments are FULL alm arrays present on every rank (the filtered alms are a few hundred MB); a rank only
  reads its own m rows.  Analysis returns the full alm on every rank (`reduce=True`: one NCCL all-reduce of the
  zero-padded per-rank rows) or only the rank's rows (`reduce=False`, alm stays m-distributed).
* maps are full-size RING arrays; a rank reads / writes only its own rings (`pixel_ranges()`), which is all the
  per-pixel QE products need.

`SimGroup` runs N simulated ranks inside ONE process on ONE GPU (peers set by pointer instead of IPC, barriers
are no-ops because every stage is executed for all ranks before the next one starts): the same kernels, partition
and exchange pattern, testable on a single-GPU box.
"""
import ctypes

import numpy as np
import torch

from . import _lib, sht
from ._lib import check, vp
from .sht import _ptr, _stream

MBLK = 64


def partition(nside, mmax, nranks, mblk=MBLK):
    """-> (pair_lo[nranks + 1], m_owner[mmax + 1]); host arithmetic of the library, usable without a GPU."""
    lib = _lib.load()
    pair_lo = (ctypes.c_int * (nranks + 1))()
    owner = (ctypes.c_int * (mmax + 1))()
    check(lib.plk_dist_partition(int(nside), int(mmax), int(nranks), int(mblk), pair_lo, owner))
    return np.array(pair_lo[:]), np.array(owner[:])


class _DistHandle:
    """plk_dist object of one rank"""

    def __init__(self, plan, rank, nranks, mblk=MBLK):
        self.lib = _lib.load()
        self.plan, self.rank, self.nranks = plan, rank, nranks
        h = vp()
        check(self.lib.plk_dist_create(ctypes.byref(h), plan._h, rank, nranks, mblk))
        self._h = h

    def __del__(self):
        try:
            if getattr(self, '_h', None):
                self.lib.plk_dist_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def export(self):
        buf = ctypes.create_string_buffer(128)
        check(self.lib.plk_dist_export(self._h, buf))
        return buf.raw

    def import_peer(self, peer, handles):
        check(self.lib.plk_dist_import(self._h, peer, ctypes.create_string_buffer(handles, 128)))

    def phase_ptrs(self):
        a, b = vp(), vp()
        check(self.lib.plk_dist_phase_ptrs(self._h, ctypes.byref(a), ctypes.byref(b)))
        return a, b

    def set_peer(self, peer, x1, x2):
        check(self.lib.plk_dist_set_peer(self._h, peer, x1, x2))

    def pixel_ranges(self):
        r = (ctypes.c_longlong * 4)()
        check(self.lib.plk_dist_pixel_ranges(self._h, r))
        return [(int(r[0]), int(r[1])), (int(r[2]), int(r[3]))]

    # stages
    def legendre_synth(self, spin, a1, a2, fl1, fl2):
        check(self.lib.plk_dist_legendre_synth(self._h, spin, _ptr(a1), _ptr(a2), _ptr(fl1), _ptr(fl2), _stream()))

    def ring_synth(self, spin, m1, m2):
        check(self.lib.plk_dist_ring_synth(self._h, spin, _ptr(m1), _ptr(m2), _stream()))

    def ring_anal(self, spin, m1, m2):
        check(self.lib.plk_dist_ring_anal(self._h, spin, _ptr(m1), _ptr(m2), _stream()))

    def legendre_anal(self, spin, fl1, fl2, a1, a2, add=None):
        """add = (x1, afl1[, x2, afl2]): out rows of this rank += afl[l] * x (m-distributed CG forward operator)"""
        if add is None:
            check(self.lib.plk_dist_legendre_anal(self._h, spin, _ptr(fl1), _ptr(fl2), _ptr(a1), _ptr(a2), _stream()))
        else:
            x1, f1 = add[0], add[1]
            x2, f2 = (add[2], add[3]) if spin else (None, None)
            check(self.lib.plk_dist_legendre_anal_add(self._h, spin, _ptr(fl1), _ptr(fl2), _ptr(x1), _ptr(f1), _ptr(x2), _ptr(f2),
                                                      _ptr(a1), _ptr(a2), _stream()))


class DistPlan:
    """One (nside, lmax) plan split over the ranks of a torch.distributed (NCCL) process group."""

    def __init__(self, nside, lmax, group=None, mblk=MBLK):
        import torch.distributed as dist
        assert dist.is_initialized(), "DistPlan needs an initialised torch.distributed process group (torchrun)"
        self.dist, self.group = dist, group
        self.rank, self.nranks = dist.get_rank(group), dist.get_world_size(group)
        self.plan = sht.get_plan(nside, lmax)
        self.nside, self.lmax, self.npix, self.nalm = self.plan.nside, self.plan.lmax, self.plan.npix, self.plan.nalm
        self.h = _DistHandle(self.plan, self.rank, self.nranks, mblk)
        # CUDA IPC handle exchange (128 bytes per rank) -- set-up only
        mine = torch.frombuffer(bytearray(self.h.export()), dtype=torch.uint8).cuda()
        allh = [torch.empty_like(mine) for _ in range(self.nranks)]
        dist.all_gather(allh, mine, group=group)
        for q, t in enumerate(allh):
            if q != self.rank:
                self.h.import_peer(q, bytes(t.cpu().numpy().tobytes()))
        self._tok = torch.zeros(1, device='cuda')
        self._ev = None          # stage timing (enable_timing())
        self.barrier()

    def enable_timing(self, on=True):
        """CUDA-event timing of the stages of every transform; read with stage_times()."""
        self._ev = [] if on else None

    def _mark(self, name):
        if self._ev is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self._ev.append((name, e))

    def stage_times(self):
        """-> {stage: total ms since enable_timing()} ('barrier1' includes waiting for the slowest rank)"""
        torch.cuda.synchronize()
        out = {}
        for (n0, e0), (n1, e1) in zip(self._ev[:-1], self._ev[1:]):
            if n1 != 'start':
                out[n1] = out.get(n1, 0.0) + e0.elapsed_time(e1)
        self._ev = []
        return out

    def barrier(self):
        """stream-ordered rendez-vous of all ranks (a one-element NCCL all-reduce on the current stream)"""
        self.dist.all_reduce(self._tok, group=self.group)

    def pixel_ranges(self):
        return self.h.pixel_ranges()

    # ---- sht.Plan call signatures
    def alm2map(self, alm, fl=None, out=None):
        out = torch.empty(self.npix, dtype=torch.float64, device='cuda') if out is None else out
        self._mark('start')
        self.barrier()
        self._mark('barrier0')
        self.h.legendre_synth(0, alm, None, fl, None)
        self._mark('legendre_synth')
        self.barrier()
        self._mark('barrier1')
        self.h.ring_synth(0, out, None)
        self._mark('ring_synth')
        return out

    def alm2map_spin(self, glm, clm, spin, flg=None, flc=None, out=None):
        if out is None:
            out = (torch.empty(self.npix, dtype=torch.float64, device='cuda'),
                   torch.empty(self.npix, dtype=torch.float64, device='cuda'))
        self._mark('start')
        self.barrier()
        self._mark('barrier0')
        self.h.legendre_synth(spin, glm, clm, flg, flc)
        self._mark('legendre_synth')
        self.barrier()
        self._mark('barrier1')
        self.h.ring_synth(spin, out[0], out[1])
        self._mark('ring_synth')
        return out

    def _reduce(self, *alms):
        for a in alms:
            self.dist.all_reduce(torch.view_as_real(a), group=self.group)

    def map2alm(self, m, fl=None, out=None, reduce=True, add=None):
        out = torch.empty(self.nalm, dtype=torch.complex128, device='cuda') if out is None else out
        self._mark('start')
        self.barrier()
        self._mark('barrier0')
        self.h.ring_anal(0, m, None)
        self._mark('ring_anal')
        self.barrier()
        self._mark('barrier1')
        self.h.legendre_anal(0, fl, None, out, None, add=add)
        self._mark('legendre_anal')
        if reduce:
            self._reduce(out)
            self._mark('allreduce_alm')
        return out

    def map2alm_spin(self, m1, m2, spin, flg=None, flc=None, out=None, reduce=True, add=None):
        if out is None:
            out = (torch.empty(self.nalm, dtype=torch.complex128, device='cuda'),
                   torch.empty(self.nalm, dtype=torch.complex128, device='cuda'))
        self._mark('start')
        self.barrier()
        self._mark('barrier0')
        self.h.ring_anal(spin, m1, m2)
        self._mark('ring_anal')
        self.barrier()
        self._mark('barrier1')
        self.h.legendre_anal(spin, flg, flc, out[0], out[1], add=add)
        self._mark('legendre_anal')
        if reduce:
            self._reduce(out[0], out[1])
            self._mark('allreduce_alm')
        return out


class SimGroup:
    """N simulated ranks in one process on one GPU: every stage is run for all ranks in turn (that ordering is the
    barrier).  Maps: one full-size array shared by all simulated ranks -- each fills / reads only its own rings."""

    def __init__(self, nside, lmax, nranks, mblk=MBLK):
        self.plan = sht.get_plan(nside, lmax)
        self.nside, self.lmax, self.npix, self.nalm = self.plan.nside, self.plan.lmax, self.plan.npix, self.plan.nalm
        self.nranks = nranks
        self.h = [_DistHandle(self.plan, r, nranks, mblk) for r in range(nranks)]
        ptrs = [h.phase_ptrs() for h in self.h]
        for h in self.h:
            for q, (a, b) in enumerate(ptrs):
                h.set_peer(q, a, b)

    def pixel_ranges(self, rank):
        return self.h[rank].pixel_ranges()

    def alm2map(self, alm, fl=None, out=None):
        out = torch.empty(self.npix, dtype=torch.float64, device='cuda') if out is None else out
        for h in self.h:
            h.legendre_synth(0, alm, None, fl, None)
        for h in self.h:
            h.ring_synth(0, out, None)
        return out

    def alm2map_spin(self, glm, clm, spin, flg=None, flc=None, out=None):
        if out is None:
            out = (torch.empty(self.npix, dtype=torch.float64, device='cuda'),
                   torch.empty(self.npix, dtype=torch.float64, device='cuda'))
        for h in self.h:
            h.legendre_synth(spin, glm, clm, flg, flc)
        for h in self.h:
            h.ring_synth(spin, out[0], out[1])
        return out

    def map2alm(self, m, fl=None, out=None):
        for h in self.h:
            h.ring_anal(0, m, None)
        tot = torch.zeros(self.nalm, dtype=torch.complex128, device='cuda')
        part = torch.empty_like(tot)
        for h in self.h:
            h.legendre_anal(0, fl, None, part, None)
            sht.alm_axpy(tot, part, 1.0)          # the all-reduce of the real thing
        if out is not None:
            out.copy_(tot)
            return out
        return tot

    def map2alm_spin(self, m1, m2, spin, flg=None, flc=None, out=None):
        for h in self.h:
            h.ring_anal(spin, m1, m2)
        tg = torch.zeros(self.nalm, dtype=torch.complex128, device='cuda')
        tc = torch.zeros_like(tg)
        pg, pc = torch.empty_like(tg), torch.empty_like(tg)
        for h in self.h:
            h.legendre_anal(spin, flg, flc, pg, pc)
            sht.alm_axpy(tg, pg, 1.0)
            sht.alm_axpy(tc, pc, 1.0)
        if out is not None:
            out[0].copy_(tg)
            out[1].copy_(tc)
            return out
        return tg, tc
