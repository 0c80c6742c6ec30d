"""Ghost (halo) exchange between ranks: the GPU counterpart of IceModelVec::update_ghosts
(src/util/iceModelVec.cc:630-643, DMLocalToLocalBegin/End on a periodic BOX-stencil DMDA,
src/util/IceGrid.cc:863-885).

One process per GPU; `torch.distributed` (NCCL over NVLink on the GPU box, gloo in the CPU tests)
moves the strips.  Two stages fill the BOX corners: stage x exchanges `w` columns over the owned rows,
stage y exchanges `w` rows over all columns including the x ghosts just received.  A rank that is its
own periodic neighbour in a direction wraps locally instead.  Strips are packed / unpacked by the
library's copy kernel on CUDA tensors (siafd_b200_halo_pack / _unpack) and by slicing on CPU tensors
(host-logic tests).
"""
import ctypes as C

import torch
import torch.distributed as dist

from .capi import F, lib


def _strip(a, w_field, xm, ym, dx, dy, width, ghost):
    """View of the send (ghost=False) or receive (ghost=True) strip of a local array a[j, i, ...]."""
    W = w_field
    if dx != 0:
        if ghost:
            i0 = W - width if dx < 0 else W + xm
        else:
            i0 = W if dx < 0 else W + xm - width
        return a[W:W + ym, i0:i0 + width]
    if ghost:
        j0 = W - width if dy < 0 else W + ym
    else:
        j0 = W if dy < 0 else W + ym - width
    return a[j0:j0 + width, :]


class HaloExchanger:
    """Exchanges ghosts of fields held as torch tensors in PISM's local layout."""

    def __init__(self, patch, sia=None, group=None):
        self.patch, self.sia, self.group = patch, sia, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._bufs = {}
        self.bytes_sent = 0

    def _buffers(self, key, shape, like):
        b = self._bufs.get(key)
        if b is None:
            b = tuple(torch.empty(shape, dtype=like.dtype, device=like.device) for _ in range(4))
            self._bufs[key] = b
        return b

    def _pack(self, name, a, w_field, dx, dy, width, out):
        if a.is_cuda and self.sia is not None:
            st = lib.siafd_b200_halo_pack(self.sia.handle, F[name], dx, dy, width, out.data_ptr())
            assert st == 0, st
        else:
            out.copy_(_strip(a, w_field, self.patch.xm, self.patch.ym, dx, dy, width, False))

    def _unpack(self, name, a, w_field, dx, dy, width, buf):
        if a.is_cuda and self.sia is not None:
            st = lib.siafd_b200_halo_unpack(self.sia.handle, F[name], dx, dy, width, buf.data_ptr())
            assert st == 0, st
        else:
            _strip(a, w_field, self.patch.xm, self.patch.ym, dx, dy, width, True).copy_(buf)

    def _stage(self, name, a, w_field, width, axis):
        p = self.patch
        lo = (-1, 0) if axis == 0 else (0, -1)
        hi = (1, 0) if axis == 0 else (0, 1)
        n_lo, n_hi = p.neighbor(*lo), p.neighbor(*hi)
        if n_lo == p.rank and n_hi == p.rank:  # own periodic neighbour: local wrap
            if a.is_cuda and self.sia is not None:
                st = lib.siafd_b200_wrap_ghosts_dir(self.sia.handle, F[name], axis)
                assert st == 0, st
            else:
                for d in (lo, hi):
                    src = _strip(a, w_field, p.xm, p.ym, -d[0], -d[1], width, False)
                    _strip(a, w_field, p.xm, p.ym, d[0], d[1], width, True).copy_(src.clone())
            return
        shape = _strip(a, w_field, p.xm, p.ym, lo[0], lo[1], width, False).shape
        s_hi, s_lo, r_lo, r_hi = self._buffers((name, axis, width, tuple(shape), str(a.device)), shape, a)
        self._pack(name, a, w_field, hi[0], hi[1], width, s_hi)
        self._pack(name, a, w_field, lo[0], lo[1], width, s_lo)
        if a.is_cuda:
            # the packs ran on the handle's stream == torch's current stream (bench.py sets it so)
            pass
        ops = [dist.P2POp(dist.isend, s_hi, n_hi, self.group), dist.P2POp(dist.isend, s_lo, n_lo, self.group),
               dist.P2POp(dist.irecv, r_lo, n_lo, self.group), dist.P2POp(dist.irecv, r_hi, n_hi, self.group)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        self.bytes_sent += 2 * s_hi.numel() * s_hi.element_size()
        self._unpack(name, a, w_field, lo[0], lo[1], width, r_lo)
        self._unpack(name, a, w_field, hi[0], hi[1], width, r_hi)

    def exchange(self, name, a, w_field, width=None):
        """update_ghosts() of one field: a is the torch tensor of the local ghosted array."""
        width = w_field if width is None else width
        self._stage(name, a, w_field, width, 0)
        self._stage(name, a, w_field, width, 1)


def global_max(value, device, group=None):
    """GlobalMax (src/util/pism_utilities.cc:140-142, SIAFD.cc:748)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


class PeerHalo:
    """Ghost exchange by direct stores into the neighbours' arrays (CUDA IPC peer memory over NVLink): a thin wrapper of
    the library's own communicator (siafd_b200_comm_init / _comm_exchange / _comm_allreduce, include/siafd_b200.h), which
    needs no torch.distributed: the IPC handles travel through files under `prefix`.  torch.distributed is only used here
    to agree on a unique prefix when the caller gives none.  `names` are touched first so that optional fields take part.
    A phase is ONE launch (strip copies, then the last CTA raises the neighbours' arrival counters and waits for this
    rank's own); fields must live in the handle's own storage (not bound tensors)."""

    def __init__(self, patch, patches, sia, names=(), group=None, prefix=None):
        self.patch, self.sia = patch, sia
        h = sia.handle
        for name in names:
            assert lib.siafd_b200_device_ptr(h, F[name]), lib.siafd_b200_last_error(h)
        if prefix is None:
            box = [None]
            if dist.get_rank(group) == 0:
                import os
                import uuid
                d = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
                box[0] = os.path.join(d, "siafd_b200_%s" % uuid.uuid4().hex)
            dist.broadcast_object_list(box, src=0, group=group)
            prefix = box[0]
        st = lib.siafd_b200_comm_init(h, patch.rank, len(patches), prefix.encode(), 120.0)
        if st != 0:
            raise RuntimeError(lib.siafd_b200_last_error(h).decode())
        self.bytes_sent = 0

    def exchange(self, names_widths, phase=None):
        """update_ghosts() of several fields at once: [(name, width), ...]."""
        n = len(names_widths)
        fa = (C.c_int * n)(*[F[nm] for nm, _ in names_widths])
        wa = (C.c_int * n)(*[w for _, w in names_widths])
        h = self.sia.handle
        st = lib.siafd_b200_comm_exchange(h, n, fa, wa)
        assert st == 0, lib.siafd_b200_last_error(h)
        p = self.patch
        for nm, w in names_widths:
            dof = lib.siafd_b200_field_dof(h, F[nm])
            self.bytes_sent += 8 * dof * (2 * w * p.ym + 2 * w * (p.xm + 2 * w))

    def allreduce(self, values, op="max"):
        """GlobalMax / GlobalMin / GlobalSum (util/pism_utilities.cc:140-167) of up to 8 doubles, through the pads."""
        n = len(values)
        v = (C.c_double * n)(*values)
        st = lib.siafd_b200_comm_allreduce(self.sia.handle, {"max": 0, "min": 1, "sum": 2}[op], n, v)
        assert st == 0, lib.siafd_b200_last_error(self.sia.handle)
        return list(v)


def device_view(sia, name, shape, device):
    """torch view of the handle-owned device storage of a field (no copy)."""

    class _Buf:
        pass

    b = _Buf()
    ptr = lib.siafd_b200_device_ptr(sia.handle, F[name])
    assert ptr, lib.siafd_b200_last_error(sia.handle)
    b.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f8", "data": (int(ptr), False), "version": 2,
                                  "strides": None}
    return torch.as_tensor(b, device=device)
