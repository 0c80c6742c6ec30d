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


DIRS = [(-1, -1), (0, -1), (1, -1), (-1, 0), (1, 0), (-1, 1), (0, 1), (1, 1)]  # siafd_b200.h: dir = 0..7


class PeerHalo:
    """Ghost exchange by direct stores into the neighbours' arrays (CUDA IPC peer memory over NVLink), the GPU
    path of IceModelVec::update_ghosts for one process per GPU on one node.  Setup maps every neighbour's
    arrays once (handles travel through torch.distributed); afterwards a phase is two stream-ordered calls,
    `siafd_b200_halo_push` + `siafd_b200_halo_wait`: three small launches, no host synchronisation, BOX
    corners included.  Fields must live in the handle's own storage (not bound tensors)."""

    def __init__(self, patch, patches, sia, names, group=None):
        self.patch, self.sia = patch, sia
        h = sia.handle
        mine = {}
        for name in list(names) + [None]:
            buf = (C.c_ubyte * 64)()
            st = lib.siafd_b200_ipc_export(h, -1 if name is None else F[name], buf)
            assert st == 0, lib.siafd_b200_last_error(h)
            mine[name] = bytes(buf)
        world = dist.get_world_size(group)
        everyone = [None] * world
        dist.all_gather_object(everyone, mine, group=group)
        mapped = {}  # (rank, name) -> peer pointer: a handle is opened once per process
        self.bytes_per_push = {}
        for name in list(names) + [None]:
            f = -1 if name is None else F[name]
            for d, (dx, dy) in enumerate(DIRS):
                nb = patch.neighbor(dx, dy)
                if nb == patch.rank:
                    ptr = None
                else:
                    if (nb, name) not in mapped:
                        out = C.c_void_p()
                        st = lib.siafd_b200_ipc_open(h, everyone[nb][name], C.byref(out))
                        assert st == 0, lib.siafd_b200_last_error(h)
                        mapped[(nb, name)] = out.value
                    ptr = mapped[(nb, name)]
                st = lib.siafd_b200_halo_attach(h, f, d, ptr, patches[nb].xm, patches[nb].ym)
                assert st == 0, lib.siafd_b200_last_error(h)
        dist.barrier(group=group)  # nobody pushes before everybody has mapped
        self.bytes_sent = 0

    def exchange(self, names_widths, phase):
        """update_ghosts() of several fields at once: [(name, width), ...]."""
        n = len(names_widths)
        fa = (C.c_int * n)(*[F[nm] for nm, _ in names_widths])
        wa = (C.c_int * n)(*[w for _, w in names_widths])
        h = self.sia.handle
        st = lib.siafd_b200_halo_push(h, n, fa, wa, phase)
        assert st == 0, lib.siafd_b200_last_error(h)
        st = lib.siafd_b200_halo_wait(h, phase)
        assert st == 0, lib.siafd_b200_last_error(h)
        p = self.patch
        for nm, w in names_widths:
            dof = lib.siafd_b200_field_dof(h, F[nm])
            self.bytes_sent += 8 * dof * (2 * w * p.ym + 2 * w * (p.xm + 2 * w))


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
