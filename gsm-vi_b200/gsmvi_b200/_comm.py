"""Peer-memory exchange buffers for the batch-sharded GSM fit (one process per GPU): CUDA IPC plumbing around
gsmvi_comm_* of libgsmvi_b200.so.  torch.distributed is used only to swap the 64-byte IPC handles and for barriers."""
import ctypes

import torch

from . import _lib as L


class CommLayoutC(ctypes.Structure):
    """gsmvi_comm_layout of include/gsmvi_b200.h"""
    _fields_ = [("stage_off", ctypes.c_longlong), ("s_off", ctypes.c_longlong * 2), ("dmu_off", ctypes.c_longlong),
                ("cnt_off", ctypes.c_longlong), ("lds", ctypes.c_longlong), ("tiles_m", ctypes.c_int), ("tpo", ctypes.c_int)]


class _RawCuda:
    """Expose a raw device allocation to torch through __cuda_array_interface__ (no copy, no ownership)."""

    def __init__(self, ptr, nfloats):
        self.__cuda_array_interface__ = {"shape": (nfloats,), "typestr": "<f4", "data": (ptr, False), "version": 3}


def _declare(lib):
    if getattr(lib, "_gsmvi_comm_declared", False):
        return
    lib.gsmvi_comm_layout_bytes.restype = ctypes.c_longlong
    lib.gsmvi_comm_layout_bytes.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(CommLayoutC)]
    lib.gsmvi_comm_alloc.restype = ctypes.c_int
    lib.gsmvi_comm_alloc.argtypes = [ctypes.c_longlong, ctypes.POINTER(ctypes.c_void_p), ctypes.c_char_p]
    lib.gsmvi_comm_open.restype = ctypes.c_int
    lib.gsmvi_comm_open.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]
    lib.gsmvi_comm_close.restype = ctypes.c_int
    lib.gsmvi_comm_close.argtypes = [ctypes.c_void_p]
    lib.gsmvi_comm_free.restype = ctypes.c_int
    lib.gsmvi_comm_free.argtypes = [ctypes.c_void_p]
    hp = ctypes.POINTER(L.H3OperandC)
    lib.gsmvi_gsm_update_h3_fused.restype = ctypes.c_int
    lib.gsmvi_gsm_update_h3_fused.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_longlong, hp,
                                              ctypes.c_void_p, hp, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                              ctypes.POINTER(CommLayoutC), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_uint, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                              ctypes.c_void_p]
    lib._gsmvi_comm_declared = True


class CommBuffer:
    """This rank's exchange buffer, mapped into every peer, plus the peers' buffers mapped here."""

    def __init__(self, D, group, dist):
        lib = L.lib()
        _declare(lib)
        self.lib, self.dist, self.group = lib, dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.D = D
        self.lay = CommLayoutC()
        self.nbytes = lib.gsmvi_comm_layout_bytes(D, self.world, ctypes.byref(self.lay))
        own = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        L.check(lib.gsmvi_comm_alloc(self.nbytes, ctypes.byref(own), handle), "gsmvi_comm_alloc")
        self.own = own.value
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle.raw), group=group)
        self.ptrs, self.opened = [], []
        for r, h in enumerate(handles):
            if r == self.rank:
                self.ptrs.append(self.own)
            else:
                p = ctypes.c_void_p()
                L.check(lib.gsmvi_comm_open(h, ctypes.byref(p)), "gsmvi_comm_open")
                self.ptrs.append(p.value)
                self.opened.append(p.value)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.peer_base = torch.tensor(self.ptrs, dtype=torch.int64, device=dev)  # device array of world pointers
        self.flat = torch.as_tensor(_RawCuda(self.own, self.nbytes // 4), device=dev)
        lds = int(self.lay.lds)
        self.S = [self.flat[int(self.lay.s_off[k]): int(self.lay.s_off[k]) + D * lds].view(D, lds) for k in (0, 1)]
        self.step = 0
        torch.cuda.synchronize()
        dist.barrier(group=group)  # every buffer is mapped everywhere before anyone pushes

    def update_fused(self, X, G, Gh, mu, Sh, mu_out, cur, B, D, B_total, ws):
        """One fused update + exchange (gsmvi_gsm_update_h3_fused); the new Sigma lands in buffer 1 - cur of every rank."""
        L.check(self.lib.gsmvi_gsm_update_h3_fused(L.ptr(X), X.stride(0), L.ptr(G), G.stride(0), Gh.ref, L.ptr(mu), Sh.ref,
                                                   L.ptr(mu_out), L.ptr(self.peer_base), ctypes.c_void_p(self.own),
                                                   ctypes.byref(self.lay), self.rank,
                                                   self.world, cur, self.step, B, D, B_total, L.ptr(ws), L.stream_ptr()),
                "gsmvi_gsm_update_h3_fused")
        self.step += 1

    def close(self, collective=True):
        """collective=True: every rank calls it (barriers order unmapping and free).  collective=False: emergency /
        interpreter-exit teardown of this rank alone - unmap the peers' buffers and free the own one without waiting for
        anybody (a peer that still has this buffer mapped keeps a valid mapping until it unmaps: CUDA IPC reference-counts
        the allocation; what is lost is only the guarantee that nobody writes into it any more, which a dying fit has
        given up anyway)."""
        if self.own is None:
            return
        try:
            torch.cuda.synchronize()
        except Exception:
            pass
        if collective:
            self.dist.barrier(group=self.group)  # nobody is still writing into a buffer that is about to be unmapped
        for p in self.opened:
            self.lib.gsmvi_comm_close(p)
        self.opened = []
        self.flat = None
        self.S = None
        # an exported allocation must outlive every importer's mapping (cudaFree before the peers' cudaIpcCloseMemHandle is
        # undefined behaviour): free only after every rank has closed its mappings
        if collective:
            self.dist.barrier(group=self.group)
        self.lib.gsmvi_comm_free(self.own)
        self.own = None


class BamShardC(ctypes.Structure):
    """gsmvi_bam_shard of include/gsmvi_b200.h"""
    _fields_ = [("rank", ctypes.c_int), ("world", ctypes.c_int), ("peer_ws", ctypes.c_void_p * 8),
                ("epoch_host", ctypes.POINTER(ctypes.c_uint))]


class PeerAlloc:
    """A zero-initialised device allocation of `nbytes` on every rank of `group`, each mapped into every other rank's
    address space through CUDA IPC (gsmvi_comm_alloc / gsmvi_comm_open).  `ptrs[r]` is rank r's allocation as seen from
    this process (own at [rank]); `flat` is this rank's allocation as a torch uint8 tensor (no ownership)."""

    def __init__(self, nbytes, group, dist):
        lib = L.lib()
        _declare(lib)
        self.lib, self.dist, self.group = lib, dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.nbytes = int(nbytes)
        own = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        L.check(lib.gsmvi_comm_alloc(self.nbytes, ctypes.byref(own), handle), "gsmvi_comm_alloc")
        self.own = own.value
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle.raw), group=group)
        self.ptrs, self.opened = [], []
        for r, h in enumerate(handles):
            if r == self.rank:
                self.ptrs.append(self.own)
            else:
                p = ctypes.c_void_p()
                L.check(lib.gsmvi_comm_open(h, ctypes.byref(p)), "gsmvi_comm_open")
                self.ptrs.append(p.value)
                self.opened.append(p.value)

        class _Raw:
            def __init__(self, ptr, n):
                self.__cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (ptr, False), "version": 3}

        self.flat = torch.as_tensor(_Raw(self.own, self.nbytes), device=torch.device("cuda", torch.cuda.current_device()))
        torch.cuda.synchronize()
        dist.barrier(group=group)  # every allocation is mapped everywhere before anyone stores into a peer

    def close(self, collective=True):
        if self.own is None:
            return
        try:
            torch.cuda.synchronize()
        except Exception:
            pass
        if collective:
            self.dist.barrier(group=self.group)
        for p in self.opened:
            self.lib.gsmvi_comm_close(p)
        self.opened = []
        self.flat = None
        if collective:
            self.dist.barrier(group=self.group)  # free only after every importer has unmapped
        self.lib.gsmvi_comm_free(self.own)
        self.own = None
