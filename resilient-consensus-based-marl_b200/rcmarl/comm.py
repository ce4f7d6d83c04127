"""Peer-memory exchange context for data-parallel training on one node (include/rcmarl.h, csrc/comm.cuh):
each rank allocates an exchange buffer, the CUDA IPC handles are swapped through torch.distributed (plumbing), and the
context is bound so that rcmarl_grad / rcmarl_team / rcmarl_minibatch_sgd reduce across GPUs inside their own kernels
over NVLink -- no NCCL call per optimisation step."""
import ctypes as C

import torch

from . import _lib as L


class PeerComm:
    def __init__(self, rank, world, group=None, max_floats=L.MAX_JOBS * 1536):
        import torch.distributed as dist
        lib = L.lib()
        self.rank, self.world = rank, world
        h = C.c_void_p()
        L.check(lib.rcmarl_comm_create(rank, world, max_floats, C.byref(h)), "rcmarl_comm_create")
        self.handle = h
        nb = lib.rcmarl_comm_handle_bytes()
        buf = C.create_string_buffer(nb)
        L.check(lib.rcmarl_comm_export(h, buf), "rcmarl_comm_export")
        gathered = [None] * world
        dist.all_gather_object(gathered, bytes(buf.raw), group=group)
        status = lib.rcmarl_comm_connect(h, b"".join(gathered))
        torch.cuda.synchronize()
        dist.barrier(group=group)                     # every rank has mapped (or failed to map) every buffer
        if status != 0:
            lib.rcmarl_comm_destroy(h)
            self.handle = None
            L.check(status, "rcmarl_comm_connect")
        L.check(lib.rcmarl_comm_bind(h), "rcmarl_comm_bind")
        self.bound = True

    def check(self):
        """Raise if a peer wait timed out inside a kernel (synchronises the device)."""
        if L.lib().rcmarl_comm_error(self.handle) != 0:
            raise L.RcmarlError("peer-memory all-reduce timed out waiting for another rank")

    def close(self):
        if self.handle is not None:
            lib = L.lib()
            torch.cuda.synchronize()
            lib.rcmarl_comm_bind(None)
            lib.rcmarl_comm_destroy(self.handle)
            self.handle, self.bound = None, False
