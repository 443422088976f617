"""Host side of the peer-memory gradient exchange (csrc/peer_reduce.cu): allocate this rank's
IPC-exported region, swap the 64-byte handles through torch.distributed, map the other ranks' regions, and hand the
trainer (a) a torch view of the gradient vector that lives inside the region and (b) the fused "reduce-scatter over NVLink
loads + all-gather + Adam" call that replaces `all_reduce` + two `spn_adam_step` launches.

OPT-IN: set SPN_P2P_ALLREDUCE=1 (Trainer picks it up on multi-GPU runs).  Written after the round's GPU budget was spent and
not yet run on hardware — the default multi-GPU path is NCCL (dist.allreduce_sum_)."""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.distributed as dist

from . import _lib as L


def enabled():
    return os.environ.get("SPN_P2P_ALLREDUCE", "0") == "1"


class _DeviceArray:
    """Minimal __cuda_array_interface__ carrier so torch can view memory the C library allocated."""

    def __init__(self, ptr, n_floats):
        self.__cuda_array_interface__ = {"shape": (int(n_floats),), "typestr": "<f4", "data": (int(ptr), False), "version": 3,
                                         "strides": None}


class PeerGradExchange:
    def __init__(self, n_params, device, group=None):
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerGradExchange needs an initialised torch.distributed process group")
        self.group, self.device = group, torch.device(device)
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if not 2 <= self.world <= 8:
            raise RuntimeError(f"PeerGradExchange: world size {self.world} (one NVSwitch box: 2..8 ranks)")
        self.n_params = int(n_params)
        self.stride = (self.n_params + 127) // 128 * 128
        self.n_floats = 2 * self.stride
        lib = L.lib()
        nbytes = int(lib.spn_peer_region_bytes(self.n_floats))
        own, handle = C.c_void_p(), C.create_string_buffer(64)
        with torch.cuda.device(self.device):
            L.check(lib.spn_peer_alloc(nbytes, C.byref(own), handle), "spn_peer_alloc")
            handles = [None] * self.world
            dist.all_gather_object(handles, handle.raw, group=group)
            self.regions = (C.c_void_p * self.world)()
            for r, h in enumerate(handles):
                if r == self.rank:
                    self.regions[r] = own.value
                else:
                    p = C.c_void_p()
                    L.check(lib.spn_peer_open(h, C.byref(p)), f"spn_peer_open(rank {r})")
                    self.regions[r] = p.value
        self.own = own.value
        grad_ptr = lib.spn_peer_grad_ptr(self.own)
        self.grad_all = torch.as_tensor(_DeviceArray(grad_ptr, self.n_floats), device=self.device)
        self.grads = [self.grad_all[:self.n_params], self.grad_all[self.stride:self.stride + self.n_params]]
        dist.barrier(group=group)          # every region zeroed and mapped before the first flag is written

    def allreduce_adam(self, epoch, net_c, net_f, m, v, lr, betas, eps, step):
        """One optimisation step on the summed gradients of all ranks (mean = grad_scale 1/world), both networks."""
        L.check(L.lib().spn_peer_allreduce_adam(self.regions, self.world, self.rank, int(epoch), self.n_floats, self.n_params,
                                                self.stride, L.ptr(net_c.flat_params()), L.ptr(m[0]), L.ptr(v[0]),
                                                L.ptr(net_f.flat_params()), L.ptr(m[1]), L.ptr(v[1]), float(lr), float(betas[0]),
                                                float(betas[1]), float(eps), int(step), 1.0 / self.world, L.stream()),
                "spn_peer_allreduce_adam")

    def close(self):
        lib = L.lib()
        for r in range(self.world):
            if r != self.rank and self.regions[r]:
                lib.spn_peer_close(self.regions[r])
        dist.barrier(group=self.group)     # nobody unmaps a region another rank may still be reading
        lib.spn_peer_free(self.own)
