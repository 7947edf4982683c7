"""Host-side logic of the disparity-sharded (multi-GPU) path: which global disparity slices a
rank owns, and the one exchange step (all-gather of packed per-pixel minima).

One process per GPU; `torch.distributed` supplies the collective (NCCL on GPUs, gloo in the CPU
tests).  Device work stays in the C-ABI: psm_disp_select_keys / psm_disp_reduce_keys.
"""
import ctypes as C

from . import capi


def shard_range(max_disp, world, rank):
    """Contiguous slices [d_begin, d_begin + d_count) of rank `rank`; the last rank takes the remainder.
    Every rank owns at least one slice (world <= max_disp)."""
    if not (0 <= rank < world) or world > max_disp:
        raise ValueError(f"bad shard request rank={rank} world={world} D={max_disp}")
    base = max_disp // world
    d_begin = rank * base
    d_count = base if rank < world - 1 else max_disp - d_begin
    return d_begin, d_count


def gather_and_reduce(de, keys, gathered, world, lmap=None, rmap=None, group=None):
    """keys: int64 CUDA tensor [2, H*W] (this rank's packed minima, written by
    psm_disp_select_keys); gathered: int64 CUDA tensor [2, world, H*W].  Runs the single
    all-gather per view and the final min -> u8 maps (optionally copied to host arrays)."""
    import torch.distributed as dist
    L = capi.lib()
    capi.check(L.psm_disp_select_keys(de.handle, keys[0].data_ptr(), keys[1].data_ptr()), de.handle)
    dist.all_gather_into_tensor(gathered[0].view(-1), keys[0], group=group)
    dist.all_gather_into_tensor(gathered[1].view(-1), keys[1], group=group)
    lp = lmap.ctypes.data_as(C.c_void_p) if lmap is not None else None
    rp = rmap.ctypes.data_as(C.c_void_p) if rmap is not None else None
    capi.check(L.psm_disp_reduce_keys(de.handle, gathered[0].data_ptr(), gathered[1].data_ptr(), world,
                                      lp, de.wid, rp, de.wid), de.handle)


class P2PExchange:
    """Sharded WTA fused with its exchange over NVLink peer memory (see include/prime_stereo_b200.h):
    select() = WTA + scatter of packed minima into the reducers' blocks, reduce() = chunk min +
    gather of the u8 result into every rank's map, fetch() = D2H.  barrier() (a 1-element NCCL
    all-reduce on the current stream) must run between select/reduce and reduce/fetch.
    Exchange blocks are shared between the per-GPU processes through CUDA IPC handles exchanged once."""

    def __init__(self, de, world, rank, group=None):
        import torch
        import torch.distributed as dist
        self.de, self.world, self.rank, self.group = de, world, rank, group
        L = capi.lib()
        own = C.c_void_p()
        capi.check(L.psm_p2p_create_buffer(de.handle, world, C.byref(own)), de.handle)
        handle = C.create_string_buffer(64)
        capi.check(L.psm_ipc_export(de.handle, own, handle), de.handle)
        mine = torch.frombuffer(bytearray(handle.raw), dtype=torch.uint8).cuda()
        allh = torch.empty((world, 64), dtype=torch.uint8, device="cuda")
        dist.all_gather_into_tensor(allh.view(-1), mine, group=group)
        allh = allh.cpu().numpy()
        ptrs = (C.c_void_p * world)()
        for r in range(world):
            if r == rank:
                ptrs[r] = own.value
            else:
                p = C.c_void_p()
                capi.check(L.psm_ipc_import(de.handle, allh[r].tobytes(), C.byref(p)), de.handle)
                ptrs[r] = p.value
        capi.check(L.psm_p2p_set_peers(de.handle, ptrs, world, rank), de.handle)
        self._flag = torch.zeros(1, device="cuda")
        dist.barrier(group=group)

    def select(self):
        capi.check(capi.lib().psm_disp_select_keys_p2p(self.de.handle), self.de.handle)

    def barrier(self):
        import torch.distributed as dist
        dist.all_reduce(self._flag, group=self.group)  # stream-ordered cross-rank barrier

    def reduce(self):
        capi.check(capi.lib().psm_disp_reduce_p2p(self.de.handle), self.de.handle)

    def fetch(self, lmap_ptr, rmap_ptr):
        capi.check(capi.lib().psm_disp_fetch_p2p(self.de.handle, lmap_ptr, self.de.wid, rmap_ptr, self.de.wid),
                   self.de.handle)

    def frame(self, lmap_ptr=None, rmap_ptr=None):
        """select -> barrier -> reduce -> barrier [-> fetch]"""
        self.select(); self.barrier(); self.reduce(); self.barrier()
        if lmap_ptr is not None:
            self.fetch(lmap_ptr, rmap_ptr)
