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
