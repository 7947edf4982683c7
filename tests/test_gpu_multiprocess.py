"""Multi-process GPU test of the disparity-sharded path: one process per GPU, the REAL exchange over CUDA IPC
peer memory (P2PExchange with device-side ARRIVE/DONE flags, then with caller-side barriers, then the NCCL
all-gather baseline), BASELINE config C5's depth (D = 256).  The maps every rank ends up with must equal the
unsharded CPU oracle's.  Skipped when the box has fewer than 2 GPUs."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _worker(rank, world, port, W, H, D, q):
    import ctypes as C
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from primestereomatch_b200 import DispEst, capi, synth
    from primestereomatch_b200.sharding import BandedUpload, P2PExchange, gather_and_reduce, shard_range
    l8, r8, _ = synth.stereo_pair_u8(W, H, D, seed=77)
    rng = np.random.default_rng(6)
    r8 = np.clip(r8.astype(np.int16) + rng.integers(-4, 5, r8.shape), 0, 255).astype(np.uint8)
    l, r = synth.to_f32(l8), synth.to_f32(r8)
    b, n = shard_range(D, world, rank)
    results = {}
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        de = DispEst(l, r, D, device=rank, d_begin=b, d_count=n)
        for name, sync in (("flags", True), ("barriers", False)):
            ex = P2PExchange(de, world, rank, device_sync=sync)
            for frame in range(3):                       # the blocks and the sequence numbers are reused across frames
                de.CostConst_GPU(); de.CostFilter_GPU()
                ld = np.zeros((H, W), np.uint8); rd = np.zeros((H, W), np.uint8)
                ex.frame(ld.ctypes.data_as(C.c_void_p), rd.ctypes.data_as(C.c_void_p))
            results[name] = (ld, rd)
        # banded upload (every rank uploads H/world rows, NVLink all-gather) feeding the same path
        up = BandedUpload(de, world, rank)
        up.upload(torch.from_numpy(l).pin_memory(), torch.from_numpy(r).pin_memory())
        capi.check(capi.lib().psm_cost_const(de.handle), de.handle)
        de.CostFilter_GPU()
        ld = np.zeros((H, W), np.uint8); rd = np.zeros((H, W), np.uint8)
        ex.frame(ld.ctypes.data_as(C.c_void_p), rd.ctypes.data_as(C.c_void_p))
        results["banded"] = (ld, rd)
        # library baseline: local WTA keys + ncclAllGather + final min
        keys = torch.empty((2, H * W), dtype=torch.int64, device="cuda")
        gathered = torch.empty((2, world, H * W), dtype=torch.int64, device="cuda")
        de.CostConst_GPU(); de.CostFilter_GPU()
        ld = np.zeros((H, W), np.uint8); rd = np.zeros((H, W), np.uint8)
        gather_and_reduce(de, keys, gathered, world, ld, rd)
        results["nccl"] = (ld, rd)
        de.close()
    ok = True
    msg = ""
    if rank in (0, world - 1):
        from oracle import oracle as O
        ref = O.pipeline(l, r, D, threads=16)
        for name, (ld, rd) in results.items():
            if not (np.array_equal(ld, ref["lDis"]) and np.array_equal(rd, ref["rDis"])):
                ok = False
                msg += f"rank {rank}: {name} maps differ from the oracle ({int((ld != ref['lDis']).sum())} / {int((rd != ref['rDis']).sum())} pixels); "
    q.put((rank, ok, msg))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 8])
def test_sharded_exchange_across_processes(world):
    n = _ngpu()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(world, n)
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 1500) + world
    W, H, D = 640, 120, 256
    procs = [ctx.Process(target=_worker, args=(r, world, port, W, H, D, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0, f"rank process exited with {p.exitcode}"
    got = [q.get(timeout=10) for _ in range(world)]
    bad = [m for _, ok, m in got if not ok]
    assert not bad, bad
