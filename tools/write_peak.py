"""Write-only and copy ceilings of this GPU's HBM, measured the same way bench.py times kernels (CUDA events, warm-up,
buffers larger than L2).  A yardstick for the store-bound kernels (CVC writes 4 bytes per voxel and reads almost nothing):
torch's fill / copy kernels are used as the measuring stick only -- nothing in the product calls them.
Usage: python tools/write_peak.py"""
import torch

n = 2_123_366_400 // 4   # the bytes one C4 frame's CVC kernel stores (2 views x 128 slices x 1080 x 1920 x 4 B)
a = torch.empty(n, dtype=torch.float32, device="cuda")
b = torch.empty(n, dtype=torch.float32, device="cuda")


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ms_fill = timed(lambda: a.fill_(1.0))
ms_copy = timed(lambda: b.copy_(a))
print(f"fill  {n * 4 / 1e9:.2f} GB: {ms_fill:.3f} ms = {n * 4 / ms_fill / 1e6:.0f} GB/s written")
print(f"copy  {n * 4 / 1e9:.2f} GB: {ms_copy:.3f} ms = {2 * n * 4 / ms_copy / 1e6:.0f} GB/s read+written")
