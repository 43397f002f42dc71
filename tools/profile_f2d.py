"""ncu driver / timer for the general-angle blur (filter2d_kernel): 64 x 1360x765, 9x9 kernel at 45 degrees."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robust_object_detection_b200 import augmentations as aug
from robust_object_detection_b200.batch import CorruptionPlan
n, h, w = 64, 765, 1360
src = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device="cuda"); dst = torch.empty_like(src)
plan = CorruptionPlan.uniform(n, h, w)
for k, ang in [(9, 45), (5, 60), (11, 45)]:
    plan.set_blur_kernel(aug._motion_blur_kernel(k, ang))
    for _ in range(2): plan.blur(src, dst, k=k)
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): plan.blur(src, dst, k=k)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"filter2d k={k} angle={ang}: {ms:.3f} ms, {n / ms * 1e3:.0f} img/s, {2 * 3 * h * w * n / ms / 1e6:.0f} GB/s")
