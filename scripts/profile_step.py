"""Fixed workload for ncu: the bench step (or a smaller one), W warm-up passes + 1 profiled pass.
    python scripts/profile_step.py [n_hyp=640] [n_panos=40] [passes=2] [max_images=1480]
With the defaults one pass is exactly one bench.py step: one launch each of splat_pano_kernel, the six stage kernels of the image
pipeline (+ image_order_kernel) over 1 358 images (1 280 posed + 78 un-posed) and replicate_images_kernel.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import synth
from salve_b200.renderer import BevRenderer

n_hyp = int(sys.argv[1]) if len(sys.argv) > 1 else 640
n_panos = int(sys.argv[2]) if len(sys.argv) > 2 else 40
passes = int(sys.argv[3]) if len(sys.argv) > 3 else 2
max_images = int(sys.argv[4]) if len(sys.argv) > 4 else 1480
rgbs, depths, p1, p2, R, t = synth.synth_building(n_panos, n_hyp, 512, 1024, seed=0)
r = BevRenderer(max_panos=n_panos, max_images=max_images)
for k in range(n_panos):
    r.upload_pano(k, rgbs[k], depths[k])
out = torch.empty(n_hyp * 4 * 501 * 501 * 3, dtype=torch.uint8, device="cuda")
for _ in range(passes):
    r.render_hypotheses_device(p1, p2, R, t, out)
torch.cuda.synchronize()
print("done", r.launch_count())
