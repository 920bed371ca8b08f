"""Small fixed workload for ncu: N hypotheses of one synthetic building, W warm-up passes + 1 profiled pass.
    python scripts/profile_step.py [n_hyp] [n_panos]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import synth
from salve_b200.renderer import BevRenderer

n_hyp = int(sys.argv[1]) if len(sys.argv) > 1 else 148
n_panos = int(sys.argv[2]) if len(sys.argv) > 2 else 8
passes = int(sys.argv[3]) if len(sys.argv) > 3 else 2
rgbs, depths, p1, p2, R, t = synth.synth_building(n_panos, n_hyp, 512, 1024, seed=0)
r = BevRenderer(max_panos=n_panos, max_images=592)
for k in range(n_panos):
    r.upload_pano(k, rgbs[k], depths[k])
out = torch.empty(n_hyp * 4 * 501 * 501 * 3, dtype=torch.uint8, device="cuda")
for _ in range(passes):
    r.render_hypotheses_device(p1, p2, R, t, out)
torch.cuda.synchronize()
print("done", r.launch_count())
