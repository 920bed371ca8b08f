"""Per-phase SM-clock breakdown of image_kernel on one chunk (developer tool).  python scripts/phase_breakdown.py [n_hyp]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import synth
from salve_b200.renderer import BevRenderer
n_hyp = int(sys.argv[1]) if len(sys.argv) > 1 else 148
rgbs, depths, p1, p2, R, t = synth.synth_building(8, n_hyp, 512, 1024, seed=0)
r = BevRenderer(max_panos=8, max_images=592)
for k in range(8): r.upload_pano(k, rgbs[k], depths[k])
out = torch.empty(n_hyp * 4 * 501 * 501 * 3, dtype=torch.uint8, device="cuda")
for _ in range(2): r.render_hypotheses_device(p1, p2, R, t, out)
torch.cuda.synchronize()
n_img = min(n_hyp * 4, 592)
c = r.last_phase_clocks(n_img)
names = ["A sites", "B/C/D hull+masks", "F list", "pass1", "shade1", "list1b", "pass1b", "shade1b", "list2", "pass2", "H end"]
d = np.diff(c[:, :12], axis=1).astype(np.float64)
tot = (c[:, 11] - c[:, 0]).astype(np.float64)
print("images", n_img, "mean total cycles %.0f (%.1f us at 1.9 GHz)" % (tot.mean(), tot.mean() / 1900))
for i, nme in enumerate(names):
    print("  %-18s mean %9.0f  max %9.0f  share %5.1f%%" % (nme, d[:, i].mean(), d[:, i].max(), 100 * d[:, i].sum() / tot.sum()))
print("queries: pass1 %.0f  pass1b %.0f  pass2 %.0f" % (c[:, 12].mean(), c[:, 13].mean(), c[:, 14].mean()))
