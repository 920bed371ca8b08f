"""Per-phase SM-clock breakdown of image_kernel on one chunk (developer tool).  python scripts/phase_breakdown.py [n_hyp]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import synth
from salve_b200.renderer import BevRenderer
n_hyp = int(sys.argv[1]) if len(sys.argv) > 1 else 148
rgbs, depths, p1, p2, R, t = synth.synth_building(8, n_hyp, 512, 1024, seed=0)
r = BevRenderer(max_panos=8, max_images=592)
r.set_dedup_unposed(False)
for k in range(8): r.upload_pano(k, rgbs[k], depths[k])
out = torch.empty(n_hyp * 4 * 501 * 501 * 3, dtype=torch.uint8, device="cuda")
for _ in range(2): r.render_hypotheses_device(p1, p2, R, t, out)
torch.cuda.synchronize()
n_img = min(n_hyp * 4, 592)
c = r.last_phase_clocks(n_img)
# slot order in time: 0 start, 1 sites, 2 hull+masks, 18 edge rule+list, 16 window pass, 3 shade+list, 4 pass1, 5 shade, 6 list, 7 pass1b, 8 shade, 9 list, 10 pass2, 11 end
order = [0, 1, 2, 18, 16, 3, 4, 5, 6, 7, 8, 9, 10, 11]
names = ["A sites", "B/C/D hull+masks", "F edge rule+list", "pass0 window", "shade0+list", "pass1 small", "shade1", "list1b", "pass1b", "shade1b", "list2", "pass2 coop", "H end"]
ts = c[:, order].astype(np.float64)
d = np.diff(ts, axis=1)
tot = ts[:, -1] - ts[:, 0]
print("images", n_img, "mean total cycles %.0f (%.1f us at 1.9 GHz)" % (tot.mean(), tot.mean() / 1900))
for i, nme in enumerate(names):
    print("  %-18s mean %9.0f  max %9.0f  share %5.1f%%" % (nme, d[:, i].mean(), d[:, i].max(), 100 * d[:, i].sum() / tot.sum()))
print("queries: window %.0f  pass1 %.0f  pass1b %.0f  pass2 %.0f" % (c[:, 17].mean(), c[:, 12].mean(), c[:, 13].mean(), c[:, 14].mean()))
d15 = c[:, 15]
print("pass2 per image: descents %.0f  waves %.0f  flips %.0f" % ((d15 >> 40).mean(), ((d15 >> 20) & 0xFFFFF).mean(), (d15 & 0xFFFFF).mean()))
