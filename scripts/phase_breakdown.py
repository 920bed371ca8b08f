"""Per-phase SM-clock breakdown of image_kernel on one chunk (developer tool).  python scripts/phase_breakdown.py [n_hyp]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import synth
from salve_b200.renderer import BevRenderer
n_hyp = int(sys.argv[1]) if len(sys.argv) > 1 else 148
n_panos = int(sys.argv[2]) if len(sys.argv) > 2 else 8
max_images = int(sys.argv[3]) if len(sys.argv) > 3 else 592
dedup = (sys.argv[4] != "0") if len(sys.argv) > 4 else False
rgbs, depths, p1, p2, R, t = synth.synth_building(n_panos, n_hyp, 512, 1024, seed=0)
r = BevRenderer(max_panos=n_panos, max_images=max_images)
r.set_dedup_unposed(dedup)
for k in range(n_panos): r.upload_pano(k, rgbs[k], depths[k])
out = torch.empty(n_hyp * 4 * 501 * 501 * 3, dtype=torch.uint8, device="cuda")
for _ in range(2): r.render_hypotheses_device(p1, p2, R, t, out)
torch.cuda.synchronize()
n_img = min(n_hyp * 4, max_images) if not dedup else min(n_hyp * 2 + 2 * n_panos, max_images)
c = r.last_phase_clocks(n_img)
# slot order in time: 0 start, 1 sites, 2 hull+masks, 18 edge rule+list, 16 window pass, 3 shade, 9 list, 10 cooperative pass, 11 end
order = [0, 1, 2, 18, 16, 3, 9, 10, 11]
names = ["A sites", "B/C/D hull+masks", "F edge rule+list", "G0 window", "shade", "list", "G2 cooperative", "H end"]
ts = c[:, order].astype(np.float64)
d = np.diff(ts, axis=1)
tot = ts[:, -1] - ts[:, 0]
print("images", n_img, "mean total cycles %.0f (%.1f us at 1.965 GHz)" % (tot.mean(), tot.mean() / 1965))
for i, nme in enumerate(names):
    print("  %-18s mean %9.0f  max %9.0f  share %5.1f%%" % (nme, d[:, i].mean(), d[:, i].max(), 100 * d[:, i].sum() / tot.sum()))
print("queries: window %.0f  cooperative %.0f" % (c[:, 17].mean(), c[:, 14].mean()))
d15 = c[:, 15]
print("flips per image: all %.0f, window pass %.0f = %.2f per window query; filled px %.0f" % (c[:, 6].mean(), (c[:, 6] - (d15 & 0xFFFFF)).mean(), (c[:, 6] - (d15 & 0xFFFFF)).mean() / max(c[:, 17].mean(), 1), c[:, 7].mean()))
print("cooperative pass per image: descents %.0f  waves %.0f  flips %.0f" % ((d15 >> 40).mean(), ((d15 >> 20) & 0xFFFFF).mean(), (d15 & 0xFFFFF).mean()))
print("cooperative pass, cycles: phase %.0f, busiest warp %.0f, mean warp %.0f, longest single descent %.0f" % (
    (c[:, 10] - c[:, 9]).mean(), c[:, 4].mean(), c[:, 5].mean() / 16, c[:, 13].mean()))
# timeline of the launch from the global timer: how well the persistent CTAs are packed
c = c[c[:, 19] > 0]
t0, t1, slot = c[:, 19].astype(np.float64), c[:, 20].astype(np.float64), c[:, 21]
os.makedirs("gpurun_out", exist_ok=True)
np.save("gpurun_out/phase_clocks.npy", c)
span = (t1.max() - t0.min()) / 1e3
busy = (t1 - t0).sum() / 1e3
ns = len(np.unique(slot))
print("launch span %.0f us, %d CTA slots, mean image %.0f us (max %.0f), packing efficiency %.1f%% (sum of image times / (slots x span))" % (
    span, ns, (t1 - t0).mean() / 1e3, (t1 - t0).max() / 1e3, 100 * busy / (ns * span)))
last = np.array([t1[slot == k].max() for k in np.unique(slot)])
print("CTA finish times relative to the end of the launch (us): mean %.0f  p10 %.0f  p50 %.0f  max %.0f" % (
    ((t1.max() - last) / 1e3).mean(), np.percentile((t1.max() - last) / 1e3, 90), np.percentile((t1.max() - last) / 1e3, 50), ((t1.max() - last) / 1e3).max()))
