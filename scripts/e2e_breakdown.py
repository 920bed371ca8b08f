"""Where the host-buffer (e2e) step spends its time (developer tool): upload only, render to device buffers in e2e-sized chunks,
the plain device->host copy of the same bytes, and the full call.   python scripts/e2e_breakdown.py [chunk ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import synth
from salve_b200.renderer import BevRenderer

N_PANOS, N_HYP, IMG_BYTES = 40, 640, 501 * 501 * 3
chunks = [int(a) for a in sys.argv[1:]] or [444]
rgbs, depths, p1, p2, R, t = synth.synth_building(N_PANOS, N_HYP, 512, 1024, seed=0)
h_rgb = torch.from_numpy(rgbs).pin_memory(); h_depth = torch.from_numpy(depths.view(np.int16)).pin_memory()
h_posed = torch.empty(N_HYP * 2 * IMG_BYTES, dtype=torch.uint8).pin_memory()
h_unposed = torch.empty(N_PANOS * 2 * IMG_BYTES, dtype=torch.uint8).pin_memory()
sh = torch.cuda.current_stream().cuda_stream


def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


d_buf = torch.empty(h_posed.numel() + h_unposed.numel(), dtype=torch.uint8, device="cuda")
h_all = torch.empty(d_buf.numel(), dtype=torch.uint8).pin_memory()
print("plain D2H of %.2f GB: %.2f ms" % (d_buf.numel() / 1e9, timeit(lambda: h_all.copy_(d_buf, non_blocking=True))))
for ch in chunks:
    r = BevRenderer(max_panos=N_PANOS, max_images=ch)
    def up():
        for k in range(N_PANOS): r.upload_pano_ptr(k, h_rgb[k].data_ptr(), h_depth[k].data_ptr(), stream=sh)
    t_up = timeit(up)
    d_posed = torch.empty(h_posed.numel(), dtype=torch.uint8, device="cuda"); d_unposed = torch.empty(h_unposed.numel(), dtype=torch.uint8, device="cuda")
    t_dev = timeit(lambda: r.render_hypotheses_compact_device(p1, p2, R, t, d_posed, d_unposed, stream=sh)) if hasattr(r, "render_hypotheses_compact_device") else float("nan")
    def full():
        up(); r.render_hypotheses_compact(p1, p2, R, t, posed_out=h_posed.numpy(), unposed_out=h_unposed.numpy(), stream=sh)
    t_full = timeit(full)
    t_nou = timeit(lambda: r.render_hypotheses_compact(p1, p2, R, t, posed_out=h_posed.numpy(), unposed_out=h_unposed.numpy(), stream=sh))
    print("chunk %4d: upload %.2f ms, device-only render %.2f ms, host call without upload %.2f ms, full e2e step %.2f ms -> %.0f hyp/s" % (ch, t_up, t_dev, t_nou, t_full, N_HYP / t_full * 1e3))
    r.close()
