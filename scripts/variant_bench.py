"""Time differently compiled builds of libsalve_bev.so on the bench step and check that their output is identical (developer tool).

    python scripts/variant_bench.py [--steps K] lib_a.so lib_b.so ...      # driver: one subprocess per library
    python scripts/variant_bench.py --one lib.so [--steps K]                # worker

Variants are built here with `salve_b200.build.build(defines=[...], out="scratch/lib_<tag>.so")` (scripts/make_variants.py) and travel
to the GPU box with the snapshot.  Every worker prints ms per step (CUDA events, L2 flush between steps), the per-stage times of
the library and a SHA-1 of all images, statuses and the schedule-independent counters; the driver flags any hash that differs
from the first library's.
"""
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker(lib, steps, n_hyp, n_panos):
    os.environ["SALVE_BEV_LIB"] = os.path.abspath(lib)
    import numpy as np
    import torch

    from oracle import synth
    from salve_b200.renderer import BevRenderer

    rgbs, depths, p1, p2, R, t = synth.synth_building(n_panos, n_hyp, 512, 1024, seed=0)
    r = BevRenderer(max_panos=n_panos, max_images=1480)
    for k in range(n_panos):
        r.upload_pano(k, rgbs[k], depths[k])
    n_img = n_hyp * 4
    out = torch.empty(n_img * 501 * 501 * 3, dtype=torch.uint8, device="cuda")
    counts = torch.zeros(n_img * 8, dtype=torch.int32, device="cuda")
    status = torch.zeros(n_img, dtype=torch.int32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream()
    for _ in range(3):
        r.render_hypotheses_device(p1, p2, R, t, out, counts, status)
    torch.cuda.synchronize()
    r.enable_timing(True)
    tot, stage = 0.0, {"splat": 0.0, "image": 0.0, "total": 0.0}
    for k in range(steps):
        flush.fill_(k & 0xFF)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        r.render_hypotheses_device(p1, p2, R, t, out, counts, status)
        b.record(st)
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
        tm = r.last_timings()
        for key in stage:
            stage[key] += tm[key]
    h = hashlib.sha1()
    h.update(out.cpu().numpy().tobytes())
    h.update(status.cpu().numpy().tobytes())
    h.update(np.ascontiguousarray(counts.cpu().numpy().reshape(n_img, 8)[:, :6]).tobytes())
    if os.environ.get("VB_DBG"):
        c = counts.cpu().numpy().reshape(n_img, 8).astype(np.int64)
        sel = c[:, 2] > 0
        print("DBG", os.path.basename(lib), "images", int(sel.sum()), "small descents", float((c[sel, 6] & 0x3FF).mean()), "their waves", float(((c[sel, 6] >> 10) & 0x3FF).mean()), "their flips", float((c[sel, 6] >> 20).mean()), "mean descents", float((c[sel, 7] & 0xFFF).mean()),
              "mean waves", float((c[sel, 7] >> 12).mean()), "mean filled", float(c[sel, 5].mean()), file=sys.stderr)
    print(json.dumps({"lib": os.path.basename(lib), "ms_per_step": tot / steps, "hyp_per_s": n_hyp * steps / (tot / 1e3),
                      "splat_ms": stage["splat"] / steps, "image_ms": stage["image"] / steps, "sha1": h.hexdigest()}))


def main():
    a = sys.argv[1:]
    steps, n_hyp, n_panos = 10, 640, 40
    if "--steps" in a:
        i = a.index("--steps"); steps = int(a[i + 1]); del a[i:i + 2]
    if "--hyp" in a:
        i = a.index("--hyp"); n_hyp = int(a[i + 1]); del a[i:i + 2]
    if "--one" in a:
        i = a.index("--one"); worker(a[i + 1], steps, n_hyp, n_panos); return
    ref = None
    for lib in a:
        res = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", lib, "--steps", str(steps), "--hyp", str(n_hyp)], capture_output=True, text=True)
        line = res.stdout.strip().splitlines()[-1] if res.stdout.strip() else ""
        try:
            d = json.loads(line)
        except Exception:
            print("FAILED", lib, res.stderr[-800:]); continue
        for ln in res.stderr.splitlines():
            if ln.startswith("DBG"):
                print(ln)
        ref = ref or d["sha1"]
        print("%-28s %8.3f ms/step  %8.0f hyp/s  splat %6.3f  image %7.3f  %s" % (d["lib"], d["ms_per_step"], d["hyp_per_s"], d["splat_ms"], d["image_ms"],
                                                                                   "same" if d["sha1"] == ref else "OUTPUT DIFFERS"))
        sys.stdout.flush()


if __name__ == "__main__":
    main()
