"""Turn an `ncu --set full` report of one bench step into the summaries kept under profiles/ (developer tool).

    python scripts/make_profiles.py gpurun_out/<tag>.ncu-rep r02
writes profiles/<round>_ncu_raw.txt      selected raw metrics per kernel
       profiles/<round>_ncu_lines.txt    the hottest source lines per kernel (stall samples, warp instructions, lanes per instruction)
       profiles/<round>_traffic.json     DRAM bytes per launch and kernel (bench.py's roofline.traffic)
       profiles/<round>_sass.txt         SASS mnemonic histogram of the kernels on the path (cuobjdump of libsalve_bev.so)
"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, rnd = sys.argv[1], sys.argv[2]
N_IMAGES = int(sys.argv[3]) if len(sys.argv) > 3 else 1358   # images rendered in the profiled step (1 280 posed + 78..80 un-posed)
N_PASSES = int(sys.argv[4]) if len(sys.argv) > 4 else 679
WANT = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed']


def ncu(*a):
    return subprocess.run(["ncu", "-i", rep, *a], capture_output=True, text=True).stdout


rows = list(csv.reader(io.StringIO(ncu("--page", "raw", "--csv"))))
hdr, units = rows[0], rows[1]
traffic = {}
with open(os.path.join(ROOT, "profiles", f"{rnd}_ncu_raw.txt"), "w") as f:
    f.write(f"# {os.path.basename(rep)}: ncu --set full --clock-control none, one launch of each kernel of a bench step (c2: 640 hypotheses, 40 panos)\n")
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        name = d.get("Kernel Name", "")
        f.write(f"--- {name}\n")
        for h, u in zip(hdr, units):
            stall = "issue_stalled" in h and "per_issue_active" in h and float(d[h] or 0) > 0.3
            if h in WANT or stall:
                f.write("  %-82s %-16s %s\n" % (h, u, d[h]))
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        rd = float(d["dram__bytes_read.sum"]) * scale[units[hdr.index("dram__bytes_read.sum")]]
        wr = float(d["dram__bytes_write.sum"]) * scale[units[hdr.index("dram__bytes_write.sum")]]
        key = re.sub(r"^void ", "", name).split("(")[0].split("<")[0]
        e = {"dram_bytes_read": rd, "dram_bytes_write": wr, "time_ms": float(d["gpu__time_duration.sum"])}
        if "splat" in key:
            e["pano_passes_in_launch"] = N_PASSES
        else:
            e["images_in_launch"] = N_IMAGES
        traffic[key] = e
with open(os.path.join(ROOT, "profiles", f"{rnd}_traffic.json"), "w") as f:
    json.dump(traffic, f, indent=1)

src = list(csv.reader(io.StringIO(ncu("--page", "source", "--csv", "--print-source", "cuda,sass"))))
per = collections.OrderedDict()
cur = fn = None
for r in src:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        fn = r[1]; continue
    if r[0] in ("Line No", ""):
        continue
    try:
        o = per.setdefault(fn, {}).setdefault((cur, int(r[0])), [r[1].strip()[:110], 0, 0, 0])
        o[1] += int(r[4]); o[2] += int(r[7]); o[3] += int(r[8]) if r[8].isdigit() else 0
    except Exception:
        pass
with open(os.path.join(ROOT, "profiles", f"{rnd}_ncu_lines.txt"), "w") as f:
    for fn, out in per.items():
        ts = sum(o[1] for o in out.values()) or 1; ti = sum(o[2] for o in out.values()) or 1
        f.write(f"=== {fn}: {ts} stall samples, {ti} warp instructions, {sum(o[3] for o in out.values()) / ti:.1f} threads per instruction\n")
        for (fl, ln), o in sorted(out.items(), key=lambda kv: -kv[1][1])[:25]:
            f.write("%-14s %4d smp %5.1f%% inst %5.1f%% thr %4.1f | %s\n" % (fl[:14], ln, 100 * o[1] / ts, 100 * o[2] / ti, o[3] / max(o[2], 1), o[0]))

sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "salve_b200", "libsalve_bev.so")], capture_output=True, text=True).stdout
hist, fn = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        hist[fn] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and fn:
        hist[fn][m.group(1).split(".")[0] + ("." + ".".join(m.group(1).split(".")[1:3]) if m.group(1).startswith(("UBLKCP", "SYNCS", "ATOM", "RED", "REDUX", "LDG", "STG", "LDS", "STS")) else "")] += 1
want = ("splat_pano", "sites_stage", "prep_stage", "window_stage", "shade_stage", "finish_stage", "image_order", "replicate_images", "verifier_preprocess", "layout_raster")
with open(os.path.join(ROOT, "profiles", f"{rnd}_sass.txt"), "w") as f:
    f.write("# cuobjdump -sass salve_b200/libsalve_bev.so (sm_100a): instructions per kernel and the mnemonics that matter for this path\n")
    f.write("# (UBLKCP = cp.async.bulk through the TMA engine, SYNCS = mbarrier, REDUX = warp reductions, ATOM/RED = key-grid atomics; no tensor-core op: none of this is a contraction)\n")
    for fn, c in hist.items():
        if not any(w in fn for w in want):
            continue
        tot = sum(c.values())
        top = ", ".join(f"{k} {v}" for k, v in c.most_common(14))
        spec = ", ".join(f"{k} {v}" for k, v in c.items() if k.startswith(("UBLKCP", "SYNCS", "ATOM", "RED", "REDUX", "DFMA", "DMUL", "DADD", "F2I", "MUFU")))
        f.write(f"{fn}: {tot} instructions\n   most frequent: {top}\n   of note: {spec}\n")
print("wrote profiles/%s_{ncu_raw.txt,ncu_lines.txt,traffic.json,sass.txt}" % rnd)
