"""BASELINE configs[3]: test-split-scale sweep sharded by building, with parity on a seeded sample.

    python scripts/sweep_c4.py [--buildings 158] [--hyp 633] [--panos 8] [--sample 0.01]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/sweep_c4.py ...

Every rank computes the same building -> rank assignment (salve_b200.sharding.assign_buildings, greedy by hypothesis count; the
reference deals buildings to processes, scripts/render_dataset_bev.py:186-191), renders its buildings with no collective on the data
path, and checks a seeded sample of its hypotheses against the CPU oracle (bit-exact against the canonical-tie interpolation).
Building b has `--hyp` +- 20 % hypotheses over `--panos` synthetic panos (fewer panos than ZInD's ~40 per building: their synthesis on
the host, not the render, is what takes time here).  Rank 0 prints one JSON line.
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch

from oracle import bev_oracle as bo, synth
import parity_utils as pu
from salve_b200.renderer import BevRenderer
from salve_b200.sharding import assign_buildings, gather_totals


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--buildings", type=int, default=158)
    ap.add_argument("--hyp", type=int, default=633)
    ap.add_argument("--panos", type=int, default=8)
    ap.add_argument("--sample", type=float, default=0.01)
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rng = np.random.default_rng(2024)
    counts = [int(a.hyp * (0.8 + 0.4 * rng.random())) for _ in range(a.buildings)]
    mine = assign_buildings(counts, world)[rank]
    r = BevRenderer(max_panos=a.panos, max_images=1480, device=local)
    ib = 501 * 501 * 3
    cap = max(counts)
    d_posed = torch.empty(cap * 2 * ib, dtype=torch.uint8, device="cuda")
    d_unposed = torch.empty(a.panos * 2 * ib, dtype=torch.uint8, device="cuda")
    n_hyp = n_checked = n_bad = 0
    t_render = 0.0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for b in mine:
        rgbs, depths, p1, p2, R, t = synth.synth_building(a.panos, counts[b], 512, 1024, seed=1000 + b)
        for k in range(a.panos):
            r.upload_pano(k, rgbs[k], depths[k])
        ev0.record()
        idx, nu = r.render_hypotheses_compact_device(p1, p2, R, t, d_posed, d_unposed)
        ev1.record(); torch.cuda.synchronize()
        t_render += ev0.elapsed_time(ev1) / 1e3
        n_hyp += counts[b]
        srng = np.random.default_rng(77 + b)
        for h in np.nonzero(srng.random(counts[b]) < a.sample)[0]:
            posed = d_posed[h * 2 * ib:(h + 1) * 2 * ib].cpu().numpy().reshape(2, 501, 501, 3)
            unposed = d_unposed[idx[h] * 2 * ib:(idx[h] + 1) * 2 * ib].cpu().numpy().reshape(2, 501, 501, 3)
            for si, surf in enumerate(("floor", "ceiling")):
                s1, s2 = bo.render_pair(rgbs[p1[h]], depths[p1[h]], rgbs[p2[h]], depths[p2[h]], R[h], t[h], surf, densify=False)
                for img, st in ((posed[si], s1), (unposed[si], s2)):
                    n_bad += int(not np.array_equal(img, pu.canonical_final(st, pu.oracle_canonical(st))))
            n_checked += 1
    tot, per_rank = gather_totals(n_hyp, t_render)
    bad = torch.tensor([n_bad, n_checked], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(bad)
    if rank == 0:
        print(json.dumps({"config": "C4 sweep sharded by building", "n_gpus": world, "buildings": a.buildings, "hypotheses": tot["units"],
                          "hypotheses_per_rank": per_rank, "render_s_max_over_ranks": tot["max_elapsed_s"],
                          "hyp_per_s": tot["units"] / tot["max_elapsed_s"], "sampled_hypotheses_checked": int(bad[1]),
                          "images_not_bit_exact": int(bad[0])}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
