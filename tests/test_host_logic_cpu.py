"""Host-side mirror types and the multi-GPU sharding logic (gloo, world_size 2)."""

import json
import os
import socket
import sys

import numpy as np
import pytest

from salve_b200.common.bevparams import BEVParams, get_line_width_by_resolution
from salve_b200.common.sim2 import Sim2
from salve_b200.utils.mesh_grid import get_mesh_grid_as_point_cloud
from salve_b200 import sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bevparams_kat():
    """reference tests/common/test_bevparams.py:9-50."""
    p = BEVParams(img_h=20, img_w=20, meters_per_px=0.5)
    got = p.bevimg_Sim2_world.transform_from(np.array([[2, 2], [-5, -5], [5, 5]]))
    assert np.allclose(got, [[14, 14], [0, 0], [20, 20]])
    assert [get_line_width_by_resolution(r) for r in (0.005, 0.01, 0.02)] == [30, 15, 8]
    d = BEVParams()
    assert d.xlims == [-5, 5] and d.ylims == [-5, 5] and d.img_h == 500


def test_sim2_contract(tmp_path):
    """reference tests/common/test_sim2.py: constructor errors, float32 storage, transform, JSON, compose/inverse."""
    with pytest.raises(ValueError):
        Sim2(np.eye(3), np.zeros(2), 1.0)
    with pytest.raises(ValueError):
        Sim2(np.eye(2), np.zeros(3), 1.0)
    with pytest.raises(ValueError):
        Sim2([[1, 0], [0, 1]], np.zeros(2), 1.0)
    with pytest.raises(ZeroDivisionError):
        Sim2(np.eye(2), np.zeros(2), 0.0)
    th = np.deg2rad(30)
    R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    s = Sim2(R, np.array([1.0, 2.0]), 3.0)
    assert s.rotation.dtype == np.float32 and s.translation.dtype == np.float32 and isinstance(s.scale, float)
    assert abs(s.theta_deg - 30) < 1e-4
    pts = np.array([[1.0, 0.0], [0.0, 1.0]])
    assert np.allclose(s.transform_from(pts), (pts @ R.T + [1, 2]) * 3, atol=1e-6)
    with pytest.raises(ValueError):
        s.transform_from(np.zeros((3, 3)))
    ident = s.compose(s.inverse())
    assert np.allclose(ident.rotation, np.eye(2), atol=1e-6) and np.allclose(ident.translation, 0, atol=1e-5) and np.isclose(ident.scale, 1)
    f = tmp_path / "a_Sim2_b.json"
    s.save_as_json(str(f))
    d = json.load(open(f))
    assert set(d) == {"R", "t", "s"} and len(d["R"]) == 4 and len(d["t"]) == 2
    assert Sim2.from_json(str(f)) == s
    assert Sim2.from_matrix(s.matrix) == s


def test_mesh_grid_order():
    g = get_mesh_grid_as_point_cloud(0, 2, 0, 1)
    assert g.tolist() == [[0, 0], [1, 0], [2, 0], [0, 1], [1, 1], [2, 1]]  # x fastest


def test_shard_assignment_balances_and_partitions():
    counts = [633, 10, 400, 399, 50, 700, 1, 1, 300, 300]
    for world in (1, 2, 4, 8):
        parts = sharding.assign_buildings(counts, world)
        flat = sorted(b for p in parts for b in p)
        assert flat == list(range(len(counts)))
        loads = [sum(counts[b] for b in p) for p in parts]
        assert max(loads) - min(loads) <= max(counts)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist

    dist.init_process_group("gloo", rank=rank, world_size=world)
    counts = [5, 3, 8, 1, 9, 2]
    mine = sharding.assign_buildings(counts, world)[rank]
    local = {b: counts[b] for b in mine}  # pretend each hypothesis was rendered
    total, per_rank = sharding.gather_totals(sum(local.values()), elapsed_s=0.5 + rank)
    q.put((rank, sorted(mine), total, per_rank))
    dist.destroy_process_group()


def test_gloo_world_size_2_shards_without_data_collective():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    (r0, b0, tot0, pr0), (r1, b1, tot1, pr1) = res
    assert sorted(b0 + b1) == [0, 1, 2, 3, 4, 5] and not set(b0) & set(b1)
    assert tot0 == tot1 == dict(units=28, max_elapsed_s=1.5)
    assert pr0 == pr1 and sum(pr0) == 28


def test_driver_enumeration_follows_the_reference_order(tmp_path):
    """scripts/render_dataset_bev.py:29-31, 87-91: pano id from the file name; label types in the reference's order,
    JSONs sorted, pair_idx restarting per label type."""
    from salve_b200 import driver

    assert driver.panoid_from_fpath("/x/panos/floor_01_partial_room_07_pano_19.jpg") == 19
    for lt, names in (("incorrect_alignment", ["5_9__b.json", "2_5__a.json"]), ("gt_alignment_approx", ["2_9__z.json", "2_5__y.json"])):
        d = tmp_path / "0007" / "floor_02" / lt
        d.mkdir(parents=True)
        for n in names:
            (d / n).write_text("{}")
    got = [(lt, k, os.path.basename(p)) for lt, k, p in driver.enumerate_floor_hypotheses(str(tmp_path), "0007", "floor_02")]
    assert got == [("gt_alignment_approx", 0, "2_5__y.json"), ("gt_alignment_approx", 1, "2_9__z.json"),
                   ("incorrect_alignment", 0, "2_5__a.json"), ("incorrect_alignment", 1, "5_9__b.json")]


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the reference's render_bev_pair from oracle/_ref, or the oracle port when those copies are absent,
    on all host cores) needs no GPU and prints one JSON line with the keys the driver reads."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "hypotheses/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["steps"] == 1 and line["n_gpus"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    have_ref = os.path.isdir(os.path.join(root, "oracle", "_ref", "salve"))
    assert line["cpu_baseline"]["kind"] == ("reference" if have_ref else "port") and line["cpu_baseline"]["cores"] >= 1
    assert "workload" in line["config"] and line["gpu_launches"] == 0


def test_local_rule_tables_against_an_independent_enumeration():
    """The host-built tables of the local rule (include/salve_bev.h: salve_bev_local_rule_tables; k_image.cuh LocalRule) restated in
    plain Python: the candidate triangles, their inside / on-circle masks over the 5 x 5 neighbourhood and, for every 12-bit
    neighbour pattern, the first four candidates whose vertices are present and whose circle holds none of the pattern's sites."""
    import ctypes
    import itertools

    from salve_b200 import _native as nat

    lib = nat.load()
    n = lib.salve_bev_local_rule_tables(None, 0)
    NPAT, MAXC = 4096, 128
    assert n == NPAT + 3 * MAXC
    tab = np.zeros(n, np.uint32)
    assert lib.salve_bev_local_rule_tables(tab.ctypes.data_as(ctypes.c_void_p), n) == n

    pos = [(0, -2), (-1, -1), (0, -1), (1, -1), (-2, 0), (-1, 0), (1, 0), (2, 0), (-1, 1), (0, 1), (1, 1), (0, 2)]  # (dx, dy) of pattern bit k
    bit = lambda p: (p[1] + 2) * 5 + p[0] + 2
    orient = lambda a, b, c: (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0])

    def incircle(a, b, c, d):
        ax, ay, bx, by, cx, cy = a[0] - d[0], a[1] - d[1], b[0] - d[0], b[1] - d[1], c[0] - d[0], c[1] - d[1]
        return (ax * ax + ay * ay) * (bx * cy - by * cx) - (bx * bx + by * by) * (ax * cy - ay * cx) + (cx * cx + cy * cy) * (ax * by - ay * bx)

    cands = []
    for i, j, k in itertools.combinations(range(12), 3):
        a, b, c = pos[i], pos[j], pos[k]
        o = orient(a, b, c)
        if o == 0:
            continue
        if o < 0:
            b, c = c, b
        if min(orient(a, b, (0, 0)), orient(b, c, (0, 0)), orient(c, a, (0, 0))) < 0:
            continue
        inside = on = 0
        fits = True
        for y in range(-8, 9):
            for x in range(-8, 9):
                if (x, y) in (a, b, c):
                    continue
                inc = incircle(a, b, c, (x, y))
                if inc < 0:
                    continue
                if max(abs(x), abs(y)) > 2:
                    fits = False
                elif (x, y) != (0, 0):
                    if inc > 0:
                        inside |= 1 << bit((x, y))
                    else:
                        on |= 1 << bit((x, y))
        if fits:
            cands.append(((a, b, c), (1 << i) | (1 << j) | (1 << k), inside, on))
    assert len(cands) == 96
    for cid, ((a, b, c), vmask, inside, on) in enumerate(cands):
        assert int(tab[NPAT + cid]) == inside and int(tab[NPAT + MAXC + cid]) == on
        assert int(tab[NPAT + 2 * MAXC + cid]) == bit(a) | (bit(b) << 5) | (bit(c) << 10)
    in12 = [sum(1 << m for m in range(12) if inside & (1 << bit(pos[m]))) for _, _, inside, _ in cands]
    n_truncated = 0
    for pat in range(NPAT):
        want = [cid for cid, (_, vmask, _, _) in enumerate(cands) if (pat & vmask) == vmask and not (pat & in12[cid])]
        got = [(int(tab[pat]) >> (8 * s)) & 0xFF for s in range(4)]
        assert got == (want[:4] + [0xFF] * 4)[:4], pat
        n_truncated += len(want) > 4
    # more than four viable candidates: 224 patterns with co-circular sites at distance 2 (the table keeps the first four; a query
    # whose triangle is one of the others goes to the window pass -- coverage, not correctness)
    assert n_truncated == 224
