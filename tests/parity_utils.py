"""Shared helpers for the GPU parity tests: compare every stage of one rendered image with the oracle."""

from __future__ import annotations

import numpy as np

from oracle import bev_oracle as bo
from oracle import canonical_dt as cdt


def oracle_canonical(st: "bo.Stages"):
    """Canonical-tie Delaunay interpolation of an oracle Stages (sites must be non-degenerate).
    Returns dict(rc, rgb, tri_v, interp, hull, tri_id, stats, check)."""
    rc, rgb = cdt.sort_sites(st.site_rc, st.site_rgb)
    tri_v, stats = cdt.triangulate(rc[:, 0], rc[:, 1], bo.IMG)
    chk = cdt.check_delaunay(rc[:, 0], rc[:, 1], tri_v)
    interp, hull, tri_id = cdt.rasterize(rc[:, 0], rc[:, 1], rgb, tri_v, bo.IMG, bo.IMG)
    return dict(rc=rc, rgb=rgb, tri_v=tri_v, interp=interp, hull=hull, tri_id=tri_id, stats=stats, check=chk)


def canonical_final(st: "bo.Stages", can) -> np.ndarray:
    """Final image built from the canonical interpolation, mask and flip as in the reference."""
    if st.degenerate:
        return np.zeros((bo.IMG, bo.IMG, 3), np.uint8)
    return np.flipud(can["interp"] * st.keep[:, :, None].astype(np.uint8))


def tri_pixel_set(tri_v_pix: np.ndarray) -> np.ndarray:
    """Real triangles (vertex = pixel id) as a lexicographically sorted array of sorted triples."""
    real = tri_v_pix[(tri_v_pix >= 0).all(1)]
    t = np.sort(real, axis=1)
    return t[np.lexsort((t[:, 2], t[:, 1], t[:, 0]))]


def oracle_tri_pixel_set(can) -> np.ndarray:
    rc = can["rc"]
    pix = rc[:, 0].astype(np.int64) * bo.IMG + rc[:, 1]
    tv = can["tri_v"]
    real = tv[(tv >= 0).all(1)]
    return tri_pixel_set(pix[real].astype(np.int32))


def check_query_triangles(qtri: np.ndarray, queries: np.ndarray, oracle_tris: np.ndarray) -> None:
    """Each query pixel's triangle (vertex pixel ids) is a triangle of the canonical mesh and contains the pixel."""
    W = queries.shape[1]
    rows, cols = np.nonzero(queries)
    t = np.sort(qtri[rows, cols].astype(np.int64), axis=1)
    base = np.int64(W) * queries.shape[0]
    code = lambda a: (a[:, 0] * base + a[:, 1]) * base + a[:, 2]
    assert np.isin(code(t), code(oracle_tris.astype(np.int64))).all(), "a query ended in a triangle that is not in the canonical mesh"
    v = qtri[rows, cols].astype(np.int64)
    vx, vy = v % W, v // W
    def orient(i, j, px, py):
        return (vx[:, j] - vx[:, i]) * (py - vy[:, i]) - (vy[:, j] - vy[:, i]) * (px - vx[:, i])
    a2 = orient(0, 1, vx[:, 2], vy[:, 2])
    assert (a2 != 0).all(), "degenerate query triangle"
    sgn = np.sign(a2)  # the tap stores vertex ids in ascending order, either orientation
    for i, j in ((0, 1), (1, 2), (2, 0)):
        assert (sgn * orient(i, j, cols, rows) >= 0).all(), "query pixel outside its triangle"


def tie_independent_mask(st: "bo.Stages", can) -> np.ndarray:
    """Pixels whose interpolated value is the same in every Delaunay triangulation: sites and
    pixels inside strict (tie-free) triangles."""
    tid = can["tri_id"]
    ok = tid >= 0
    strict_px = np.zeros_like(can["hull"])
    strict_px[ok] = can["check"]["strict_flag"][tid[ok]]
    site_px = np.zeros_like(can["hull"])
    site_px[can["rc"][:, 0], can["rc"][:, 1]] = True
    return strict_px | site_px


def hull_check(exact_hull: np.ndarray, scipy_hull: np.ndarray) -> int:
    """The exact closed convex hull vs SciPy's `defined` mask.

    They must agree except, at most, on pixels lying on the hull *boundary*: SciPy decides
    inside/outside with float64 barycentrics and a 100-ulp tolerance, which misclassifies a pixel
    that sits exactly on the long hull edge of a needle triangle (observed: twice-area 3 over a
    ~90 px edge).  Returns the number of such pixels; asserts everything else.
    """
    diff = exact_hull != scipy_hull
    if not diff.any():
        return 0
    assert not (scipy_hull & ~exact_hull).any(), "SciPy defines a pixel outside the exact hull"
    pad = np.pad(exact_hull, 1)
    interior = np.ones_like(exact_hull)
    for dy in (0, 1, 2):
        for dx in (0, 1, 2):
            interior &= pad[dy : dy + exact_hull.shape[0], dx : dx + exact_hull.shape[1]]
    assert not (diff & interior).any(), "hull masks differ away from the hull boundary"
    n = int(diff.sum())
    assert n <= max(4, 1e-4 * exact_hull.sum()), f"{n} boundary pixels differ"
    return n


def rgb_report(gpu_final: np.ndarray, st: "bo.Stages", can) -> dict:
    """Fractions for the RGB contract, on the un-flipped grid."""
    ref = np.flipud(st.final).astype(int)
    got = np.flipud(gpu_final).astype(int)
    d = np.abs(ref - got).max(2)
    d[can["hull"] & ~st.hull] = 0  # hull-boundary pixels SciPy's float tolerance drops (see hull_check)
    kept = st.keep & st.hull
    safe = kept & tie_independent_mask(st, can)
    return dict(
        kept=int(kept.sum()),
        safe_frac=float(safe.sum() / max(kept.sum(), 1)),
        all_gt0=float((d[kept] > 0).mean()),
        all_gt1=float((d[kept] > 1).mean()),
        safe_gt0=float((d[safe] > 0).mean()),
        safe_gt1=float((d[safe] > 1).mean()),
        safe_max=int(d[safe].max()) if safe.any() else 0,
        outside_kept_diff=int((d[~kept] > 0).sum()),
    )
