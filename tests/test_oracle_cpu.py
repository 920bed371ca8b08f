"""CPU tests of the oracle: against the golden vectors frozen from the reference, against the
reference's own known-answer tests for this path, and (in the build container) against the
imported reference itself."""

import hashlib

import os

import numpy as np
import pytest

from oracle import bev_oracle as bo
from oracle import canonical_dt as cdt
from oracle import synth


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_synthetic_inputs_are_stable(golden, golden_inputs):
    _, meta = golden
    rgb1, d1, rgb2, d2, R, t = golden_inputs
    assert sha(d1) == meta["inputs_sha"]["d1"] and sha(d2) == meta["inputs_sha"]["d2"]
    assert sha(rgb1) == meta["inputs_sha"]["rgb1"] and sha(rgb2) == meta["inputs_sha"]["rgb2"]
    g, _ = golden
    assert np.array_equal(R, g["R"]) and np.array_equal(t, g["t"])


@pytest.mark.parametrize("surf", ["floor", "ceiling"])
def test_oracle_matches_golden_reference_output(golden, golden_inputs, surf):
    """Integer stages must match the reference bit-for-bit.  The final image additionally depends on
    SciPy/Qhull tie-breaks, so it is compared exactly only under the SciPy that froze the goldens."""
    import scipy

    g, meta = golden
    rgb1, d1, rgb2, d2, R, t = golden_inputs
    s1, s2 = bo.render_pair(rgb1, d1, rgb2, d2, R, t, surf)
    for name, st in ((f"{surf}_1", s1), (f"{surf}_2", s2)):
        m = meta["images"][name]
        assert st.count_crop == m["count_crop"] and st.count_bbox == m["count_bbox"] and len(st.site_rc) == m["n_sites"]
        assert sha(st.key_grid.astype(np.int64)) == m["key_grid_sha"]
        assert np.array_equal(np.packbits(st.nonempty), g[f"{name}_nonempty"])
        assert np.array_equal(np.packbits(st.keep), g[f"{name}_keep"])
        assert np.array_equal(np.packbits(st.hull), g[f"{name}_hull"])
        if scipy.__version__ == meta["generated_with"]["scipy"]:
            assert np.array_equal(st.final, g[f"{name}_final"])
        else:
            d = np.abs(st.final.astype(int) - g[f"{name}_final"].astype(int)).max(2)
            assert (d > 1).mean() < 0.05


@pytest.mark.needs_reference
def test_oracle_equals_imported_reference():
    from oracle import ref_import

    rgb1, d1 = synth.synth_pano(512, 1024, 7, "smooth")
    rgb2, d2 = synth.synth_pano(512, 1024, 8, "iid")
    R, t = synth.synth_pose(3)
    r1, r2 = ref_import.render_bev_pair(rgb1, d1, rgb2, d2, R, t, "ceiling")
    o1, o2 = bo.render_pair_images(rgb1, d1, rgb2, d2, R, t, "ceiling")
    assert np.array_equal(r1, o1) and np.array_equal(r2, o2)


# ---- the reference's own known-answer tests for this path, replayed on the oracle ---------------------
ZORDER_KATS = [  # tests/utils/test_zorder_utils.py:8-66
    ([[0, 1, 0], [1, 2, 4], [0, 1, 5], [5, 6, 1]], 5, [False, True, True, True]),
    ([[0, 1, 0], [1, 2, 4], [2, 3, 5], [3, 4, 1]], 5, [True, True, True, True]),
    ([[0, 1, 0], [0, 1, 1], [0, 1, 2], [0, 1, 3]], 5, [False, False, False, True]),
    ([[0, 1, 0], [0, 1, 1], [0, 1, 10], [0, 1, 11]], 5, [False, True, False, False]),
    ([[0, 1, 0], [0, 1, 1], [0, 1, 2], [0, 1, 3]], 2, [False, False, False, True]),
]


@pytest.mark.parametrize("xyz,slices,expected", ZORDER_KATS)
def test_zorder_kats(xyz, slices, expected):
    xyz = np.array(xyz)
    got = bo.choose_elevated(xyz[:, 0], xyz[:, 1], xyz[:, 2], zmin=0, zmax=10, num_slices=slices)
    assert got.tolist() == expected


def test_hallucination_kat():
    """tests/utils/test_interpolation_utils.py:85-127 (K=3)."""
    sparse = np.zeros((6, 6), np.int64)
    sparse[0, 1] = 2; sparse[0, 3] = 4; sparse[2, 1] = 2; sparse[4, 1] = 2
    interp = np.tile(np.arange(1, 7), (6, 1))
    keep = bo.keep_mask(sparse > 0, 3)
    expected = np.array([[1, 2, 3, 4, 5, 0], [1, 2, 3, 4, 5, 0], [1, 2, 3, 0, 0, 0], [1, 2, 3, 0, 0, 0], [1, 2, 3, 0, 0, 0], [1, 2, 3, 0, 0, 0]])
    assert np.array_equal(keep * interp, expected)


def test_uint8_wrap_counts_as_empty():
    sp = np.zeros((1, 2, 3), np.uint8)
    sp[0, 0] = (16, 16, 1)  # 256 -> 0 in uint8
    sp[0, 1] = (3, 5, 7)
    assert bo.nonempty_mask(sp).tolist() == [[False, True]]


def test_sphere_directions():
    """tests/test_hohonet_pano_utils.py:8-24: -x at the pano centre, z up."""
    s = bo.uni_sphere_xyz(512, 1024)
    assert np.allclose(s[256, 512], [-1, 0, 0], atol=4e-3)
    assert np.allclose(s[0, 0], [0, 0, 1], atol=4e-3)
    assert np.allclose(s[511, 0], [0, 0, -1], atol=4e-3)
    assert np.allclose(np.linalg.norm(s, axis=2), 1.0)


def test_degenerate_guards():
    """tests/utils/test_interpolation_utils.py:8-59."""
    assert bo.is_degenerate(np.array([[1, 1], [1, 5], [1, 7], [1, 9]]))
    assert bo.is_degenerate(np.array([[1, 3], [5, 3], [7, 3], [9, 3]]))
    assert bo.is_degenerate(np.array([[0, 0], [3, 3]]))
    assert not bo.is_degenerate(np.array([[0, 0], [3, 0], [3, 3], [0, 3]]))


# ---- canonical Delaunay checker pinned against SciPy ------------------------------------------------------
def _random_sites(rng, h, w, dens):
    occ = rng.random((h, w)) < dens
    return np.nonzero(occ)


@pytest.mark.parametrize("seed", range(6))
def test_canonical_dt_against_scipy(seed):
    import scipy.interpolate
    import scipy.spatial

    rng = np.random.default_rng(seed)
    h, w = 40 + 7 * seed, 60 - 5 * seed
    rows, cols = _random_sites(rng, h, w, [0.05, 0.2, 0.5, 0.9, 0.3, 0.1][seed])
    tri_v, stats = cdt.triangulate(rows, cols, w)
    assert stats["init_check"] == 0 and stats["final_check"] == 0 and stats["residual_ties"] == 0
    chk = cdt.check_delaunay(rows, cols, tri_v)
    assert chk["violations"] == 0
    pts = np.stack([cols, rows], 1).astype(np.float64)
    dl = scipy.spatial.Delaunay(pts)
    assert (tri_v >= 0).all(1).sum() == len(dl.simplices)
    simp = dl.simplices.astype(np.int32)
    a, b, c = simp[:, 0], simp[:, 1], simp[:, 2]
    o = (cols[b] - cols[a]).astype(np.int64) * (rows[c] - rows[a]) - (rows[b] - rows[a]).astype(np.int64) * (cols[c] - cols[a])
    simp[o < 0] = simp[o < 0][:, [0, 2, 1]]
    chk_s = cdt.check_delaunay(rows, cols, simp)
    assert chk_s["violations"] == 0
    real = tri_v[(tri_v >= 0).all(1)]
    mine = set(map(tuple, np.sort(real[chk["strict_flag"][(tri_v >= 0).all(1)]], 1)))
    theirs = set(map(tuple, np.sort(simp[chk_s["strict_flag"]], 1)))
    assert mine == theirs  # tie-independent triangles are in every Delaunay triangulation
    rgb = rng.integers(1, 256, (len(rows), 3)).astype(np.uint8)
    interp, hull, tid = cdt.rasterize(rows, cols, rgb, tri_v, h, w)
    xg, yg = np.meshgrid(np.arange(w), np.arange(h))
    vals = scipy.interpolate.griddata(pts, rgb.astype(np.float64), np.stack([xg.ravel(), yg.ravel()], 1).astype(np.float64), "linear")
    vals = vals.reshape(h, w, 3)
    import parity_utils as pu

    pu.hull_check(hull, ~np.isnan(vals[:, :, 0]))
    safe = np.zeros((h, w), bool)
    ok = tid >= 0
    safe[ok] = chk["strict_flag"][tid[ok]]
    safe[rows, cols] = True
    with np.errstate(invalid="ignore"):
        d = np.abs(interp.astype(int) - np.nan_to_num(vals).astype(np.uint8).astype(int)).max(2)
    assert d[safe].max() <= 1


def test_canonical_dt_is_order_and_start_independent():
    """Same site set, mirrored: the canonical triangulation must map onto itself only where ties are
    absent; but re-running on the same set must be identical (pure function of the set)."""
    rng = np.random.default_rng(11)
    rows, cols = _random_sites(rng, 50, 50, 0.3)
    t1, _ = cdt.triangulate(rows, cols, 50)
    t2, _ = cdt.triangulate(rows.copy(), cols.copy(), 50)
    assert np.array_equal(t1, t2)


def test_single_point_rows_and_collinear():
    rows = np.arange(8, dtype=np.int32)
    cols = np.array([3, 1, 4, 1, 5, 9, 2, 6], np.int32)
    tri_v, stats = cdt.triangulate(rows, cols, 12)
    assert stats["final_check"] == 0 and cdt.check_delaunay(rows, cols, tri_v)["violations"] == 0
    assert (tri_v >= 0).all(1).sum() > 0
    cols = rows.copy()  # oblique line: no real triangle
    tri_v, stats = cdt.triangulate(rows, cols, 12)
    assert (tri_v >= 0).all(1).sum() == 0


# ---- verifier pre-processing ---------------------------------------------------------------------
def _golden_preprocess_inputs():
    import importlib.util

    spec = importlib.util.spec_from_file_location("mgp", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts", "make_golden_preprocess.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.golden_inputs()


def test_preprocess_oracle_equals_reference_golden():
    """tests/golden/preprocess_c1.npz was produced by the reference's own transform classes (scripts/make_golden_preprocess.py)."""
    import hashlib

    from oracle import preprocess_oracle as po

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "preprocess_c1.npz"))
    out = po.preprocess_quadruplet(*_golden_preprocess_inputs())
    assert out.shape == (12, 224, 224) and out.dtype == np.float32
    assert np.array_equal(out[:, g["rows"], :], g["values"])
    assert np.array_equal(np.frombuffer(hashlib.sha256(np.ascontiguousarray(out).tobytes()).digest(), np.uint8), g["sha256"])


def test_preprocess_oracle_resize_equals_cv2():
    """The restated fixed-point bilinear against OpenCV itself (third-party arithmetic of the chain), down-scaling sizes."""
    cv2 = pytest.importorskip("cv2")
    from oracle import preprocess_oracle as po

    rng = np.random.default_rng(3)
    for shape, out in (((501, 501, 3), (234, 234)), ((501, 501, 3), (224, 224)), ((64, 37, 3), (50, 29)), ((300, 200, 3), (299, 101))):
        img = rng.integers(0, 256, size=shape, dtype=np.int64).astype(np.uint8)
        assert np.array_equal(cv2.resize(img, (out[1], out[0]), interpolation=cv2.INTER_LINEAR), po.resize_linear_u8(img, out[0], out[1]))


def test_resize_at_scale_two_is_the_rounded_box_mean():
    """cv2.resize(INTER_LINEAR) at a scale of exactly 2 == (sum of the 2x2 block + 2) >> 2: the identity the
    fused full-resolution colour gather relies on."""
    from oracle import preprocess_oracle as po

    rng = np.random.default_rng(4)
    big = rng.integers(0, 256, size=(64, 96, 3), dtype=np.int64).astype(np.uint8)
    box = (big[0::2, 0::2].astype(np.int32) + big[0::2, 1::2] + big[1::2, 0::2] + big[1::2, 1::2] + 2) >> 2
    assert np.array_equal(po.resize_linear_u8(big, 32, 48), box.astype(np.uint8))
    cv2 = pytest.importorskip("cv2")
    assert np.array_equal(cv2.resize(big, (48, 32), interpolation=cv2.INTER_LINEAR), box.astype(np.uint8))


@pytest.mark.needs_reference
def test_layout_golden_is_what_the_reference_draws():
    """tests/golden/layout_c1.npz (scripts/make_golden_layout.py) against the imported, unmodified reference and the real cv2."""
    import os

    from oracle import layout_synth, ref_import

    ref = ref_import.load()
    import salve.common.pano_data as pano_data

    g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "layout_c1.npz"))
    for k, (s1, s2, ps) in enumerate([(0, 1, 3), (2, 3, 4), (4, 5, 6)]):
        graph = layout_synth.nodes([s1, s2], wdo_cls=pano_data.WDO, sim2_cls=ref.sim2.Sim2)
        R, t = synth.synth_pose(ps)
        i1, i2 = ref.bru.rasterize_room_layout_pair(ref.sim2.Sim2(R, (t * 0.25).astype(np.float32), 1.0), graph, "b", "f", 0, 1)
        assert np.array_equal(i1, g[f"case{k}_img1"]) and np.array_equal(i2, g[f"case{k}_img2"])


def test_numpy_matmul_rounding_order_assumed_by_the_cuda_path():
    """rot_pose() (salve_b200/csrc/k_splat.cuh) spells out how numpy evaluates the (N,2) @ (2,2) products of
    bev_rendering_utils.py:443-451 on the strided view xyzrgb[:, :2]: out_j = fma(p1, M[j][1], round(p0 * M[j][0])) -- first product
    rounded, second fused.  That order belongs to the BLAS / numpy build, not to the reference; this test pins it
    on the host that runs the oracle, so that a different build shows up here and not as a pixel flip next to a .5 boundary."""
    from fractions import Fraction

    rng = np.random.default_rng(123)
    n = 4000
    xyz = np.zeros((n, 6))
    xyz[:, :2] = rng.uniform(-7, 7, (n, 2))
    th = rng.uniform(-np.pi, np.pi)
    M = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]]).astype(np.float32)
    got = xyz[:, :2] @ M.T  # the reference's expression: float64 strided view times the transposed float32 matrix
    Md = M.astype(np.float64)
    same = 0
    for i in range(n):
        p0, p1 = float(xyz[i, 0]), float(xyz[i, 1])
        ok = True
        for j in range(2):
            first = p0 * float(Md[j, 0])  # rounded product
            fused = float(Fraction(p1) * Fraction(float(Md[j, 1])) + Fraction(first))  # exact, rounded once
            ok &= fused == got[i, j]
        same += ok
    frac = same / n
    if frac < 1.0:
        pytest.skip(f"this numpy/BLAS build rounds (N,2)@(2,2) differently on {100 * (1 - frac):.2f} % of rows: pixel indices within 1 ulp of a "
                    ".5 boundary may differ from a reference run on this host (the CUDA path follows the build SURVEY.md measured)")
    assert frac == 1.0
