"""Grid build parity: exact EDT (bit-exact integer distances), fused Gaussian, splat, .grid files."""
import numpy as np
import pytest

from conftest import bits

pytestmark = pytest.mark.gpu


def _small_map(O, seed=3, extent=(6.4, 5.0, 3.1), resol=0.1, n=900, lattice="centre"):
    rng = np.random.RandomState(seed)
    size = np.round(np.array(extent) / resol).astype(int)
    vox = np.stack([rng.randint(0, size[a], n) for a in range(3)], 1)
    if lattice == "centre":
        pts = (vox + 0.5) * resol
    elif lattice == "corner":
        pts = vox * resol
    else:  # mixed half-lattice
        pts = (vox + rng.randint(0, 2, (n, 3)) * 0.5) * resol
    pts = np.concatenate([[[0, 0, 0], list(extent)], pts]).astype(np.float32)
    mn, mx = O.compute_bounds(pts)
    return pts, O.Map(mn, mx, resol, 0.05)


@pytest.mark.parametrize("lattice,sub", [("centre", 2), ("corner", 1), ("corner", 2), ("mixed", 2)])
def test_edt_bit_exact_small(O, lattice, sub):
    from amcl3d_test_b200 import Grid3d

    pts, m = _small_map(O, lattice=lattice)
    ref = O.edt_exact(m, pts, sub=sub)
    brute = O.edt_brute(m, pts, sub=sub)
    assert np.array_equal(ref, brute)
    g = Grid3d(0)
    g.computePointCloud(pts, m.resol)
    assert g.info()["size"] == m.size
    g.computeGrid(pts, 0.05, mode=O.GRID_QUIRK, sub=sub, keep_edt=True)
    assert np.array_equal(g.edt(), ref)
    prob = g.downloadGrid()
    ref_prob = O.grid_from_edt(m, ref, sub, O.GRID_QUIRK)
    assert np.allclose(prob, ref_prob, rtol=1e-6, atol=1e-30)
    assert np.array_equal(prob == 0, ref_prob == 0)


@pytest.mark.parametrize("sub", [1, 2])
def test_snap_to_the_lattice_at_and_around_ties(O, sub):
    """Map points half-way between two lattice nodes and a few float steps either side of it: the device snaps with a
    reciprocal product and falls back to the real f64 divide near a tie -- same node as lrint((p - min) / pitch) always."""
    from amcl3d_test_b200 import Grid3d

    rng = np.random.RandomState(11)
    resol, extent = 0.1, (5.0, 4.0, 2.0)
    pitch = resol / sub
    size = np.round(np.array(extent) / resol).astype(int)
    k = np.stack([rng.randint(0, size[a] * sub, 1500) for a in range(3)], 1)
    pts = ((k + 0.5) * pitch).astype(np.float32)  # ties up to the float rounding of the coordinate
    for step in range(1, 4):  # ... and up to three float steps above / below on random axes
        up = np.nextafter(pts, np.float32(np.inf)) if step % 2 else np.nextafter(pts, np.float32(-np.inf))
        pts = np.where(rng.randint(0, 2, pts.shape).astype(bool), up, pts).astype(np.float32)
    pts = np.concatenate([[[0, 0, 0], list(extent)], pts]).astype(np.float32)
    mn, mx = O.compute_bounds(pts)
    m = O.Map(mn, mx, resol, 0.05)
    ref = O.edt_exact(m, pts, sub=sub)
    g = Grid3d(0)
    g.computePointCloud(pts, resol)
    g.computeGrid(pts, 0.05, sub=sub, keep_edt=True)
    assert np.array_equal(g.edt(), ref)
    exact_field = g.downloadGrid()
    g.computeGrid(pts, 0.05, sub=sub)  # the fused fast build snaps through the same function
    assert np.array_equal(bits(g.downloadGrid()), bits(exact_field))


def test_edt_c1_map(O, world_c1):
    from amcl3d_test_b200 import Grid3d

    w = world_c1
    g = Grid3d(0)
    mn, mx = g.computePointCloud(w["synth"].points, 0.1)
    assert np.array_equal(mn, w["map"].mn) and np.array_equal(mx, w["map"].mx)
    g.computeGrid(w["synth"].points, 0.05, sub=2, keep_edt=True)
    assert np.array_equal(g.edt(), w["d2"])
    assert np.allclose(g.downloadGrid(), w["prob"], rtol=1e-6, atol=1e-30)
    # gaussian (non-quirk) mode
    g.computeGrid(w["synth"].points, 0.05, mode=O.GRID_GAUSSIAN, sub=2)
    assert np.allclose(g.downloadGrid(), O.grid_from_edt(w["map"], w["d2"], 2, O.GRID_GAUSSIAN), rtol=1e-6, atol=1e-30)


def test_edt_empty_rows_and_single_site(O):
    from amcl3d_test_b200 import Grid3d

    pts = np.array([[0, 0, 0], [3.2, 2.0, 1.0], [1.05, 0.55, 0.35]], np.float32)  # two corners + one centre site
    mn, mx = O.compute_bounds(pts)
    m = O.Map(mn, mx, 0.1, 0.05)
    ref = O.edt_exact(m, pts, sub=2)
    g = Grid3d(0)
    g.computePointCloud(pts, 0.1)
    g.computeGrid(pts, 0.05, sub=2, keep_edt=True)
    assert np.array_equal(g.edt(), ref)
    # dense single class with whole empty rows/planes
    rng = np.random.RandomState(0)
    wall = np.stack([np.full(400, 7), rng.randint(0, 20, 400), rng.randint(0, 10, 400)], 1)
    pts2 = np.concatenate([pts[:2], (wall + 0.5) * 0.1]).astype(np.float32)
    ref2 = O.edt_exact(m, pts2, sub=2)
    g.computeGrid(pts2, 0.05, sub=2, keep_edt=True)
    assert np.array_equal(g.edt(), ref2)


def test_splat_matches_compute_grid2(O, world_c1):
    from amcl3d_test_b200 import Grid3d

    rng = np.random.RandomState(9)
    pts = (rng.uniform(0, 1, (5000, 3)) * [8.0, 6.0, 3.0]).astype(np.float32)  # arbitrary (off-lattice) points
    mn, mx = O.compute_bounds(pts)
    m = O.Map(mn, mx, 0.3, 0.05)  # Grid3d.cpp:156 hard-codes 0.3
    ref = O.compute_grid2(m, pts)
    g = Grid3d(0)
    g.computePointCloud(pts, 0.3)
    g.computeGrid(pts, 0.05, mode=O.GRID_SPLAT)
    got = g.downloadGrid()
    assert np.array_equal(got == 0, ref == 0)
    assert np.allclose(got, ref, rtol=1e-6, atol=1e-30)


def test_grid_file_roundtrip(O, world_c1, gpu_c1, tmp_path):
    from amcl3d_test_b200 import Grid3d

    w = world_c1
    path = tmp_path / "cloud.grid"
    assert gpu_c1.saveGrid(path)
    # byte-identical to the oracle's writer (Grid3d.cpp:388-395 format)
    ref_path = tmp_path / "ref.grid"
    assert O.save_grid(ref_path, w["map"]) == 0
    assert path.read_bytes() == ref_path.read_bytes()
    g = Grid3d(0)
    g.setMap(w["map"].mn, w["map"].mx, 0.1)
    assert g.loadGrid(path, 0.05)
    assert np.array_equal(bits(g.downloadGrid()), bits(w["prob"]))
    assert g.info()["size"] == w["map"].size
    assert not g.loadGrid(path, 0.06)  # Grid3d.cpp:422-427
    assert not g.loadGrid(tmp_path / "missing.grid", 0.05)
    with pytest.raises(RuntimeError):
        g.computePointCloud(np.zeros((1, 3), np.float32), 0.1)  # PointCloudTools.cpp:89-91


@pytest.mark.parametrize("lattice,sub", [("centre", 2), ("corner", 1), ("corner", 2), ("mixed", 2)])
@pytest.mark.parametrize("sigma,mode", [(0.05, 0), (0.12, 0), (0.05, 1), (0.2, 1), (0.6, 0)])
def test_capped_window_passes_give_the_same_field(O, lattice, sub, sigma, mode):
    """When the exact transform is not kept the y/z passes search only the window the expf underflow allows
    (min(d2, cap)); the float grid must be the same bits as the one the exact envelope passes write -- for small and
    wide windows, both Gaussians, and the fallback when the window is too wide (sigma 0.6)."""
    from amcl3d_test_b200 import Grid3d

    pts, m = _small_map(O, lattice=lattice)
    mode = O.GRID_QUIRK if mode == 0 else O.GRID_GAUSSIAN
    g = Grid3d(0)
    g.computePointCloud(pts, m.resol)
    g.computeGrid(pts, sigma, mode=mode, sub=sub, keep_edt=True)   # exact envelopes
    exact = g.downloadGrid()
    rep_exact = g.info()["rep"]
    g.computeGrid(pts, sigma, mode=mode, sub=sub)                  # capped windows (or the fallback)
    fast = g.downloadGrid()
    assert np.array_equal(bits(fast), bits(exact))
    assert g.info()["rep"] == rep_exact
    assert (exact > 0).any()


def test_capped_window_passes_c1_update_bits(O, world_c1):
    """C1 map through the fast passes: the coded grid drives update() to the oracle's bits."""
    from amcl3d_test_b200 import Grid3d, ParticleFilter

    w = world_c1
    g = Grid3d(0)
    g.computePointCloud(w["synth"].points, 0.1)
    g.computeGrid(w["synth"].points, 0.05, sub=2)
    prob = g.downloadGrid()
    assert np.allclose(prob, w["prob"], rtol=1e-6, atol=1e-30) and np.array_equal(prob == 0, w["prob"] == 0)
    m = O.Map(w["map"].mn, w["map"].mx, 0.1, 0.05, prob=prob)
    pf = ParticleFilter(g)
    pf.p_ = w["T"]
    pf.update(w["cloud"])
    ref = O.update(m, O.particles(w["T"]), w["cloud"])
    assert np.array_equal(bits(pf.p_[:, 6]), bits(ref[:, 6]))


@pytest.mark.parametrize("sub,lattice", [(2, "centre"), (1, "corner"), (2, "corner")])
def test_fused_build_codes_and_sparse_patch(O, sub, lattice):
    """The fused fast build (xpass8 / ypass_dpx / zpass_dpx + sparse patch, gridbuild.cu): odd sizes that are no multiple
    of the 64 x 4 x 32 tile, a half-empty map, and a few isolated sites of OTHER lattice classes inside the empty half
    (patched in after the passes, extending the dictionary).  The float grid must equal the exact-envelope build bit for
    bit, and the bricked codes must drive update() to the same weights as the linear float grid."""
    from amcl3d_test_b200 import Grid3d, ParticleFilter

    resol, sigma = 0.1, 0.05
    extent = (7.25, 4.5, 4.1)  # 73 x 45 x 41 voxels: odd x size (scalar float stores), no multiple of the 64 x 4 x 32 tile
    size = np.ceil(np.array(extent) / resol).astype(int)
    rng = np.random.RandomState(31)
    vox = np.stack([rng.randint(0, 30, 1500), rng.randint(0, size[1], 1500), rng.randint(0, size[2], 1500)], 1)
    dense = ((vox + 0.5) if lattice == "centre" else vox) * resol
    lone = np.array([[5.0, 2.0, 1.0], [5.55, 2.05, 1.3], [6.0, 3.05, 2.0], [4.45, 1.0, 3.05], [6.85, 0.5, 0.25]])
    if sub == 1:
        lone = np.round(lone / resol) * resol  # sub = 1 has a single class: the lone sites are then ordinary dense sites
    pts = np.concatenate([[[0, 0, 0], list(extent)], dense, lone]).astype(np.float32)
    g = Grid3d(0)
    mn, mx = g.computePointCloud(pts, resol)
    size = np.array(g.info()["size"])  # (74, 45, 41): 7.3f / 0.1 rounds up
    assert size[0] % 2 == 0 or size[0] % 64 != 0
    g.computeGrid(pts, sigma, sub=sub, keep_edt=True)          # exact envelope passes + legacy coder
    exact = g.downloadGrid()
    g.computeGrid(pts, sigma, sub=sub)                         # fused fast build
    fast = g.downloadGrid()
    assert g.info()["rep"] == 1
    assert np.array_equal(bits(fast), bits(exact))
    m = O.Map(mn, mx, resol, sigma, prob=fast)
    lin = np.ravel_multi_index((20, 10, 50), (size[2], size[1], size[0]))  # voxel (50, 10, 20) = corner (5.0, 1.0, 2.0)
    assert fast[lin] >= 0
    # particles around the lone sites, cloud = all map points seen from there: the patched codes are what they read
    n = 4096
    P = np.zeros((n, 7), np.float32)
    P[:, :6] = np.array([5.2, 2.0, 1.5, 0, 0, 0.2], np.float32) + rng.normal(0, 1, (n, 6)).astype(np.float32) * np.array(
        [0.4, 0.4, 0.3, 0.1, 0.1, 0.3], np.float32)
    world = np.concatenate([dense[rng.choice(dense.shape[0], 600, replace=False)], np.repeat(lone, 40, 0)])
    cloud = (world - np.array([5.2, 2.0, 1.5])).astype(np.float32)
    cloud += rng.normal(0, 0.02, cloud.shape).astype(np.float32)
    pf = ParticleFilter(g)
    pf.p_ = P
    pf.update(cloud)
    coded = pf.p_[:, 6].copy()
    ref = O.update(m, O.particles(P), cloud)
    assert (ref[:, 6] > 0).mean() > 0.5
    assert np.array_equal(bits(coded), bits(ref[:, 6]))
    g.useCodes(False)
    pf2 = ParticleFilter(g)
    pf2.p_ = P
    pf2.update(cloud)
    assert np.array_equal(bits(pf2.p_[:, 6]), bits(coded))


def test_fused_build_gaussian_and_wider_windows(O):
    """Window half-widths 12 / 20 / 28 of the DPX passes (sigma 0.03, 0.05, 0.09) and the plain-Gaussian mode."""
    from amcl3d_test_b200 import Grid3d

    pts, m = _small_map(O, seed=8, extent=(9.7, 3.3, 4.9), n=700)
    g = Grid3d(0)
    g.computePointCloud(pts, m.resol)
    for sigma, mode in ((0.03, O.GRID_QUIRK), (0.05, O.GRID_QUIRK), (0.09, O.GRID_QUIRK), (0.03, O.GRID_GAUSSIAN)):
        g.computeGrid(pts, sigma, mode=mode, sub=2, keep_edt=True)
        exact = g.downloadGrid()
        g.computeGrid(pts, sigma, mode=mode, sub=2)
        assert np.array_equal(bits(g.downloadGrid()), bits(exact)), (sigma, mode)


def test_off_lattice_map_points_stay_within_the_stated_snap_bound(O):
    """computeGrid on the device is an exact EDT on the lattice of pitch resol/sub: map points that do not lie on it are
    snapped first (include/amcl3d_b200.h).  The reference queries a kd-tree with the true float positions
    (PointCloudTools.cpp:152-162).  For arbitrary clouds the two therefore differ, by a bounded amount: a point moves by at
    most half a lattice pitch per axis, so the nearest-point DISTANCE of every voxel is off by at most
    sqrt(3)/2 * resol/sub, and the field -- monotone in the distance -- lies between the reference's values at
    d - bound and d + bound.  Both bounds are asserted against the compiled reference computeGrid run on the true points."""
    from amcl3d_test_b200 import Grid3d

    resol, sub, sigma = 0.1, 2, 0.05
    rng = np.random.RandomState(17)
    extent = np.array([3.0, 2.4, 1.6])
    pts = np.concatenate([[[0, 0, 0], list(extent)], rng.uniform(0, 1, (350, 3)) * extent]).astype(np.float32)
    g = Grid3d(0)
    mn, mx = g.computePointCloud(pts, resol)
    g.computeGrid(pts, sigma, sub=sub, keep_edt=True)
    size = np.array(g.info()["size"])
    d_edt = np.sqrt(g.edt().astype(np.float64)) * (resol / sub)
    # true nearest-point distance of every voxel corner (brute force, f64)
    ax = [mn[a] + np.arange(size[a]) * resol for a in range(3)]
    Z, Y, X = np.meshgrid(ax[2], ax[1], ax[0], indexing="ij")
    q = np.stack([X.ravel(), Y.ravel(), Z.ravel()], 1)
    d_true = np.full(q.shape[0], np.inf)
    for p in pts.astype(np.float64):
        d_true = np.minimum(d_true, np.sqrt(((q - p) ** 2).sum(1)))
    bound = np.sqrt(3.0) / 2.0 * resol / sub
    assert np.abs(d_edt - d_true).max() <= bound * (1 + 1e-6)
    assert np.abs(d_edt - d_true).max() > 0.1 * bound  # the cloud really is off the lattice
    # the field against the reference's own computeGrid (compiled from /root/reference here; oracle port elsewhere)
    m = O.Map(mn, mx, resol, sigma)
    ref = O.ref_compute_grid(pts, resol, sigma, which=0)["prob"] if O.ref_available() else O.compute_grid_nn(m, pts, O.GRID_QUIRK)
    c1 = 1.0 / (sigma * np.sqrt(2 * np.pi))
    c2 = 1.0 / (2 * sigma * sigma)
    f = lambda d: c1 * np.exp(-np.minimum((d * d) ** 2 * c2, 700.0))  # the d^4 quirk: dist is already squared  # noqa: E731
    got = g.downloadGrid().astype(np.float64)
    hi = f(np.maximum(d_true - bound, 0.0)) * (1 + 1e-4) + 1e-30
    lo = f(d_true + bound) * (1 - 1e-4) - 1e-30
    assert np.all(got <= hi) and np.all(got >= lo)
    assert np.allclose(ref, f(d_true), rtol=1e-4, atol=1e-30)  # and the reference is what the bound is centred on
