"""The reference-compatible C++ classes (include/amcl3d/{Grid3d,ParticleFilter,PointCloudTools}.h) over the C-ABI:
built here, exercised end to end on the GPU by tests/cpp/host_api_test.cpp and checked against the CPU oracle."""
import struct
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import bits

ROOT = Path(__file__).resolve().parents[1]


def test_host_library_builds_and_exports_the_reference_surface():
    from amcl3d_test_b200 import build

    lib = build.build_host()
    out = subprocess.run(["nm", "-DC", "--defined-only", str(lib)], capture_output=True, text=True).stdout
    for sym in [
        "amcl3d::Grid3d::open(", "amcl3d::Grid3d::computeCloudWeight(", "amcl3d::Grid3d::isIntoMap(float, float, float) const",
        "amcl3d::Grid3d::loadPCD(", "amcl3d::Grid3d::point2grid(", "amcl3d::ParticleFilter::relocPose(",
        "amcl3d::ParticleFilter::predict(", "amcl3d::ParticleFilter::update(", "amcl3d::ParticleFilter::particleNormalize()",
        "amcl3d::ParticleFilter::resample()", "amcl3d::ParticleFilter::uniformSample(int&)", "amcl3d::ParticleFilter::getMean()",
        "amcl3d::ParticleFilter::getBestParticle()", "amcl3d::ParticleFilter::setParticleNum(int, int)",
        "amcl3d::computePointCloud(", "amcl3d::computeGrid(", "amcl3d::computeGrid2(",
        "amcl3d::MonteCarloLocalization::PFResample(bool)", "amcl3d::MonteCarloLocalization::computeSampleNum(int)",
        "amcl3d::MonteCarloLocalization::updateSampleHeight()", "amcl3d::MonteCarloLocalization::RelocPose(",
        "amcl3d::MonteCarloLocalization::PFMove(", "amcl3d::MonteCarloLocalization::getMapHeight(",
        "amcl3d::MonteCarloLocalization::downsampleInput(", "amcl3d::MonteCarloLocalization::updateDownsampled()",
    ]:
        assert sym in out, sym


def _read_dump(path):
    b = Path(path).read_bytes()
    off, out = 0, {}
    while off < len(b):
        (ln,) = struct.unpack_from("<I", b, off)
        off += 4
        name = b[off:off + ln].decode()
        off += ln
        rows, cols = struct.unpack_from("<QQ", b, off)
        off += 16
        out[name] = np.frombuffer(b, "<f4", rows * cols, off).reshape(rows, cols).copy()
        off += rows * cols * 4
    return out


@pytest.mark.gpu
def test_cpp_classes_match_the_oracle(O, tmp_path):
    from amcl3d_test_b200 import build, synth

    build.build_host()
    sm = synth.make_map((8.0, 7.0, 3.0), 0.1)
    cloud = synth.make_cloud(sm, 700, 3.2)
    P = synth.make_particles(sm, 500, "T")
    P[::9, :3] += 20.0  # some particles outside the map
    seed, S = 4242, 350
    inp = tmp_path / "in.bin"
    with open(inp, "wb") as f:
        f.write(struct.pack("<IIIIi", sm.points.shape[0], cloud.shape[0], P.shape[0], seed, S))
        f.write(struct.pack("<dd", 0.1, 0.05))
        f.write(np.ascontiguousarray(sm.points, np.float32).tobytes())
        f.write(np.ascontiguousarray(cloud, np.float32).tobytes())
        f.write(np.ascontiguousarray(P, np.float32).tobytes())
    work = tmp_path / "mapdir"
    work.mkdir()
    r = subprocess.run([str(build.HOST_TEST), str(inp), str(tmp_path / "out.bin"), str(work)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    d = _read_dump(tmp_path / "out.bin")

    assert d["weight_before_open"][0, 0] == 0 and d["into_before_open"][0, 0] == 0  # Grid3dTest.cpp:160-161
    assert d["compute_point_cloud_throws"][0, 0] == 3  # PointCloudTools.cpp:86-91
    assert d["into_map_111"][0, 0] == 1 and d["into_map_far"][0, 0] == 0  # Grid3dTest.cpp:260-265

    mn, mx = O.compute_bounds(sm.points)
    m = O.Map(mn, mx, 0.1, 0.05)
    assert list(d["info"][0, 2:5]) == list(m.size)
    grid = d["grid"].reshape(-1)
    ref_grid = O.grid_from_edt(m, O.edt_exact(m, sm.points, sub=2), 2, O.GRID_QUIRK)
    assert np.allclose(grid, ref_grid, rtol=1e-6, atol=1e-30)
    m.set_prob(grid)  # identical grid on both sides from here on

    W = O.update(m, O.particles(P), cloud)
    assert np.array_equal(bits(d["after_update"]), bits(W))
    assert bits(d["single_pose_weight"][0, 0]) == bits(O.cloud_weight(m, cloud, P[0, :6]))
    assert np.array_equal(bits(d["after_add_particle_weight"]), bits(O.scale_by_max(W.copy())))
    N = W.copy()
    O.normalize(m, N)
    assert np.array_equal(bits(d["after_normalize"]), bits(N))
    assert np.array_equal(bits(d["mean"][0]), bits(O.mean(N)))
    assert np.array_equal(bits(d["best"][0]), bits(O.best(N)))
    out, _ = O.resample(N, d["resample_u01"][0, 0])
    assert np.array_equal(bits(d["after_resample"]), bits(out))
    out, _, k = O.uniform_sample(m, W.copy(), S)
    assert d["after_uniform_sample"].shape[0] == S
    assert np.array_equal(bits(d["after_uniform_sample"]), bits(out))

    init = np.array([P[0, 0], P[0, 1], P[0, 2], 0, 0, 0.3], np.float32)
    dev = np.array([0.5, 0.5, 0.4, 0.2, 0.2, 0.4], np.float32)
    R, mean = O.reloc_pose(P.shape[0], init, dev, d["reloc_noise"].reshape(-1))
    assert np.array_equal(bits(d["after_init"]), bits(R))
    assert np.array_equal(bits(d["init_mean"][0, :6]), bits(mean[:6]))
    Q = O.predict(R.copy(), np.array([0.31, -0.12, 0.02, 0.001, -0.002, 0.05]), d["predict_noise"].reshape(-1))
    assert np.array_equal(bits(d["after_predict"]), bits(Q))

    # Grid3d::open on a PCD directory: voxel-filtered union -> computeGrid2 at 0.3 m -> cloud.grid, then the cache
    assert d["open_missing_dir"][0, 0] == 0 and d["open_ok"][0, 0] == 1 and d["open_cached_ok"][0, 0] == 1
    fc = d["open_filtered_cloud"]
    assert 10 < fc.shape[0] < sm.points.shape[0]
    # loadPCD (Grid3d.cpp:88-130): points bucketed by keyframe id (intensity = i % 5), concatenated per keyframe as surf,
    # outlier, corner, then pcl::VoxelGrid at 0.3 m -- the oracle's restatement of that filter on the same order
    n_map = sm.points.shape[0]
    ids = np.arange(n_map)
    kind = ids % 3  # the driver deals the map points out as corner (0), surf (1), outlier (2)
    order = np.concatenate([ids[(ids % 5 == k) & (kind == c)] for k in range(5) for c in (1, 2, 0)])
    merged = np.concatenate([sm.points[order], (order % 5).astype(np.float32)[:, None]], 1)
    ref_fc = O.voxel_filter(merged, 0.3, ioff=3)
    assert np.array_equal(bits(fc), bits(ref_fc[:, :3]))
    fmn, fmx = O.compute_bounds(fc)
    fm = O.Map(fmn, fmx, 0.3, 0.05)
    assert list(d["open_size"][0]) == list(fm.size)
    ref_open = O.compute_grid2(fm, fc)
    assert np.array_equal(d["open_grid"].reshape(-1) == 0, ref_open == 0)
    assert np.allclose(d["open_grid"].reshape(-1), ref_open, rtol=1e-6, atol=1e-30)
    assert np.array_equal(bits(d["open_cached_grid"]), bits(d["open_grid"]))
    assert (work / "cloud.grid").stat().st_size == 20 + 4 * fm.n_cells
    rc, size, cached, devv = O.load_grid(work / "cloud.grid", 0.05)
    assert rc == 0 and size == fm.size and np.array_equal(bits(cached), bits(d["open_grid"].reshape(-1)))
    # MonteCarloLocalization steps (amcl3d.cpp) through the C++ class
    kp = d["mcl_keyposes"]
    assert int(d["mcl_sample_num"][0, 0]) == O.compute_sample_num(W, 3, 150, 400)
    Z = W.copy()
    O.update_sample_height(Z, kp)
    assert np.array_equal(bits(d["mcl_after_height"]), bits(Z))
    assert np.array_equal(bits(d["mcl_after_pfresample"]), bits(O.pf_resample_step(m, W.copy(), 1, 150, 400, kp)))
    # update -> PFResample -> update, on one device (against the oracle) and sharded over two ranks (same bits, no download)
    G = O.pf_resample_step(m, O.update(m, O.particles(P), cloud), 1, 150, 400, kp)
    G = O.update(m, O.particles(G), cloud)
    assert np.array_equal(bits(d["mcl_sd_after_frame"]), bits(G))
    assert np.array_equal(bits(d["mcl_md_after_frame"]), bits(G))
    assert d["mcl_sd_frame_downloads"][0, 0] == 0 and d["mcl_md_frame_downloads"][0, 0] == 0
    # RelocPose: z of the nearest keypose replaces the given z (amcl3d.cpp:169-174), then ParticleFilter::relocPose
    q = np.array([P[0, 0], P[0, 1], 0.0], np.float32)
    nearest = kp[np.argmin(((kp - q) ** 2).astype(np.float32).sum(1))]
    init = np.array([P[0, 0], P[0, 1], nearest[2], 0.0, 0.0, 0.3], np.float32)
    R2, _ = O.reloc_pose(300, init, np.array([0.5, 0.5, 0.4, 0.2, 0.2, 0.4], np.float32), d["mcl_reloc_noise"].reshape(-1))
    assert np.array_equal(bits(d["mcl_after_reloc"]), bits(R2))
    Q2 = O.predict(R2.copy(), np.array([0.31, -0.12, 0.02, 0.001, -0.002, 0.05], np.float32).astype(np.float64), d["mcl_move_noise"].reshape(-1))
    assert np.array_equal(bits(d["mcl_after_move"]), bits(Q2))
    # input downsample through the C++ class (to_cloud() of the driver sets intensity = i % 5)
    raw = np.concatenate([cloud, (np.arange(cloud.shape[0]) % 5).astype(np.float32)[:, None]], 1)
    ds = O.voxel_filter(raw, 0.3, ioff=3)
    assert np.array_equal(bits(d["mcl_cloud_ds"]), bits(ds))
    assert np.array_equal(bits(d["mcl_after_update_ds"][:, 6]), bits(O.update(m, O.particles(P), ds[:, :3].copy())[:, 6]))
    # lazily synced host mirror: update -> particleNormalize -> resample -> getMean costs ONE upload and no
    # download until the host looks at p_; a host edit makes the next device step upload again
    assert list(d["lazy_before_read"][0]) == [1, 0]
    assert list(d["lazy_after_read"][0]) == [1, 1]
    assert list(d["lazy_after_edit"][0]) == [2, 1]
    assert np.array_equal(bits(d["lazy_after_frame"]), bits(d["after_resample"]))
    # the same class sharded (amcl3d::ParticleFilter::setDevices): two and three ranks on device 0 -- the exchange protocol
    # of csrc/shard.cu with regions of one device as "peers", so this runs on a one-GPU box -- and two devices when the
    # box has them: bit-identical to one device.  frame2 changes the particle count between two sharded steps; the set
    # moves to the first device and back device to device (one gather, one scatter, no host copy).
    assert list(d["sd_frame2_moves"][0]) == [1, 0, 0, 0]
    F, _, _ = O.uniform_sample(m, O.scale_by_max(W.copy()), S)
    F = O.update(m, O.particles(F), cloud)
    O.normalize(m, F)
    F, _ = O.resample(F, d["resample_u01"][0, 0])
    assert F.shape[0] == S != P.shape[0]
    assert np.array_equal(bits(d["sd_frame2"]), bits(F))
    for tag in ["r2", "r3"] + (["md"] if d["md_devices"][0, 0] >= 2 else []):
        assert np.array_equal(bits(d[tag + "_after_update"]), bits(d["after_update"])), tag
        assert np.array_equal(bits(d[tag + "_after_normalize"]), bits(d["after_normalize"])), tag
        assert np.array_equal(bits(d[tag + "_after_resample"]), bits(d["after_resample"])), tag
        assert np.array_equal(bits(d[tag + "_after_uniform_sample"]), bits(d["after_uniform_sample"])), tag
        assert np.array_equal(bits(d[tag + "_mean"][0]), bits(d["mean"][0])), tag  # the chain carried from rank to rank
        assert np.allclose(d[tag + "_mean_fast"][0, :6], d["mean"][0, :6], rtol=1e-5, atol=1e-6), tag
        assert d[tag + "_mean_fast"][0, 6] == d["mean"][0, 6], tag
        assert list(d[tag + "_frame2_moves"][0]) == [1, 0, 1, 1], tag
        assert np.array_equal(bits(d[tag + "_frame2"]), bits(d["sd_frame2"])), tag
    # free-function builders
    ref2 = O.compute_grid2(m, sm.points)
    assert np.allclose(d["free_compute_grid2"].reshape(-1), ref2, rtol=1e-6, atol=1e-30)
