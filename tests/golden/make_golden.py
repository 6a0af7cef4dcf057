#!/usr/bin/env python
"""Generates the committed golden vectors under tests/golden/ from the REFERENCE itself.

Run in the build container (needs /root/reference and oracle/_ref = the reference's own
Grid3d.cpp / ParticleFilter.cpp / PointCloudTools.cpp compiled where they lie, see oracle/Makefile):

    python tests/golden/make_golden.py

Outputs
  fixture_cloud.npy        the 25 809-point lidar cloud of amcl3d/tests/data/grid_info.bin (ROS1
                           geometry_msgs/PoseArray, positions only), float32 like Grid3dTest.cpp:150-157
  fixture_map_half.npz     the 179 551 map points of amcl3d/tests/data/mappointcloud_msg.bin (PointCloud2,
                           x,y,z f32) as uint16 indices on the 0.025 m half-voxel lattice + f32 residuals (< 1e-5 m)
  ref_pf_case.npz          seeded small world + every output of the compiled reference on it: update weights,
                           particleNormalize, resample (with the u01 the reference drew), uniformSample,
                           getMean, getBestParticle, relocPose, predict (with the noise the reference drew)
  ref_grid_case.npz        tiny map + computeGrid (kd-tree NN, quirk) and computeGrid2 (splat) outputs
"""
import struct
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
HERE = Path(__file__).resolve().parent
REF_DATA = Path("/root/reference/amcl3d/tests/data")

from amcl3d_test_b200 import synth  # noqa: E402
from oracle import loader as O  # noqa: E402


def decode_pose_array(path):
    b = path.read_bytes()
    off = 12  # Header: seq, stamp.sec, stamp.nsec
    (n,) = struct.unpack_from("<I", b, off)
    off += 4 + n  # frame_id
    (cnt,) = struct.unpack_from("<I", b, off)
    off += 4
    a = np.frombuffer(b, "<f8", cnt * 7, off).reshape(cnt, 7)
    return a[:, :3].astype(np.float32)


def decode_pointcloud2(path):
    b = path.read_bytes()
    off = 12
    (n,) = struct.unpack_from("<I", b, off)
    off += 4 + n
    height, width = struct.unpack_from("<II", b, off)
    off += 8
    (nf,) = struct.unpack_from("<I", b, off)
    off += 4
    fields = {}
    for _ in range(nf):
        (ln,) = struct.unpack_from("<I", b, off)
        off += 4
        name = b[off:off + ln].decode()
        off += ln
        foff, dtype, count = struct.unpack_from("<IBI", b, off)
        off += 9
        fields[name] = foff
    is_bigendian, point_step, row_step = struct.unpack_from("<BII", b, off)
    off += 9
    (dlen,) = struct.unpack_from("<I", b, off)
    off += 4
    raw = np.frombuffer(b, np.uint8, dlen, off).reshape(-1, point_step)
    xyz = np.stack([raw[:, fields[k]:fields[k] + 4].copy().view("<f4")[:, 0] for k in ("x", "y", "z")], 1)
    return xyz


def main():
    assert O.ref_available() or O.REF_SRC.exists(), "needs /root/reference"
    O.build(ref=True)
    assert O.ref().ref_trig_width() == 8, "expected the ::sin(double) binding"

    # ---- the reference's own fixtures ---------------------------------------------------------------
    cloud = decode_pose_array(REF_DATA / "grid_info.bin")
    assert cloud.shape == (25809, 3), cloud.shape
    np.save(HERE / "fixture_cloud.npy", cloud)
    mp = decode_pointcloud2(REF_DATA / "mappointcloud_msg.bin")
    assert mp.shape == (179551, 3), mp.shape  # PointCloudToolsTest.cpp:166
    half = np.rint(mp.astype(np.float64) / 0.025).astype(np.int64)
    assert np.abs(half * 0.025 - mp).max() < 1e-4 and half.min() >= 0 and half.max() < 65536
    # the f32 coordinates deviate from k*0.025 only by f32 rounding of the octomap arithmetic (< 1e-5 m): keep
    # the lattice indices plus the exact residual so the floats are reproduced bit for bit
    resid = mp - (half * 0.025).astype(np.float32)
    assert np.abs(resid).max() < 1e-5
    np.savez_compressed(HERE / "fixture_map_half.npz", half=half.astype(np.uint16), resid=resid.astype(np.float32))

    # ---- particle-filter case -------------------------------------------------------------------------
    sm = synth.make_map((8.0, 8.0, 4.0), 0.1)
    mn, mx = O.compute_bounds(sm.points)
    g2 = O.ref_compute_grid(sm.points, 0.1, 0.05, which=2)  # the fork-live builder, reference code
    m = O.Map(mn, mx, 0.1, 0.05, prob=g2["prob"])
    assert np.array_equal(O.compute_grid2(m, sm.points).view(np.uint32), g2["prob"].view(np.uint32))
    cloud_s = synth.make_cloud(sm, 600, 3.5)
    out = {}
    rg = O.RefGrid(m)
    for kind in ("T", "G"):
        P = synth.make_particles(sm, 300, kind)
        pf = O.RefPF(rg)
        pf.set(P)
        pf.update(cloud_s)
        W = pf.get()
        out[f"{kind}_update"] = W[:, 6].copy()
        pf.normalize()
        out[f"{kind}_normalized"] = pf.get()[:, 6].copy()
        pf.seed(4004 + (kind == "G"))
        u01 = pf.resample()
        out[f"{kind}_resample_u01"] = np.float32(u01)
        out[f"{kind}_resampled"] = pf.get()
        out[f"{kind}_mean_after_resample"] = pf.mean()
        for S in (200, 300, 450):
            pf.set(W)
            n = pf.uniform_sample(S)
            assert n == S
            out[f"{kind}_uniform_{S}"] = pf.get()
        pf.set(W)
        pf.normalize()
        out[f"{kind}_mean"] = pf.mean()
        out[f"{kind}_best"] = pf.best()
    # relocPose / predict with the noise the reference drew
    pf = O.RefPF(rg)
    init = np.array([4.0, 4.0, 1.2, 0.0, 0.0, 0.3], np.float32)
    dev = np.array([0.5, 0.5, 0.4, 0.2, 0.2, 0.4], np.float32)
    noise = pf.reloc_pose(77, 200, init, dev)
    out["reloc_init"], out["reloc_dev"], out["reloc_noise"] = init, dev, noise
    out["reloc_particles"] = pf.get()
    delta = np.array([0.31, -0.12, 0.02, 0.001, -0.002, 0.05])
    ndev = np.array([0.1, 0.1, 0.05, 0.05, 0.05, 0.05])
    pnoise = pf.predict(78, delta, ndev)
    out["predict_delta"], out["predict_noise"] = delta, pnoise
    out["predict_particles"] = pf.get()
    # single-pose computeCloudWeight incl. a pose outside the map and the "before open" zero
    poses = np.array([[4.0, 4.0, 1.2, 0, 0, 0.3], [3.3, 5.1, 1.0, 0.05, -0.1, 1.9], [-5.0, 3.0, 1.0, 0, 0, 0]], np.float32)
    out["poses"] = poses
    out["pose_weights"] = np.array([rg.cloud_weight(cloud_s, p) for p in poses], np.float32)
    out["weight_before_open"] = np.float32(O.RefGrid().cloud_weight(cloud_s, poses[0]))  # Grid3dTest.cpp:160-161
    np.savez_compressed(HERE / "ref_pf_case.npz", **out)

    # ---- grid builders ----------------------------------------------------------------------------------
    rng = np.random.RandomState(12)
    vox = np.stack([rng.randint(0, 24, 150), rng.randint(0, 20, 150), rng.randint(0, 12, 150)], 1)
    pts = np.concatenate([[[0, 0, 0], [2.4, 2.0, 1.2]], (vox + 0.5) * 0.1]).astype(np.float32)
    g0 = O.ref_compute_grid(pts, 0.1, 0.05, which=0)
    g2 = O.ref_compute_grid(pts, 0.1, 0.05, which=2)
    off = (rng.uniform(0, 1, (300, 3)) * [2.4, 2.0, 1.2]).astype(np.float32)  # off-lattice points for the splat
    g2b = O.ref_compute_grid(off, 0.3, 0.05, which=2)
    np.savez_compressed(HERE / "ref_grid_case.npz", points=pts, size=np.array(g0["size"]), min=g0["min"], max=g0["max"],
                        compute_grid=g0["prob"], compute_grid2=g2["prob"], off_points=off, off_size=np.array(g2b["size"]),
                        off_min=g2b["min"], off_max=g2b["max"], off_compute_grid2=g2b["prob"])
    # ---- callers of the path: amcl3d.cpp addParticleWeight / computeSampleNum / updateSampleHeight / PFResample
    out = {}
    kp = (np.random.RandomState(3).uniform(0, 1, (40, 3)) * [8, 8, 2]).astype(np.float32)
    out["keyposes"] = kp
    gt = synth.gt_pose(sm.extent, sm.resol)
    for kind, spread in (("T", 1.0), ("T", 0.2), ("T", 0.05), ("G", 1.0)):
        P = synth.make_particles(sm, 300, kind)
        P[:, :6] = gt + (P[:, :6] - gt) * np.float32(spread)
        pf = O.RefPF(rg)
        pf.set(P)
        pf.update(cloud_s)
        W = pf.get()
        tag = f"{kind}_{spread}"
        mcl = O.RefMCL(rg, kp)
        mcl.set(W)
        mcl.add_particle_weight()
        out[f"{tag}_scaled"] = mcl.get()[:, 6].copy()
        for cpg in (3, 9):
            mcl.set(W)
            mcl.set_particle_num(800, 150)
            out[f"{tag}_sample_num_{cpg}"] = np.int32(mcl.compute_sample_num(cpg))
        mcl.set(W)
        mcl.update_sample_height()
        out[f"{tag}_height"] = mcl.get()[:, 2].copy()
        for wm in (0, 1):
            mcl.set(W)
            mcl.set_particle_num(800, 150)
            mcl.pf_resample(wm)
            out[f"{tag}_pfresample_{wm}"] = mcl.get()
    np.savez_compressed(HERE / "ref_mcl_case.npz", **out)
    for f in sorted(HERE.glob("*.np*")):
        print(f.name, f.stat().st_size)


if __name__ == "__main__":
    main()
