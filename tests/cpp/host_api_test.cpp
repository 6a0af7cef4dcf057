// host_api_test.cpp -- drives the reference-compatible C++ classes (include/amcl3d/*.h) end to end and dumps
// every result for tests/test_host_cpp.py to compare against the CPU oracle.  Usage: host_api_test in.bin out.bin tmpdir
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "amcl3d/amcl3d.h"
#include "../../amcl3d_test_b200/host/map_io.h"

using namespace amcl3d;

static std::ofstream g_out;

static void dump(const std::string& name, const float* data, uint64_t rows, uint64_t cols)
{
  const uint32_t len = static_cast<uint32_t>(name.size());
  g_out.write(reinterpret_cast<const char*>(&len), 4);
  g_out.write(name.data(), len);
  g_out.write(reinterpret_cast<const char*>(&rows), 8);
  g_out.write(reinterpret_cast<const char*>(&cols), 8);
  g_out.write(reinterpret_cast<const char*>(data), static_cast<std::streamsize>(rows * cols * 4));
}
static void dump_particles(const std::string& name, const std::vector<Particle>& p)
{
  dump(name, p.empty() ? nullptr : &p[0].x, p.size(), 7);
}
static void dump_scalar(const std::string& name, float v)
{
  dump(name, &v, 1, 1);
}

static pcl::PointCloud<PointType>::Ptr to_cloud(const std::vector<float>& xyz)
{
  pcl::PointCloud<PointType>::Ptr c(new pcl::PointCloud<PointType>());
  c->points.resize(xyz.size() / 3);
  for (size_t i = 0; i < c->points.size(); ++i)
  {
    c->points[i].x = xyz[3 * i];
    c->points[i].y = xyz[3 * i + 1];
    c->points[i].z = xyz[3 * i + 2];
    c->points[i].intensity = static_cast<float>(i % 5);  // keyframe id, like the LeGO-LOAM maps the fork loads
  }
  return c;
}
static void dump_cloud(const std::string& name, const pcl::PointCloud<PointType>& c)
{
  std::vector<float> xyz(c.size() * 3);
  for (size_t i = 0; i < c.size(); ++i)
  {
    xyz[3 * i] = c.points[i].x;
    xyz[3 * i + 1] = c.points[i].y;
    xyz[3 * i + 2] = c.points[i].z;
  }
  dump(name, xyz.data(), c.size(), 3);
}

int main(int argc, char** argv)
{
  if (argc < 4)
    return 2;
  std::ifstream in(argv[1], std::ios::binary);
  uint32_t n_map, n_cloud, n_part, seed;
  int32_t sample_num;
  double resol, sensor_dev;
  in.read(reinterpret_cast<char*>(&n_map), 4);
  in.read(reinterpret_cast<char*>(&n_cloud), 4);
  in.read(reinterpret_cast<char*>(&n_part), 4);
  in.read(reinterpret_cast<char*>(&seed), 4);
  in.read(reinterpret_cast<char*>(&sample_num), 4);
  in.read(reinterpret_cast<char*>(&resol), 8);
  in.read(reinterpret_cast<char*>(&sensor_dev), 8);
  std::vector<float> map_xyz(3ull * n_map), cloud_xyz(3ull * n_cloud), parts(7ull * n_part);
  in.read(reinterpret_cast<char*>(map_xyz.data()), static_cast<std::streamsize>(map_xyz.size() * 4));
  in.read(reinterpret_cast<char*>(cloud_xyz.data()), static_cast<std::streamsize>(cloud_xyz.size() * 4));
  in.read(reinterpret_cast<char*>(parts.data()), static_cast<std::streamsize>(parts.size() * 4));
  if (!in)
    return 3;
  g_out.open(argv[2], std::ios::binary);
  const std::string tmp = argv[3];

  auto map_cloud = to_cloud(map_xyz);
  auto cloud = to_cloud(cloud_xyz);
  std::vector<Particle> P(n_part);
  std::memcpy(P.data(), parts.data(), parts.size() * 4);

  VSCOMMON::LoggerPtr log(new VSCOMMON::Logger("TEST"));
  ParticleFilter pf(log);

  // Grid3dTest.cpp:160-161 / 260-265 behaviour before a map is opened
  dump_scalar("weight_before_open", pf.grid3d_->computeCloudWeight(cloud, 1, 1, 1, 0, 0, 0.1f));
  dump_scalar("into_before_open", pf.grid3d_->isIntoMap(1, 1, 1) ? 1.f : 0.f);

  // computePointCloud error behaviour (PointCloudTools.cpp:86-91)
  int threw = 0;
  try
  {
    computePointCloud(pcl::PointCloud<PointType>::Ptr(), resol);
  }
  catch (std::runtime_error&)
  {
    threw |= 1;
  }
  try
  {
    pcl::PointCloud<PointType>::Ptr one(new pcl::PointCloud<PointType>());
    one->push_back(PointType());
    computePointCloud(one, resol);
  }
  catch (std::runtime_error&)
  {
    threw |= 2;
  }
  dump_scalar("compute_point_cloud_throws", static_cast<float>(threw));

  // the path
  if (!pf.grid3d_->openCloud(map_cloud, sensor_dev, resol, "", AMCL3D_B200_GRID_QUIRK))
    return 4;
  const float info[7] = { (float)pf.grid3d_->pc_info_->octo_min_x, (float)pf.grid3d_->pc_info_->octo_max_x,
                          (float)pf.grid3d_->grid_info_->size_x, (float)pf.grid3d_->grid_info_->size_y,
                          (float)pf.grid3d_->grid_info_->size_z, (float)pf.grid3d_->grid_info_->step_y,
                          (float)pf.grid3d_->grid_info_->step_z };
  dump("info", info, 1, 7);
  pf.grid3d_->mirrorGridToHost();
  dump("grid", reinterpret_cast<const float*>(pf.grid3d_->grid_info_->grid.data()), pf.grid3d_->grid_info_->grid.size(), 1);
  dump_scalar("into_map_111", pf.grid3d_->isIntoMap(1, 1, 1) ? 1.f : 0.f);
  dump_scalar("into_map_far", pf.grid3d_->isIntoMap(-100, -100, -100) ? 1.f : 0.f);
  dump_scalar("single_pose_weight", pf.grid3d_->computeCloudWeight(cloud, P[0].x, P[0].y, P[0].z, P[0].roll, P[0].pitch, P[0].yaw));

  pf.p_ = P;
  pf.update(cloud);
  dump_particles("after_update", pf.p_);
  const std::vector<Particle> weighted = pf.p_;
  pf.addParticleWeight();
  dump_particles("after_add_particle_weight", pf.p_);
  pf.p_ = weighted;
  pf.particleNormalize();
  dump_particles("after_normalize", pf.p_);
  Particle m = pf.getMean();
  dump("mean", &m.x, 1, 7);
  Particle b = pf.getBestParticle();
  dump("best", &b.x, 1, 7);
  pf.setSeed(seed);
  pf.resample();
  dump_scalar("resample_u01", pf.lastResampleDraw());
  dump_particles("after_resample", pf.p_);
  pf.p_ = weighted;
  int S = sample_num;
  pf.uniformSample(S);
  dump_particles("after_uniform_sample", pf.p_);

  // init / predict with a replayable RNG
  {
    std::mt19937 peek(seed + 1);
    const float dev[6] = { 0.5f, 0.5f, 0.4f, 0.2f, 0.2f, 0.4f };
    std::vector<float> noise(6ull * (n_part - 1));
    for (size_t i = 0; i + 1 < n_part; ++i)
      for (int k = 0; k < 6; ++k)
      {
        std::normal_distribution<float> d(0, (double)dev[k]);
        noise[6 * i + k] = d(peek);
      }
    dump("reloc_noise", noise.data(), n_part - 1, 6);
    pf.setSeed(seed + 1);
    pf.init((int)n_part, P[0].x, P[0].y, P[0].z, 0.f, 0.f, 0.3f, dev[0], dev[1], dev[2], dev[3], dev[4], dev[5]);
    dump_particles("after_init", pf.p_);
    dump("init_mean", &pf.mean_.x, 1, 7);
    const double nd[6] = { 0.1, 0.1, 0.05, 0.05, 0.05, 0.05 };
    std::vector<float> pnoise(6ull * n_part);
    for (size_t i = 0; i < n_part; ++i)
      for (int k = 0; k < 6; ++k)
      {
        std::normal_distribution<float> d(0, nd[k]);
        pnoise[6 * i + k] = d(peek);
      }
    dump("predict_noise", pnoise.data(), n_part, 6);
    pf.predict(0.31, -0.12, 0.02, 0.001, -0.002, 0.05, nd[0], nd[1], nd[2], nd[3], nd[4], nd[5]);
    dump_particles("after_predict", pf.p_);
  }

  // Grid3d::open on a map directory: four PCDs -> voxel filter 0.3 -> computeGrid2 -> cloud.grid; then the cached path
  {
    pcl::PointCloud<PointType> key, corner, surf, outlier;
    for (size_t i = 0; i < map_cloud->size(); ++i)
      (i % 3 == 0 ? corner : (i % 3 == 1 ? surf : outlier)).push_back(map_cloud->points[i]);
    for (int k = 0; k < 5; ++k)
    {
      PointType kp;
      kp.x = 1.f + k;
      kp.y = 1.f;
      kp.z = 0.5f;
      kp.intensity = (float)k;
      key.push_back(kp);
    }
    const std::string dir = tmp + "/";
    io::savePCDFile(dir + "keypose_b.pcd", key);
    io::savePCDFile(dir + "corner_b.pcd", corner);
    io::savePCDFile(dir + "surf_b.pcd", surf);
    io::savePCDFile(dir + "outlier_b.pcd", outlier);
    Grid3d g(log);
    dump_scalar("open_missing_dir", g.open(tmp + "/nope/", sensor_dev) ? 1.f : 0.f);
    dump_scalar("open_ok", g.open(dir, sensor_dev) ? 1.f : 0.f);
    dump_cloud("open_filtered_cloud", *g.pc_info_->cloud);
    g.mirrorGridToHost();
    dump("open_grid", reinterpret_cast<const float*>(g.grid_info_->grid.data()), g.grid_info_->grid.size(), 1);
    const float sz[3] = { (float)g.grid_info_->size_x, (float)g.grid_info_->size_y, (float)g.grid_info_->size_z };
    dump("open_size", sz, 1, 3);
    Grid3d g2(log);
    dump_scalar("open_cached_ok", g2.open(dir, sensor_dev) ? 1.f : 0.f);  // loads cloud.grid this time
    g2.mirrorGridToHost();
    dump("open_cached_grid", reinterpret_cast<const float*>(g2.grid_info_->grid.data()), g2.grid_info_->grid.size(), 1);
    // free builders
    PointCloudInfo::Ptr pc = computePointCloud(map_cloud, resol);
    Grid3dInfo::Ptr gi2 = computeGrid2(pc, sensor_dev);
    dump("free_compute_grid2", reinterpret_cast<const float*>(gi2->grid.data()), gi2->grid.size(), 1);
  }
  // MonteCarloLocalization: the callers of the path (amcl3d.cpp:165-321)
  {
    MonteCarloLocalization mcl(log);
    if (!mcl.grid3d_->openCloud(map_cloud, sensor_dev, resol, "", AMCL3D_B200_GRID_QUIRK))
      return 5;
    mcl.grid3d_->keyposes_3d.reset(new pcl::PointCloud<PointType>());
    for (int k = 0; k < 12; ++k)
    {
      PointType kp;
      kp.x = 0.7f * k;
      kp.y = 6.5f - 0.5f * k;
      kp.z = 0.1f * (k % 4);
      mcl.grid3d_->keyposes_3d->push_back(kp);
    }
    dump_cloud("mcl_keyposes", *mcl.grid3d_->keyposes_3d);
    mcl.setParticleNum(400, 150);
    mcl.p_ = weighted;
    dump_scalar("mcl_sample_num", (float)mcl.computeSampleNum());
    mcl.p_ = weighted;
    mcl.updateSampleHeight();
    dump_particles("mcl_after_height", mcl.p_);
    mcl.p_ = weighted;
    mcl.PFResample(true);
    dump_particles("mcl_after_pfresample", mcl.p_);
    // the frame of the localisation loop (update, then PFResample: amcl3d.cpp:125,160,192-203) on one device and with the
    // particles sharded over two ranks: PFResample has no sharded form, the set visits the first device and comes back
    for (int pass = 0; pass < 2; ++pass)
    {
      if (pass == 1)
        mcl.setDevices({ 0, 0 });
      mcl.p_ = P;
      const uint64_t down0 = mcl.p_.downloads();
      mcl.update(cloud);
      mcl.PFResample(true);
      mcl.update(cloud);
      dump_scalar(pass ? "mcl_md_frame_downloads" : "mcl_sd_frame_downloads", (float)(mcl.p_.downloads() - down0));
      dump_particles(pass ? "mcl_md_after_frame" : "mcl_sd_after_frame", mcl.p_);
    }
    mcl.setDevices({ 0 });
    PointType h = mcl.getMapHeight(P[0].x, P[0].y);
    const float hv[3] = { h.x, h.y, h.z };
    dump("mcl_map_height", hv, 1, 3);
    // RelocPose (amcl3d.cpp:165-181: z from the nearest keypose, init deviations, global particle window) and PFMove
    {
      auto& lp = mcl.localization_params_;
      lp.init_x_dev_ = 0.5;
      lp.init_y_dev_ = 0.5;
      lp.init_z_dev_ = 0.4;
      lp.init_roll_dev_ = 0.2;
      lp.init_pitch_dev_ = 0.2;
      lp.init_yaw_dev_ = 0.4;
      lp.max_particle_num_global = 3000;
      lp.min_particle_num_global = 900;
      const int n_rel = 300;
      const float dev[6] = { 0.5f, 0.5f, 0.4f, 0.2f, 0.2f, 0.4f };
      std::mt19937 peek(seed + 7);
      std::vector<float> noise(6ull * (n_rel - 1));
      for (int i = 0; i + 1 < n_rel; ++i)
        for (int k = 0; k < 6; ++k)
        {
          std::normal_distribution<float> d(0, (double)dev[k]);
          noise[6 * i + k] = d(peek);
        }
      dump("mcl_reloc_noise", noise.data(), n_rel - 1, 6);
      mcl.setSeed(seed + 7);
      mcl.RelocPose(n_rel, P[0].x, P[0].y, /*ignored*/ 123.f, 0.f, 0.f, 0.3f);
      dump_particles("mcl_after_reloc", mcl.p_);
      if (mcl.globalLocalizationDone())
        return 7;
      const float nd[6] = { 0.1f, 0.1f, 0.05f, 0.05f, 0.05f, 0.05f };
      std::vector<float> pnoise(6ull * n_rel);
      for (int i = 0; i < n_rel; ++i)
        for (int k = 0; k < 6; ++k)
        {
          std::normal_distribution<float> d(0, (double)nd[k]);
          pnoise[6 * i + k] = d(peek);
        }
      dump("mcl_move_noise", pnoise.data(), n_rel, 6);
      mcl.PFMove(Particle(0.31f, -0.12f, 0.02f, 0.001f, -0.002f, 0.05f), Particle(nd[0], nd[1], nd[2], nd[3], nd[4], nd[5]));
      dump_particles("mcl_after_move", mcl.p_);
    }
    // the input downsample (amcl3d.cpp:37,51-52) and the update on the resident filtered cloud
    mcl.setVoxelSize(0.3);
    pcl::PointCloud<PointType> cloud_ds;
    const int kept = mcl.downsampleInput(cloud, &cloud_ds);
    if (kept <= 0 || kept >= (int)cloud->size() || (int)cloud_ds.size() != kept)
      return 6;
    {
      std::vector<float> xyzi(cloud_ds.size() * 4);
      for (size_t i = 0; i < cloud_ds.size(); ++i)
      {
        xyzi[4 * i] = cloud_ds.points[i].x;
        xyzi[4 * i + 1] = cloud_ds.points[i].y;
        xyzi[4 * i + 2] = cloud_ds.points[i].z;
        xyzi[4 * i + 3] = cloud_ds.points[i].intensity;
      }
      dump("mcl_cloud_ds", xyzi.data(), cloud_ds.size(), 4);
    }
    mcl.p_ = P;
    mcl.updateDownsampled();
    dump_particles("mcl_after_update_ds", mcl.p_);
  }
  // lazily synced host mirror: one frame of the reference's call sequence moves the set once each way
  {
    pf.p_ = P;
    const uint64_t up0 = pf.p_.uploads(), down0 = pf.p_.downloads();
    pf.update(cloud);
    pf.particleNormalize();
    pf.setSeed(seed);
    pf.resample();
    Particle mm = pf.getMean();
    (void)mm;
    const float before_read[2] = { (float)(pf.p_.uploads() - up0), (float)(pf.p_.downloads() - down0) };
    dump("lazy_before_read", before_read, 1, 2);
    dump_particles("lazy_after_frame", pf.p_);  // the host looks: one download
    const float after_read[2] = { (float)(pf.p_.uploads() - up0), (float)(pf.p_.downloads() - down0) };
    dump("lazy_after_read", after_read, 1, 2);
    pf.p_[0].w = pf.p_[0].w;  // a host edit: the next device step uploads again
    pf.getMean();
    const float after_edit[2] = { (float)(pf.p_.uploads() - up0), (float)(pf.p_.downloads() - down0) };
    dump("lazy_after_edit", after_edit, 1, 2);
  }
  // a frame that changes the particle count between two sharded steps (the PFResample pattern, amcl3d.cpp:192-203):
  // update, addParticleWeight, uniformSample, update, particleNormalize, resample.  Transfers counted before the host looks.
  auto frame2 = [&](const std::string& tag) {
    pf.p_ = P;
    const uint64_t up0 = pf.p_.uploads(), down0 = pf.p_.downloads(), g0 = pf.deviceGathers(), s0 = pf.deviceScatters();
    pf.update(cloud);
    pf.addParticleWeight();
    int S3 = sample_num;
    pf.uniformSample(S3);
    pf.update(cloud);
    pf.particleNormalize();
    pf.setSeed(seed);
    pf.resample();
    const float moves[4] = { (float)(pf.p_.uploads() - up0), (float)(pf.p_.downloads() - down0), (float)(pf.deviceGathers() - g0),
                             (float)(pf.deviceScatters() - s0) };
    dump((tag + "_frame2_moves").c_str(), moves, 1, 4);
    dump_particles((tag + "_frame2").c_str(), pf.p_);
  };
  frame2("sd");
  // particles sharded behind the same class: two and three ranks on device 0 (runs on a one-GPU box: the exchange protocol
  // is the same, the "peers" are regions of the same device), and two devices when the box has them
  {
    int ndev = 0;
    amcl3d_b200_device_info(0, &ndev, nullptr, nullptr, nullptr, nullptr, nullptr);
    dump_scalar("md_devices", (float)ndev);
    auto run_md = [&](const std::vector<int>& devs, const std::string& tag) {
      pf.setDevices(devs);
      pf.p_ = P;
      pf.update(cloud);
      dump_particles((tag + "_after_update").c_str(), pf.p_);
      pf.particleNormalize();
      dump_particles((tag + "_after_normalize").c_str(), pf.p_);
      Particle mm = pf.getMean();  // one f32 chain carried from rank to rank
      dump((tag + "_mean").c_str(), &mm.x, 1, 7);
      pf.setExactMean(false);
      mm = pf.getMean();  // per-rank partial sums combined in f64
      dump((tag + "_mean_fast").c_str(), &mm.x, 1, 7);
      pf.setExactMean(true);
      pf.setSeed(seed);
      pf.resample();
      dump_particles((tag + "_after_resample").c_str(), pf.p_);
      // a step that only exists on one device: the set goes to the first device and carries on there
      pf.p_ = weighted;
      int S2 = sample_num;
      pf.uniformSample(S2);
      dump_particles((tag + "_after_uniform_sample").c_str(), pf.p_);
      frame2(tag);
      pf.setDevices({ 0 });
    };
    run_md({ 0, 0 }, "r2");
    run_md({ 0, 0, 0 }, "r3");
    if (ndev >= 2)
      run_md({ 0, 1 }, "md");
  }
  g_out.close();
  std::printf("host_api_test ok\n");
  return 0;
}
