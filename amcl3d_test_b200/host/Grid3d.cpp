// Grid3d.cpp / PointCloudTools -- host shims of the B200 likelihood-field map over the C-ABI.
#include "amcl3d/Grid3d.h"

#include <cmath>
#include <cstddef>
#include <fstream>
#include <sstream>
#include <vector>

#include "map_io.h"

namespace amcl3d
{
namespace
{
constexpr uint32_t kStride = sizeof(PointType) / sizeof(float);  // 8 floats per pcl::PointXYZI

const float* xyz_of(const pcl::PointCloud<PointType>& c)
{
  return c.points.empty() ? nullptr : &c.points[0].x;
}

bool file_exists(const std::string& path)
{
  std::ifstream f(path, std::ios::binary);
  return static_cast<bool>(f);
}

void throw_last(const char* what)
{
  throw std::runtime_error(std::string(what) + ": " + amcl3d_b200_last_error());
}

void fill_geometry(amcl3d_b200_grid_t h, PointCloudInfo& pc, Grid3dInfo* gi)
{
  amcl3d_b200_grid_info_t info;
  if (amcl3d_b200_grid_get_info(h, &info) != AMCL3D_B200_OK)
    throw_last("grid_get_info");
  pc.octo_min_x = info.min[0];
  pc.octo_min_y = info.min[1];
  pc.octo_min_z = info.min[2];
  pc.octo_max_x = info.max[0];
  pc.octo_max_y = info.max[1];
  pc.octo_max_z = info.max[2];
  pc.octo_resol = info.resol;
  if (gi)
  {
    gi->sensor_dev = info.sensor_dev;
    gi->size_x = info.size[0];
    gi->size_y = info.size[1];
    gi->size_z = info.size[2];
    gi->step_y = info.size[0];
    gi->step_z = info.size[0] * info.size[1];
  }
}

// shared body of the two free builders: a temporary device grid, result mirrored to the host
Grid3dInfo::Ptr build_free(PointCloudInfo::Ptr pc_info, double sensor_dev, int mode)
{
  if (!pc_info)
    throw std::runtime_error("PointCloudInfo is NULL");  // reference: PointCloudTools.cpp:113-114,178-179
  if (!pc_info->cloud)
    throw std::runtime_error("PointCloudInfo has no cloud");
  amcl3d_b200_grid_t h = nullptr;
  if (amcl3d_b200_grid_create(0, &h) != AMCL3D_B200_OK)
    throw_last("grid_create");
  Grid3dInfo::Ptr gi(new Grid3dInfo());
  try
  {
    const double mn[3] = { pc_info->octo_min_x, pc_info->octo_min_y, pc_info->octo_min_z };
    const double mx[3] = { pc_info->octo_max_x, pc_info->octo_max_y, pc_info->octo_max_z };
    if (amcl3d_b200_grid_set_map(h, mn, mx, pc_info->octo_resol) != AMCL3D_B200_OK)
      throw_last("grid_set_map");
    if (amcl3d_b200_grid_build(h, xyz_of(*pc_info->cloud), kStride, pc_info->cloud->size(), sensor_dev, mode, 2, 0) !=
        AMCL3D_B200_OK)
      throw_last("grid_build");
    PointCloudInfo scratch;
    fill_geometry(h, scratch, gi.get());
    amcl3d_b200_grid_info_t info;
    amcl3d_b200_grid_get_info(h, &info);
    gi->grid.resize(info.n_cells);
    static_assert(sizeof(Grid3dCell) == sizeof(float), "Grid3dCell is one float");
    if (amcl3d_b200_grid_download(h, reinterpret_cast<float*>(gi->grid.data()), gi->grid.size()) != AMCL3D_B200_OK)
      throw_last("grid_download");
  }
  catch (...)
  {
    amcl3d_b200_grid_destroy(h);
    throw;
  }
  amcl3d_b200_grid_destroy(h);
  return gi;
}
}  // namespace

// ---- PointCloudTools ----------------------------------------------------------------------------------
PointCloudInfo::Ptr computePointCloud(pcl::PointCloud<PointType>::Ptr inputCloud, double resolution)
{
  if (!inputCloud)
    throw std::runtime_error("OcTree is NULL");  // reference: PointCloudTools.cpp:86-87
  if (inputCloud->size() <= 1)
    throw std::runtime_error("OcTree is empty");  // reference: PointCloudTools.cpp:89-91
  amcl3d_b200_grid_t h = nullptr;
  if (amcl3d_b200_grid_create(0, &h) != AMCL3D_B200_OK)
    throw_last("grid_create");
  double mn[3], mx[3];
  const int rc = amcl3d_b200_grid_compute_bounds(h, xyz_of(*inputCloud), kStride, inputCloud->size(), resolution, mn, mx);
  amcl3d_b200_grid_destroy(h);
  if (rc != AMCL3D_B200_OK)
    throw_last("compute_bounds");
  PointCloudInfo::Ptr pc(new PointCloudInfo());
  pc->octo_min_x = mn[0];
  pc->octo_min_y = mn[1];
  pc->octo_min_z = mn[2];
  pc->octo_max_x = mx[0];
  pc->octo_max_y = mx[1];
  pc->octo_max_z = mx[2];
  pc->octo_resol = resolution;
  pc->cloud = inputCloud;
  return pc;
}

Grid3dInfo::Ptr computeGrid(PointCloudInfo::Ptr pc_info, const double sensor_dev)
{
  return build_free(pc_info, sensor_dev, AMCL3D_B200_GRID_QUIRK);
}

Grid3dInfo::Ptr computeGrid2(PointCloudInfo::Ptr pc_info, const double sensor_dev)
{
  return build_free(pc_info, sensor_dev, AMCL3D_B200_GRID_SPLAT);
}

// ---- Grid3d ---------------------------------------------------------------------------------------------
Grid3d::Grid3d(VSCOMMON::LoggerPtr& log, int device) : g_log(log)
{
  if (amcl3d_b200_grid_create(device, &handle_) != AMCL3D_B200_OK)
    throw_last("Grid3d: no usable CUDA device (there is no CPU fallback)");
}

Grid3d::Grid3d(int device) : g_log(new VSCOMMON::Logger("GRID3D"))
{
  if (amcl3d_b200_grid_create(device, &handle_) != AMCL3D_B200_OK)
    throw_last("Grid3d: no usable CUDA device (there is no CPU fallback)");
}

Grid3d::Grid3d(amcl3d_b200_grid_t borrowed) : handle_(borrowed), owns_(false), g_log(new VSCOMMON::Logger("GRID3D"))
{
  if (!handle_)
    throw std::runtime_error("Grid3d: NULL handle");
  refreshInfo();
}

Grid3d::~Grid3d()
{
  if (owns_)
    amcl3d_b200_grid_destroy(handle_);
}

void Grid3d::refreshInfo()
{
  ++generation_;
  if (!pc_info_)
    pc_info_.reset(new PointCloudInfo());
  amcl3d_b200_grid_info_t info;
  amcl3d_b200_grid_get_info(handle_, &info);
  if (info.has_grid && !grid_info_)
    grid_info_.reset(new Grid3dInfo());
  fill_geometry(handle_, *pc_info_, info.has_grid ? grid_info_.get() : nullptr);
}

bool Grid3d::loadPCD(std::string file_path, pcl::PointCloud<PointType>::Ptr& input)
{
  pcl::PointCloud<PointType> corner, surf, outlier, merged;
  keyposes_3d.reset(new pcl::PointCloud<PointType>());
  if (io::loadPCDFile(file_path + "keypose_b.pcd", *keyposes_3d) == -1 || io::loadPCDFile(file_path + "corner_b.pcd", corner) == -1 ||
      io::loadPCDFile(file_path + "surf_b.pcd", surf) == -1 || io::loadPCDFile(file_path + "outlier_b.pcd", outlier) == -1)
  {
    g_log->warn("couldn't load pcd file under " + file_path);
    return false;
  }
  // Grid3d.cpp:88-121: the three clouds are bucketed by keyframe id (the intensity channel) and concatenated per
  // keyframe as surf, outlier, corner -- the order the voxel filter then sees (it decides the f32 summation order inside
  // a leaf).  A point whose id is not a valid keyframe index is undefined behaviour in the reference; it is dropped here.
  const size_t nkey = keyposes_3d->points.size();
  std::vector<std::vector<uint32_t>> bucket[3];
  const pcl::PointCloud<PointType>* src[3] = { &surf, &outlier, &corner };
  for (int c = 0; c < 3; ++c)
  {
    bucket[c].resize(nkey);
    for (size_t i = 0; i < src[c]->points.size(); ++i)
    {
      const float id = src[c]->points[i].intensity;
      if (id >= 0.f && static_cast<size_t>(static_cast<int>(id)) < nkey)
        bucket[c][static_cast<size_t>(static_cast<int>(id))].push_back(static_cast<uint32_t>(i));
    }
  }
  merged.points.reserve(surf.size() + outlier.size() + corner.size());
  for (size_t k = 0; k < nkey; ++k)
    for (int c = 0; c < 3; ++c)
      for (uint32_t i : bucket[c][k])
        merged.points.push_back(src[c]->points[i]);
  if (!input)
    input.reset(new pcl::PointCloud<PointType>());
  // Grid3d.cpp:127-130: pcl::VoxelGrid at 0.3 m -- the device filter of the input downsample (csrc/voxelfilter.cu: f32 sums in
  // input order, non-finite points skipped like PCL does for a non-dense cloud)
  input->points.clear();
  if (!merged.points.empty())
  {
    amcl3d_b200_pf_t tmp = nullptr;
    if (amcl3d_b200_pf_create(handle_, &tmp) != AMCL3D_B200_OK)
      throw_last("loadPCD: pf_create");
    uint32_t kept = 0;
    const int ioff = static_cast<int>(offsetof(PointType, intensity) / sizeof(float));
    int rc = amcl3d_b200_pf_set_cloud_downsampled(tmp, &merged.points[0].x, kStride, ioff, static_cast<uint32_t>(merged.points.size()),
                                                  0.3f, /*skip_nonfinite=*/1, &kept);
    std::vector<float> xyzi(static_cast<size_t>(kept) * 4);
    if (rc == AMCL3D_B200_OK && kept)
      rc = amcl3d_b200_pf_get_cloud(tmp, xyzi.data(), kept, &kept);
    amcl3d_b200_pf_destroy(tmp);
    if (rc != AMCL3D_B200_OK)
      throw_last("loadPCD: voxel filter");
    input->points.resize(kept);
    for (uint32_t i = 0; i < kept; ++i)
    {
      PointType& q = input->points[i];
      q.x = xyzi[4 * i];
      q.y = xyzi[4 * i + 1];
      q.z = xyzi[4 * i + 2];
      q.intensity = xyzi[4 * i + 3];
    }
  }
  input->width = static_cast<uint32_t>(input->points.size());
  input->height = 1;
  return true;
}

bool Grid3d::open(const std::string& map_path, const double sensor_dev)
{
  pcl::PointCloud<PointType>::Ptr cloud(new pcl::PointCloud<PointType>());
  if (!loadPCD(map_path, cloud))
    return false;
  return openCloud(cloud, sensor_dev, 0.3 /* Grid3d.cpp:156 */, map_path + "cloud.grid", AMCL3D_B200_GRID_SPLAT);
}

bool Grid3d::openCloud(pcl::PointCloud<PointType>::Ptr cloud, const double sensor_dev, const double resolution,
                       const std::string& grid_path, int builder)
{
  try
  {
    if (!cloud)
      throw std::runtime_error("OcTree is NULL");
    if (cloud->size() <= 1)
      throw std::runtime_error("OcTree is empty");
    double mn[3], mx[3];
    if (amcl3d_b200_grid_compute_bounds(handle_, xyz_of(*cloud), kStride, cloud->size(), resolution, mn, mx) != AMCL3D_B200_OK)
      throw_last("compute_bounds");
    pc_info_.reset(new PointCloudInfo());
    pc_info_->cloud = cloud;
    refreshInfo();
  }
  catch (std::exception& e)
  {
    g_log->warn(std::string("open: ") + e.what());  // reference: Grid3d.cpp:169-173 returns false
    return false;
  }
  if (!grid_path.empty() && file_exists(grid_path) && loadGrid(grid_path, sensor_dev))
    return true;
  g_log->info("Computing 3D occupancy grid on the GPU");
  if (amcl3d_b200_grid_build(handle_, xyz_of(*cloud), kStride, cloud->size(), sensor_dev, builder, 2, 0) != AMCL3D_B200_OK)
  {
    g_log->warn(std::string("grid build failed: ") + amcl3d_b200_last_error());
    return false;
  }
  refreshInfo();
  if (!grid_path.empty())
    saveGrid(grid_path);
  return true;
}

float Grid3d::computeCloudWeight(const pcl::PointCloud<pcl::PointXYZI>::Ptr& cloud, const float tx, const float ty,
                                 const float tz, const float roll, const float pitch, const float yaw) const
{
  if (!grid_info_ || !pc_info_)
    return 0;  // reference: Grid3d.cpp:237-238
  float w = 0.f;
  const uint32_t n = cloud ? static_cast<uint32_t>(cloud->size()) : 0u;
  if (amcl3d_b200_grid_cloud_weight(handle_, n ? xyz_of(*cloud) : nullptr, kStride, n, tx, ty, tz, roll, pitch, yaw,
                                    AMCL3D_B200_TRIG_F64, &w) != AMCL3D_B200_OK)
    throw_last("computeCloudWeight");
  return w;
}

bool Grid3d::isIntoMap(const float x, const float y, const float z) const
{
  if (!pc_info_)
    return false;
  int inside = 0;
  amcl3d_b200_grid_is_into_map(handle_, x, y, z, &inside);
  return inside != 0;
}

uint32_t Grid3d::point2grid(const float x, const float y, const float z) const
{
  // reference: Grid3d.cpp:444-449
  return static_cast<uint32_t>((x - pc_info_->octo_min_x) / pc_info_->octo_resol) +
         static_cast<uint32_t>((y - pc_info_->octo_min_y) / pc_info_->octo_resol) * grid_info_->step_y +
         static_cast<uint32_t>((z - pc_info_->octo_min_z) / pc_info_->octo_resol) * grid_info_->step_z;
}

bool Grid3d::mirrorGridToHost()
{
  if (!grid_info_)
    return false;
  amcl3d_b200_grid_info_t info;
  amcl3d_b200_grid_get_info(handle_, &info);
  grid_info_->grid.resize(info.n_cells);
  return amcl3d_b200_grid_download(handle_, reinterpret_cast<float*>(grid_info_->grid.data()), grid_info_->grid.size()) ==
         AMCL3D_B200_OK;
}

bool Grid3d::saveGrid(const std::string& grid_path)
{
  if (!grid_info_)
    return false;  // reference: Grid3d.cpp:376-377
  if (amcl3d_b200_grid_save_file(handle_, grid_path.c_str()) != AMCL3D_B200_OK)
  {
    g_log->warn(std::string("saveGrid: ") + amcl3d_b200_last_error());
    return false;
  }
  return true;
}

bool Grid3d::loadGrid(const std::string& grid_path, const double sensor_dev)
{
  if (amcl3d_b200_grid_load_file(handle_, grid_path.c_str(), sensor_dev) != AMCL3D_B200_OK)
  {
    g_log->warn(std::string("loadGrid: ") + amcl3d_b200_last_error());
    return false;
  }
  refreshInfo();
  return true;
}

}  // namespace amcl3d
