// Grid3d.h -- likelihood-field map of amcl3d, B200 edition.
//
// Source-compatible with amcl3d/src/Grid3d.h of the reference for the weighting path:
//   open / computeCloudWeight (7-arg) / isIntoMap / loadPCD / point2grid, public pc_info_ / grid_info_.
// The cells live in HBM behind an amcl3d_b200_grid_t handle; every query launches CUDA kernels through
// the C-ABI (include/amcl3d_b200.h).  There is no CPU fallback.
#pragma once

#include <string>

#include "amcl3d/PointCloudTools.h"
#include "amcl3d/compat/vs_common.h"

namespace amcl3d
{
class Grid3d
{
public:
  /*! Reference constructor shape (Grid3d.h:38). */
  explicit Grid3d(VSCOMMON::LoggerPtr& log, int device = 0);
  /*! Upstream (fada-catec) default construction, used by the reference's stale tests. */
  explicit Grid3d(int device = 0);
  /*! A view of a grid that lives behind an existing C-ABI handle (not owned, not destroyed): resident / benchmark callers. */
  explicit Grid3d(amcl3d_b200_grid_t borrowed);
  virtual ~Grid3d();
  Grid3d(const Grid3d&) = delete;
  Grid3d& operator=(const Grid3d&) = delete;

  /*! Loads <map_path>{keypose_b,corner_b,surf_b,outlier_b}.pcd, voxel-filters them at 0.3 m, takes the
   *  bounds, then loads <map_path>cloud.grid if it exists and matches sensor_dev, else builds the grid
   *  with computeGrid2 on the GPU and saves it (reference: Grid3d.cpp:139-198).  false on any failure. */
  bool open(const std::string& map_path, const double sensor_dev);

  /*! Same flow starting from an in-memory map cloud (what open() does after loadPCD).  builder:
   *  AMCL3D_B200_GRID_SPLAT (fork-live computeGrid2), _QUIRK (computeGrid / exact EDT) or _GAUSSIAN. */
  bool openCloud(pcl::PointCloud<PointType>::Ptr cloud, const double sensor_dev, const double resolution,
                 const std::string& grid_path = std::string(), int builder = AMCL3D_B200_GRID_SPLAT);

  /*! Particle weight of one pose (reference: Grid3d.cpp:234-301); 0 before a map is open. */
  float computeCloudWeight(const pcl::PointCloud<pcl::PointXYZI>::Ptr& cloud, const float tx, const float ty,
                           const float tz, const float roll, const float pitch, const float yaw) const;

  /*! min <= v < max on every axis; false before a map is open (reference: Grid3d.cpp:365-372). */
  bool isIntoMap(const float x, const float y, const float z) const;

  /*! Reads and merges the four PCD files of a map directory, voxel-filters at 0.3 m (Grid3d.cpp:73-137). */
  bool loadPCD(std::string file_path, pcl::PointCloud<pcl::PointXYZI>::Ptr& input);

  /*! Linear cell index of a point (reference: Grid3d.cpp:444-449). */
  uint32_t point2grid(const float x, const float y, const float z) const;

  /*! Copies the cells from HBM into grid_info_->grid (not done by default: multi-GB at benchmark sizes). */
  bool mirrorGridToHost();

  bool saveGrid(const std::string& grid_path);
  bool loadGrid(const std::string& grid_path, const double sensor_dev);

  /*! The C-ABI handle, for resident / multi-GPU callers. */
  amcl3d_b200_grid_t handle() const { return handle_; }
  /*! Bumped whenever the cells change (open / load / build): replicas on other devices are rebuilt when it moves. */
  uint64_t generation() const { return generation_; }

  PointCloudInfo::Ptr pc_info_;
  Grid3dInfo::Ptr grid_info_;
  pcl::PointCloud<PointType>::Ptr keyposes_3d;

private:
  void refreshInfo();
  amcl3d_b200_grid_t handle_{ nullptr };
  bool owns_{ true };
  uint64_t generation_{ 0 };
  VSCOMMON::LoggerPtr g_log;
};

}  // namespace amcl3d
