// PointCloudTools.h -- map metadata + grid builders of the amcl3d weighting path, B200 edition.
//
// Source-compatible with amcl3d/src/PointCloudTools.h of the reference (same type and member names):
//   Grid3dCell, Grid3dInfo, PointCloudInfo, computePointCloud, computeGrid, computeGrid2.
// The arithmetic runs on the GPU behind the C-ABI of include/amcl3d_b200.h; these types are the host
// mirrors the reference's callers read.
#pragma once

#include <cstdint>
#include <stdexcept>
#include <vector>

#include "amcl3d/compat/pcl_compat.h"
#include "amcl3d_b200.h"

namespace amcl3d
{
using PointType = pcl::PointXYZI;

/*! One cell of the likelihood field (reference: PointCloudTools.h:30-35). */
class Grid3dCell
{
public:
  float prob{ 0 };
};

/*! Grid geometry + (optionally mirrored) cells (reference: PointCloudTools.h:39-52).
 *  `grid` is filled on demand only (Grid3d::mirrorGridToHost): at the benchmark sizes it is 6-26 GB. */
class Grid3dInfo
{
public:
  typedef boost::shared_ptr<Grid3dInfo> Ptr;
  typedef boost::shared_ptr<const Grid3dInfo> ConstPtr;

  std::vector<Grid3dCell> grid;
  double sensor_dev{ 0 };
  uint32_t size_x{ 0 };
  uint32_t size_y{ 0 };
  uint32_t size_z{ 0 };
  uint32_t step_y{ 0 };
  uint32_t step_z{ 0 };
};

/*! Bounds + resolution + the map cloud (reference: PointCloudTools.h:56-70). */
class PointCloudInfo
{
public:
  typedef boost::shared_ptr<PointCloudInfo> Ptr;
  typedef boost::shared_ptr<const PointCloudInfo> ConstPtr;

  pcl::PointCloud<PointType>::Ptr cloud;
  double octo_min_x{ 0 };
  double octo_min_y{ 0 };
  double octo_min_z{ 0 };
  double octo_max_x{ 0 };
  double octo_max_y{ 0 };
  double octo_max_z{ 0 };
  double octo_resol{ 0 };
};

/*! Bounds of the map cloud (device min/max reduction).  Throws std::runtime_error for a null cloud or
 *  one with <= 1 point, like the reference (PointCloudTools.cpp:86-91). */
PointCloudInfo::Ptr computePointCloud(pcl::PointCloud<PointType>::Ptr inputCloud, double resolution);

/*! Nearest-occupied-point Gaussian field (reference: kd-tree query per voxel, PointCloudTools.cpp:111-174).
 *  Here: exact Euclidean distance transform on the half-voxel lattice + fused Gaussian writer on the GPU.
 *  The returned info carries geometry; cells are mirrored to the host because this free function has no
 *  handle to keep them on (use Grid3d for the resident path). */
Grid3dInfo::Ptr computeGrid(PointCloudInfo::Ptr pc_info, const double sensor_dev);

/*! 27-neighbour max splat, the builder the fork's open() runs (PointCloudTools.cpp:176-254). */
Grid3dInfo::Ptr computeGrid2(PointCloudInfo::Ptr pc_info, const double sensor_dev);

}  // namespace amcl3d
