// map_io.h -- the data formats either side of the path: a minimal PCD reader (ASCII and binary, float32
// x y z [intensity]) and the centroid voxel filter applied to the merged map cloud, standing in for
// pcl::io::loadPCDFile and pcl::VoxelGrid which the reference reaches through PCL (Grid3d.cpp:80-81,127-130).
#pragma once

#include <string>

#include "amcl3d/PointCloudTools.h"

namespace amcl3d
{
namespace io
{
/*! 0 on success, -1 if the file cannot be opened or has no x/y/z float fields. */
int loadPCDFile(const std::string& path, pcl::PointCloud<PointType>& cloud);
/*! Writes a binary PCD with fields x y z intensity (used by tests / tools). */
int savePCDFile(const std::string& path, const pcl::PointCloud<PointType>& cloud);
/*! One centroid per occupied leaf, leaves ordered by index (x fastest), like pcl::VoxelGrid. */
void voxelGridFilter(const pcl::PointCloud<PointType>& in, float leaf, pcl::PointCloud<PointType>& out);
}  // namespace io
}  // namespace amcl3d
