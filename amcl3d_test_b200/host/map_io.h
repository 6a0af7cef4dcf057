// map_io.h -- the data formats either side of the path: a minimal PCD reader / writer (ASCII and binary, float32
// x y z [intensity]) standing in for pcl::io::loadPCDFile, which the reference reaches through PCL (Grid3d.cpp:80-81).
// The 0.3 m voxel filter of the merged map cloud (Grid3d.cpp:127-130) runs on the device (csrc/voxelfilter.cu).
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
}  // namespace io
}  // namespace amcl3d
