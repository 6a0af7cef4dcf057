// Stand-in for <pcl/common/common.h> (test infrastructure, see oracle/stubs/README.md).
#pragma once
#include <cfloat>
#include <pcl/point_cloud.h>

namespace pcl
{
// per-axis f32 min/max over the finite points (PCL skips non-finite points of non-dense clouds)
template <typename PointT>
void getMinMax3D(const PointCloud<PointT>& cloud, PointT& min_pt, PointT& max_pt)
{
  float lo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, hi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
  for (const auto& p : cloud.points)
  {
    if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z))
      continue;
    const float v[3] = { p.x, p.y, p.z };
    for (int a = 0; a < 3; ++a)
    {
      if (v[a] < lo[a]) lo[a] = v[a];
      if (v[a] > hi[a]) hi[a] = v[a];
    }
  }
  min_pt.x = lo[0]; min_pt.y = lo[1]; min_pt.z = lo[2];
  max_pt.x = hi[0]; max_pt.y = hi[1]; max_pt.z = hi[2];
}
}  // namespace pcl
