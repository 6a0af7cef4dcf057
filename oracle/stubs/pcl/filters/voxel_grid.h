// Stand-in for <pcl/filters/voxel_grid.h> (test infrastructure, see oracle/stubs/README.md).
// Centroid voxel filter; output ordered by voxel index (x fastest) like PCL's sorted leaf layout.
#pragma once
#include <algorithm>
#include <cfloat>
#include <pcl/point_cloud.h>

namespace pcl
{
template <typename PointT>
class VoxelGrid
{
public:
  void setLeafSize(float lx, float ly, float lz) { leaf_[0] = lx; leaf_[1] = ly; leaf_[2] = lz; }
  void setInputCloud(const typename PointCloud<PointT>::Ptr& c) { in_ = c; }
  void filter(PointCloud<PointT>& out)
  {
    out.points.clear();
    if (!in_ || in_->points.empty()) return;
    float lo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, hi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    for (const auto& p : in_->points)
    {
      const float v[3] = { p.x, p.y, p.z };
      for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], v[a]); hi[a] = std::max(hi[a], v[a]); }
    }
    long mn[3], dv[3];
    for (int a = 0; a < 3; ++a)
    {
      mn[a] = (long)std::floor(lo[a] / leaf_[a]);
      dv[a] = (long)std::floor(hi[a] / leaf_[a]) - mn[a] + 1;
    }
    std::vector<std::pair<long, uint32_t>> keyed(in_->points.size());
    for (size_t i = 0; i < in_->points.size(); ++i)
    {
      const auto& p = in_->points[i];
      const long ix = (long)std::floor(p.x / leaf_[0]) - mn[0];
      const long iy = (long)std::floor(p.y / leaf_[1]) - mn[1];
      const long iz = (long)std::floor(p.z / leaf_[2]) - mn[2];
      keyed[i] = { ix + iy * dv[0] + iz * dv[0] * dv[1], (uint32_t)i };
    }
    std::sort(keyed.begin(), keyed.end());
    size_t i = 0;
    while (i < keyed.size())
    {
      size_t j = i;
      double s[4] = { 0, 0, 0, 0 };
      while (j < keyed.size() && keyed[j].first == keyed[i].first)
      {
        const auto& p = in_->points[keyed[j].second];
        s[0] += p.x; s[1] += p.y; s[2] += p.z; s[3] += p.intensity;
        ++j;
      }
      PointT c;
      const double n = (double)(j - i);
      c.x = (float)(s[0] / n); c.y = (float)(s[1] / n); c.z = (float)(s[2] / n); c.intensity = (float)(s[3] / n);
      out.points.push_back(c);
      i = j;
    }
    out.width = (uint32_t)out.points.size();
  }

private:
  float leaf_[3] = { 1, 1, 1 };
  typename PointCloud<PointT>::Ptr in_;
};
}  // namespace pcl
