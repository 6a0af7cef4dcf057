// Stand-in for <pcl/filters/voxel_grid.h> (test infrastructure, see oracle/stubs/README.md).
// Follows the published algorithm of pcl::VoxelGrid<PointT>::applyFilter (PCL 1.7.2 / 1.8,
// filters/include/pcl/filters/impl/voxel_grid.hpp) for a PointXYZI cloud with the default settings
// (downsample_all_data_ = true, min_points_per_voxel_ = 0, no filter field): f32 bounds, f32 inverse leaf,
// idx = ijk . (1, div_x, div_x*div_y), std::sort on idx alone (so the order inside a cell is whatever the
// unstable sort leaves), f32 running sums in that order, f32 divide by the count.
#pragma once
#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <pcl/point_cloud.h>

namespace pcl
{
template <typename PointT>
class VoxelGrid
{
  struct cloud_point_index_idx
  {
    unsigned int idx;
    unsigned int cloud_point_index;
    cloud_point_index_idx(unsigned int idx_, unsigned int cloud_point_index_) : idx(idx_), cloud_point_index(cloud_point_index_) {}
    bool operator<(const cloud_point_index_idx& p) const { return (idx < p.idx); }
  };

public:
  void setLeafSize(float lx, float ly, float lz)
  {
    leaf_[0] = lx; leaf_[1] = ly; leaf_[2] = lz;
    for (int a = 0; a < 3; ++a) inv_[a] = 1.0f / leaf_[a];
  }
  void setInputCloud(const typename PointCloud<PointT>::Ptr& c) { in_ = c; }
  void filter(PointCloud<PointT>& out)
  {
    out.points.clear();
    out.width = 0;
    if (!in_ || in_->points.empty()) return;
    const bool dense = in_->is_dense;
    float lo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, hi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    for (const auto& p : in_->points)
    {
      if (!dense && (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z))) continue;
      const float v[3] = { p.x, p.y, p.z };
      for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], v[a]); hi[a] = std::max(hi[a], v[a]); }
    }
    const int64_t dx = static_cast<int64_t>((hi[0] - lo[0]) * inv_[0]) + 1;
    const int64_t dy = static_cast<int64_t>((hi[1] - lo[1]) * inv_[1]) + 1;
    const int64_t dz = static_cast<int64_t>((hi[2] - lo[2]) * inv_[2]) + 1;
    if ((dx * dy * dz) > static_cast<int64_t>(INT_MAX))
    {
      out = *in_;  // "Leaf size is too small for the input dataset"
      return;
    }
    int min_b[3], max_b[3], div_b[3], mul[3];
    for (int a = 0; a < 3; ++a)
    {
      min_b[a] = static_cast<int>(std::floor(lo[a] * inv_[a]));
      max_b[a] = static_cast<int>(std::floor(hi[a] * inv_[a]));
      div_b[a] = max_b[a] - min_b[a] + 1;
    }
    mul[0] = 1; mul[1] = div_b[0]; mul[2] = div_b[0] * div_b[1];
    std::vector<cloud_point_index_idx> index_vector;
    index_vector.reserve(in_->points.size());
    for (size_t i = 0; i < in_->points.size(); ++i)
    {
      const auto& p = in_->points[i];
      if (!dense && (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z))) continue;
      const int ijk0 = static_cast<int>(std::floor(p.x * inv_[0]) - static_cast<float>(min_b[0]));
      const int ijk1 = static_cast<int>(std::floor(p.y * inv_[1]) - static_cast<float>(min_b[1]));
      const int ijk2 = static_cast<int>(std::floor(p.z * inv_[2]) - static_cast<float>(min_b[2]));
      const int idx = ijk0 * mul[0] + ijk1 * mul[1] + ijk2 * mul[2];
      index_vector.push_back(cloud_point_index_idx(static_cast<unsigned int>(idx), static_cast<unsigned int>(i)));
    }
    std::sort(index_vector.begin(), index_vector.end(), std::less<cloud_point_index_idx>());
    size_t index = 0;
    while (index < index_vector.size())
    {
      size_t i = index + 1;
      while (i < index_vector.size() && index_vector[i].idx == index_vector[index].idx) ++i;
      float s[4] = { 0.f, 0.f, 0.f, 0.f };
      for (size_t li = index; li < i; ++li)
      {
        const auto& p = in_->points[index_vector[li].cloud_point_index];
        s[0] += p.x; s[1] += p.y; s[2] += p.z; s[3] += p.intensity;
      }
      const float n = static_cast<float>(i - index);
      PointT c;
      c.x = s[0] / n; c.y = s[1] / n; c.z = s[2] / n; c.intensity = s[3] / n;
      out.points.push_back(c);
      index = i;
    }
    out.width = (uint32_t)out.points.size();
    out.height = 1;
    out.is_dense = true;
  }

private:
  float leaf_[3] = { 1, 1, 1 }, inv_[3] = { 1, 1, 1 };
  typename PointCloud<PointT>::Ptr in_;
};
}  // namespace pcl
