// Stand-in for <pcl/kdtree/kdtree_flann.h> (test infrastructure, see oracle/stubs/README.md).
// Exact nearest-neighbour search by uniform buckets.  Distances follow FLANN 1.8.4
// L2_Simple<float>: result += diff*diff accumulated in f32 over x, y, z.
#pragma once
#include <algorithm>
#include <cfloat>
#include <pcl/point_cloud.h>

namespace pcl
{
template <typename PointT>
class KdTreeFLANN
{
public:
  typedef boost::shared_ptr<KdTreeFLANN<PointT>> Ptr;
  typedef typename PointCloud<PointT>::Ptr CloudPtr;

  void setInputCloud(const CloudPtr& cloud)
  {
    cloud_ = cloud;
    start_.clear();
    items_.clear();
    if (!cloud_ || cloud_->points.empty())
      return;
    float hi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    lo_[0] = lo_[1] = lo_[2] = FLT_MAX;
    for (const auto& p : cloud_->points)
    {
      const float v[3] = { p.x, p.y, p.z };
      for (int a = 0; a < 3; ++a)
      {
        lo_[a] = std::min(lo_[a], v[a]);
        hi[a] = std::max(hi[a], v[a]);
      }
    }
    // bucket edge: aim at ~4 points per bucket, never below 1e-3
    double vol = 1;
    for (int a = 0; a < 3; ++a)
      vol *= std::max(1e-3, (double)hi[a] - lo_[a]);
    cell_ = (float)std::max(1e-3, std::cbrt(vol * 4.0 / cloud_->points.size()));
    size_t cells = 1;
    for (int a = 0; a < 3; ++a)
    {
      dim_[a] = (int)std::floor((hi[a] - lo_[a]) / cell_) + 1;
      cells *= (size_t)dim_[a];
    }
    start_.assign(cells + 1, 0);
    for (const auto& p : cloud_->points)
      start_[bucket(p.x, p.y, p.z) + 1]++;
    for (size_t c = 0; c < cells; ++c)
      start_[c + 1] += start_[c];
    std::vector<uint32_t> fill(start_.begin(), start_.end() - 1);
    items_.resize(cloud_->points.size());
    for (size_t i = 0; i < cloud_->points.size(); ++i)
    {
      const auto& p = cloud_->points[i];
      items_[fill[bucket(p.x, p.y, p.z)]++] = (uint32_t)i;
    }
  }

  int nearestKSearch(const PointT& q, int k, std::vector<int>& k_indices, std::vector<float>& k_sqr_distances) const
  {
    if (!cloud_ || cloud_->points.empty() || k != 1)
      return 0;
    k_indices.resize(1);
    k_sqr_distances.resize(1);
    int c0[3] = { axis_cell(q.x, 0), axis_cell(q.y, 1), axis_cell(q.z, 2) };
    const bool inside = in_box(q.x, 0) && in_box(q.y, 1) && in_box(q.z, 2);
    float best = FLT_MAX;
    int best_i = -1;
    const int max_ring = dim_[0] + dim_[1] + dim_[2];
    for (int ring = 0; ring <= max_ring; ++ring)
    {
      if (ring >= 2 && inside && best < FLT_MAX)
      {
        const double lb = (double)(ring - 1) * cell_;
        if (lb * lb > (double)best * (1.0 + 1e-5))
          break;
      }
      for (int cz = c0[2] - ring; cz <= c0[2] + ring; ++cz)
      {
        if (cz < 0 || cz >= dim_[2]) continue;
        for (int cy = c0[1] - ring; cy <= c0[1] + ring; ++cy)
        {
          if (cy < 0 || cy >= dim_[1]) continue;
          const bool shell = (std::abs(cz - c0[2]) == ring) || (std::abs(cy - c0[1]) == ring);
          const int xstep = shell ? 1 : std::max(1, 2 * ring);
          for (int cx = c0[0] - ring; cx <= c0[0] + ring; cx += xstep)
          {
            if (cx < 0 || cx >= dim_[0]) continue;
            const size_t c = ((size_t)cz * dim_[1] + cy) * dim_[0] + cx;
            for (uint32_t t = start_[c]; t < start_[c + 1]; ++t)
            {
              const auto& p = cloud_->points[items_[t]];
              float result = 0.f, diff;
              diff = q.x - p.x; result += diff * diff;
              diff = q.y - p.y; result += diff * diff;
              diff = q.z - p.z; result += diff * diff;
              if (result < best) { best = result; best_i = (int)items_[t]; }
            }
          }
        }
      }
    }
    k_indices[0] = best_i;
    k_sqr_distances[0] = best;
    return best_i >= 0 ? 1 : 0;
  }

  int radiusSearch(const PointT& q, double radius, std::vector<int>& k_indices, std::vector<float>& k_sqr_distances,
                   unsigned int max_nn = 0) const
  {
    k_indices.clear();
    k_sqr_distances.clear();
    if (!cloud_) return 0;
    const float r2 = (float)(radius * radius);
    for (size_t i = 0; i < cloud_->points.size(); ++i)
    {
      const auto& p = cloud_->points[i];
      float result = 0.f, diff;
      diff = q.x - p.x; result += diff * diff;
      diff = q.y - p.y; result += diff * diff;
      diff = q.z - p.z; result += diff * diff;
      if (result <= r2) { k_indices.push_back((int)i); k_sqr_distances.push_back(result); }
      if (max_nn && k_indices.size() >= max_nn) break;
    }
    return (int)k_indices.size();
  }

private:
  int axis_cell(float v, int a) const
  {
    int c = (int)std::floor((v - lo_[a]) / cell_);
    return std::min(std::max(c, 0), dim_[a] - 1);
  }
  bool in_box(float v, int a) const
  {
    const float rel = (v - lo_[a]) / cell_;
    return rel >= 0 && rel < dim_[a];
  }
  size_t bucket(float x, float y, float z) const
  {
    return ((size_t)axis_cell(z, 2) * dim_[1] + axis_cell(y, 1)) * dim_[0] + axis_cell(x, 0);
  }
  CloudPtr cloud_;
  float lo_[3] = { 0, 0, 0 };
  float cell_ = 1.f;
  int dim_[3] = { 1, 1, 1 };
  std::vector<uint32_t> start_, items_;
};
}  // namespace pcl
