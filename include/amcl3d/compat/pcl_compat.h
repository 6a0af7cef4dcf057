// pcl_compat.h -- the handful of PCL names the amcl3d weighting path touches, for builds without PCL.
//
// With PCL installed (and AMCL3D_B200_USE_PCL defined) the real headers are used and nothing here is
// compiled; otherwise these minimal stand-ins keep amcl3d::Grid3d / ParticleFilter source-compatible:
// pcl::PointXYZI (PCL's 32-byte layout: x,y,z,pad | intensity,pad,pad,pad) and pcl::PointCloud<T> with
// points / size() / begin() / end() / push_back / Ptr.
#pragma once

#if defined(AMCL3D_B200_USE_PCL)
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#else
#include <cstddef>
#include <cstdint>
#include <memory>
#include <vector>

namespace boost
{
using std::shared_ptr;  // the reference spells its smart pointers boost::shared_ptr
}

namespace pcl
{
struct alignas(16) PointXYZI
{
  float x = 0.f, y = 0.f, z = 0.f, data_pad = 1.f;
  float intensity = 0.f, pad1 = 0.f, pad2 = 0.f, pad3 = 0.f;
};
static_assert(sizeof(PointXYZI) == 32, "PCL layout");

struct alignas(16) PointXYZ
{
  float x = 0.f, y = 0.f, z = 0.f, data_pad = 1.f;
};

template <typename PointT>
class PointCloud
{
public:
  typedef std::shared_ptr<PointCloud<PointT>> Ptr;
  typedef std::shared_ptr<const PointCloud<PointT>> ConstPtr;
  typedef typename std::vector<PointT>::iterator iterator;
  typedef typename std::vector<PointT>::const_iterator const_iterator;

  std::vector<PointT> points;
  std::uint32_t width = 0, height = 1;
  bool is_dense = true;

  std::size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void clear() { points.clear(); }
  void push_back(const PointT& p) { points.push_back(p); }
  iterator begin() { return points.begin(); }
  iterator end() { return points.end(); }
  const_iterator begin() const { return points.begin(); }
  const_iterator end() const { return points.end(); }
  PointT& operator[](std::size_t i) { return points[i]; }
  const PointT& operator[](std::size_t i) const { return points[i]; }
  PointCloud& operator+=(const PointCloud& other)
  {
    points.insert(points.end(), other.points.begin(), other.points.end());
    width = static_cast<std::uint32_t>(points.size());
    return *this;
  }
};
}  // namespace pcl
#endif
