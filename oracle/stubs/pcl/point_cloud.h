// Stand-in for <pcl/point_cloud.h> (test infrastructure, see oracle/stubs/README.md).
#pragma once
#include <pcl/point_types.h>

namespace pcl
{
template <typename PointT>
class PointCloud
{
public:
  typedef boost::shared_ptr<PointCloud<PointT>> Ptr;
  typedef boost::shared_ptr<const PointCloud<PointT>> ConstPtr;
  typedef typename std::vector<PointT>::iterator iterator;
  typedef typename std::vector<PointT>::const_iterator const_iterator;

  std::vector<PointT> points;
  uint32_t width = 0, height = 1;
  bool is_dense = true;

  size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void clear() { points.clear(); }
  void push_back(const PointT& p) { points.push_back(p); }
  iterator begin() { return points.begin(); }
  iterator end() { return points.end(); }
  const_iterator begin() const { return points.begin(); }
  const_iterator end() const { return points.end(); }
  PointT& operator[](size_t i) { return points[i]; }
  const PointT& operator[](size_t i) const { return points[i]; }
  PointCloud& operator+=(const PointCloud& o)
  {
    points.insert(points.end(), o.points.begin(), o.points.end());
    return *this;
  }
};
}  // namespace pcl
