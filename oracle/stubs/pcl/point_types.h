// Stand-in for <pcl/point_types.h> (test infrastructure, see oracle/stubs/README.md).
#pragma once
// real PCL pulls these in transitively (Eigen, boost, <iostream>); the reference sources rely on that
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace boost
{
// the reference spells boost::shared_ptr (PointCloudTools.h:42,59); PCL 1.7 used boost pointers
using std::shared_ptr;
using std::make_shared;
}  // namespace boost

namespace pcl
{
// PCL's PointXYZI: 16-byte aligned xyz + padding, then intensity + padding = 32 bytes
struct alignas(16) PointXYZI
{
  float x = 0, y = 0, z = 0, pad0 = 1.f;
  float intensity = 0, pad1 = 0, pad2 = 0, pad3 = 0;
};
struct alignas(16) PointXYZ
{
  float x = 0, y = 0, z = 0, pad0 = 1.f;
};
}  // namespace pcl
