// Stand-in for <pcl/registration/icp.h> (test infrastructure, see oracle/stubs/README.md): amcl3d.h includes it
// but nothing of ICP is used.  Provides the two names amcl3d.cpp needs from PCL/Eigen beyond the containers:
// Eigen::Affine3f and pcl::getTranslationAndEulerAngles.
#pragma once
#include <cmath>
#include <pcl/point_cloud.h>

namespace Eigen
{
// row-major 4x4 homogeneous transform, only what the harness and getTranslationAndEulerAngles touch
struct Affine3f
{
  float m[4][4];
  Affine3f()
  {
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c)
        m[r][c] = (r == c) ? 1.f : 0.f;
  }
  float operator()(int r, int c) const { return m[r][c]; }
  float& operator()(int r, int c) { return m[r][c]; }
};
}  // namespace Eigen

namespace pcl
{
// pcl/common/eigen.hpp: x,y,z from the last column; roll = atan2(t(2,1), t(2,2)), pitch = asin(-t(2,0)),
// yaw = atan2(t(1,0), t(0,0))
inline void getTranslationAndEulerAngles(const Eigen::Affine3f& t, float& x, float& y, float& z, float& roll, float& pitch,
                                         float& yaw)
{
  x = t(0, 3);
  y = t(1, 3);
  z = t(2, 3);
  roll = atan2f(t(2, 1), t(2, 2));
  pitch = asinf(-t(2, 0));
  yaw = atan2f(t(1, 0), t(0, 0));
}
}  // namespace pcl
