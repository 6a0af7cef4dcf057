// Stand-in for <pcl_conversions/pcl_conversions.h> (test infrastructure, see oracle/stubs/README.md):
// provides the two PCL free functions Grid3d.cpp reaches through it: pcl::io::loadPCDFile and
// pcl::copyPointCloud.  The PCD reader handles ASCII and binary files with float32 x y z [intensity].
#pragma once
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <pcl/point_cloud.h>

namespace pcl
{
template <typename A, typename B>
void copyPointCloud(const PointCloud<A>& in, PointCloud<B>& out)
{
  out.points.resize(in.points.size());
  for (size_t i = 0; i < in.points.size(); ++i)
  {
    out.points[i].x = in.points[i].x;
    out.points[i].y = in.points[i].y;
    out.points[i].z = in.points[i].z;
    out.points[i].intensity = in.points[i].intensity;
  }
  out.width = (uint32_t)out.points.size();
}

namespace io
{
template <typename PointT>
int loadPCDFile(const std::string& path, PointCloud<PointT>& cloud)
{
  std::ifstream f(path, std::ios::binary);
  if (!f)
    return -1;
  std::vector<std::string> fields;
  std::vector<int> sizes;
  size_t npoints = 0;
  std::string data = "ascii", line;
  while (std::getline(f, line))
  {
    if (line.empty() || line[0] == '#') continue;
    std::istringstream ss(line);
    std::string key;
    ss >> key;
    if (key == "FIELDS") { std::string s; while (ss >> s) fields.push_back(s); }
    else if (key == "SIZE") { int s; while (ss >> s) sizes.push_back(s); }
    else if (key == "POINTS") ss >> npoints;
    else if (key == "DATA") { ss >> data; break; }
  }
  int fx = -1, fy = -1, fz = -1, fi = -1;
  for (size_t i = 0; i < fields.size(); ++i)
  {
    if (fields[i] == "x") fx = (int)i;
    if (fields[i] == "y") fy = (int)i;
    if (fields[i] == "z") fz = (int)i;
    if (fields[i] == "intensity") fi = (int)i;
  }
  if (fx < 0 || fy < 0 || fz < 0) return -1;
  cloud.points.clear();
  cloud.points.reserve(npoints);
  if (data == "ascii")
  {
    for (size_t n = 0; n < npoints && std::getline(f, line); ++n)
    {
      std::istringstream ss(line);
      std::vector<float> v(fields.size(), 0.f);
      for (size_t i = 0; i < fields.size(); ++i) ss >> v[i];
      PointT p;
      p.x = v[fx]; p.y = v[fy]; p.z = v[fz];
      if (fi >= 0) p.intensity = v[fi];
      cloud.points.push_back(p);
    }
  }
  else if (data == "binary")
  {
    size_t step = 0;
    std::vector<size_t> off(fields.size(), 0);
    for (size_t i = 0; i < fields.size(); ++i) { off[i] = step; step += (i < sizes.size() ? sizes[i] : 4); }
    std::vector<char> rec(step);
    for (size_t n = 0; n < npoints && f.read(rec.data(), (std::streamsize)step); ++n)
    {
      PointT p;
      std::memcpy(&p.x, rec.data() + off[fx], 4);
      std::memcpy(&p.y, rec.data() + off[fy], 4);
      std::memcpy(&p.z, rec.data() + off[fz], 4);
      if (fi >= 0) std::memcpy(&p.intensity, rec.data() + off[fi], 4);
      cloud.points.push_back(p);
    }
  }
  else
    return -1;
  cloud.width = (uint32_t)cloud.points.size();
  return 0;
}
}  // namespace io
}  // namespace pcl
