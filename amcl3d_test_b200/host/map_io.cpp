#include "map_io.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <limits>
#include <sstream>
#include <vector>

namespace amcl3d
{
namespace io
{
namespace
{
struct Field
{
  std::string name;
  int size = 4;
  char type = 'F';
  int count = 1;
  size_t offset = 0;
};
}  // namespace

int loadPCDFile(const std::string& path, pcl::PointCloud<PointType>& cloud)
{
  std::ifstream f(path, std::ios::binary);
  if (!f)
    return -1;
  std::vector<Field> fields;
  size_t npoints = 0, width = 0, height = 1;
  std::string data_kind, line;
  while (std::getline(f, line))
  {
    if (!line.empty() && line.back() == '\r')
      line.pop_back();
    if (line.empty() || line[0] == '#')
      continue;
    std::istringstream ss(line);
    std::string key;
    ss >> key;
    if (key == "FIELDS")
    {
      std::string s;
      while (ss >> s)
      {
        Field fd;
        fd.name = s;
        fields.push_back(fd);
      }
    }
    else if (key == "SIZE")
      for (auto& fd : fields)
        ss >> fd.size;
    else if (key == "TYPE")
      for (auto& fd : fields)
        ss >> fd.type;
    else if (key == "COUNT")
      for (auto& fd : fields)
        ss >> fd.count;
    else if (key == "WIDTH")
      ss >> width;
    else if (key == "HEIGHT")
      ss >> height;
    else if (key == "POINTS")
      ss >> npoints;
    else if (key == "DATA")
    {
      ss >> data_kind;
      break;
    }
  }
  if (npoints == 0)
    npoints = width * height;
  // a malformed header (SIZE / COUNT of 0 or below, absurd sizes) is rejected like a missing file
  for (const auto& fd : fields)
    if (fd.size <= 0 || fd.size > 8 || fd.count <= 0 || fd.count > (1 << 20))
      return -1;
  if (npoints > (size_t(1) << 33))
    return -1;
  size_t step = 0;
  int fx = -1, fy = -1, fz = -1, fi = -1;
  for (size_t i = 0; i < fields.size(); ++i)
  {
    fields[i].offset = step;
    step += static_cast<size_t>(fields[i].size) * fields[i].count;
    const bool f32 = fields[i].type == 'F' && fields[i].size == 4;
    if (fields[i].name == "x" && f32)
      fx = static_cast<int>(i);
    if (fields[i].name == "y" && f32)
      fy = static_cast<int>(i);
    if (fields[i].name == "z" && f32)
      fz = static_cast<int>(i);
    if (fields[i].name == "intensity" && f32)
      fi = static_cast<int>(i);
  }
  if (fx < 0 || fy < 0 || fz < 0)
    return -1;
  cloud.points.clear();
  cloud.points.reserve(npoints);
  if (data_kind == "ascii")
  {
    while (cloud.points.size() < npoints && std::getline(f, line))
    {
      std::istringstream ss(line);
      std::vector<double> v;
      double t;
      while (ss >> t)
        v.push_back(t);
      size_t col = 0;
      std::vector<size_t> first(fields.size());
      for (size_t i = 0; i < fields.size(); ++i)
      {
        first[i] = col;
        col += fields[i].count;
      }
      if (v.size() < col)
        continue;
      PointType p;
      p.x = static_cast<float>(v[first[fx]]);
      p.y = static_cast<float>(v[first[fy]]);
      p.z = static_cast<float>(v[first[fz]]);
      if (fi >= 0)
        p.intensity = static_cast<float>(v[first[fi]]);
      cloud.points.push_back(p);
    }
  }
  else if (data_kind == "binary")
  {
    std::vector<char> rec(step);
    for (size_t n = 0; n < npoints && f.read(rec.data(), static_cast<std::streamsize>(step)); ++n)
    {
      PointType p;
      std::memcpy(&p.x, rec.data() + fields[fx].offset, 4);
      std::memcpy(&p.y, rec.data() + fields[fy].offset, 4);
      std::memcpy(&p.z, rec.data() + fields[fz].offset, 4);
      if (fi >= 0)
        std::memcpy(&p.intensity, rec.data() + fields[fi].offset, 4);
      cloud.points.push_back(p);
    }
  }
  else
    return -1;  // binary_compressed is not supported
  cloud.width = static_cast<uint32_t>(cloud.points.size());
  cloud.height = 1;
  return 0;
}

int savePCDFile(const std::string& path, const pcl::PointCloud<PointType>& cloud)
{
  std::ofstream f(path, std::ios::binary);
  if (!f)
    return -1;
  f << "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\n"
       "COUNT 1 1 1 1\nWIDTH "
    << cloud.size() << "\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS " << cloud.size() << "\nDATA binary\n";
  for (const auto& p : cloud.points)
  {
    const float rec[4] = { p.x, p.y, p.z, p.intensity };
    f.write(reinterpret_cast<const char*>(rec), sizeof(rec));
  }
  return f ? 0 : -1;
}

}  // namespace io
}  // namespace amcl3d
