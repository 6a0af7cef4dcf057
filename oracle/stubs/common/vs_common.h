// Stand-in for vscommon's <common/vs_common.h> (test infrastructure, see oracle/stubs/README.md).
// The library is not vendored in the reference tree; only the names the three compiled sources
// touch are provided.  Logging is swallowed (arguments are never evaluated).
#pragma once
#include <chrono>
#include <map>
#include <memory>
#include <string>
#include <sys/stat.h>

namespace VSCOMMON
{
struct Logger
{
};
typedef std::shared_ptr<Logger> LoggerPtr;

inline bool existFile(const std::string& path)
{
  struct stat st;
  return ::stat(path.c_str(), &st) == 0;
}

inline std::map<std::string, std::chrono::steady_clock::time_point>& tic_table()
{
  static std::map<std::string, std::chrono::steady_clock::time_point> t;
  return t;
}
inline void tic(const char* name) { tic_table()[name] = std::chrono::steady_clock::now(); }
inline double toc(const char* name)
{
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - tic_table()[name]).count();
}
template <typename T>
inline T clip(T v, T lo, T hi)
{
  return v < lo ? lo : (v > hi ? hi : v);
}
inline double rad2deg(double r)
{
  return r * 57.29577951308232;
}
}  // namespace VSCOMMON

#define LOG_INFO(...) do { } while (0)
#define LOG_COUT_INFO(...) do { } while (0)
#define LOG_COUT_WARN(...) do { } while (0)
#define LOG_COUT_ERROR(...) do { } while (0)
