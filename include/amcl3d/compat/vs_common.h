// vs_common.h -- stand-in for the un-vendored vscommon logger type the reference constructors take
// (amcl3d/src/Grid3d.h:38, ParticleFilter.h:109).  Only the constructor shape is kept; messages go to
// stderr when AMCL3D_B200_VERBOSE is set in the environment.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>

namespace VSCOMMON
{
class Logger
{
public:
  explicit Logger(const std::string& name = "MAIN") : name_(name), verbose_(std::getenv("AMCL3D_B200_VERBOSE") != nullptr) {}
  void info(const std::string& msg) const
  {
    if (verbose_)
      std::fprintf(stderr, "[%s] %s\n", name_.c_str(), msg.c_str());
  }
  void warn(const std::string& msg) const { std::fprintf(stderr, "[%s] WARN %s\n", name_.c_str(), msg.c_str()); }

private:
  std::string name_;
  bool verbose_;
};
typedef std::shared_ptr<Logger> LoggerPtr;
}  // namespace VSCOMMON
