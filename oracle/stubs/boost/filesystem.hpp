// empty stand-in: PointCloudTools.cpp:18 includes <boost/filesystem.hpp> but uses nothing from it
