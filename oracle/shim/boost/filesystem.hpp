/* Test-infrastructure shim (oracle build only): boost is not in this image. The reference's
 * src/core/filesystem.cpp:31-67 uses seven boost::filesystem calls; forward them to std::filesystem. */
#ifndef SHKZ_ORACLE_SHIM_BOOST_FILESYSTEM_HPP
#define SHKZ_ORACLE_SHIM_BOOST_FILESYSTEM_HPP
#include <cstdint>
#include <filesystem>
namespace boost {
namespace filesystem {
using std::filesystem::path;
using std::filesystem::exists;
using std::filesystem::create_directory;
using std::filesystem::create_directories;
using std::filesystem::remove;
using std::filesystem::remove_all;
inline void rename(const path &from, const path &to) { std::filesystem::rename(from, to); }
}
}
#endif
