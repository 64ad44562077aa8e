// io/filesystem.h -- directory helpers of /root/reference/src/filesystem.h:30-33 (src/filesystem.cpp:36-90),
// on std::filesystem instead of Boost.Filesystem.
#pragma once

#include <string>
#include <vector>

namespace paris
{
    // canonical paths of every entry of the directory, sorted; throws std::runtime_error if `path` is a regular
    // file, does not exist or cannot be read (same messages as the reference)
    auto read_directory(const std::string& path) -> std::vector<std::string>;
    // true if the directory exists or was created (parents included); throws if `path` exists and is no directory
    auto create_directory(const std::string& path) -> bool;
}
