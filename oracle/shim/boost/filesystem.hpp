/*
 * oracle/shim/boost/filesystem.hpp -- TEST INFRASTRUCTURE, not product code.
 *
 * Boost.Filesystem is not installed in this image.  The reference's src/filesystem.cpp uses only
 * path, exists, is_regular_file, is_directory, directory_iterator, canonical, create_directories and
 * filesystem_error, all of which std::filesystem (C++17) provides with the same names and semantics,
 * so the reference source compiles unmodified against this alias.
 */
#pragma once
#include <filesystem>
namespace boost { namespace filesystem { using namespace std::filesystem; } }
