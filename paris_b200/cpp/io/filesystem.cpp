// io/filesystem.cpp -- see filesystem.h.
#include "filesystem.h"

#include <algorithm>
#include <filesystem>
#include <stdexcept>

#include "log.h"

namespace paris
{
    namespace fs = std::filesystem;

    auto read_directory(const std::string& path) -> std::vector<std::string>
    {
        auto entries = std::vector<std::string>{};
        auto ec = std::error_code{};
        const auto status = fs::status(path, ec);
        if(ec && ec != std::errc::no_such_file_or_directory)
        {
            log::fatal() << path << " could not be read: " << ec.message();
            throw std::runtime_error{path + " could not be read"};
        }
        if(!fs::exists(status))
            throw std::runtime_error{path + " does not exist."};
        if(fs::is_regular_file(status))
            throw std::runtime_error{path + " is not a directory."};
        if(!fs::is_directory(status))
            throw std::runtime_error{path + " exists but is neither a regular file nor a directory."};
        try
        {
            for(const auto& e : fs::directory_iterator{path})
                entries.push_back(fs::canonical(e.path()).string());
        }
        catch(const fs::filesystem_error& err)
        {
            log::fatal() << path << " could not be read: " << err.what();
            throw std::runtime_error{path + " could not be read"};
        }
        std::sort(entries.begin(), entries.end());
        return entries;
    }

    auto create_directory(const std::string& path) -> bool
    {
        auto ec = std::error_code{};
        if(fs::exists(path, ec))
        {
            if(fs::is_directory(path, ec))
                return true;
            throw std::runtime_error{path + " exists but is not a directory."};
        }
        fs::create_directories(path, ec);
        if(ec)
        {
            log::fatal() << path << " could not be created: " << ec.message();
            return false;
        }
        return true;
    }
}
