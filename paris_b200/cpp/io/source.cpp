// io/source.cpp -- see source.h.
#include "source.h"

#include <algorithm>
#include <cstdlib>
#include <exception>
#include <fstream>
#include <sstream>
#include <utility>

#include "filesystem.h"
#include "his.h"
#include "log.h"

namespace paris
{
    auto read_angles(const std::string& path) -> std::vector<float>
    {
        auto angles = std::vector<float>{};
        auto file = std::ifstream{path.c_str()};
        if(!file.is_open())
        {
            log::warning() << "Could not open angle file at " << path << ", using default values.";
            return angles;
        }
        auto text = std::stringstream{};
        text << file.rdbuf();
        auto content = text.str();
        const auto first_line = content.substr(0, content.find('\n'));
        if(first_line.find(',') != std::string::npos)
            std::replace(content.begin(), content.end(), ',', '.');   // decimal comma
        auto tokens = std::istringstream{content};
        auto token = std::string{};
        while(tokens >> token)
        {
            char* end = nullptr;
            const auto value = std::strtof(token.c_str(), &end);
            if(end == token.c_str())
                break;   // not a number: the reference's stream extraction stops here as well
            angles.push_back(value);
        }
        return angles;
    }

    auto make_scan_index(const std::string& proj_dir, bool enable_angles, const std::string& angle_file, std::uint16_t quality)
        -> scan_index
    {
        auto index = scan_index{};
        const auto q = quality == 0 ? std::uint16_t{1} : quality;
        auto angles = std::vector<float>{};
        if(enable_angles)
            angles = read_angles(angle_file);
        auto counter = 0u;
        for(const auto& path : read_directory(proj_dir))
        {
            auto info = his::file_info{};
            try
            {
                info = his::probe(path);
            }
            catch(const std::exception&)
            {
                info.valid = false;
            }
            if(!info.valid || info.frames == 0)
            {
                log::warning() << "Skipping invalid file at " << path;
                continue;
            }
            index.paths.push_back(path);
            index.infos.push_back(info);
            if(index.dim_x == 0)
            {
                index.dim_x = info.width;
                index.dim_y = info.height;
            }
            for(auto f = 0u; f < info.frames; ++f)
            {
                if(counter % q == 0u)
                {
                    const auto has = enable_angles && counter < angles.size();
                    index.frames.push_back(scan_frame{index.paths.size() - 1u, f, counter, has, has ? angles[counter] : 0.f});
                }
                ++counter;
            }
        }
        return index;
    }

    auto load_scan_frame(const scan_index& index, std::size_t i, float* dst) -> bool
    {
        if(i >= index.frames.size())
            return false;
        const auto& fr = index.frames[i];
        const auto& info = index.infos[fr.file];
        if(info.width != index.dim_x || info.height != index.dim_y)
            return false;   // (a scan is one detector: frames of another size cannot share the stack)
        return his::read_frame(index.paths[fr.file], info, fr.frame, dst);
    }

    auto load_scan_frame_u16(const scan_index& index, std::size_t i, std::uint16_t* dst) -> bool
    {
        if(i >= index.frames.size())
            return false;
        const auto& fr = index.frames[i];
        const auto& info = index.infos[fr.file];
        if(info.width != index.dim_x || info.height != index.dim_y)
            return false;
        return his::read_frame_u16(index.paths[fr.file], info, fr.frame, dst);
    }

    auto scan_is_u16(const scan_index& index) -> bool
    {
        if(index.frames.empty())
            return false;
        for(const auto& fr : index.frames)
            if(index.infos[fr.file].number_type != 4)
                return false;
        return true;
    }

    source::source(const std::string& proj_dir, bool enable_angles, const std::string& angle_file,
                   std::uint16_t quality) noexcept
    : enable_angles_{enable_angles}, quality_{quality == 0 ? std::uint16_t{1} : quality}
    {
        try
        {
            paths_ = read_directory(proj_dir);
            if(enable_angles_)
                angles_ = read_angles(angle_file);
            refill();
        }
        catch(const std::exception& e)
        {
            log::error() << "source: " << e.what();
            paths_.clear();
        }
        drained_ = queue_.empty();
    }

    // queue the kept frames of the next file(s) until at least one projection is waiting or the files run out
    auto source::refill() -> void
    {
        while(queue_.empty() && next_path_ < paths_.size())
        {
            const auto& path = paths_[next_path_++];
            auto frames = his::load(path);
            if(frames.empty())
            {
                log::warning() << "Skipping invalid file at " << path;
                continue;
            }
            for(auto& p : frames)
            {
                if(counter_ % quality_ == 0u)   // src/source.cpp:105
                {
                    p.idx = counter_;
                    if(has_angle_for(counter_))
                        p.phi = angles_[counter_];
                    queue_.push(std::move(p));
                }
                ++counter_;
            }
        }
    }

    auto source::load_next() -> output_type
    {
        if(queue_.empty())
            refill();
        if(queue_.empty())
        {
            drained_ = true;
            return output_type{};
        }
        auto p = std::move(queue_.front());
        queue_.pop();
        if(queue_.empty())
            refill();   // look ahead so that drained() is exact even when the remaining files hold nothing
        drained_ = queue_.empty();
        return p;
    }

    auto source::drained() const noexcept -> bool
    {
        return drained_;
    }

    auto source::has_angle_for(std::uint32_t idx) const noexcept -> bool
    {
        return enable_angles_ && idx < angles_.size();
    }
}
