// io/source.h -- the projection source of /root/reference/src/source.h:34-57 (src/source.cpp:38-135): walks the
// sorted HIS files of a directory, numbers the frames, keeps every quality-th one and attaches the angle read
// from the optional angle file.
//
// Deliberate differences (SURVEY F9 and the undefined behaviour of src/source.cpp:96-131):
//   * the frame counter belongs to the source object; the reference keeps it in a `thread_local static`, so the
//     second task a device thread picks up continues counting where the first one stopped;
//   * running out of files while skipping invalid ones, or a file that contributes no frame after the quality
//     filter, ends the stream (load_next returns a projection with an empty buffer) instead of indexing an empty
//     vector / popping an empty queue;
//   * an angle file shorter than the scan leaves the remaining projections on idx*delta_phi (has_angle_for()).
#pragma once

#include <cstdint>
#include <queue>
#include <string>
#include <vector>

#include "../b200/backend.h"

namespace paris
{
    // one float per whitespace-separated token; a decimal comma is accepted when the first line contains a comma
    // (the reference switches to the de_DE locale in that case, src/source.cpp:57-62)
    auto read_angles(const std::string& path) -> std::vector<float>;

    class source
    {
        private:
            using output_type = b200::projection_host_type;

        public:
            source(const std::string& proj_dir, bool enable_angles = false, const std::string& angle_file = "",
                   std::uint16_t quality = 1) noexcept;

            auto load_next() -> output_type;
            auto drained() const noexcept -> bool;
            // false for projections the angle file does not cover (callers then use idx*delta_phi)
            auto has_angle_for(std::uint32_t idx) const noexcept -> bool;

        private:
            auto refill() -> void;

            std::vector<std::string> paths_;
            std::size_t next_path_ = 0;
            std::queue<output_type> queue_;
            bool drained_ = true;
            std::uint32_t counter_ = 0;

            bool enable_angles_ = false;
            std::vector<float> angles_;
            std::uint16_t quality_ = 1;
    };
}
