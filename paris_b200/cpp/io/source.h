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
#include "his.h"

namespace paris
{
    // one float per whitespace-separated token; a decimal comma is accepted when the first line contains a comma
    // (the reference switches to the de_DE locale in that case, src/source.cpp:57-62)
    auto read_angles(const std::string& path) -> std::vector<float>;

    // The kept frames of a scan without decoding any: which file and frame holds projection i of the scan, its
    // frame counter idx (src/source.cpp:105: every quality-th frame is kept, idx is the counter of ALL frames) and
    // its angle.  What several readers of one scan -- the members of a reconstruction group -- agree on.
    struct scan_frame
    {
        std::size_t file;          // index into scan_index::paths
        std::uint32_t frame;       // frame inside that file
        std::uint32_t idx;         // frame counter of the scan (drives idx * delta_phi)
        bool has_angle;
        float phi;                 // from the angle file, if has_angle
    };

    struct scan_index
    {
        std::vector<std::string> paths;
        std::vector<his::file_info> infos;          // per file
        std::vector<scan_frame> frames;             // the projections of the scan, in order
        std::uint32_t dim_x = 0, dim_y = 0;          // of the first valid file
    };

    // walks the directory exactly like source does (sorted files, invalid ones skipped, frame counter, quality, angles)
    auto make_scan_index(const std::string& proj_dir, bool enable_angles, const std::string& angle_file, std::uint16_t quality)
        -> scan_index;
    // decodes projection i of the index into dst (dim_x * dim_y floats); false if the frame cannot be read
    auto load_scan_frame(const scan_index& index, std::size_t i, float* dst) -> bool;

    // every projection of the scan sits in a file of 16-bit samples: the group can take them as they are
    auto scan_is_u16(const scan_index& index) -> bool;
    // projection i of such a scan, not widened (dim_x * dim_y uint16 values)
    auto load_scan_frame_u16(const scan_index& index, std::size_t i, std::uint16_t* dst) -> bool;

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
