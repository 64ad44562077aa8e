// task.h -- one z-slab of the reconstruction region (/root/reference/src/task.h:33-57, src/task.cpp:33-51).
#pragma once

#include <cstdint>
#include <queue>
#include <string>

#include "paris_types.h"
#include "program_options.h"

namespace paris
{
    struct task
    {
        std::uint32_t id;
        std::uint32_t num;

        std::string input_path;

        detector_geometry det_geo;
        volume_geometry vol_geo;       // the FULL volume (src/task.cpp:41 passes vol_geo, not the ROI geometry)
        subvolume_geometry subvol_geo;

        bool enable_roi;
        region_of_interest roi;

        bool enable_angles;
        std::string angle_path;

        std::uint16_t quality;
    };

    auto make_tasks(const program_options& po, const volume_geometry& vol_geo, const subvolume_info& subvol_info)
        -> std::queue<task>;
}
