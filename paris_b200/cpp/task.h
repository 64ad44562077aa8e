// task.h -- one z-slab of the reconstruction region: what /root/reference/src/task.h:33-57 carries per task, with
// the settings every task of a run shares (paths, detector, ROI, angle file, quality) grouped in one member instead
// of being repeated field by field (src/task.cpp:38-48 copies them into every task).
#pragma once

#include <cstdint>
#include <queue>
#include <string>

#include "paris_types.h"
#include "program_options.h"

namespace paris
{
    // the part of the program options a device thread needs to run a task
    struct scan_settings
    {
        std::string input_path;
        detector_geometry det_geo;
        bool enable_roi;
        region_of_interest roi;
        bool enable_angles;
        std::string angle_path;
        std::uint16_t quality;
    };

    struct task
    {
        std::uint32_t id;              // slab index; the slab starts at slice id * subvol_geo.dim_z of the region
        std::uint32_t num;             // number of slabs (the last one also takes subvol_geo.remainder)
        scan_settings scan;
        volume_geometry vol_geo;       // the FULL volume (src/task.cpp:41 passes vol_geo, not the ROI geometry)
        subvolume_geometry subvol_geo;
    };

    auto make_tasks(const program_options& po, const volume_geometry& vol_geo, const subvolume_info& subvol_info)
        -> std::queue<task>;
}
