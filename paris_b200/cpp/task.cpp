// task.cpp -- see task.h.
#include "task.h"

namespace paris
{
    auto make_tasks(const program_options& po, const volume_geometry& vol_geo, const subvolume_info& subvol_info)
        -> std::queue<task>
    {
        const auto scan = scan_settings{po.input_path, po.det_geo, po.enable_roi, po.roi, po.enable_angles,
                                        po.angle_path, po.quality};
        const auto slabs = static_cast<std::uint32_t>(subvol_info.num > 0 ? subvol_info.num : 0);
        auto tasks = std::queue<task>{};
        for(auto slab = 0u; slab < slabs; ++slab)
            tasks.push(task{slab, slabs, scan, vol_geo, subvol_info.geo});
        return tasks;
    }
}
