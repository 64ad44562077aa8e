// task.cpp -- see task.h.
#include "task.h"

namespace paris
{
    auto make_tasks(const program_options& po, const volume_geometry& vol_geo, const subvolume_info& subvol_info)
        -> std::queue<task>
    {
        auto q = std::queue<task>{};
        const auto num = static_cast<std::uint32_t>(subvol_info.num > 0 ? subvol_info.num : 0);
        for(auto id = 0u; id < num; ++id)
            q.push(task{id, num, po.input_path, po.det_geo, vol_geo, subvol_info.geo, po.enable_roi, po.roi,
                        po.enable_angles, po.angle_path, po.quality});
        return q;
    }
}
