// pipeline.h -- the backend-agnostic stage wrappers of PARIS, re-stated over namespace paris::b200.
//
// Same names, argument meaning and derived constants as the reference's
//   load         src/loader.cpp:28-33          weight       src/weighting.cpp:32-45
//   filter       src/filtering.cpp:32-45       backproject  src/backprojection.cpp:37-69
//   make_volume  src/make_volume.cpp:30-37
// so the reconstruction loop of src/main.cpp:79-109 can be written against them unchanged.  Unlike the
// reference, nothing is frozen in function-local statics: constants are derived per call
// (a handful of float operations) and the filter table is cached per thread and (size, tau).
//
// Inside the PARIS tree these wrappers are NOT needed -- the reference's own weighting.cpp, filtering.cpp,
// backprojection.cpp, loader.cpp and make_volume.cpp compile unmodified against paris::b200
// (oracle/Makefile target ref_b200).
#pragma once

#include <cstdint>

#include "b200/backend.h"

namespace paris
{
    namespace backend = b200;

    auto load(const backend::projection_host_type& p) -> backend::projection_device_type;
    auto weight(backend::projection_device_type& p, const detector_geometry& det_geo) -> void;
    auto filter(backend::projection_device_type& p, const detector_geometry& det_geo) -> void;
    auto backproject(const backend::projection_device_type& p, backend::volume_device_type& v,
                     std::uint32_t v_offset, const detector_geometry& det_geo, const volume_geometry& vol_geo,
                     bool enable_angles, bool enable_roi, const region_of_interest& roi) -> void;
    auto make_volume(const subvolume_geometry& subvol_geo, bool last) -> backend::volume_device_type;

    // One task = one z-slab of the (ROI-)region, as src/task.h:33-54 describes it, minus the file paths.
    struct slab_task
    {
        std::uint32_t id;
        std::uint32_t num;
        detector_geometry det_geo;
        volume_geometry vol_geo;       // FULL volume geometry (src/task.cpp:41: vol_geo, not roi_geo)
        subvolume_geometry subvol_geo;
        bool enable_roi;
        region_of_interest roi;
    };

    // The body of reconstruct() (src/main.cpp:93-107) for one task over an in-memory stack of raw
    // projections (pinned host memory, n_proj x n_col x n_row floats): load -> weight -> filter ->
    // backproject per projection, then the slab is copied to h_region at its z offset
    // (slices [id*dim_z, id*dim_z + slab_dim_z) of the region, x fastest) -- the host reassembly the
    // reference's sink intended (src/sink.cpp:72-93; its volume.off is never set).
    auto reconstruct_task(const slab_task& t, const float* h_stack, std::uint32_t n_proj, std::uint32_t first_idx,
                          std::uint32_t idx_stride, float* h_region) -> void;
}
