// b200/backend.h -- namespace paris::b200: the B200 backend behind PARIS's backend contract.
//
// Provides, in one namespace, every name src/backend.h:26-46 expects of a backend (canonical text
// src/generic/backend.h:42-88; CUDA flavour src/cuda/backend.h:46-102): the ten buffer/projection/
// volume types, make_*/copy_*, make_subvolume_information, weight, make_filter/apply_filter,
// backproject, get_devices/set_device.  Everything forwards to the C ABI of include/paris_b200.h.
//
// Two deliberate differences from the legacy CUDA backend, both invisible to the callers:
//   * stages are LAZY.  weight() and apply_filter() only record what has to happen to the projection;
//     backproject() hands the raw projection plus those records to one fused weight+filter kernel that
//     writes straight into the filtered stack, and the backprojection itself runs once per batch of
//     projections.  The pending work is materialised on demand when a projection is observed
//     (copy_d2h) and the pending batch is flushed when the volume is (copy_d2h of a volume is the only
//     observer, src/sink.cpp:77).
//   * nothing synchronises per call (the legacy backend syncs its stream in every stage,
//     src/cuda/weighting.cu:72, filtering.cu:260, backprojection.cu:236, memory.cpp:66).
#pragma once

#include <cstdint>
#include <memory>
#include <vector>

#include "../../../include/paris_b200.h"
#include "../paris_types.h"

namespace paris
{
    namespace b200
    {
        struct host_deleter { auto operator()(float* p) const noexcept -> void; };
        struct device_deleter { auto operator()(float* p) const noexcept -> void; };
        struct volume_deleter { auto operator()(float* p) const noexcept -> void; };
        struct filter_deleter { auto operator()(paris_b200_filter* p) const noexcept -> void; };

        using projection_host_buffer_type = std::unique_ptr<float[], host_deleter>;
        using projection_device_buffer_type = std::unique_ptr<float[], device_deleter>;
        using volume_host_buffer_type = std::unique_ptr<float[], host_deleter>;
        using volume_device_buffer_type = std::unique_ptr<float[], volume_deleter>;

        // What still has to happen to a device projection before it may be observed.
        struct metadata
        {
            bool weight_pending = false;
            paris_b200_weighting weighting{};
            bool filter_pending = false;
            const paris_b200_filter* filter = nullptr;
            std::uint32_t filter_size = 0;
        };

        using projection_host_type = projection<projection_host_buffer_type, metadata>;
        using projection_device_type = projection<projection_device_buffer_type, metadata>;
        using volume_host_type = volume<volume_host_buffer_type>;
        using volume_device_type = volume<volume_device_buffer_type>;

        auto make_projection_host(std::uint32_t dim_x, std::uint32_t dim_y) -> projection_host_type;
        auto make_projection_device(std::uint32_t dim_x, std::uint32_t dim_y) -> projection_device_type;

        auto make_volume_host(std::uint32_t dim_x, std::uint32_t dim_y, std::uint32_t dim_z) -> volume_host_type;
        auto make_volume_device(std::uint32_t dim_x, std::uint32_t dim_y, std::uint32_t dim_z) -> volume_device_type;

        auto copy_h2d(const projection_host_type& h_p, projection_device_type& d_p) -> void;
        auto copy_d2h(const projection_device_type& d_p, projection_host_type& h_p) -> void;

        auto copy_h2d(const volume_host_type& h_v, volume_device_type& d_v) -> void;
        auto copy_d2h(const volume_device_type& d_v, volume_host_type& h_v) -> void;

        auto make_subvolume_information(const volume_geometry& vol_geo, const detector_geometry& det_geo)
            -> subvolume_info;

        auto weight(projection_device_type& p, float h_min, float v_min, float d_sd, float l_px_row, float l_px_col)
            -> void;

        using filter_buffer_type = std::unique_ptr<paris_b200_filter, filter_deleter>;
        auto make_filter(std::uint32_t size, float tau) -> filter_buffer_type;
        auto apply_filter(projection_device_type& p, const filter_buffer_type& k, std::uint32_t filter_size,
                          std::uint32_t n_col) -> void;

        auto backproject(const projection_device_type& p, volume_device_type& v, std::uint32_t v_offset,
                         const detector_geometry& det_geo, const volume_geometry& vol_geo,
                         bool enable_roi, const region_of_interest& roi,
                         float sin, float cos, float delta_s, float delta_t) -> void;

        /**
         * Device management
         * */
        using device_handle = int;
        auto get_devices() -> std::vector<device_handle>;
        auto set_device(device_handle& device) -> void;

        // ---- beyond the contract (used by this repository's own driver and tests) ----------------------

        // The calling thread's context (created on first use for device 0 unless set_device ran first).
        auto context() -> paris_b200_ctx*;
        // Number of equal z-slabs make_subvolume_information returns; 0 (default) = size by free memory.
        auto set_slab_count(int num) noexcept -> void;
        // Wrap caller-owned pinned memory as a host projection without copying (never freed by the wrapper).
        auto borrow_projection_host(float* pinned, std::uint32_t dim_x, std::uint32_t dim_y) -> projection_host_type;
        // Run every deferred backprojection batch now.
        auto flush() -> void;
    }
}
