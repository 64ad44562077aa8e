// paris_types.h -- the plain data carriers that cross the backend boundary.
//
// Inside the PARIS source tree these come from the reference's own headers (src/geometry.h,
// src/projection.h, src/volume.h, src/region_of_interest.h, src/subvolume_information.h,
// src/exception.h) and this file only forwards to them.  Stand-alone (this repository's own build,
// the GPU box) the same types are declared here with the same names and members, because that IS the
// interface: field order matters for aggregate initialisation in the reference's callers.
#pragma once

#if defined(PARIS_B200_IN_PARIS_TREE)

#include "exception.h"
#include "geometry.h"
#include "projection.h"
#include "region_of_interest.h"
#include "subvolume_information.h"
#include "volume.h"

#else

#include <cstdint>
#include <stdexcept>
#include <utility>

namespace paris
{
    // src/geometry.h:30-47
    struct detector_geometry
    {
        std::uint32_t n_row;
        std::uint32_t n_col;
        float l_px_row;
        float l_px_col;
        float delta_s;
        float delta_t;
        float d_so;
        float d_od;
        float delta_phi;
    };

    // src/geometry.h:49-58
    struct volume_geometry
    {
        std::uint32_t dim_x;
        std::uint32_t dim_y;
        std::uint32_t dim_z;
        float l_vx_x;
        float l_vx_y;
        float l_vx_z;
    };

    // src/geometry.h:60-69
    struct subvolume_geometry
    {
        std::uint32_t dim_x;
        std::uint32_t dim_y;
        std::uint32_t dim_z;
        std::uint32_t remainder;
    };

    // src/region_of_interest.h:30-38
    struct region_of_interest
    {
        std::uint32_t x1;
        std::uint32_t x2;
        std::uint32_t y1;
        std::uint32_t y2;
        std::uint32_t z1;
        std::uint32_t z2;
    };

    // src/subvolume_information.h:30-34
    struct subvolume_info
    {
        subvolume_geometry geo;
        int num;
    };

    // src/projection.h:31-46
    template <typename BufferType, typename Metadata>
    struct projection
    {
        projection() noexcept = default;
        projection(BufferType b, std::uint32_t x, std::uint32_t y, std::uint32_t i, float ph, Metadata m) noexcept
        : buf(std::move(b)), dim_x{x}, dim_y{y}, idx{i}, phi{ph}, meta(std::move(m))
        {
        }

        BufferType buf = BufferType{};
        std::uint32_t dim_x = 0;
        std::uint32_t dim_y = 0;
        std::uint32_t idx = 0;
        float phi = 0.f;
        Metadata meta = Metadata{};
    };

    // src/volume.h:31-45
    template <class BufferType>
    struct volume
    {
        volume() noexcept = default;
        volume(BufferType b, std::uint32_t x, std::uint32_t y, std::uint32_t z, std::uint32_t o) noexcept
        : buf{std::move(b)}, dim_x{x}, dim_y{y}, dim_z{z}, off{o}
        {
        }

        BufferType buf = BufferType{};
        std::uint32_t dim_x = 0;
        std::uint32_t dim_y = 0;
        std::uint32_t dim_z = 0;
        std::uint32_t off = 0;
    };

    // src/exception.h:31-41
    class stage_construction_error : public std::runtime_error
    {
        public:
            using std::runtime_error::runtime_error;
    };

    class stage_runtime_error : public std::runtime_error
    {
        public:
            using std::runtime_error::runtime_error;
    };

    // src/geometry.h:71-76 (implemented in paris_b200/cpp/pipeline.cpp on top of the C ABI)
    auto calculate_volume_geometry(const detector_geometry& det_geo) noexcept -> volume_geometry;
    auto apply_roi(const volume_geometry& vol_geo, std::uint32_t roi_x1, std::uint32_t roi_x2,
                   std::uint32_t roi_y1, std::uint32_t roi_y2, std::uint32_t roi_z1, std::uint32_t roi_z2) noexcept
        -> volume_geometry;
}

#endif
