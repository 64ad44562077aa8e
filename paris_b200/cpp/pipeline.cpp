// pipeline.cpp -- see pipeline.h.
#include "pipeline.h"

#include <cmath>
#include <map>
#include <utility>

namespace paris
{
#if !defined(PARIS_B200_IN_PARIS_TREE)
    // src/geometry.cpp:71-84 / :86-130 -- the arithmetic lives behind the C ABI
    auto calculate_volume_geometry(const detector_geometry& det_geo) noexcept -> volume_geometry
    {
        const auto det = paris_b200_detector_geometry{det_geo.n_row, det_geo.n_col, det_geo.l_px_row, det_geo.l_px_col,
                                                      det_geo.delta_s, det_geo.delta_t, det_geo.d_so, det_geo.d_od,
                                                      det_geo.delta_phi};
        auto v = paris_b200_volume_geometry{};
        paris_b200_calculate_volume_geometry(&det, &v);
        return volume_geometry{v.dim_x, v.dim_y, v.dim_z, v.l_vx_x, v.l_vx_y, v.l_vx_z};
    }

    auto apply_roi(const volume_geometry& vol_geo, std::uint32_t x1, std::uint32_t x2, std::uint32_t y1,
                   std::uint32_t y2, std::uint32_t z1, std::uint32_t z2) noexcept -> volume_geometry
    {
        const auto v = paris_b200_volume_geometry{vol_geo.dim_x, vol_geo.dim_y, vol_geo.dim_z,
                                                  vol_geo.l_vx_x, vol_geo.l_vx_y, vol_geo.l_vx_z};
        const auto r = paris_b200_roi{x1, x2, y1, y2, z1, z2};
        auto o = paris_b200_volume_geometry{};
        paris_b200_apply_roi(&v, &r, &o);
        return volume_geometry{o.dim_x, o.dim_y, o.dim_z, o.l_vx_x, o.l_vx_y, o.l_vx_z};
    }
#endif

    auto load(const backend::projection_host_type& p) -> backend::projection_device_type
    {
        auto d_p = backend::make_projection_device(p.dim_x, p.dim_y);
        backend::copy_h2d(p, d_p);
        return d_p;
    }

    auto weight(backend::projection_device_type& p, const detector_geometry& det_geo) -> void
    {
        const auto n_row_f = static_cast<float>(det_geo.n_row);
        const auto n_col_f = static_cast<float>(det_geo.n_col);
        const auto h_min = (det_geo.delta_s * det_geo.l_px_row) - ((n_row_f * det_geo.l_px_row) / 2);
        const auto v_min = (det_geo.delta_t * det_geo.l_px_col) - ((n_col_f * det_geo.l_px_col) / 2);
        const auto d_sd = std::abs(det_geo.d_so) + std::abs(det_geo.d_od);
        backend::weight(p, h_min, v_min, d_sd, det_geo.l_px_row, det_geo.l_px_col);
    }

    auto filter(backend::projection_device_type& p, const detector_geometry& det_geo) -> void
    {
        const auto filter_size = paris_b200_filter_size(det_geo.n_row);
        const auto tau = det_geo.l_px_row;
        // per thread (= per device), like the reference's thread_local static k
        thread_local std::map<std::pair<std::uint32_t, float>, backend::filter_buffer_type> tables;
        auto it = tables.find({filter_size, tau});
        if(it == tables.end())
            it = tables.emplace(std::make_pair(filter_size, tau), backend::make_filter(filter_size, tau)).first;
        backend::apply_filter(p, it->second, filter_size, det_geo.n_col);
    }

    auto backproject(const backend::projection_device_type& p, backend::volume_device_type& v,
                     std::uint32_t v_offset, const detector_geometry& det_geo, const volume_geometry& vol_geo,
                     bool enable_angles, bool enable_roi, const region_of_interest& roi) -> void
    {
        const auto delta_s = det_geo.delta_s * det_geo.l_px_row;
        const auto delta_t = det_geo.delta_t * det_geo.l_px_col;

        auto phi = enable_angles ? p.phi : static_cast<float>(p.idx) * det_geo.delta_phi;
        phi *= static_cast<float>(M_PI) / 180.f;

        backend::backproject(p, v, v_offset, det_geo, vol_geo, enable_roi, roi, std::sin(phi), std::cos(phi),
                             delta_s, delta_t);
    }

    auto make_volume(const subvolume_geometry& subvol_geo, bool last) -> backend::volume_device_type
    {
        const auto dim_z = subvol_geo.dim_z + (last ? subvol_geo.remainder : 0u);
        return backend::make_volume_device(subvol_geo.dim_x, subvol_geo.dim_y, dim_z);
    }

    auto reconstruct_task(const slab_task& t, const float* h_stack, std::uint32_t n_proj, std::uint32_t first_idx,
                          std::uint32_t idx_stride, float* h_region) -> void
    {
        const auto last = (t.num - t.id) <= 1u;
        auto v = make_volume(t.subvol_geo, last);
        const auto offset = t.id * t.subvol_geo.dim_z;
        const auto px = static_cast<std::size_t>(t.det_geo.n_row) * t.det_geo.n_col;

        for(auto i = 0u; i < n_proj; ++i)
        {
            // source.load_next(): the stack is already in pinned memory, so the host projection aliases it
            auto p = backend::borrow_projection_host(const_cast<float*>(h_stack) + i * px, t.det_geo.n_row,
                                                     t.det_geo.n_col);
            p.idx = first_idx + i * idx_stride;
            auto d_p = load(p);
            weight(d_p, t.det_geo);
            filter(d_p, t.det_geo);
            backproject(d_p, v, offset, t.det_geo, t.vol_geo, false, t.enable_roi, t.roi);
        }

        // sink.save(v): device -> host, straight to the slab's place in the region
        const auto slice = static_cast<std::size_t>(v.dim_x) * v.dim_y;
        auto h_v = backend::volume_host_type{backend::volume_host_buffer_type{h_region + slice * offset}, v.dim_x,
                                             v.dim_y, v.dim_z, offset};
        backend::copy_d2h(v, h_v);
        (void)h_v.buf.release(); // caller's memory
    }
}
