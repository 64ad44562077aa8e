/*
 * oracle/ref_b200_api.cpp -- TEST INFRASTRUCTURE: the drop-in proof.
 *
 * Compiled together with the reference's UNMODIFIED stage wrappers
 *   /root/reference/src/{weighting,filtering,backprojection,loader,make_volume,geometry}.cpp
 * but with paris::backend = paris::b200 (oracle/b200_select.h), i.e. the reference's own hot loop
 * (src/main.cpp:93-107) running on the sm_100a backend.  tests/test_dropin.py compares the result with
 * the same loop on the reference's OpenMP backend (oracle/_ref/libparis_ref.so).
 */
#include <cstdint>
#include <cstring>
#include <exception>

#include "backend.h"
#include "backprojection.h"
#include "filtering.h"
#include "geometry.h"
#include "loader.h"
#include "make_volume.h"
#include "weighting.h"

extern "C" int paris_ref_b200_reconstruct(const float* stack, std::uint32_t n_proj, float* vol,
                                          std::uint32_t v_dim_x, std::uint32_t v_dim_y, std::uint32_t v_dim_z,
                                          const paris_b200_detector_geometry* det,
                                          const paris_b200_volume_geometry* vol_full, int enable_roi,
                                          const paris_b200_roi* roi)
{
    try
    {
        const auto det_geo = paris::detector_geometry{det->n_row, det->n_col, det->l_px_row, det->l_px_col,
                                                      det->delta_s, det->delta_t, det->d_so, det->d_od,
                                                      det->delta_phi};
        const auto vol_geo = paris::volume_geometry{vol_full->dim_x, vol_full->dim_y, vol_full->dim_z,
                                                    vol_full->l_vx_x, vol_full->l_vx_y, vol_full->l_vx_z};
        const auto r = paris::region_of_interest{roi->x1, roi->x2, roi->y1, roi->y2, roi->z1, roi->z2};
        const auto px = static_cast<std::size_t>(det->n_row) * det->n_col;

        auto devices = paris::backend::get_devices();
        paris::backend::set_device(devices.at(0));

        // src/main.cpp:95: one task covering the whole region
        auto v = paris::make_volume(paris::subvolume_geometry{v_dim_x, v_dim_y, v_dim_z, 0u}, true);
        for(auto i = 0u; i < n_proj; ++i)
        {
            auto p = paris::backend::make_projection_host(det->n_row, det->n_col);   // his::load, src/his.cpp:161
            std::memcpy(p.buf.get(), stack + i * px, px * sizeof(float));
            p.idx = i;
            auto d_p = paris::load(p);                                                // src/main.cpp:101
            paris::weight(d_p, det_geo);                                              // :102
            paris::filter(d_p, det_geo);                                              // :103
            paris::backproject(d_p, v, 0u, det_geo, vol_geo, false, enable_roi != 0, r); // :104
        }
        // sink.save, src/sink.cpp:76-77
        auto h_v = paris::backend::make_volume_host(v.dim_x, v.dim_y, v.dim_z);
        paris::backend::copy_d2h(v, h_v);
        std::memcpy(vol, h_v.buf.get(), static_cast<std::size_t>(v_dim_x) * v_dim_y * v_dim_z * sizeof(float));
        return 0;
    }
    catch(const std::exception&)
    {
        return -1;
    }
}
