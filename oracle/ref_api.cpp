/*
 * oracle/ref_api.cpp -- TEST INFRASTRUCTURE, not product code.
 *
 * A C ABI around the UNMODIFIED reference pipeline so that tests/ and the
 * bench's CPU-baseline leg can drive it through ctypes.  This file is compiled
 * together with the reference's own sources, taken where they lie under
 * /root/reference/src (see oracle/Makefile; nothing is copied into this repo):
 *
 *   src/openmp/{weighting,filtering,backprojection,memory,subvolume_information}.cpp
 *   src/{weighting,filtering,backprojection,loader,make_volume,geometry}.cpp
 *
 * and calls them exactly the way the reference's hot loop does
 * (src/main.cpp:98-105): load -> weight -> filter -> backproject.
 *
 * NB: the reference freezes geometry in function-local statics at
 * the first call, so ONE loaded copy of this library serves ONE geometry.
 * Python loads a private copy of the .so per geometry (oracle/__init__.py).
 */
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <memory>

#include <omp.h>

#include "backend.h"
#include "backprojection.h"
#include "filtering.h"
#include "geometry.h"
#include "loader.h"
#include "make_volume.h"
#include "weighting.h"

namespace
{
    using dev_proj = paris::backend::projection_device_type;
    using dev_vol = paris::backend::volume_device_type;

    // Wrap caller memory in the reference's owning buffer type without copying;
    // released (not freed) before the wrapper returns.
    struct borrowed_projection
    {
        dev_proj p;
        borrowed_projection(float* data, std::uint32_t dim_x, std::uint32_t dim_y, std::uint32_t idx, float phi)
        {
            p.buf = paris::backend::projection_device_buffer_type{data};
            p.dim_x = dim_x;
            p.dim_y = dim_y;
            p.idx = idx;
            p.phi = phi;
        }
        ~borrowed_projection() { (void)p.buf.release(); }
    };

    struct borrowed_volume
    {
        dev_vol v;
        borrowed_volume(float* data, std::uint32_t dx, std::uint32_t dy, std::uint32_t dz)
        {
            v.buf = paris::backend::volume_device_buffer_type{data};
            v.dim_x = dx;
            v.dim_y = dy;
            v.dim_z = dz;
            v.off = 0;
        }
        ~borrowed_volume() { (void)v.buf.release(); }
    };
}

extern "C"
{
    struct ref_detector_geometry
    {
        std::uint32_t n_row, n_col;
        float l_px_row, l_px_col, delta_s, delta_t, d_so, d_od, delta_phi;
    };

    struct ref_volume_geometry
    {
        std::uint32_t dim_x, dim_y, dim_z;
        float l_vx_x, l_vx_y, l_vx_z;
    };

    struct ref_roi
    {
        std::uint32_t x1, x2, y1, y2, z1, z2;
    };

    static paris::detector_geometry to_det(const ref_detector_geometry* g)
    {
        return paris::detector_geometry{g->n_row, g->n_col, g->l_px_row, g->l_px_col, g->delta_s, g->delta_t,
                                        g->d_so, g->d_od, g->delta_phi};
    }

    static paris::volume_geometry to_vol(const ref_volume_geometry* g)
    {
        return paris::volume_geometry{g->dim_x, g->dim_y, g->dim_z, g->l_vx_x, g->l_vx_y, g->l_vx_z};
    }

    int paris_ref_num_threads(void) { return omp_get_max_threads(); }
    // (torchrun exports OMP_NUM_THREADS=1 to its children: the CPU arm of the bench sets the count itself)
    void paris_ref_set_num_threads(int n) { omp_set_num_threads(n > 0 ? n : 1); }

    // src/geometry.cpp:71
    void paris_ref_calculate_volume_geometry(const ref_detector_geometry* det, ref_volume_geometry* out)
    {
        const auto v = paris::calculate_volume_geometry(to_det(det));
        *out = ref_volume_geometry{v.dim_x, v.dim_y, v.dim_z, v.l_vx_x, v.l_vx_y, v.l_vx_z};
    }

    // src/geometry.cpp:86
    void paris_ref_apply_roi(const ref_volume_geometry* vol, const ref_roi* r, ref_volume_geometry* out)
    {
        const auto v = paris::apply_roi(to_vol(vol), r->x1, r->x2, r->y1, r->y2, r->z1, r->z2);
        *out = ref_volume_geometry{v.dim_x, v.dim_y, v.dim_z, v.l_vx_x, v.l_vx_y, v.l_vx_z};
    }

    // src/weighting.cpp:32 -> src/openmp/weighting.cpp:32 ; in place
    void paris_ref_weight(float* proj, const ref_detector_geometry* det)
    {
        borrowed_projection b{proj, det->n_row, det->n_col, 0u, 0.f};
        paris::weight(b.p, to_det(det));
    }

    // src/filtering.cpp:32 -> src/openmp/filtering.cpp:139,167 ; in place
    void paris_ref_filter(float* proj, const ref_detector_geometry* det)
    {
        borrowed_projection b{proj, det->n_row, det->n_col, 0u, 0.f};
        paris::filter(b.p, to_det(det));
    }

    // The filter table the reference builds (src/openmp/filtering.cpp:139-165): K[x] for x = 0..size/2
    void paris_ref_make_filter(std::uint32_t size, float tau, float* k_out)
    {
        auto k = paris::backend::make_filter(size, tau);
        for(auto x = 0u; x < size / 2 + 1; ++x)
            k_out[x] = k[x][0];
    }

    // src/backprojection.cpp:37 -> src/openmp/backprojection.cpp:156 ; accumulates into vol
    void paris_ref_backproject(const float* proj, std::uint32_t idx, float phi_deg, int enable_angles,
                               float* vol, std::uint32_t v_dim_x, std::uint32_t v_dim_y, std::uint32_t v_dim_z,
                               std::uint32_t v_offset,
                               const ref_detector_geometry* det, const ref_volume_geometry* vol_full,
                               int enable_roi, const ref_roi* roi)
    {
        borrowed_projection b{const_cast<float*>(proj), det->n_row, det->n_col, idx, phi_deg};
        borrowed_volume v{vol, v_dim_x, v_dim_y, v_dim_z};
        const auto r = paris::region_of_interest{roi->x1, roi->x2, roi->y1, roi->y2, roi->z1, roi->z2};
        paris::backproject(b.p, v.v, v_offset, to_det(det), to_vol(vol_full), enable_angles != 0, enable_roi != 0, r);
    }

    /*
     * The reference's hot loop (src/main.cpp:98-105) over an in-memory stack of raw projections.
     * stack: n_proj x (n_col x n_row) floats, untouched (each projection is copied through
     * paris::load, as the reference does).  first_idx/idx_stride let the caller run a bounded
     * sample (projection i of the sample carries idx = first_idx + i*idx_stride).
     * times[3]: accumulated seconds in weight / filter / backproject, excluding the first
     * projection (FFT plans and function-local statics are built there).
     */
    void paris_ref_reconstruct(const float* stack, std::uint32_t n_proj, std::uint32_t first_idx,
                               std::uint32_t idx_stride,
                               float* vol, std::uint32_t v_dim_x, std::uint32_t v_dim_y, std::uint32_t v_dim_z,
                               const ref_detector_geometry* det, const ref_volume_geometry* vol_full,
                               int enable_roi, const ref_roi* roi, double* times)
    {
        const auto det_geo = to_det(det);
        const auto vol_geo = to_vol(vol_full);
        const auto r = paris::region_of_interest{roi->x1, roi->x2, roi->y1, roi->y2, roi->z1, roi->z2};
        borrowed_volume v{vol, v_dim_x, v_dim_y, v_dim_z};
        const auto px = static_cast<std::size_t>(det->n_row) * det->n_col;
        double t_w = 0.0, t_f = 0.0, t_b = 0.0;

        for(auto i = 0u; i < n_proj; ++i)
        {
            auto h_p = paris::backend::make_projection_host(det->n_row, det->n_col);
            std::memcpy(h_p.buf.get(), stack + i * px, px * sizeof(float));
            h_p.idx = first_idx + i * idx_stride;

            auto d_p = paris::load(h_p);
            const auto t0 = omp_get_wtime();
            paris::weight(d_p, det_geo);
            const auto t1 = omp_get_wtime();
            paris::filter(d_p, det_geo);
            const auto t2 = omp_get_wtime();
            paris::backproject(d_p, v.v, 0u, det_geo, vol_geo, false, enable_roi != 0, r);
            const auto t3 = omp_get_wtime();
            if(i > 0 || n_proj == 1)
            {
                t_w += t1 - t0;
                t_f += t2 - t1;
                t_b += t3 - t2;
            }
        }
        if(times != nullptr)
        {
            times[0] = t_w;
            times[1] = t_f;
            times[2] = t_b;
        }
    }
}
