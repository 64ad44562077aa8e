// backproject.cuh -- parameter blocks shared by the backprojection kernels.
#pragma once

#include "common.cuh"

namespace pb
{
    struct bp_geometry
    {
        uint32_t v_dim_x, v_dim_y, v_dim_z;   // slab being updated
        uint32_t full_x, full_y, full_z;      // full volume (centre of rotation = centre of this box)
        uint32_t off_x, off_y, off_z;         // ROI origin (+ slab offset on z)
        float l_vx_x, l_vx_y, l_vx_z;
        uint32_t p_dim_x, p_dim_y, pitch;     // n_row, n_col, slot line pitch
        uint32_t layout;                      // kLayoutPlain / kLayoutSplit2
        float l_px_x, l_px_y;
        float d_so, d_sd, delta_s, delta_t;   // delta_* in millimetres
        // derived once on the host for the table builder of the TMA kernel
        float so_over_sd;                     // d_so / d_sd
        double min_v_d;                       // -(dim_y * l_px_y / 2) - delta_t
        double inv_l_px_y_d;                  // 1 / l_px_y
        double dv_scale_d;                    // l_vx_z / l_px_y
        float zero;                           // 0.f that the assembler cannot see (keeps z_m * factor a rounded product)
    };

    struct bp_angles
    {
        int count;
        float sn[kMaxBatch];
        float cs[kMaxBatch];
    };

    bp_geometry make_bp_geometry(const bp_target& t, uint32_t pitch, uint32_t layout);

    // *handled = false when the geometry does not fit the kernel's tiles (caller falls back to the exact
    // kernel) unless `required`, in which case that is an error.
    int launch_bp_tma(paris_b200_ctx* ctx, const float* d_stack, size_t slot_floats, uint32_t first,
                      const bp_geometry& g, const bp_angles& a, float* d_vol, bool required, bool* handled);
}
