// dropin_api.cpp -- a C entry point over the C++ pipeline (namespace paris, backend = paris::b200) so that
// tests/ and bench.py can drive the reference-shaped per-projection loop (src/main.cpp:79-109) through
// ctypes without paying a Python call per stage.
#include <cstdio>
#include <cstring>
#include <exception>

#include "pipeline.h"

extern "C"
{
    // Reconstruct slab `slab_id` of `num_slabs` equal z-slabs (remainder on the last) of the region
    // (region = ROI box if enable_roi, else the full volume; region_dim_* are its dimensions) from an
    // in-memory stack of raw projections and write it into h_region at its z offset.
    // Returns 0, or -1 with the message in err (if non-null).
    int paris_b200_dropin_reconstruct(const float* h_stack, std::uint32_t n_proj, std::uint32_t first_idx,
                                      std::uint32_t idx_stride, const paris_b200_detector_geometry* det,
                                      const paris_b200_volume_geometry* vol_full, int enable_roi,
                                      const paris_b200_roi* roi, std::uint32_t region_dim_x,
                                      std::uint32_t region_dim_y, std::uint32_t region_dim_z, std::uint32_t slab_id,
                                      std::uint32_t num_slabs, int device, float* h_region, char* err,
                                      std::size_t err_len)
    {
        try
        {
            auto dev = paris::b200::device_handle{device};
            paris::b200::set_device(dev);

            auto t = paris::slab_task{};
            t.id = slab_id;
            t.num = num_slabs;
            t.det_geo = paris::detector_geometry{det->n_row, det->n_col, det->l_px_row, det->l_px_col, det->delta_s,
                                                 det->delta_t, det->d_so, det->d_od, det->delta_phi};
            t.vol_geo = paris::volume_geometry{vol_full->dim_x, vol_full->dim_y, vol_full->dim_z,
                                               vol_full->l_vx_x, vol_full->l_vx_y, vol_full->l_vx_z};
            const auto region = paris::volume_geometry{region_dim_x, region_dim_y, region_dim_z,
                                                       vol_full->l_vx_x, vol_full->l_vx_y, vol_full->l_vx_z};
            paris::b200::set_slab_count(static_cast<int>(num_slabs));
            t.subvol_geo = paris::b200::make_subvolume_information(region, t.det_geo).geo;
            t.enable_roi = enable_roi != 0;
            if(t.enable_roi)
                t.roi = paris::region_of_interest{roi->x1, roi->x2, roi->y1, roi->y2, roi->z1, roi->z2};
            else
                t.roi = paris::region_of_interest{0u, 0u, 0u, 0u, 0u, 0u};

            paris::reconstruct_task(t, h_stack, n_proj, first_idx, idx_stride, h_region);
            return 0;
        }
        catch(const std::exception& e)
        {
            if(err != nullptr && err_len > 0)
            {
                std::strncpy(err, e.what(), err_len - 1);
                err[err_len - 1] = '\0';
            }
            return -1;
        }
    }

    // paris::b200::set_device for the calling thread.
    int paris_b200_dropin_set_device(int device)
    {
        try
        {
            auto dev = paris::b200::device_handle{device};
            paris::b200::set_device(dev);
            return 0;
        }
        catch(const std::exception&)
        {
            return -1;
        }
    }

    // The calling thread's context, so callers can time / count launches on the same streams.
    paris_b200_ctx* paris_b200_dropin_context(void)
    {
        try
        {
            return paris::b200::context();
        }
        catch(const std::exception&)
        {
            return nullptr;
        }
    }
}
