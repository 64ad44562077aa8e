// backproject.cu -- K2: voxel-driven backprojection of a BATCH of filtered projections.
//
// Replaces backend::backproject of the reference (/root/reference/src/openmp/backprojection.cpp:86-199;
// legacy CUDA src/cuda/backprojection.cu:64-130, one launch + one volume read-modify-write per projection,
// hardware 8-bit bilinear texture fetch).
//
// Two kernels read the same filtered stack (slots of n_row lines x pitch floats, detector-row index v
// fastest -- i.e. each projection TRANSPOSED, written that way by the filter kernel):
//
//   bp_exact_kernel   the reference's arithmetic, operation for operation, with round-to-nearest
//                     intrinsics so no FMA contraction happens.  Per voxel the projections of a batch are
//                     added in projection order to the value loaded from the volume, which is the
//                     reference's own summation order -- the result is bit-identical to the OpenMP
//                     backend for any batch size.  One thread per voxel; used for pinning the fast
//                     kernel, and as the fallback for geometries the fast kernel's tiles do not cover.
//
//   bp_tma_kernel     the production kernel (backproject_tma.cuh): lanes along z, per-column projective
//                     terms computed once per (column, projection) and shared through shared memory,
//                     filtered-projection tiles staged by TMA, a whole batch accumulated in registers.
#include "common.cuh"
#include "backproject.cuh"

namespace pb
{
    // src/openmp/backprojection.cpp:39-43
    __device__ __forceinline__ float vol_centered_coordinate(uint32_t coord, uint32_t dim, float size)
    {
        const float size2 = size / 2.f;
        return __fadd_rn(__fadd_rn(-__fmul_rn(static_cast<float>(dim), size2), size2),
                         __fmul_rn(static_cast<float>(coord), size));
    }

    // src/openmp/backprojection.cpp:45-50
    __device__ __forceinline__ float proj_real_coordinate(float coord, uint32_t dim, float size, float offset)
    {
        const float size2 = size / 2.f;
        const float min = __fsub_rn(-__fmul_rn(static_cast<float>(dim), size2), offset);
        return __fsub_rn(__fdiv_rn(__fsub_rn(coord, min), size), 0.5f);
    }

    // src/openmp/backprojection.cpp:52-84, on a transposed slot: p[x + y*dim_x] lives at slot[x*pitch + y]
    __device__ __forceinline__ float interpolate_exact(const float* __restrict__ slot, uint32_t pitch, uint32_t layout,
                                                       float x, float y, uint32_t dim_x, uint32_t dim_y)
    {
        const float x1 = floorf(x);
        const float x2 = __fadd_rn(x1, 1.f);
        const float y1 = floorf(y);
        const float y2 = __fadd_rn(y1, 1.f);
        float interp = 0.f;
        if(x1 >= 0.f && x2 < static_cast<float>(dim_x) && y1 >= 0.f && y2 < static_cast<float>(dim_y))
        {
            const uint32_t x1u = static_cast<uint32_t>(x1), y1u = static_cast<uint32_t>(y1);
            const float* c0 = slot + static_cast<size_t>(x1u) * pitch;
            const uint32_t o1 = line_offset(y1u, pitch, layout), o2 = line_offset(y1u + 1u, pitch, layout);
            const float q11 = __ldg(c0 + o1);
            const float q12 = __ldg(c0 + o2);
            const float q21 = __ldg(c0 + pitch + o1);
            const float q22 = __ldg(c0 + pitch + o2);
            const float dx = __fsub_rn(x2, x1), dy = __fsub_rn(y2, y1);
            const float wx1 = __fdiv_rn(__fsub_rn(x2, x), dx), wx2 = __fdiv_rn(__fsub_rn(x, x1), dx);
            const float interp_y1 = __fadd_rn(__fmul_rn(wx1, q11), __fmul_rn(wx2, q21));
            const float interp_y2 = __fadd_rn(__fmul_rn(wx1, q12), __fmul_rn(wx2, q22));
            const float wy1 = __fdiv_rn(__fsub_rn(y2, y), dy), wy2 = __fdiv_rn(__fsub_rn(y, y1), dy);
            interp = __fadd_rn(__fmul_rn(wy1, interp_y1), __fmul_rn(wy2, interp_y2));
        }
        return interp;
    }

    __global__ void __launch_bounds__(256)
    bp_exact_kernel(const float* __restrict__ stack, size_t slot_floats, float* __restrict__ vol, bp_geometry g,
                    bp_angles a)
    {
        const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
        const uint32_t l = blockIdx.y * blockDim.y + threadIdx.y;
        const uint32_t m = blockIdx.z;
        if(k >= g.v_dim_x || l >= g.v_dim_y)
            return;

        const float x_k = vol_centered_coordinate(k + g.off_x, g.full_x, g.l_vx_x);
        const float y_l = vol_centered_coordinate(l + g.off_y, g.full_y, g.l_vx_y);
        const float z_m = vol_centered_coordinate(m + g.off_z, g.full_z, g.l_vx_z);

        const size_t coord = static_cast<size_t>(k) + static_cast<size_t>(l) * g.v_dim_x
                           + static_cast<size_t>(m) * g.v_dim_x * g.v_dim_y;
        float acc = vol[coord];

        for(int p = 0; p < a.count; ++p)
        {
            const float sn = a.sn[p], cs = a.cs[p];
            const float s = __fadd_rn(__fmul_rn(x_k, cs), __fmul_rn(y_l, sn));
            const float t = __fadd_rn(__fmul_rn(-x_k, sn), __fmul_rn(y_l, cs));
            const float denom = __fadd_rn(s, g.d_so);
            const float factor = __fdiv_rn(g.d_sd, denom);
            const float h = proj_real_coordinate(__fmul_rn(t, factor), g.p_dim_x, g.l_px_x, g.delta_s);
            const float v = proj_real_coordinate(__fmul_rn(z_m, factor), g.p_dim_y, g.l_px_y, g.delta_t);
            const float det = interpolate_exact(stack + slot_floats * p, g.pitch, g.layout, h, v, g.p_dim_x, g.p_dim_y);
            const float u = -__fdiv_rn(g.d_so, denom);
            // 0.5f * det * u * u, left to right
            acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(__fmul_rn(0.5f, det), u), u));
        }
        vol[coord] = acc;
    }

    bp_geometry make_bp_geometry(const bp_target& t, uint32_t pitch, uint32_t layout)
    {
        bp_geometry g{};
        g.layout = layout;
        g.v_dim_x = t.v_dim_x;
        g.v_dim_y = t.v_dim_y;
        g.v_dim_z = t.v_dim_z;
        g.full_x = t.vol_full.dim_x;
        g.full_y = t.vol_full.dim_y;
        g.full_z = t.vol_full.dim_z;
        g.off_x = t.enable_roi ? t.roi.x1 : 0u;
        g.off_y = t.enable_roi ? t.roi.y1 : 0u;
        g.off_z = (t.enable_roi ? t.roi.z1 : 0u) + t.v_offset;
        g.l_vx_x = t.vol_full.l_vx_x;
        g.l_vx_y = t.vol_full.l_vx_y;
        g.l_vx_z = t.vol_full.l_vx_z;
        g.p_dim_x = t.det.n_row;
        g.p_dim_y = t.det.n_col;
        g.pitch = pitch;
        g.l_px_x = t.det.l_px_row;
        g.l_px_y = t.det.l_px_col;
        g.d_so = t.det.d_so;                                        // NOT abs: src/openmp/backprojection.cpp:176
        g.d_sd = std::fabs(t.det.d_so) + std::fabs(t.det.d_od);     // :177
        g.delta_s = t.delta_s_mm;
        g.delta_t = t.delta_t_mm;
        g.so_over_sd = g.d_so / g.d_sd;
        g.min_v_d = -(static_cast<double>(g.p_dim_y) * (static_cast<double>(g.l_px_y) / 2.0)) - static_cast<double>(g.delta_t);
        g.inv_l_px_y_d = 1.0 / static_cast<double>(g.l_px_y);
        g.dv_scale_d = static_cast<double>(g.l_vx_z) / static_cast<double>(g.l_px_y);
        g.zero = 0.f;
        return g;
    }

    uint32_t choose_stack_layout(const paris_b200_detector_geometry& det, const paris_b200_volume_geometry& vol_full)
    {
        // detector rows per voxel step in z on the rotation axis: l_vx_z * (d_sd / d_so) / l_px_col
        const double d_so = det.d_so, d_sd = std::fabs(det.d_so) + std::fabs(det.d_od);
        if(!(d_so > 0.0) || det.n_row <= 64u)
            return kLayoutPlain;
        const double dv = static_cast<double>(vol_full.l_vx_z) * (d_sd / d_so) / static_cast<double>(det.l_px_col);
        return (dv >= 1.45 && dv <= 2.75) ? kLayoutSplit2 : kLayoutPlain;
    }

    static int launch_exact(paris_b200_ctx* ctx, const float* d_first_slot, size_t slot_floats, const bp_geometry& g,
                            const bp_angles& a, float* d_vol)
    {
        const dim3 block(64, 4);
        const dim3 grid((g.v_dim_x + 63u) / 64u, (g.v_dim_y + 3u) / 4u, g.v_dim_z);
        if(grid.y > 65535u || grid.z > 65535u)
        {
            set_error("volume too large for the exact kernel's grid");
            return PARIS_B200_EINVAL;
        }
        bp_exact_kernel<<<grid, block, 0, ctx->compute>>>(d_first_slot, slot_floats, d_vol, g, a);
        PB_CUDA(cudaGetLastError());
        ++ctx->launches;
        ++ctx->bp_launches_exact;
        std::snprintf(ctx->bp_last_kernel, sizeof(ctx->bp_last_kernel), "bp_exact_kernel");
        return PARIS_B200_OK;
    }

    void preload_bp_tma_kernels();

    void preload_backprojection_kernels()
    {
        cudaFuncAttributes a{};
        (void)cudaFuncGetAttributes(&a, bp_exact_kernel);
        preload_bp_tma_kernels();
        (void)cudaGetLastError();
    }

    int launch_backproject(paris_b200_ctx* ctx, const float* d_stack, size_t slot_floats, uint32_t pitch,
                           uint32_t first, uint32_t count, const float* sn, const float* cs, const bp_target& t,
                           uint32_t layout)
    {
        if(count == 0)
            return PARIS_B200_OK;
        if(count > static_cast<uint32_t>(kMaxBatch))
        {
            set_error("batch of %u exceeds %d", count, kMaxBatch);
            return PARIS_B200_EINVAL;
        }
        const bp_geometry g = make_bp_geometry(t, pitch, layout);
        bp_angles a{};
        a.count = static_cast<int>(count);
        for(uint32_t i = 0; i < count; ++i)
        {
            a.sn[i] = sn[i];
            a.cs[i] = cs[i];
        }
        const float* d_first = d_stack + slot_floats * first;

        if(ctx->bp_kernel != 1)
        {
            bool handled = false;
            PB_TRY(launch_bp_tma(ctx, d_stack, slot_floats, first, g, a, t.d_vol, ctx->bp_kernel == 2, &handled));
            if(handled)
                return PARIS_B200_OK;
        }
        return launch_exact(ctx, d_first, slot_floats, g, a, t.d_vol);
    }
}
