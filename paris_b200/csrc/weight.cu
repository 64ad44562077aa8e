// weight.cu -- stand-alone cosine pre-weighting (backend::weight, /root/reference/src/openmp/weighting.cpp:32-57)
// and the row-major -> stack-slot transpose.  In the pipeline proper the weighting is fused into the filter
// kernel (filter.cu); this kernel exists because the backend contract exposes weight() on its own.
#include "common.cuh"
#include "device_math.cuh"

namespace pb
{
    __global__ void __launch_bounds__(256)
    weight_kernel(float* __restrict__ p, uint32_t dim_x, uint32_t dim_y, float h_min, float v_min, float d_sd,
                  float l_px_row, float l_px_col)
    {
        const uint32_t t = blockIdx.y;
        for(uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < dim_x; s += gridDim.x * blockDim.x)
        {
            const size_t i = static_cast<size_t>(t) * dim_x + s;
            p[i] = __fmul_rn(p[i], pixel_weight(s, t, h_min, v_min, d_sd, l_px_row, l_px_col));
        }
    }

    int launch_weight(paris_b200_ctx* ctx, float* d_proj, uint32_t dim_x, uint32_t dim_y, float h_min, float v_min,
                      float d_sd, float l_px_row, float l_px_col)
    {
        const dim3 block(256);
        const dim3 grid((dim_x + 255u) / 256u, dim_y);
        weight_kernel<<<grid, block, 0, ctx->compute>>>(d_proj, dim_x, dim_y, h_min, v_min, d_sd, l_px_row, l_px_col);
        PB_CUDA(cudaGetLastError());
        ++ctx->launches;
        return PARIS_B200_OK;
    }

    // dst[s * pitch + line_offset(t)] = src[t * dim_x + s]   (32x32 tiles through shared memory, both sides coalesced)
    __global__ void __launch_bounds__(256)
    transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, uint32_t dim_x, uint32_t dim_y,
                     uint32_t pitch, uint32_t layout)
    {
        __shared__ float tile[32][33];
        const uint32_t s0 = blockIdx.x * 32u, t0 = blockIdx.y * 32u;
        for(uint32_t r = threadIdx.y; r < 32u; r += 8u)
        {
            const uint32_t s = s0 + threadIdx.x, t = t0 + r;
            tile[r][threadIdx.x] = (s < dim_x && t < dim_y) ? src[static_cast<size_t>(t) * dim_x + s] : 0.f;
        }
        __syncthreads();
        for(uint32_t r = threadIdx.y; r < 32u; r += 8u)
        {
            const uint32_t s = s0 + r, t = t0 + threadIdx.x;
            if(s < dim_x && t < dim_y)
                dst[static_cast<size_t>(s) * pitch + line_offset(t, pitch, layout)] = tile[threadIdx.x][r];
        }
    }

    int launch_transpose_to_slot(paris_b200_ctx* ctx, const float* d_src, float* d_slot, uint32_t dim_x,
                                 uint32_t dim_y, uint32_t pitch, uint32_t layout)
    {
        const dim3 block(32, 8);
        const dim3 grid((dim_x + 31u) / 32u, (dim_y + 31u) / 32u);
        transpose_kernel<<<grid, block, 0, ctx->compute>>>(d_src, d_slot, dim_x, dim_y, pitch, layout);
        PB_CUDA(cudaGetLastError());
        ++ctx->launches;
        return PARIS_B200_OK;
    }
}
