// filter_small.cuh -- the original all-shared-memory radix-4 formulation of K1, kept for transform sizes
// below 256 points (detector rows of <= 64 samples) where the register-grouped kernel of filter.cu has
// too few threads per transform to be worth specialising.  Same mathematics as filter.cu, one pass per
// shared-memory round trip.
#pragma once

#include "common.cuh"
#include "device_math.cuh"

namespace pb
{
    constexpr int kStageRows = 8;     // detector rows per CTA in transposed-output mode

    // frequency index held at storage position p after the forward passes (mixed-radix digit reversal)
    template <int LOG2N>
    __device__ __forceinline__ int small_frequency_of_position(int p)
    {
        constexpr int N = 1 << LOG2N;
        int k = 0;
        int rem = p;
        #pragma unroll
        for(int m = LOG2N; m >= 2; m -= 2)
        {
            // pass with span M = 2^m: digit = rem / (M/4), weight N/M
            const int d = rem >> (m - 2);
            rem &= (1 << (m - 2)) - 1;
            k += d << (LOG2N - m);
        }
        if(LOG2N & 1)
            k += rem * (N / 2);
        return k;
    }

    template <int LOG2N>
    __device__ __forceinline__ void forward_passes(float2* x, const float2* __restrict__ tw, int tid, int nt)
    {
        constexpr int N = 1 << LOG2N;
        #pragma unroll 1
        for(int m = LOG2N; m >= 2; m -= 2)
        {
            const int q_log = m - 2;           // Q = M/4
            const int Q = 1 << q_log;
            const int tw_shift = LOG2N - m;     // twiddle stride N/M
            for(int b = tid; b < N / 4; b += nt)
            {
                const int j = b & (Q - 1);
                const int base = ((b >> q_log) << m) + j;
                const float2 a0 = x[base], a1 = x[base + Q], a2 = x[base + 2 * Q], a3 = x[base + 3 * Q];
                const float2 t0 = make_float2(a0.x + a2.x, a0.y + a2.y);
                const float2 t1 = make_float2(a0.x - a2.x, a0.y - a2.y);
                const float2 t2 = make_float2(a1.x + a3.x, a1.y + a3.y);
                // (a1 - a3) * (-i)
                const float2 t3 = make_float2(a1.y - a3.y, a3.x - a1.x);
                const float2 y0 = make_float2(t0.x + t2.x, t0.y + t2.y);
                const float2 y1 = make_float2(t1.x + t3.x, t1.y + t3.y);
                const float2 y2 = make_float2(t0.x - t2.x, t0.y - t2.y);
                const float2 y3 = make_float2(t1.x - t3.x, t1.y - t3.y);
                const int e = j << tw_shift;
                x[base] = y0;
                x[base + Q] = cmul(y1, __ldg(tw + e));
                x[base + 2 * Q] = cmul(y2, __ldg(tw + 2 * e));
                x[base + 3 * Q] = cmul(y3, __ldg(tw + 3 * e));
            }
            __syncthreads();
        }
        if(LOG2N & 1)
        {
            for(int b = tid; b < N / 2; b += nt)
            {
                const float2 a0 = x[2 * b], a1 = x[2 * b + 1];
                x[2 * b] = make_float2(a0.x + a1.x, a0.y + a1.y);
                x[2 * b + 1] = make_float2(a0.x - a1.x, a0.y - a1.y);
            }
            __syncthreads();
        }
    }

    template <int LOG2N>
    __device__ __forceinline__ void inverse_passes(float2* x, const float2* __restrict__ tw, int tid, int nt)
    {
        constexpr int N = 1 << LOG2N;
        if(LOG2N & 1)
        {
            for(int b = tid; b < N / 2; b += nt)
            {
                const float2 a0 = x[2 * b], a1 = x[2 * b + 1];
                x[2 * b] = make_float2(a0.x + a1.x, a0.y + a1.y);
                x[2 * b + 1] = make_float2(a0.x - a1.x, a0.y - a1.y);
            }
            __syncthreads();
        }
        #pragma unroll 1
        for(int m = 2 + (LOG2N & 1); m <= LOG2N; m += 2)
        {
            const int q_log = m - 2;
            const int Q = 1 << q_log;
            const int tw_shift = LOG2N - m;
            for(int b = tid; b < N / 4; b += nt)
            {
                const int j = b & (Q - 1);
                const int base = ((b >> q_log) << m) + j;
                const int e = j << tw_shift;
                const float2 b0 = x[base];
                const float2 b1 = cmul_conj(x[base + Q], __ldg(tw + e));
                const float2 b2 = cmul_conj(x[base + 2 * Q], __ldg(tw + 2 * e));
                const float2 b3 = cmul_conj(x[base + 3 * Q], __ldg(tw + 3 * e));
                const float2 t0 = make_float2(b0.x + b2.x, b0.y + b2.y);
                const float2 t1 = make_float2(b0.x - b2.x, b0.y - b2.y);
                const float2 t2 = make_float2(b1.x + b3.x, b1.y + b3.y);
                // (b1 - b3) * (+i)
                const float2 t3 = make_float2(b3.y - b1.y, b1.x - b3.x);
                x[base] = make_float2(t0.x + t2.x, t0.y + t2.y);
                x[base + Q] = make_float2(t1.x + t3.x, t1.y + t3.y);
                x[base + 2 * Q] = make_float2(t0.x - t2.x, t0.y - t2.y);
                x[base + 3 * Q] = make_float2(t1.x - t3.x, t1.y - t3.y);
            }
            __syncthreads();
        }
    }

    // One CTA filters PAIRS pairs of detector rows.  TRANSPOSED: results are staged in shared memory and
    // written as dst[s * dst_pitch + t] in 32-byte runs (8 consecutive t per sample s); otherwise each row is
    // written back in place / row-major right after its transform.
    template <int LOG2N, bool TRANSPOSED>
    __global__ void __launch_bounds__(((1 << LOG2N) / 4 > 1024) ? 1024 : (1 << LOG2N) / 4)
    filter_small_kernel(const float* src, float* dst, uint32_t dim_x, uint32_t dim_y,
                  const float* __restrict__ kn, const float2* __restrict__ tw, weight_params w, uint32_t dst_pitch,
                  uint32_t src_u16)
    {
        // 16-bit source samples (detector counts) are widened by the load, like the grouped kernel does
        const auto sample = [&](size_t at) -> float {
            return src_u16 ? static_cast<float>(__ldg(reinterpret_cast<const unsigned short*>(src) + at)) : __ldg(src + at);
        };
        constexpr int N = 1 << LOG2N;
        constexpr int PAIRS = TRANSPOSED ? kStageRows / 2 : 1;
        extern __shared__ __align__(16) unsigned char smem_raw[];
        float2* x = reinterpret_cast<float2*>(smem_raw);
        float* stage = reinterpret_cast<float*>(smem_raw + sizeof(float2) * N); // [dim_x][kStagePitch], TRANSPOSED only

        const int tid = threadIdx.x;
        const int nt = blockDim.x;
        const uint32_t row_base = blockIdx.x * (2u * PAIRS);

        #pragma unroll 1
        for(int pair = 0; pair < PAIRS; ++pair)
        {
            const uint32_t row0 = row_base + 2u * pair;
            const uint32_t row1 = row0 + 1u;
            const bool has0 = row0 < dim_y;
            const bool has1 = row1 < dim_y;
            if(!has0)
                break; // uniform for the CTA; staged rows >= dim_y are never written out

            // load + weight + zero-pad
            for(int i = tid; i < N; i += nt)
            {
                float a = 0.f, b = 0.f;
                if(static_cast<uint32_t>(i) < dim_x)
                {
                    if(has0)
                    {
                        a = sample(static_cast<size_t>(row0) * dim_x + i);
                        if(w.enable)
                            a = __fmul_rn(a, pixel_weight(i, row0, w.h_min, w.v_min, w.d_sd, w.l_px_row, w.l_px_col));
                    }
                    if(has1)
                    {
                        b = sample(static_cast<size_t>(row1) * dim_x + i);
                        if(w.enable)
                            b = __fmul_rn(b, pixel_weight(i, row1, w.h_min, w.v_min, w.d_sd, w.l_px_row, w.l_px_col));
                    }
                }
                x[i] = make_float2(a, b);
            }
            __syncthreads();

            forward_passes<LOG2N>(x, tw, tid, nt);

            // scale by K/N (real, even): position p holds frequency k
            for(int p = tid; p < N; p += nt)
            {
                const int k = small_frequency_of_position<LOG2N>(p);
                const float s = __ldg(kn + (k <= N / 2 ? k : N - k));
                float2 v = x[p];
                v.x *= s;
                v.y *= s;
                x[p] = v;
            }
            __syncthreads();

            inverse_passes<LOG2N>(x, tw, tid, nt);

            // keep the first dim_x samples
            if(TRANSPOSED)
            {
                for(int i = tid; i < static_cast<int>(dim_x); i += nt)
                    *reinterpret_cast<float2*>(stage + i * kStagePitch + 2 * pair) = x[i];
            }
            else
            {
                for(int i = tid; i < static_cast<int>(dim_x); i += nt)
                {
                    const float2 v = x[i];
                    dst[static_cast<size_t>(row0) * dim_x + i] = v.x;
                    if(has1)
                        dst[static_cast<size_t>(row1) * dim_x + i] = v.y;
                }
            }
            __syncthreads();
        }

        if(TRANSPOSED)
        {
            // 4 lanes cover the 8 staged rows of one sample: 32 contiguous bytes in the stack slot
            for(int e = tid; e < static_cast<int>(dim_x) * 4; e += nt)
            {
                const int i = e >> 2, q = e & 3;
                const uint32_t t = row_base + 2u * q;
                if(t < dim_y)
                {
                    float2 v = *reinterpret_cast<const float2*>(stage + i * kStagePitch + 2 * q);
                    if(t + 1u >= dim_y)
                        v.y = 0.f; // the slot's padding columns stay zero
                    *reinterpret_cast<float2*>(dst + static_cast<size_t>(i) * dst_pitch + t) = v;
                }
            }
        }
    }

    template <int LOG2N>
    static int launch_filter_small(paris_b200_ctx* ctx, const float* d_src, float* d_dst, uint32_t dim_x, uint32_t dim_y,
                               const paris_b200_filter* f, const weight_params& w, bool transposed, uint32_t pitch,
                               bool src_u16 = false)
    {
        constexpr int N = 1 << LOG2N;
        constexpr int threads = (N / 4 > 1024) ? 1024 : N / 4;
        if(transposed)
        {
            const size_t smem = sizeof(float2) * N + sizeof(float) * kStagePitch * dim_x;
            auto kern = filter_small_kernel<LOG2N, true>;
            PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
            const uint32_t grid = (dim_y + kStageRows - 1) / kStageRows;
            kern<<<grid, threads, smem, ctx->compute>>>(d_src, d_dst, dim_x, dim_y, f->d_kn, f->d_tw, w, pitch, src_u16 ? 1u : 0u);
        }
        else
        {
            const size_t smem = sizeof(float2) * N;
            auto kern = filter_small_kernel<LOG2N, false>;
            PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
            const uint32_t grid = (dim_y + 1) / 2;
            kern<<<grid, threads, smem, ctx->compute>>>(d_src, d_dst, dim_x, dim_y, f->d_kn, f->d_tw, w, 0u, src_u16 ? 1u : 0u);
        }
        PB_CUDA(cudaGetLastError());
        ++ctx->launches;
        return PARIS_B200_OK;
    }

    inline void preload_filter_small_kernels(uint32_t size)
    {
        cudaFuncAttributes a{};
        switch(size)
        {
            case 32: (void)cudaFuncGetAttributes(&a, filter_small_kernel<5, true>); (void)cudaFuncGetAttributes(&a, filter_small_kernel<5, false>); break;
            case 64: (void)cudaFuncGetAttributes(&a, filter_small_kernel<6, true>); (void)cudaFuncGetAttributes(&a, filter_small_kernel<6, false>); break;
            case 128: (void)cudaFuncGetAttributes(&a, filter_small_kernel<7, true>); (void)cudaFuncGetAttributes(&a, filter_small_kernel<7, false>); break;
            default: break;
        }
    }
}
