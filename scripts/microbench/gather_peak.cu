// gather_peak.cu -- the MEASURED ceiling of the backprojection kernel's shared-memory gather.
//
// The production kernel (paris_b200/csrc/backproject_tma.cu) is bound by the shared-memory data pipe: every voxel
// update reads four 4-byte samples with one LDS.32 each, 32 lanes of a warp along z.  The theoretical peak is one
// 128-byte wavefront per clock and SM; the pattern itself costs more than that wherever a warp's 32 rows span more
// than 32 banks (detector rows advance by dv = 1.75 .. 2.34 per slice in BASELINE configs 1-3, i.e. 0.87 .. 1.17
// words per lane in a parity plane of the split layout).  This program measures what the SM delivers for exactly
// that access pattern with (almost) nothing else in the instruction stream:
//
//   per "update": 4 LDS.32 (two parity planes x two adjacent staged columns, same address arithmetic and the same
//   box geometry as cfg_coarse_split_tall: 32 columns x 2 planes x 156 row pairs) + 4 FADD.
//   Addresses are computed once per (column, slice) outside the timed loop and advanced by a constant per pass.
//
// Variants: dv = 2.0 exactly (conflict-free: the hardware peak for this instruction mix), the c2/c3 distribution of dv
// (uniform over the tile columns' magnifications), and the worst case dv = 2.34.  256 threads, two CTAs per SM, 80 KB
// of staged data per CTA -- the production kernel's launch shape.  Output: one JSON object.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o gather_peak gather_peak.cu
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace
{
    constexpr int BH = 32, BV = 312, BVH = BV / 2, STAGES = 2;
    constexpr int CPW = 8, NZ = 4, THREADS = 256;
    constexpr int STAGE_FLOATS = BH * BV;

    __device__ __forceinline__ float lds_f32(uint32_t addr)
    {
        float v;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
        return v;
    }

    // dv_cols: rows per slice for each of the CTA's 64 columns (rows = detector rows; the planes hold row pairs)
    __global__ void __launch_bounds__(THREADS, 2)
    gather_kernel(const float* __restrict__ dv_cols, float* __restrict__ out, int passes)
    {
        extern __shared__ __align__(128) float stage[];
        for(int i = threadIdx.x; i < STAGES * STAGE_FLOATS; i += THREADS)
            stage[i] = static_cast<float>(i & 1023) * 1e-3f;
        __syncthreads();
        const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const uint32_t base0 = static_cast<uint32_t>(__cvta_generic_to_shared(stage));

        // byte addresses of the even-plane and odd-plane sample of (column i, slice lane + 32 j)
        uint32_t a_even[CPW][NZ], a_odd[CPW][NZ];
        #pragma unroll
        for(int i = 0; i < CPW; ++i)
        {
            const int col = warp * CPW + i;
            const float dv = dv_cols[col];
            const int x1 = (col * 5) % (BH - 2);             // staged column of this voxel column
            #pragma unroll
            for(int j = 0; j < NZ; ++j)
            {
                const float v = 3.3f + 0.1f * static_cast<float>(col) + dv * static_cast<float>(lane + 32u * j);
                const int r = static_cast<int>(floorf(v));   // rows r, r + 1: one even, one odd
                const int k = r >> 1, p = r & 1;
                a_odd[i][j] = base0 + 4u * static_cast<uint32_t>(x1 * BV + BVH + k);
                a_even[i][j] = base0 + 4u * static_cast<uint32_t>(x1 * BV + k + p);
            }
        }
        float acc[CPW][NZ];
        #pragma unroll
        for(int i = 0; i < CPW; ++i)
            #pragma unroll
            for(int j = 0; j < NZ; ++j)
                acc[i][j] = 0.f;

        #pragma unroll 1
        for(int it = 0; it < passes; ++it)
        {
            // (a different stage and a different 32-byte phase per pass, like successive projections)
            const uint32_t shift = static_cast<uint32_t>(it & 1) * (STAGE_FLOATS * 4u) + static_cast<uint32_t>((it >> 1) & 3) * 32u;
            #pragma unroll
            for(int i = 0; i < CPW; ++i)
            {
                #pragma unroll
                for(int j = 0; j < NZ; ++j)
                {
                    const uint32_t ae = a_even[i][j] + shift, ao = a_odd[i][j] + shift;
                    const float q11 = lds_f32(ae), q21 = lds_f32(ae + 4 * BV);
                    const float q12 = lds_f32(ao), q22 = lds_f32(ao + 4 * BV);
                    acc[i][j] += (q11 + q21) + (q12 + q22);
                }
            }
        }
        float s = 0.f;
        #pragma unroll
        for(int i = 0; i < CPW; ++i)
            #pragma unroll
            for(int j = 0; j < NZ; ++j)
                s += acc[i][j];
        out[blockIdx.x * THREADS + threadIdx.x] = s;
    }

    double run(const std::vector<float>& dv, int sms, int passes, float* d_out, double* clock_mhz)
    {
        float* d_dv = nullptr;
        cudaMalloc(&d_dv, dv.size() * sizeof(float));
        cudaMemcpy(d_dv, dv.data(), dv.size() * sizeof(float), cudaMemcpyHostToDevice);
        const size_t smem = sizeof(float) * STAGES * STAGE_FLOATS;
        cudaFuncSetAttribute(gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        const int grid = 2 * sms * 4;   // four waves of two CTAs per SM
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        for(int w = 0; w < 3; ++w)
            gather_kernel<<<grid, THREADS, smem>>>(d_dv, d_out, passes);
        cudaDeviceSynchronize();
        float best = 1e30f;
        for(int rep = 0; rep < 5; ++rep)
        {
            cudaEventRecord(e0);
            gather_kernel<<<grid, THREADS, smem>>>(d_dv, d_out, passes);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            best = std::fmin(best, ms);
        }
        int khz = 0;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
        *clock_mhz = khz / 1e3;
        cudaFree(d_dv);
        const double updates = static_cast<double>(grid) * THREADS * CPW * NZ * passes;
        return updates / (best * 1e-3) / 1e9;   // giga-updates per second (4 LDS.32 each)
    }
}

int main(int argc, char** argv)
{
    const int passes = argc > 1 ? std::atoi(argv[1]) : 2000;
    cudaDeviceProp prop{};
    if(cudaGetDeviceProperties(&prop, 0) != cudaSuccess)
    {
        std::fprintf(stderr, "no CUDA device\n");
        return 1;
    }
    const int sms = prop.multiProcessorCount;
    float* d_out = nullptr;
    cudaMalloc(&d_out, sizeof(float) * 2 * sms * 4 * THREADS);

    // dv per column.  BASELINE configs 1-3: d_so = d_od = 500 mm, voxel = 2 x the natural voxel, so
    // dv = l_vx_z * factor / l_px = factor = d_sd / (s + d_so) with s = x cos(phi) + y sin(phi) over the volume's
    // 101.9 mm square cross-section and all angles: the 64 columns take the 64 quantile midpoints of
    // that distribution.
    std::vector<float> exact(64, 2.0f), worst(64, 2.34f), best(64, 1.75f), dist(64);
    {
        std::vector<double> f;
        const int n = 96, na = 180;
        for(int a = 0; a < na; ++a)
        {
            const double phi = 2.0 * M_PI * (a + 0.5) / na;
            for(int iy = 0; iy < n; ++iy)
                for(int ix = 0; ix < n; ++ix)
                {
                    const double x = -50.93 + 101.86 * (ix + 0.5) / n, y = -50.93 + 101.86 * (iy + 0.5) / n;
                    f.push_back(1000.0 / (500.0 + x * std::cos(phi) + y * std::sin(phi)));
                }
        }
        std::sort(f.begin(), f.end());
        for(int c = 0; c < 64; ++c)
            dist[c] = static_cast<float>(f[static_cast<size_t>((c + 0.5) / 64.0 * f.size())]);
    }
    double mhz = 0.0;
    const double g_exact = run(exact, sms, passes, d_out, &mhz);
    const double g_best = run(best, sms, passes, d_out, &mhz);
    const double g_dist = run(dist, sms, passes, d_out, &mhz);
    const double g_worst = run(worst, sms, passes, d_out, &mhz);
    const cudaError_t e = cudaDeviceSynchronize();
    if(e != cudaSuccess)
    {
        std::fprintf(stderr, "CUDA error: %s\n", cudaGetErrorString(e));
        return 1;
    }
    std::printf("{\"device\": \"%s\", \"sms\": %d, \"clock_rate_mhz\": %.0f, \"passes\": %d, "
                "\"gather_gups\": {\"dv_2.00_conflict_free\": %.1f, \"dv_1.75\": %.1f, \"dv_c2c3_distribution\": %.1f, "
                "\"dv_2.34_worst\": %.1f}, "
                "\"note\": \"giga voxel-updates/s with 4 LDS.32 + 4 FADD per update, 256 threads x 2 CTAs/SM, addresses as "
                "in cfg_coarse_split_tall; theoretical pipe peak = sms x 8 updates/clk\"}\n",
                prop.name, sms, mhz, passes, g_exact, g_best, g_dist, g_worst);
    cudaFree(d_out);
    return 0;
}
