// filter.cu -- K1: cosine weighting fused with the ramp filter, one pass over HBM.
//
// Replaces backend::weight + backend::apply_filter of the reference
// (/root/reference/src/openmp/weighting.cpp:32-57, src/openmp/filtering.cpp:167-219; legacy CUDA chain
// src/cuda/filtering.cu:145-256: fill, expand-copy, cuFFT R2C, multiply, cuFFT C2R, shrink-copy, normalise =
// ~88 B of HBM traffic per pixel).  Here every detector row is read once and written once (8 B per pixel):
//
//   * two detector rows a, b are packed as ONE complex sequence z = a + i b.  The filter K is real and
//     even (src/openmp/filtering.cpp:155-161 stores tau*|R| in re and im), so filtering z filters both
//     rows at once and no Hermitian untangling pass is needed;
//   * zero padding to N = filter_size is implicit: the padded half is a compile-time zero in the first
//     butterfly group and the discarded half of the output is never computed in the last one;
//   * the transform is a chain of radix-4 decimation-in-frequency passes (natural order in, digit-reversed
//     out), K/N applied in digit-reversed order, then decimation-in-time passes back (digit-reversed in,
//     natural out) -- no reordering pass.  Passes are executed in GROUPS held in registers: every thread
//     owns 16 points and runs two radix-4 passes on them between shared-memory exchanges (an effective
//     radix-16 step); the last forward passes, the multiplication by K and the first inverse passes are
//     one register-resident step.  N = 2048 and 4096 need 4 exchanges in total;
//   * the first group reads its points straight from global memory (weighted on the fly), the last one
//     writes straight to global memory or, for the filtered stack, to a staging tile so that the
//     TRANSPOSED slot (detector-row index fastest) is written in 32-byte runs;
//   * one launch covers a whole batch of projections (blockIdx.y), 8 detector rows per CTA.
//
// No cuFFT, and the weighted / padded / transformed row never exists in HBM.
#include "common.cuh"
#include "device_math.cuh"

namespace pb
{
    constexpr int kRowsPerCta = 8;    // detector rows (4 row pairs) per CTA
    constexpr int kStagePitch = 10;   // floats per staged sample line: 8 rows + 2 pad (conflict-free float2 stores)
}

#include "filter_small.cuh"

namespace pb
{
    // shared-memory index of point i: one float2 of padding per 16 points keeps both the strided group
    // accesses and the contiguous 16-point accesses of the middle step conflict-free
    __host__ __device__ constexpr int pad(int i) { return i + (i >> 4); }

    // frequency index held at storage position p after the forward passes (mixed-radix digit reversal:
    // radix-4 passes with spans N, N/4, ... then one radix-2 pass if log2 N is odd)
    int frequency_of_position(int log2n, int p)
    {
        int k = 0;
        int rem = p;
        for(int m = log2n; m >= 2; m -= 2)
        {
            const int d = rem >> (m - 2);
            rem &= (1 << (m - 2)) - 1;
            k += d << (log2n - m);
        }
        if(log2n & 1)
            k += rem << (log2n - 1);
        return k;
    }

    // ---- radix-4 butterflies: output q carries X[4k+q] of the sub-transform -------------------------------------

    __device__ __forceinline__ void dif4(float2& a0, float2& a1, float2& a2, float2& a3)
    {
        const float2 t0 = make_float2(a0.x + a2.x, a0.y + a2.y);
        const float2 t1 = make_float2(a0.x - a2.x, a0.y - a2.y);
        const float2 t2 = make_float2(a1.x + a3.x, a1.y + a3.y);
        const float2 t3 = make_float2(a1.y - a3.y, a3.x - a1.x); // (a1 - a3) * (-i)
        a0 = make_float2(t0.x + t2.x, t0.y + t2.y);
        a1 = make_float2(t1.x + t3.x, t1.y + t3.y);
        a2 = make_float2(t0.x - t2.x, t0.y - t2.y);
        a3 = make_float2(t1.x - t3.x, t1.y - t3.y);
    }

    __device__ __forceinline__ void dit4(float2& b0, float2& b1, float2& b2, float2& b3)
    {
        const float2 t0 = make_float2(b0.x + b2.x, b0.y + b2.y);
        const float2 t1 = make_float2(b0.x - b2.x, b0.y - b2.y);
        const float2 t2 = make_float2(b1.x + b3.x, b1.y + b3.y);
        const float2 t3 = make_float2(b3.y - b1.y, b1.x - b3.x); // (b1 - b3) * (+i)
        b0 = make_float2(t0.x + t2.x, t0.y + t2.y);
        b1 = make_float2(t1.x + t3.x, t1.y + t3.y);
        b2 = make_float2(t0.x - t2.x, t0.y - t2.y);
        b3 = make_float2(t1.x - t3.x, t1.y - t3.y);
    }

    __device__ __forceinline__ void bfly2(float2& a0, float2& a1)
    {
        const float2 s = make_float2(a0.x + a1.x, a0.y + a1.y);
        a1 = make_float2(a0.x - a1.x, a0.y - a1.y);
        a0 = s;
    }

    // One radix-4 pass of span 2^M on 4 register points whose position inside the quarter is j:
    // forward = butterfly, then twiddle W_span^(q j) on output q; inverse = conjugate twiddle, then butterfly.
    // `twc` holds one COMPACT table per span, W_(2^m)^i for i < 3*2^m/4 at offset 3*(2^m - 8)/4, so that lanes
    // with consecutive j read consecutive entries (a strided walk through one table of W_N costs up to 32
    // cache lines per load instruction).  The kernel keeps the tables in shared memory: a miss-free 29-cycle load
    // instead of an L1/L2 round trip in front of every butterfly group.
    template <int LOG2N, int M, bool INVERSE>
    __device__ __forceinline__ void pass4(float2& e0, float2& e1, float2& e2, float2& e3, int j,
                                          const float2* __restrict__ twc)
    {
        const float2* const t = twc + 3 * ((1 << M) - 8) / 4;
        if(!INVERSE)
        {
            dif4(e0, e1, e2, e3);
            if(M > 2)
            {
                e1 = cmul(e1, t[j]);
                e2 = cmul(e2, t[2 * j]);
                e3 = cmul(e3, t[3 * j]);
            }
        }
        else
        {
            if(M > 2)
            {
                e1 = cmul_conj(e1, t[j]);
                e2 = cmul_conj(e2, t[2 * j]);
                e3 = cmul_conj(e3, t[3 * j]);
            }
            dit4(e0, e1, e2, e3);
        }
    }

    // radix-4 pass with the three twiddles given (see pass4)
    template <bool INVERSE>
    __device__ __forceinline__ void pass4w(float2& e0, float2& e1, float2& e2, float2& e3, float2 w1, float2 w2, float2 w3)
    {
        if(!INVERSE)
        {
            dif4(e0, e1, e2, e3);
            e1 = cmul(e1, w1);
            e2 = cmul(e2, w2);
            e3 = cmul(e3, w3);
        }
        else
        {
            e1 = cmul_conj(e1, w1);
            e2 = cmul_conj(e2, w2);
            e3 = cmul_conj(e3, w3);
            dit4(e0, e1, e2, e3);
        }
    }

    // exp(-2 pi i k / 16)
    __device__ __forceinline__ constexpr float2 root16(int k)
    {
        constexpr float c[16] = {1.f, 0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f, 0.f,
                                 -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f, -1.f,
                                 -0.92387953251128674f, -0.70710678118654752f, -0.38268343236508977f, 0.f,
                                 0.38268343236508977f, 0.70710678118654752f, 0.92387953251128674f};
        constexpr float sn[16] = {0.f, 0.38268343236508977f, 0.70710678118654752f, 0.92387953251128674f, 1.f,
                                  0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f, 0.f,
                                  -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f, -1.f,
                                  -0.92387953251128674f, -0.70710678118654752f, -0.38268343236508977f};
        return float2{c[k & 15], -sn[k & 15]};
    }

    // Two consecutive passes (spans 2^M and 2^(M-2)) on the 16 points e[k] = x[block*2^M + j + k*2^(M-4)].
    // Twiddles: butterfly r of the first pass sits at position j + r*Q2 of its quarter, Q2 = 2^(M-4), so
    //   W_(2^M)^(q (j + r Q2)) = W_(2^M)^(q j) * exp(-2 pi i q r / 16):
    // ONE loaded triple W^j, W^2j, W^3j times compile-time 16th roots gives all twelve twiddles of the pass, and the
    // second pass uses one triple for its four butterflies -- 6 shared-memory loads per group instead of 15 (the kernel
    // is bound by its shared-memory pipe: the multiplications by constants are cheaper than the loads they replace).
    template <int LOG2N, int M, bool INVERSE>
    __device__ __forceinline__ void group16(float2 (&e)[16], int j, const float2* __restrict__ tw)
    {
        static_assert(M >= 7, "both passes of a double group carry twiddles");
        const float2* const t1 = tw + 3 * ((1 << M) - 8) / 4;
        const float2* const t2 = tw + 3 * ((1 << (M - 2)) - 8) / 4;
        const float2 a = t1[j], b = t1[2 * j], c = t1[3 * j];
        const float2 a2 = t2[j], b2 = t2[2 * j], c2 = t2[3 * j];
        if(!INVERSE)
        {
            #pragma unroll
            for(int r = 0; r < 4; ++r)
                pass4w<false>(e[r], e[r + 4], e[r + 8], e[r + 12], r == 0 ? a : cmul(a, root16(r)),
                              r == 0 ? b : cmul(b, root16(2 * r)), r == 0 ? c : cmul(c, root16(3 * r)));
            #pragma unroll
            for(int q = 0; q < 4; ++q)
                pass4w<false>(e[4 * q], e[4 * q + 1], e[4 * q + 2], e[4 * q + 3], a2, b2, c2);
        }
        else
        {
            #pragma unroll
            for(int q = 0; q < 4; ++q)
                pass4w<true>(e[4 * q], e[4 * q + 1], e[4 * q + 2], e[4 * q + 3], a2, b2, c2);
            #pragma unroll
            for(int r = 0; r < 4; ++r)
                pass4w<true>(e[r], e[r + 4], e[r + 8], e[r + 12], r == 0 ? a : cmul(a, root16(r)),
                             r == 0 ? b : cmul(b, root16(2 * r)), r == 0 ? c : cmul(c, root16(3 * r)));
        }
    }

    // Shared-memory positions of a thread's 16 points for the double group at span 2^M: padded position of
    // point 0, and the COMPILE-TIME offset of point k from it.  pad(P0 + k*Q2) = pad(P0) + k*Q2 + (k*Q2 >> 4)
    // holds for Q2 >= 16 trivially and for Q2 = 8 because the point's index inside its block of 8 never carries
    // into bit 4 -- so every access of a group is `base + immediate`, no per-element address arithmetic.
    template <int M>
    __device__ __forceinline__ int group_base(int b)
    {
        constexpr int Q2LOG = M - 4;
        static_assert(Q2LOG >= 3, "the padded offsets of a group are constants only for strides >= 8");
        return pad(((b >> Q2LOG) << M) + (b & ((1 << Q2LOG) - 1)));
    }

    template <int M>
    __host__ __device__ constexpr int group_off(int k)
    {
        return (k << (M - 4)) + ((k << (M - 4)) >> 4);
    }

    // A lone radix-4 pass of span 2^M: thread b runs the butterflies b, b+T, b+2T, b+3T (T threads per transform);
    // butterfly t occupies e[4t .. 4t+3].
    template <int LOG2N, int M, bool INVERSE>
    __device__ __forceinline__ void lone_group(float2* x, int b, const float2* __restrict__ tw)
    {
        constexpr int T = (1 << LOG2N) / 16;
        constexpr int QLOG = M - 2;
        float2 e[16];
        int base[4], j[4];
        #pragma unroll
        for(int t = 0; t < 4; ++t)
        {
            const int bb = b + t * T;
            j[t] = bb & ((1 << QLOG) - 1);
            base[t] = pad(((bb >> QLOG) << M) + j[t]);   // (offsets of the other three points are constants, as above)
            #pragma unroll
            for(int q = 0; q < 4; ++q)
                e[4 * t + q] = x[base[t] + (q << QLOG) + ((q << QLOG) >> 4)];
        }
        #pragma unroll
        for(int t = 0; t < 4; ++t)
            pass4<LOG2N, M, INVERSE>(e[4 * t], e[4 * t + 1], e[4 * t + 2], e[4 * t + 3], j[t], tw);
        #pragma unroll
        for(int t = 0; t < 4; ++t)
        {
            #pragma unroll
            for(int q = 0; q < 4; ++q)
                x[base[t] + (q << QLOG) + ((q << QLOG) >> 4)] = e[4 * t + q];
        }
    }

    // The register-resident middle: the passes with span <= 16 on the 16 CONTIGUOUS points 16b..16b+15,
    // forward, then K/N (table permuted to storage order on the host), then the same passes backwards.
    template <int LOG2N>
    __device__ __forceinline__ void middle_group(float2* x, const float* __restrict__ knp, int b,
                                                 const float2* __restrict__ tw)
    {
        float2 e[16];
        #pragma unroll
        for(int k = 0; k < 16; ++k)
            e[k] = x[17 * b + k];   // pad(16 b + k) = 17 b + k

        if(LOG2N % 2 == 0)
        {
            #pragma unroll
            for(int r = 0; r < 4; ++r)
                pass4<LOG2N, 4, false>(e[r], e[r + 4], e[r + 8], e[r + 12], r, tw);
            #pragma unroll
            for(int q = 0; q < 4; ++q)
                pass4<LOG2N, 2, false>(e[4 * q], e[4 * q + 1], e[4 * q + 2], e[4 * q + 3], 0, tw);
        }
        else
        {
            #pragma unroll
            for(int u = 0; u < 2; ++u)
            {
                #pragma unroll
                for(int r = 0; r < 2; ++r)
                    pass4<LOG2N, 3, false>(e[8 * u + r], e[8 * u + r + 2], e[8 * u + r + 4], e[8 * u + r + 6], r, tw);
            }
            #pragma unroll
            for(int p = 0; p < 8; ++p)
                bfly2(e[2 * p], e[2 * p + 1]);
        }

        const float4* k4 = reinterpret_cast<const float4*>(knp + 16 * b);
        #pragma unroll
        for(int v = 0; v < 4; ++v)
        {
            const float4 k = __ldg(k4 + v);
            e[4 * v + 0].x *= k.x; e[4 * v + 0].y *= k.x;
            e[4 * v + 1].x *= k.y; e[4 * v + 1].y *= k.y;
            e[4 * v + 2].x *= k.z; e[4 * v + 2].y *= k.z;
            e[4 * v + 3].x *= k.w; e[4 * v + 3].y *= k.w;
        }

        if(LOG2N % 2 == 0)
        {
            #pragma unroll
            for(int q = 0; q < 4; ++q)
                pass4<LOG2N, 2, true>(e[4 * q], e[4 * q + 1], e[4 * q + 2], e[4 * q + 3], 0, tw);
            #pragma unroll
            for(int r = 0; r < 4; ++r)
                pass4<LOG2N, 4, true>(e[r], e[r + 4], e[r + 8], e[r + 12], r, tw);
        }
        else
        {
            #pragma unroll
            for(int p = 0; p < 8; ++p)
                bfly2(e[2 * p], e[2 * p + 1]);
            #pragma unroll
            for(int u = 0; u < 2; ++u)
            {
                #pragma unroll
                for(int r = 0; r < 2; ++r)
                    pass4<LOG2N, 3, true>(e[8 * u + r], e[8 * u + r + 2], e[8 * u + r + 4], e[8 * u + r + 6], r, tw);
            }
        }

        #pragma unroll
        for(int k = 0; k < 16; ++k)
            x[17 * b + k] = e[k];
    }

    // ---- group plan ------------------------------------------------------------------------------------------
    // Passes with span > 16 ("outer"), first to last: 2^LOG2N, 2^(LOG2N-2), ... down to 2^6 (LOG2N even) or
    // 2^5 (odd).  They are taken two at a time from the top; an odd one out is a lone radix-4 group.
    // WIDE: twice the row pairs side by side (512 threads).  Only the 4096-point transform uses it: its 256-thread
    // CTA needs 165 KB of shared memory (twiddles, one transform, the staging tile), so ONE CTA = 8 warps lives on
    // an SM and the kernel waits on latencies; two transforms in one 512-thread CTA share the twiddles and the
    // staging tile (181 KB) and give the schedulers 16 warps.
    template <int LOG2N, bool WIDE = false>
    struct plan
    {
        static_assert(LOG2N >= 8 && LOG2N <= 13, "register-grouped kernel covers 256..8192 points");
        static_assert(true, "lone groups have span 32 or 64: stride 8 or 16, same constant-offset argument");
        static constexpr int N = 1 << LOG2N;
        static constexpr int T = N / 16;                                  // threads per transform
        static constexpr int LOW = (LOG2N % 2 == 0) ? 6 : 5;
        static constexpr int OUTER = (LOG2N - LOW) / 2 + 1;               // 2 (256) .. 5 (8192)
        static constexpr int DOUBLES = OUTER / 2;                         // 1 or 2
        static constexpr bool LONE = (OUTER % 2) != 0;
        static constexpr int NARROW = (256 / T) < 4 ? ((256 / T) < 1 ? 1 : 256 / T) : 4;
        static constexpr int PAIRS = (WIDE && NARROW * T == 256 && NARROW < 4) ? 2 * NARROW : NARROW;  // row pairs side by side (<= 256 threads: two CTAs per SM in different phases)
        static constexpr int ROUNDS = 4 / PAIRS;
        static constexpr int THREADS = PAIRS * T;
        static constexpr int NPAD = pad(N);
        static constexpr bool TW_SMEM = LOG2N <= 12;                      // 8192 points: the tables stay in global memory
    };

    struct filter_batch
    {
        const float* src[kMaxBatch];   // raw projections: float samples, or (src_u16) detector-native 16-bit counts
        float* dst[kMaxBatch];         // row-major destinations (TRANSPOSED == false)
        uint32_t src_u16;              // the HIS reader widens u16 -> float on the host (src/his.cpp:169-185); here the
                                       // widening is the kernel's first load, so that half the bytes cross PCIe and HBM
    };

    template <int LOG2N, bool TRANSPOSED, bool WIDE = false>
    __global__ void __launch_bounds__((plan<LOG2N, WIDE>::THREADS), (plan<LOG2N, WIDE>::THREADS <= 256 ? 2 : 1))
    filter_kernel(const filter_batch io, float* dst_stack, uint32_t first_slot, size_t slot_floats, uint32_t dim_x,
                  uint32_t dim_y, uint32_t n_proj, const float* __restrict__ knp, const float2* __restrict__ tw,
                  weight_params w, uint32_t dst_pitch, uint32_t layout)
    {
        using P = plan<LOG2N, WIDE>;
        constexpr int N = P::N;
        constexpr int TWC = P::TW_SMEM ? 3 * (2 * N - 8) / 4 : 0;   // entries of the compact twiddle tables kept on chip
        extern __shared__ __align__(16) unsigned char smem_raw[];
        const float2* const twc = P::TW_SMEM ? reinterpret_cast<const float2*>(smem_raw) : tw;
        unsigned char* const work = smem_raw + sizeof(float2) * TWC;
        float* stage = reinterpret_cast<float*>(work + sizeof(float2) * P::NPAD * P::PAIRS); // [dim_x][kStagePitch]

        const int lp = threadIdx.x / P::T;                 // row pair slot inside this round
        const int b = threadIdx.x % P::T;                  // thread index inside the transform
        float2* x = reinterpret_cast<float2*>(work) + lp * P::NPAD;

        // persistent CTA: the twiddle tables are fetched once, then the CTA walks over its (projection, 8-row
        // block) work items
        for(int i = threadIdx.x; i < TWC; i += P::THREADS)
            reinterpret_cast<float2*>(smem_raw)[i] = __ldg(tw + i);
        __syncthreads();

        const uint32_t blocks_per_proj = (dim_y + kRowsPerCta - 1) / kRowsPerCta;
        const uint32_t n_items = blocks_per_proj * n_proj;

        // Raw samples of the NEXT round travel in registers: the loads are issued right after the current
        // round's first butterfly group and are consumed one round later, so their HBM latency is covered
        // by four shared-memory phases of arithmetic instead of stalling the warp at the top of every round.
        float pa[8], pc[8];
        auto prefetch = [&](uint32_t item, int round) {
            const uint32_t proj = item / blocks_per_proj;
            const uint32_t row0 = (item % blocks_per_proj) * kRowsPerCta + 2u * static_cast<uint32_t>(round * P::PAIRS + lp);
            const bool has0 = row0 < dim_y, has1 = row0 + 1u < dim_y;
            if(io.src_u16)
            {
                const unsigned short* const s0 = reinterpret_cast<const unsigned short*>(io.src[proj])
                                               + static_cast<size_t>(row0) * dim_x + b;
                const unsigned short* const s1 = s0 + dim_x;
                #pragma unroll
                for(int k = 0; k < 8; ++k)
                {
                    const uint32_t i = b + k * (N / 16);
                    // (the counts stay integers in the prefetch registers: converting here would make the warp wait
                    // for its loads right away instead of one round later, where they are used)
                    pa[k] = __uint_as_float((i < dim_x && has0) ? static_cast<uint32_t>(__ldg(s0 + k * (N / 16))) : 0u);
                    pc[k] = __uint_as_float((i < dim_x && has1) ? static_cast<uint32_t>(__ldg(s1 + k * (N / 16))) : 0u);
                }
                return;
            }
            const float* const src0 = io.src[proj] + static_cast<size_t>(row0) * dim_x + b;
            const float* const src1 = src0 + dim_x;
            #pragma unroll
            for(int k = 0; k < 8; ++k)
            {
                const uint32_t i = b + k * (N / 16);
                pa[k] = (i < dim_x && has0) ? __ldg(src0 + k * (N / 16)) : 0.f;
                pc[k] = (i < dim_x && has1) ? __ldg(src1 + k * (N / 16)) : 0.f;
            }
        };
        if(blockIdx.x < n_items)
            prefetch(blockIdx.x, 0);

        #pragma unroll 1
        for(uint32_t item = blockIdx.x; item < n_items; item += gridDim.x)
        {
        const uint32_t proj = item / blocks_per_proj;
        const uint32_t row_base = (item % blocks_per_proj) * kRowsPerCta;

        #pragma unroll 1
        for(int round = 0; round < P::ROUNDS; ++round)
        {
            const int pair = round * P::PAIRS + lp;        // 0..3 within the CTA's 8 rows
            const uint32_t row0 = row_base + 2u * pair, row1 = row0 + 1u;
            const bool has0 = row0 < dim_y, has1 = row1 < dim_y;

            float2 e[16];
            // ---- first group: points b + k*N/16 (prefetched); k >= 8 is the zero padding -----------------------------
            // cosine weight d_sd / sqrt(d_sd^2 + h_s^2 + v_t^2) (src/openmp/weighting.cpp:44-52) with one rsqrt per
            // pixel; the stand-alone weight kernel keeps the reference's exact operation sequence, here a 1-ulp
            // difference disappears in the float32 transform that follows
            const float v0 = fmaf(static_cast<float>(row0), w.l_px_col, 0.5f * w.l_px_col + w.v_min);
            const float v1 = v0 + w.l_px_col;
            const float c0 = fmaf(v0, v0, w.d_sd * w.d_sd), c1 = fmaf(v1, v1, w.d_sd * w.d_sd);
            #pragma unroll
            for(int k = 0; k < 16; ++k)
            {
                float a = 0.f, c = 0.f;
                if(k < 8)
                {
                    a = pa[k];
                    c = pc[k];
                    if(io.src_u16)
                    {
                        // src/his.cpp:98-99: the detector's 16-bit count as a float
                        a = static_cast<float>(__float_as_uint(a));
                        c = static_cast<float>(__float_as_uint(c));
                    }
                    if(w.enable)
                    {
                        const uint32_t i = b + k * (N / 16);
                        const float hs = fmaf(static_cast<float>(i), w.l_px_row, 0.5f * w.l_px_row + w.h_min);
                        a *= w.d_sd * rsqrtf(fmaf(hs, hs, c0));
                        c *= w.d_sd * rsqrtf(fmaf(hs, hs, c1));
                    }
                }
                e[k] = make_float2(a, c);
            }
            {
                // next round of this item, or the first round of this CTA's next item
                const bool last_round = round + 1 == P::ROUNDS;
                const uint32_t nitem = last_round ? item + gridDim.x : item;
                if(nitem < n_items)
                    prefetch(nitem, last_round ? 0 : round + 1);
            }
            group16<LOG2N, LOG2N, false>(e, b, twc);
            float2* const x1 = x + group_base<LOG2N>(b);
            #pragma unroll
            for(int k = 0; k < 16; ++k)
                x1[group_off<LOG2N>(k)] = e[k];
            __syncthreads();

            // ---- remaining outer groups, forward -------------------------------------------------------------------
            if constexpr(P::DOUBLES == 2)
            {
                constexpr int M = LOG2N - 4;
                float2* const x2 = x + group_base<M>(b);
                #pragma unroll
                for(int k = 0; k < 16; ++k)
                    e[k] = x2[group_off<M>(k)];
                group16<LOG2N, M, false>(e, b & ((1 << (M - 4)) - 1), twc);
                #pragma unroll
                for(int k = 0; k < 16; ++k)
                    x2[group_off<M>(k)] = e[k];
                __syncthreads();
            }
            if constexpr(P::LONE)
            {
                lone_group<LOG2N, P::LOW, false>(x, b, twc);
                __syncthreads();
            }

            // ---- middle: last forward passes, K/N, first inverse passes ------------------------------------------------
            middle_group<LOG2N>(x, knp, b, twc);
            __syncthreads();

            // ---- outer groups, inverse --------------------------------------------------------------------------------
            if constexpr(P::LONE)
            {
                lone_group<LOG2N, P::LOW, true>(x, b, twc);
                __syncthreads();
            }
            if constexpr(P::DOUBLES == 2)
            {
                constexpr int M = LOG2N - 4;
                float2* const x2 = x + group_base<M>(b);
                #pragma unroll
                for(int k = 0; k < 16; ++k)
                    e[k] = x2[group_off<M>(k)];
                group16<LOG2N, M, true>(e, b & ((1 << (M - 4)) - 1), twc);
                #pragma unroll
                for(int k = 0; k < 16; ++k)
                    x2[group_off<M>(k)] = e[k];
                __syncthreads();
            }
            #pragma unroll
            for(int k = 0; k < 16; ++k)
                e[k] = x1[group_off<LOG2N>(k)];
            group16<LOG2N, LOG2N, true>(e, b, twc);

            // ---- keep the first dim_x samples (k < 8; the rest is never computed) ------------------------------------------
            if(TRANSPOSED)
            {
                #pragma unroll
                for(int k = 0; k < 8; ++k)
                {
                    const uint32_t i = b + k * (N / 16);
                    if(i < dim_x)
                        *reinterpret_cast<float2*>(stage + i * kStagePitch + 2 * pair) = e[k];
                }
            }
            else
            {
                float* dst = io.dst[proj];
                #pragma unroll
                for(int k = 0; k < 8; ++k)
                {
                    const uint32_t i = b + k * (N / 16);
                    if(i < dim_x)
                    {
                        if(has0)
                            dst[static_cast<size_t>(row0) * dim_x + i] = e[k].x;
                        if(has1)
                            dst[static_cast<size_t>(row1) * dim_x + i] = e[k].y;
                    }
                }
            }
            __syncthreads(); // x is reused by the next round; the stage tile is complete after the last one
        }

        if(TRANSPOSED)
        {
            float* dst = dst_stack + slot_floats * (first_slot + proj);
            if(layout == kLayoutPlain)
            {
                // 4 lanes cover the 8 staged rows of one sample: 32 contiguous bytes in the stack slot
                for(int el = threadIdx.x; el < static_cast<int>(dim_x) * 4; el += P::THREADS)
                {
                    const int i = el >> 2, q = el & 3;
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
            else
            {
                // split2: the 4 even rows of a sample are 16 contiguous bytes in the even plane, the 4 odd rows
                // 16 contiguous bytes in the odd plane (row_base is a multiple of 8)
                for(int el = threadIdx.x; el < static_cast<int>(dim_x) * 2; el += P::THREADS)
                {
                    const int i = el >> 1, parity = el & 1;
                    if(row_base + parity < dim_y)
                    {
                        float4 v;
                        const float* st = stage + i * kStagePitch + parity;
                        v.x = st[0];
                        v.y = row_base + 2u + parity < dim_y ? st[2] : 0.f;
                        v.z = row_base + 4u + parity < dim_y ? st[4] : 0.f;
                        v.w = row_base + 6u + parity < dim_y ? st[6] : 0.f;
                        *reinterpret_cast<float4*>(dst + static_cast<size_t>(i) * dst_pitch + parity * (dst_pitch >> 1)
                                                   + (row_base >> 1)) = v;
                    }
                }
            }
        }
        __syncthreads(); // the staging tile and the transform buffers are reused by the next work item
        }
    }

    template <int LOG2N, bool WIDE = false>
    static int launch_grouped(paris_b200_ctx* ctx, const filter_batch& io, uint32_t count, float* d_stack,
                              uint32_t first_slot, size_t slot_floats, uint32_t dim_x, uint32_t dim_y,
                              const paris_b200_filter* f, const weight_params& w, bool transposed, uint32_t pitch,
                              uint32_t layout)
    {
        using P = plan<LOG2N, WIDE>;
        constexpr size_t twc_bytes = P::TW_SMEM ? sizeof(float2) * (3 * (2 * P::N - 8) / 4) : 0;
        const size_t smem = twc_bytes + sizeof(float2) * P::NPAD * P::PAIRS
                          + (transposed ? sizeof(float) * kStagePitch * dim_x : 0);
        if(smem > 227u * 1024u)
        {
            set_error("a detector row of %u samples needs %zu bytes of shared memory in the fused filter kernel (transform of "
                      "%d points, staging tile for the transposed slot): more than one SM has", dim_x, smem, P::N);
            return PARIS_B200_EINVAL;
        }
        const uint32_t items = ((dim_y + kRowsPerCta - 1) / kRowsPerCta) * count;
        // persistent grid: as many CTAs as can be resident (two per SM when threads/registers/shared memory allow)
        const uint32_t per_sm = (P::THREADS <= 256 && 2 * smem <= 220u * 1024u) ? 2u : 1u;
        const uint32_t grid = std::min<uint32_t>(items, per_sm * static_cast<uint32_t>(ctx->sm_count));
        if(transposed)
        {
            auto kern = filter_kernel<LOG2N, true, WIDE>;
            PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
            kern<<<grid, P::THREADS, smem, ctx->compute>>>(io, d_stack, first_slot, slot_floats, dim_x, dim_y, count, f->d_knp,
                                                          f->d_twc, w, pitch, layout);
        }
        else
        {
            auto kern = filter_kernel<LOG2N, false, WIDE>;
            PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
            kern<<<grid, P::THREADS, smem, ctx->compute>>>(io, nullptr, 0u, 0, dim_x, dim_y, count, f->d_knp, f->d_twc, w, 0u,
                                                          kLayoutPlain);
        }
        PB_CUDA(cudaGetLastError());
        ++ctx->launches;
        return PARIS_B200_OK;
    }

    template <int LOG2N>
    static void preload_grouped()
    {
        cudaFuncAttributes a{};
        (void)cudaFuncGetAttributes(&a, filter_kernel<LOG2N, true>);
        (void)cudaFuncGetAttributes(&a, filter_kernel<LOG2N, false>);
    }

    // the transposed-output kernel of this filter size (what a group step launches), loaded before the step
    void preload_filter_kernels(uint32_t size)
    {
        switch(size)
        {
            case 256: preload_grouped<8>(); break;
            case 512: preload_grouped<9>(); break;
            case 1024: preload_grouped<10>(); break;
            case 2048: preload_grouped<11>(); break;
            case 4096:
            {
                preload_grouped<12>();
                cudaFuncAttributes a{};
                (void)cudaFuncGetAttributes(&a, filter_kernel<12, true, true>);
                (void)cudaFuncGetAttributes(&a, filter_kernel<12, false, true>);
                break;
            }
            case 8192: preload_grouped<13>(); break;
            default: preload_filter_small_kernels(size); break;
        }
        (void)cudaGetLastError();
    }

    // `count` projections in one launch.  Transposed: src[i] -> slot first_slot + i of d_stack.
    // Row-major: src[i] -> dst[i] (may alias).
    int launch_filter_batch(paris_b200_ctx* ctx, const float* const* d_src, float* const* d_dst, uint32_t count,
                            float* d_stack, uint32_t first_slot, size_t slot_floats, uint32_t dim_x, uint32_t dim_y,
                            const paris_b200_filter* f, const weight_params& w, bool transposed, uint32_t pitch,
                            uint32_t layout, bool src_u16)
    {
        if(transposed && layout != kLayoutPlain && (f->size < 256 || (pitch & 7u) != 0u))
        {
            set_error("the split stack layout needs a filter size >= 256 and a pitch that is a multiple of 8");
            return PARIS_B200_EINVAL;
        }
        if(f->device != ctx->device)
        {
            set_error("filter lives on device %d, context on device %d", f->device, ctx->device);
            return PARIS_B200_EINVAL;
        }
        if(count == 0)
            return PARIS_B200_OK;
        if(count > static_cast<uint32_t>(kMaxBatch))
        {
            set_error("filter batch of %u exceeds %d", count, kMaxBatch);
            return PARIS_B200_EINVAL;
        }
        if(f->size < 256)
        {
            // small transforms: the all-shared-memory kernel, one projection per launch
            for(uint32_t i = 0; i < count; ++i)
            {
                float* dst = transposed ? d_stack + slot_floats * (first_slot + i) : d_dst[i];
                int rc = PARIS_B200_EINVAL;
                switch(f->size)
                {
                    case 32: rc = launch_filter_small<5>(ctx, d_src[i], dst, dim_x, dim_y, f, w, transposed, pitch, src_u16); break;
                    case 64: rc = launch_filter_small<6>(ctx, d_src[i], dst, dim_x, dim_y, f, w, transposed, pitch, src_u16); break;
                    case 128: rc = launch_filter_small<7>(ctx, d_src[i], dst, dim_x, dim_y, f, w, transposed, pitch, src_u16); break;
                    default: set_error("unsupported filter size %u", f->size); break;
                }
                PB_TRY(rc);
            }
            return PARIS_B200_OK;
        }
        filter_batch io{};
        io.src_u16 = src_u16 ? 1u : 0u;
        for(uint32_t i = 0; i < count; ++i)
        {
            io.src[i] = d_src[i];
            io.dst[i] = transposed ? nullptr : d_dst[i];
        }
        switch(f->size)
        {
            case 256: return launch_grouped<8>(ctx, io, count, d_stack, first_slot, slot_floats, dim_x, dim_y, f, w, transposed, pitch, layout);
            case 512: return launch_grouped<9>(ctx, io, count, d_stack, first_slot, slot_floats, dim_x, dim_y, f, w, transposed, pitch, layout);
            case 1024: return launch_grouped<10>(ctx, io, count, d_stack, first_slot, slot_floats, dim_x, dim_y, f, w, transposed, pitch, layout);
            case 2048: return launch_grouped<11>(ctx, io, count, d_stack, first_slot, slot_floats, dim_x, dim_y, f, w, transposed, pitch, layout);
            case 4096:
            {
                // two transforms per CTA when the staging tile leaves room for them ("filter_wide", default on)
                const size_t wide_smem = sizeof(float2) * (3 * (2 * 4096 - 8) / 4) + sizeof(float2) * plan<12, true>::NPAD * 2
                                       + (transposed ? sizeof(float) * kStagePitch * dim_x : 0);
                if(ctx->filter_wide != 0 && wide_smem <= 227u * 1024u)
                    return launch_grouped<12, true>(ctx, io, count, d_stack, first_slot, slot_floats, dim_x, dim_y, f, w, transposed, pitch, layout);
                return launch_grouped<12>(ctx, io, count, d_stack, first_slot, slot_floats, dim_x, dim_y, f, w, transposed, pitch, layout);
            }
            case 8192: return launch_grouped<13>(ctx, io, count, d_stack, first_slot, slot_floats, dim_x, dim_y, f, w, transposed, pitch, layout);
            default: break;
        }
        set_error("unsupported filter size %u", f->size);
        return PARIS_B200_EINVAL;
    }

    int launch_filter(paris_b200_ctx* ctx, const float* d_src, float* d_dst, uint32_t dim_x, uint32_t dim_y,
                      const paris_b200_filter* f, const weight_params& w, bool dst_transposed, uint32_t dst_pitch,
                      uint32_t layout)
    {
        // single projection; for the transposed case d_dst is the slot itself (a one-slot stack)
        const float* src[1] = {d_src};
        float* dst[1] = {d_dst};
        return launch_filter_batch(ctx, src, dst, 1u, dst_transposed ? d_dst : nullptr, 0u, 0, dim_x, dim_y, f, w,
                                   dst_transposed, dst_pitch, layout);
    }
}
