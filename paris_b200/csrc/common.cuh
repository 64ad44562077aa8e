// common.cuh -- context, error plumbing and launch helpers shared by the sm_100a kernels.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <deque>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/paris_b200.h"

namespace pb
{
    void set_error(const char* fmt, ...);

    // Layout of one line (fixed detector column, all detector rows v) of a filtered-stack slot:
    //   plain   line[v]
    //   split2  even rows first, then odd rows: line[(v & 1) * pitch/2 + (v >> 1)].  Used when a voxel step
    //           in z moves ~2 detector rows: the 32 lanes of a backprojection warp then read each parity
    //           plane with unit stride instead of hitting every other shared-memory bank twice.
    enum : uint32_t { kLayoutPlain = 0u, kLayoutSplit2 = 1u };

    __host__ __device__ inline uint32_t line_offset(uint32_t v, uint32_t pitch, uint32_t layout)
    {
        return layout == kLayoutSplit2 ? (v & 1u) * (pitch >> 1) + (v >> 1) : v;
    }

    // Maximum projections accumulated by one backprojection launch (sin/cos travel as kernel parameters).
    constexpr int kMaxBatch = 256;

    struct raw_buffer
    {
        void* ptr = nullptr;
        size_t bytes = 0;
        bool in_use = false;
        cudaEvent_t freed = nullptr;   // recorded on the compute stream at dev_free
        bool freed_valid = false;
        bool held = false;          // read by a deferred filter launch that has not been enqueued yet
        bool free_pending = false;  // dev_free arrived while held
        bool in_slab = false;       // carved out of one of the context's pool slabs (not freed on its own)
        bool touched = false;       // since dev_alloc: uploaded to, or named in a launch queued on the compute stream
    };

    // Geometry of one pending backprojection batch: everything but the angles must match for
    // projections to share a launch.
    struct bp_target
    {
        float* d_vol = nullptr;
        uint32_t v_dim_x = 0, v_dim_y = 0, v_dim_z = 0, v_offset = 0;
        paris_b200_detector_geometry det{};
        paris_b200_volume_geometry vol_full{};
        int enable_roi = 0;
        paris_b200_roi roi{};
        float delta_s_mm = 0.f, delta_t_mm = 0.f;
    };

    struct weight_params
    {
        int enable = 0;
        float h_min = 0.f, v_min = 0.f, d_sd = 0.f, l_px_row = 0.f, l_px_col = 0.f;
    };

    struct tma_desc_cache
    {
        const float* base = nullptr;
        uint32_t n_row = 0, pitch = 0, slots = 0, box_v = 0, box_h = 0, layout = 0;
        CUtensorMap map{};
        bool valid = false;
    };
}

struct paris_b200_filter
{
    int device = 0;
    uint32_t size = 0;       // N
    float tau = 0.f;
    float* d_k = nullptr;    // K[x], x = 0..N/2 (as the reference defines it)
    float* d_kn = nullptr;   // K[x] / N  (exact: N is a power of two)
    float* d_knp = nullptr;  // K/N per STORAGE position of the forward transform (digit-reversed order), N entries
    float2* d_tw = nullptr;  // exp(-2 pi i k / N), k = 0..N-1
    float2* d_twc = nullptr; // compact per-span tables: exp(-2 pi i k / 2^m), k < 3*2^m/4, at offset 3*(2^m - 8)/4, m = 3..log2 N
};

struct paris_b200_ctx
{
    int device = 0;
    int sm_count = 0;
    cudaStream_t compute = nullptr;
    cudaStream_t copy = nullptr;
    cudaEvent_t h2d_done = nullptr;  // last H2D on the copy stream
    bool h2d_any = false;
    cudaEvent_t scratch_ev = nullptr;
    uint64_t launches = 0;
    // which backprojection kernel the launches went to (paris_b200_ctx_bp_kernel_info): a geometry whose footprint
    // does not fit the TMA kernel's tiles silently takes the exact kernel, 2-3x slower -- callers can see that
    uint64_t bp_launches_tma = 0, bp_launches_exact = 0;
    char bp_last_kernel[96] = "";
    // pool statistics (paris_b200_ctx_stats)
    uint64_t stat_pool_malloc = 0, stat_pool_ready = 0, stat_pool_busy = 0, stat_flush = 0;

    // options
    int bp_batch = 256;
    int bp_kernel = 0;
    int bp_tile = 0;     // 0: half tiles (two CTAs per SM) when the footprint fits, 1: full tiles only
    int filter_wide = 1; // 4096-point transforms: two row pairs per 512-thread CTA (16 warps per SM) instead of one per 256
    int bp_swizzle = 16; // CTAs numbered in bp_swizzle x bp_swizzle super-blocks of (x, y) tiles; 0: row-major

    // pooled raw projection buffers (dev_alloc / dev_free)
    std::vector<pb::raw_buffer> pool;
    std::deque<size_t> free_fifo;   // indices into pool, in release order
    std::vector<void*> pool_slabs;  // one allocation per projection size, carved into the pool's first buffers
    // one cached volume allocation (volume_free keeps the last buffer, volume_alloc reuses it if the size matches):
    // the reference allocates and frees the slab once per task (src/main.cpp:95,107)
    std::unordered_map<float*, size_t> vol_bytes;   // every live volume allocation of this context
    float* spare_vol = nullptr;
    size_t spare_vol_bytes = 0;

    // filtered stack owned by the context (deferred backprojection)
    float* stack = nullptr;
    uint32_t stack_n_row = 0, stack_n_col = 0, stack_pitch = 0, stack_slots = 0, stack_layout = 0;
    size_t stack_slot_floats = 0;

    // pending batch
    pb::bp_target target{};
    int pending = 0;
    int flushes_for_target = 0;   // batches already launched into target.d_vol (the first one is kept short)
    float pend_sin[pb::kMaxBatch];
    float pend_cos[pb::kMaxBatch];
    // raw projections of the pending batch whose fused weight+filter launch is deferred to the flush
    const float* pend_raw[pb::kMaxBatch];
    int pend_raw_first = 0;   // they are consecutive: slots pend_raw_first .. pend_raw_first + pend_raw_count - 1
    int pend_raw_count = 0;
    pb::weight_params pend_w{};
    const paris_b200_filter* pend_filter = nullptr;

    pb::tma_desc_cache tma;
};

#define PB_CUDA(expr)                                                                                   \
    do                                                                                                  \
    {                                                                                                   \
        cudaError_t e_ = (expr);                                                                        \
        if(e_ != cudaSuccess)                                                                           \
        {                                                                                               \
            (void)cudaGetLastError(); /* reported here: the next launch check must not see it again */  \
            pb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__);  \
            return e_ == cudaErrorMemoryAllocation ? PARIS_B200_ENOMEM : PARIS_B200_ECUDA;              \
        }                                                                                               \
    } while(0)

#define PB_CHECK_ARG(cond)                                                                              \
    do                                                                                                  \
    {                                                                                                   \
        if(!(cond))                                                                                     \
        {                                                                                               \
            pb::set_error("invalid argument: %s (%s:%d)", #cond, __FILE__, __LINE__);                   \
            return PARIS_B200_EINVAL;                                                                   \
        }                                                                                               \
    } while(0)

#define PB_TRY(expr)                 \
    do                               \
    {                                \
        int rc_ = (expr);            \
        if(rc_ != PARIS_B200_OK)     \
            return rc_;              \
    } while(0)

// kernel launchers implemented in the individual .cu files ------------------------------------------------
namespace pb
{
    int launch_weight(paris_b200_ctx* ctx, float* d_proj, uint32_t dim_x, uint32_t dim_y, float h_min, float v_min,
                      float d_sd, float l_px_row, float l_px_col);

    // src rows -> (optional weight) -> ramp filter -> dst.  dst_transposed: write dst[s*dst_pitch + t]
    // (stack slot layout) instead of dst[t*dim_x + s].
    int launch_filter(paris_b200_ctx* ctx, const float* d_src, float* d_dst, uint32_t dim_x, uint32_t dim_y,
                      const paris_b200_filter* f, const weight_params& w, bool dst_transposed, uint32_t dst_pitch,
                      uint32_t layout = kLayoutPlain);

    // `count` (<= kMaxBatch) projections in one launch.  Transposed: src[i] -> slot first_slot + i of d_stack;
    // row-major: src[i] -> dst[i] (may alias).
    int launch_filter_batch(paris_b200_ctx* ctx, const float* const* d_src, float* const* d_dst, uint32_t count,
                            float* d_stack, uint32_t first_slot, size_t slot_floats, uint32_t dim_x, uint32_t dim_y,
                            const paris_b200_filter* f, const weight_params& w, bool transposed, uint32_t pitch,
                            uint32_t layout = kLayoutPlain, bool src_u16 = false);

    // frequency index stored at position p after the forward passes of the size-2^log2n transform
    int frequency_of_position(int log2n, int p);

    int launch_transpose_to_slot(paris_b200_ctx* ctx, const float* d_src, float* d_slot, uint32_t dim_x,
                                 uint32_t dim_y, uint32_t pitch, uint32_t layout = kLayoutPlain);

    int launch_backproject(paris_b200_ctx* ctx, const float* d_stack, size_t slot_floats, uint32_t pitch,
                           uint32_t first, uint32_t count, const float* sn, const float* cs, const bp_target& t,
                           uint32_t layout = kLayoutPlain);

    // the layout the backprojection wants for this geometry (from the detector rows one voxel step in z spans
    // at the rotation axis)
    uint32_t choose_stack_layout(const paris_b200_detector_geometry& det, const paris_b200_volume_geometry& vol_full);

    int launch_phantom(paris_b200_ctx* ctx, const double* h_ellipsoids, uint32_t n, const paris_b200_detector_geometry* det,
                       uint32_t first_idx, uint32_t n_proj, float* d_out);

    // CUDA loads kernels lazily, on their first launch, and that load synchronises the context.  Inside a group step
    // -- streams waiting for flags that other, not yet enqueued work sets -- such a load never returns (measured: a
    // slab shape that needed a not yet used tile instantiation hung the whole group).  Everything a step can launch
    // is therefore loaded before the first step.
    void preload_backprojection_kernels();
    void preload_filter_kernels(uint32_t size);

    inline uint32_t stack_pitch_for(uint32_t n_col) { return (n_col + 31u) & ~31u; }

    // Backproject `count` stack slots into the target and bring the volume to the host, z-chunk by z-chunk (api.cu)
    int backproject_and_download(paris_b200_ctx* ctx, const float* d_stack, size_t slot_floats, uint32_t pitch, uint32_t first,
                                 uint32_t count, const float* sn, const float* cs, const bp_target& t, uint32_t layout,
                                 float* h_dst, bool wait, uint32_t h_row_floats);
}
