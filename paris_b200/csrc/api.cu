// api.cu -- the C ABI of include/paris_b200.h: contexts, memory, geometry, filter tables and the
// deferred-backprojection bookkeeping.  The kernels live in weight.cu / filter.cu / backproject.cu.
#include "common.cuh"

#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <complex>

namespace
{
    // Runs when the library is loaded, before it makes its first CUDA call: kernels are loaded with their module instead of
    // lazily at their first launch (a lazy load synchronises the context, see common.cuh), and every stream gets a
    // hardware queue of its own (the default of 8 makes unrelated streams wait for each other).  Both are defaults only:
    // whoever set the variables keeps their values, and a process that initialised CUDA earlier is not affected --
    // group_create loads its kernels explicitly for that case.
    __attribute__((constructor)) void paris_b200_library_defaults()
    {
        setenv("CUDA_MODULE_LOADING", "EAGER", 0);
        setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    }
}

namespace pb
{
    static thread_local char g_error[512] = "";

    void set_error(const char* fmt, ...)
    {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(g_error, sizeof(g_error), fmt, ap);
        va_end(ap);
    }

    static int bind(const paris_b200_ctx* ctx)
    {
        PB_CUDA(cudaSetDevice(ctx->device));
        return PARIS_B200_OK;
    }
}

using namespace pb;

static int flush_if_held(paris_b200_ctx* ctx, const void* d_ptr);
static int run_pending_filter(paris_b200_ctx* ctx);

extern "C" const char* paris_b200_last_error(void) { return pb::g_error; }
extern "C" const char* paris_b200_version(void) { return "paris_b200 0.1 (sm_100a)"; }

// ---- devices / contexts -----------------------------------------------------------------------------------

extern "C" int paris_b200_device_count(int* count)
{
    PB_CHECK_ARG(count != nullptr);
    *count = 0;
    PB_CUDA(cudaGetDeviceCount(count));
    return PARIS_B200_OK;
}

extern "C" int paris_b200_ctx_create(int device, paris_b200_ctx** out)
{
    PB_CHECK_ARG(out != nullptr);
    *out = nullptr;
    int n = 0;
    PB_CUDA(cudaGetDeviceCount(&n));
    PB_CHECK_ARG(device >= 0 && device < n);
    PB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop{};
    PB_CUDA(cudaGetDeviceProperties(&prop, device));
    if(prop.major != 10)
    {
        set_error("device %d is sm_%d%d; this library only carries sm_100a code (no fallback)", device, prop.major,
                  prop.minor);
        return PARIS_B200_ECUDA;
    }
    auto* ctx = new paris_b200_ctx{};
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    if(const char* e = std::getenv("PARIS_B200_FILTER_WIDE"))   // (A/B of the 4096-point filter's CTA shape)
        ctx->filter_wide = std::atoi(e) != 0 ? 1 : 0;
    // (a context whose streams or events cannot all be created is taken apart again, not leaked)
    const auto build = [&]() -> int {
        PB_CUDA(cudaStreamCreateWithFlags(&ctx->compute, cudaStreamNonBlocking));
        PB_CUDA(cudaStreamCreateWithFlags(&ctx->copy, cudaStreamNonBlocking));
        PB_CUDA(cudaEventCreateWithFlags(&ctx->h2d_done, cudaEventDisableTiming));
        PB_CUDA(cudaEventCreateWithFlags(&ctx->scratch_ev, cudaEventDisableTiming));
        return PARIS_B200_OK;
    };
    const int rc = build();
    if(rc != PARIS_B200_OK)
    {
        if(ctx->scratch_ev) cudaEventDestroy(ctx->scratch_ev);
        if(ctx->h2d_done) cudaEventDestroy(ctx->h2d_done);
        if(ctx->copy) cudaStreamDestroy(ctx->copy);
        if(ctx->compute) cudaStreamDestroy(ctx->compute);
        delete ctx;
        return rc;
    }
    *out = ctx;
    return PARIS_B200_OK;
}

extern "C" int paris_b200_ctx_destroy(paris_b200_ctx* ctx)
{
    if(ctx == nullptr)
        return PARIS_B200_OK;
    PB_TRY(bind(ctx));
    cudaStreamSynchronize(ctx->compute);
    cudaStreamSynchronize(ctx->copy);
    for(auto& b : ctx->pool)
    {
        if(b.ptr && !b.in_slab) cudaFree(b.ptr);
        if(b.freed) cudaEventDestroy(b.freed);
    }
    for(void* slab : ctx->pool_slabs)
        cudaFree(slab);
    if(ctx->stack) cudaFree(ctx->stack);
    if(ctx->spare_vol) cudaFree(ctx->spare_vol);
    cudaEventDestroy(ctx->h2d_done);
    cudaEventDestroy(ctx->scratch_ev);
    cudaStreamDestroy(ctx->compute);
    cudaStreamDestroy(ctx->copy);
    delete ctx;
    return PARIS_B200_OK;
}

extern "C" int paris_b200_ctx_device(const paris_b200_ctx* ctx, int* device)
{
    PB_CHECK_ARG(ctx != nullptr && device != nullptr);
    *device = ctx->device;
    return PARIS_B200_OK;
}

extern "C" int paris_b200_ctx_bind(paris_b200_ctx* ctx)
{
    PB_CHECK_ARG(ctx != nullptr);
    return bind(ctx);
}

extern "C" int paris_b200_ctx_sync(paris_b200_ctx* ctx)
{
    PB_CHECK_ARG(ctx != nullptr);
    PB_TRY(bind(ctx));
    PB_CUDA(cudaStreamSynchronize(ctx->copy));
    PB_CUDA(cudaStreamSynchronize(ctx->compute));
    return PARIS_B200_OK;
}

extern "C" int paris_b200_ctx_stream(paris_b200_ctx* ctx, void** stream)
{
    PB_CHECK_ARG(ctx != nullptr && stream != nullptr);
    *stream = static_cast<void*>(ctx->compute);
    return PARIS_B200_OK;
}

extern "C" int paris_b200_ctx_launch_count(const paris_b200_ctx* ctx, uint64_t* launches)
{
    PB_CHECK_ARG(ctx != nullptr && launches != nullptr);
    *launches = ctx->launches;
    return PARIS_B200_OK;
}

extern "C" int paris_b200_ctx_stats(const paris_b200_ctx* ctx, uint64_t* stats, int n)
{
    PB_CHECK_ARG(ctx != nullptr && stats != nullptr && n >= 6);
    stats[0] = ctx->launches;
    stats[1] = ctx->stat_pool_malloc;
    stats[2] = ctx->stat_pool_ready;
    stats[3] = ctx->stat_pool_busy;
    stats[4] = ctx->stat_flush;
    stats[5] = ctx->pool.size();
    return PARIS_B200_OK;
}

extern "C" int paris_b200_ctx_bp_kernel_info(const paris_b200_ctx* ctx, char* name, size_t name_len, uint64_t* tma_launches,
                                             uint64_t* exact_launches)
{
    PB_CHECK_ARG(ctx != nullptr);
    if(name != nullptr && name_len > 0)
    {
        std::strncpy(name, ctx->bp_last_kernel, name_len - 1);
        name[name_len - 1] = '\0';
    }
    if(tma_launches != nullptr)
        *tma_launches = ctx->bp_launches_tma;
    if(exact_launches != nullptr)
        *exact_launches = ctx->bp_launches_exact;
    return PARIS_B200_OK;
}

struct paris_b200_event
{
    int device;
    cudaEvent_t ev;
};

extern "C" int paris_b200_event_create(paris_b200_ctx* ctx, paris_b200_event** ev)
{
    PB_CHECK_ARG(ctx != nullptr && ev != nullptr);
    PB_TRY(bind(ctx));
    auto* e = new paris_b200_event{ctx->device, nullptr};
    PB_CUDA(cudaEventCreate(&e->ev));
    *ev = e;
    return PARIS_B200_OK;
}

extern "C" int paris_b200_event_record(paris_b200_ctx* ctx, paris_b200_event* ev)
{
    PB_CHECK_ARG(ctx != nullptr && ev != nullptr);
    PB_TRY(bind(ctx));
    PB_CUDA(cudaEventRecord(ev->ev, ctx->compute));
    return PARIS_B200_OK;
}

extern "C" int paris_b200_event_elapsed_ms(paris_b200_event* start, paris_b200_event* stop, float* ms)
{
    PB_CHECK_ARG(start != nullptr && stop != nullptr && ms != nullptr);
    PB_CUDA(cudaSetDevice(stop->device));
    PB_CUDA(cudaEventSynchronize(stop->ev));
    PB_CUDA(cudaEventElapsedTime(ms, start->ev, stop->ev));
    return PARIS_B200_OK;
}

extern "C" int paris_b200_event_destroy(paris_b200_event* ev)
{
    if(ev == nullptr)
        return PARIS_B200_OK;
    cudaSetDevice(ev->device);
    cudaEventDestroy(ev->ev);
    delete ev;
    return PARIS_B200_OK;
}

extern "C" int paris_b200_ctx_set_option(paris_b200_ctx* ctx, const char* name, int64_t value)
{
    PB_CHECK_ARG(ctx != nullptr && name != nullptr);
    if(std::strcmp(name, "bp_batch") == 0)
    {
        PB_CHECK_ARG(value >= 1 && value <= kMaxBatch);
        PB_TRY(paris_b200_flush(ctx));
        if(ctx->stack != nullptr && static_cast<uint32_t>(value) > ctx->stack_slots)
        {
            PB_TRY(bind(ctx));
            PB_CUDA(cudaStreamSynchronize(ctx->compute));
            PB_CUDA(cudaFree(ctx->stack));
            ctx->stack = nullptr;
            ctx->stack_slots = 0;
            ctx->tma.valid = false;
        }
        ctx->bp_batch = static_cast<int>(value);
        return PARIS_B200_OK;
    }
    if(std::strcmp(name, "bp_kernel") == 0)
    {
        PB_CHECK_ARG(value >= 0 && value <= 2);
        PB_TRY(paris_b200_flush(ctx));
        ctx->bp_kernel = static_cast<int>(value);
        return PARIS_B200_OK;
    }
    if(std::strcmp(name, "bp_swizzle") == 0)
    {
        PB_CHECK_ARG(value >= 0 && value <= 64);
        PB_TRY(paris_b200_flush(ctx));
        ctx->bp_swizzle = static_cast<int>(value);
        return PARIS_B200_OK;
    }
    if(std::strcmp(name, "filter_wide") == 0)
    {
        PB_CHECK_ARG(value == 0 || value == 1);
        ctx->filter_wide = static_cast<int>(value);
        return PARIS_B200_OK;
    }
    if(std::strcmp(name, "bp_tile") == 0)
    {
        PB_CHECK_ARG(value >= 0 && value <= 2);
        PB_TRY(paris_b200_flush(ctx));
        ctx->bp_tile = static_cast<int>(value);
        return PARIS_B200_OK;
    }
    set_error("unknown option '%s'", name);
    return PARIS_B200_EINVAL;
}

// ---- geometry (host arithmetic; same float expressions as the reference) -----------------------------------

// src/geometry.cpp:36-67
extern "C" int paris_b200_calculate_volume_geometry(const paris_b200_detector_geometry* det,
                                                    paris_b200_volume_geometry* vol)
{
    PB_CHECK_ARG(det != nullptr && vol != nullptr);
    const float n_row = static_cast<float>(det->n_row);
    const float n_col = static_cast<float>(det->n_col);
    const float off_s = std::fabs(det->delta_s * det->l_px_row);
    const float off_t = std::fabs(det->delta_t * det->l_px_col);
    const float d_so = std::fabs(det->d_so);
    const float d_sd = std::fabs(det->d_od) + d_so;

    const float half_width = ((n_row * det->l_px_row) / 2.f) + off_s;
    const float alpha = std::atan(half_width / d_sd);
    const float r = d_so * std::sin(alpha);

    vol->l_vx_x = r / (half_width / det->l_px_row);
    vol->l_vx_y = vol->l_vx_x;
    vol->l_vx_z = vol->l_vx_x;
    vol->dim_x = static_cast<uint32_t>((2.f * r) / vol->l_vx_x);
    vol->dim_y = vol->dim_x;
    vol->dim_z = static_cast<uint32_t>(((n_col * det->l_px_col / 2.f) + off_t) * (d_so / d_sd) * (2.f / vol->l_vx_z));
    return PARIS_B200_OK;
}

// src/geometry.cpp:86-130
extern "C" int paris_b200_apply_roi(const paris_b200_volume_geometry* vol, const paris_b200_roi* roi,
                                    paris_b200_volume_geometry* out)
{
    PB_CHECK_ARG(vol != nullptr && roi != nullptr && out != nullptr);
    *out = *vol;
    if(!(roi->x1 < roi->x2 && roi->y1 < roi->y2 && roi->z1 < roi->z2))
        return PARIS_B200_OK; // "Invalid ROI coordinates. ROI NOT applied."
    const uint32_t dx = roi->x2 - roi->x1 + (roi->x1 == 0 ? 1u : 0u);
    const uint32_t dy = roi->y2 - roi->y1 + (roi->y1 == 0 ? 1u : 0u);
    const uint32_t dz = roi->z2 - roi->z1 + (roi->z1 == 0 ? 1u : 0u);
    if(dx <= vol->dim_x && dy <= vol->dim_y && dz <= vol->dim_z)
    {
        out->dim_x = dx;
        out->dim_y = dy;
        out->dim_z = dz;
    }
    return PARIS_B200_OK;
}

// src/filtering.cpp:38
extern "C" uint32_t paris_b200_filter_size(uint32_t n_row)
{
    return static_cast<uint32_t>(2 * std::pow(2.f, std::ceil(std::log2(n_row))));
}

// src/cuda/subvolume_information.cpp:63-118
extern "C" int paris_b200_make_subvolume_information(paris_b200_ctx* ctx, const paris_b200_volume_geometry* vol,
                                                     const paris_b200_detector_geometry* det, int num_slabs,
                                                     paris_b200_subvolume_info* out)
{
    PB_CHECK_ARG(vol != nullptr && det != nullptr && out != nullptr && num_slabs >= 0);
    uint32_t slabs = static_cast<uint32_t>(num_slabs);
    if(slabs == 0)
    {
        PB_CHECK_ARG(ctx != nullptr);
        PB_TRY(bind(ctx));
        // 64-bit byte counts (the reference's 32-bit product overflows at >= 4 GiB)
        const size_t vol_bytes = static_cast<size_t>(vol->dim_x) * vol->dim_y * vol->dim_z * sizeof(float);
        const size_t proj_bytes = static_cast<size_t>(det->n_row) * det->n_col * sizeof(float);
        // Besides the slab and the reference's ten projections, this backend holds the filtered stack of one batch
        // and a pool of raw projection buffers (bp_batch + 66 of them, dev_alloc); neither shrinks with the slab.
        // What the context already holds of them (and a spare slab, which volume_alloc releases) is not counted twice.
        const size_t slot_bytes = static_cast<size_t>(stack_pitch_for(det->n_col)) * det->n_row * sizeof(float);
        const size_t pool_stride = (proj_bytes + 255u) & ~static_cast<size_t>(255u);
        size_t fixed = 10u * proj_bytes;
        if(ctx->stack == nullptr)
            fixed += static_cast<size_t>(ctx->bp_batch) * slot_bytes;
        if(ctx->pool.empty())
            fixed += (static_cast<size_t>(ctx->bp_batch) + 66u) * pool_stride;
        size_t mem_free = 0, mem_total = 0;
        PB_CUDA(cudaMemGetInfo(&mem_free, &mem_total));
        mem_free += ctx->spare_vol_bytes;
        size_t need = vol_bytes;
        slabs = 1;
        // (the last slab carries the remainder: dim_z / slabs + dim_z % slabs slices)
        const size_t slice_bytes = static_cast<size_t>(vol->dim_x) * vol->dim_y * sizeof(float);
        while(slabs < vol->dim_z)
        {
            need = (vol->dim_z / slabs + vol->dim_z % slabs) * slice_bytes;
            if(need + fixed < mem_free)
                break;
            slabs *= 2;
        }
    }
    PB_CHECK_ARG(slabs >= 1 && slabs <= vol->dim_z);
    out->dim_x = vol->dim_x;
    out->dim_y = vol->dim_y;
    out->dim_z = vol->dim_z / slabs;
    out->remainder = vol->dim_z % slabs;
    out->num = static_cast<int32_t>(slabs);
    return PARIS_B200_OK;
}

// ---- memory ----------------------------------------------------------------------------------------------

extern "C" int paris_b200_host_alloc(size_t bytes, int zero, void** h_ptr)
{
    PB_CHECK_ARG(h_ptr != nullptr && bytes > 0);
    *h_ptr = nullptr;
    PB_CUDA(cudaHostAlloc(h_ptr, bytes, cudaHostAllocPortable));
    if(zero)
        std::memset(*h_ptr, 0, bytes);
    return PARIS_B200_OK;
}

// page-lock memory the caller owns (e.g. a POSIX shared-memory mapping several member processes write their slabs
// into: the host volume is then assembled by the downloads themselves)
extern "C" int paris_b200_host_register(void* h_ptr, size_t bytes)
{
    PB_CHECK_ARG(h_ptr != nullptr && bytes > 0);
    PB_CUDA(cudaHostRegister(h_ptr, bytes, cudaHostRegisterPortable));
    return PARIS_B200_OK;
}

extern "C" int paris_b200_host_unregister(void* h_ptr)
{
    if(h_ptr != nullptr)
        PB_CUDA(cudaHostUnregister(h_ptr));
    return PARIS_B200_OK;
}

extern "C" int paris_b200_host_free(void* h_ptr)
{
    if(h_ptr != nullptr)
        PB_CUDA(cudaFreeHost(h_ptr));
    return PARIS_B200_OK;
}

static pb::raw_buffer* find_buffer(paris_b200_ctx* ctx, const void* p)
{
    for(auto& b : ctx->pool)
        if(b.ptr == p)
            return &b;
    return nullptr;
}

// a launch queued on the compute stream reads or writes this (pooled) buffer: a later upload into it must wait
static void touch(paris_b200_ctx* ctx, const void* p)
{
    if(auto* b = find_buffer(ctx, p))
        b->touched = true;
}

extern "C" int paris_b200_dev_alloc(paris_b200_ctx* ctx, size_t bytes, void** d_ptr)
{
    PB_CHECK_ARG(ctx != nullptr && d_ptr != nullptr && bytes > 0);
    PB_TRY(bind(ctx));
    *d_ptr = nullptr;
    // Free buffers are recycled in the order they were released (FIFO): the front of the queue is the
    // buffer whose last reader finished -- or will finish -- first, so an upload into it can run as far
    // ahead of the compute stream as the pool is deep.
    size_t same_size = 0;
    for(const auto& b : ctx->pool)
        same_size += b.bytes == bytes ? 1u : 0u;
    // deep enough for one pending batch (its raw buffers are held until the fused filter launch) plus the uploads
    // running ahead of it
    const size_t cap = static_cast<size_t>(ctx->bp_batch) + 66u;
    for(auto it = ctx->free_fifo.begin(); it != ctx->free_fifo.end(); ++it)
    {
        pb::raw_buffer& b = ctx->pool[*it];
        if(b.bytes != bytes)
            continue;
        bool ready = !b.freed_valid;
        if(!ready)
        {
            ready = cudaEventQuery(b.freed) == cudaSuccess;
            (void)cudaGetLastError(); // cudaErrorNotReady is not an error
        }
        if(!ready && same_size < cap)
            break; // oldest candidate still busy and the pool may grow: allocate a fresh buffer instead
        // either idle, or busy with the pool at its cap: proj_h2d makes the copy stream wait for `freed`
        if(ready)
        {
            b.freed_valid = false;
            ++ctx->stat_pool_ready;
        }
        else
            ++ctx->stat_pool_busy;
        b.in_use = true;
        b.touched = false;
        *d_ptr = b.ptr;
        ctx->free_fifo.erase(it);
        return PARIS_B200_OK;
    }
    if(same_size == 0 && bytes >= (1u << 16) && bytes <= (64u << 20))   // (projection-sized requests only)
    {
        // first projection buffer of this size: one allocation for the whole pool (a cudaMalloc per buffer costs
        // about a millisecond each and would be paid during the first reconstruction's uploads)
        const size_t stride = (bytes + 255u) & ~static_cast<size_t>(255u);
        void* slab = nullptr;
        if(cudaMalloc(&slab, stride * cap) == cudaSuccess)
        {
            ctx->pool_slabs.push_back(slab);
            ++ctx->stat_pool_malloc;
            for(size_t i = 0; i < cap; ++i)
            {
                pb::raw_buffer sb{};
                sb.ptr = static_cast<unsigned char*>(slab) + i * stride;
                sb.bytes = bytes;
                sb.in_slab = true;
                PB_CUDA(cudaEventCreateWithFlags(&sb.freed, cudaEventDisableTiming));
                ctx->pool.push_back(sb);
                if(i > 0)
                    ctx->free_fifo.push_back(ctx->pool.size() - 1u);
            }
            pb::raw_buffer& first = ctx->pool[ctx->pool.size() - cap];
            first.in_use = true;
            *d_ptr = first.ptr;
            return PARIS_B200_OK;
        }
        (void)cudaGetLastError(); // not enough memory for the whole pool: grow buffer by buffer
    }
    pb::raw_buffer nb{};
    ++ctx->stat_pool_malloc;
    PB_CUDA(cudaMalloc(&nb.ptr, bytes));
    PB_CUDA(cudaEventCreateWithFlags(&nb.freed, cudaEventDisableTiming));
    nb.bytes = bytes;
    nb.in_use = true;
    ctx->pool.push_back(nb);
    *d_ptr = nb.ptr;
    return PARIS_B200_OK;
}

extern "C" int paris_b200_dev_free(paris_b200_ctx* ctx, void* d_ptr)
{
    PB_CHECK_ARG(ctx != nullptr);
    if(d_ptr == nullptr)
        return PARIS_B200_OK;
    PB_TRY(bind(ctx));
    auto* b = find_buffer(ctx, d_ptr);
    if(b == nullptr || !b->in_use)
    {
        set_error("dev_free: %p was not allocated by this context", d_ptr);
        return PARIS_B200_EINVAL;
    }
    if(b->held)
    {
        b->free_pending = true; // released for real once the deferred filter launch that reads it is enqueued
        return PARIS_B200_OK;
    }
    PB_CUDA(cudaEventRecord(b->freed, ctx->compute));
    b->freed_valid = true;
    b->in_use = false;
    ctx->free_fifo.push_back(static_cast<size_t>(b - ctx->pool.data()));
    return PARIS_B200_OK;
}

extern "C" int paris_b200_volume_alloc(paris_b200_ctx* ctx, uint32_t dim_x, uint32_t dim_y, uint32_t dim_z,
                                       float** d_vol)
{
    PB_CHECK_ARG(ctx != nullptr && d_vol != nullptr && dim_x > 0 && dim_y > 0 && dim_z > 0);
    PB_TRY(bind(ctx));
    const size_t bytes = static_cast<size_t>(dim_x) * dim_y * dim_z * sizeof(float);
    *d_vol = nullptr;
    if(ctx->spare_vol != nullptr && ctx->spare_vol_bytes == bytes)
    {
        *d_vol = ctx->spare_vol;   // (its last use was ordered on the compute stream, as is the memset below)
        ctx->spare_vol = nullptr;
        ctx->spare_vol_bytes = 0;
    }
    else
    {
        // A spare of another size (the last slab of a run carries the remainder, src/make_volume.cpp:32-34) is let go
        // FIRST: slabs are sized to fill the device, two of them do not fit (the reference never holds two either).
        if(ctx->spare_vol != nullptr)
        {
            PB_CUDA(cudaStreamSynchronize(ctx->compute));
            PB_CUDA(cudaFree(ctx->spare_vol));
            ctx->vol_bytes.erase(ctx->spare_vol);
            ctx->spare_vol = nullptr;
            ctx->spare_vol_bytes = 0;
        }
        PB_CUDA(cudaMalloc(reinterpret_cast<void**>(d_vol), bytes));
        ctx->vol_bytes[*d_vol] = bytes;
    }
    PB_CUDA(cudaMemsetAsync(*d_vol, 0, bytes, ctx->compute));
    return PARIS_B200_OK;
}

extern "C" int paris_b200_volume_clear(paris_b200_ctx* ctx, float* d_vol, uint32_t dim_x, uint32_t dim_y,
                                       uint32_t dim_z)
{
    PB_CHECK_ARG(ctx != nullptr && d_vol != nullptr);
    PB_TRY(bind(ctx));
    if(ctx->target.d_vol == d_vol)
    {
        // the pending batch would be zeroed right after being added: drop it (the deferred filter launch still runs,
        // which is how the raw buffers it holds are let go) and start the short-first-batch ramp again
        if(ctx->pending > 0)
            PB_TRY(run_pending_filter(ctx));
        ctx->pending = 0;
        ctx->flushes_for_target = 0;
    }
    PB_CUDA(cudaMemsetAsync(d_vol, 0, static_cast<size_t>(dim_x) * dim_y * dim_z * sizeof(float), ctx->compute));
    return PARIS_B200_OK;
}

extern "C" int paris_b200_volume_free(paris_b200_ctx* ctx, float* d_vol)
{
    PB_CHECK_ARG(ctx != nullptr);
    if(d_vol == nullptr)
        return PARIS_B200_OK;
    PB_TRY(bind(ctx));
    if(ctx->pending > 0 && ctx->target.d_vol == d_vol)
        PB_TRY(paris_b200_flush(ctx)); // (simplest way to let go of the batch's raw buffers)
    if(ctx->target.d_vol == d_vol)
    {
        ctx->target = bp_target{};
        ctx->flushes_for_target = 0;
    }
    // keep the most recent buffer for the next volume_alloc of the same size; release the older spare
    const auto it = ctx->vol_bytes.find(d_vol);
    if(it == ctx->vol_bytes.end())
    {
        set_error("volume_free: %p was not allocated by this context", static_cast<void*>(d_vol));
        return PARIS_B200_EINVAL;
    }
    float* old = ctx->spare_vol;
    ctx->spare_vol = d_vol;
    ctx->spare_vol_bytes = it->second;
    if(old != nullptr)
        ctx->vol_bytes.erase(old);
    if(old != nullptr)
    {
        PB_CUDA(cudaStreamSynchronize(ctx->compute));
        PB_CUDA(cudaFree(old));
    }
    return PARIS_B200_OK;
}

extern "C" int paris_b200_proj_h2d(paris_b200_ctx* ctx, const float* h_src, float* d_dst, uint32_t dim_x,
                                   uint32_t dim_y)
{
    PB_CHECK_ARG(ctx != nullptr && h_src != nullptr && d_dst != nullptr && dim_x > 0 && dim_y > 0);
    PB_TRY(bind(ctx));
    PB_TRY(flush_if_held(ctx, d_dst));
    const size_t bytes = static_cast<size_t>(dim_x) * dim_y * sizeof(float);
    auto* b = find_buffer(ctx, d_dst);
    if(b != nullptr && b->freed_valid)
    {
        // pooled buffer recycled while its previous reader may still run
        PB_CUDA(cudaStreamWaitEvent(ctx->copy, b->freed, 0));
    }
    if(b != nullptr && b->touched)
    {
        // a live pooled buffer that was uploaded to before, or that a kernel queued on the compute stream reads or
        // writes (weight, apply_filter, filter_to_stack, the deferred launch flush_if_held just enqueued): the new
        // upload must not overtake that work.  (The first upload into a freshly allocated buffer -- the
        // reference's loop, src/main.cpp:100-101 -- does not take this path and keeps running ahead of compute.)
        PB_CUDA(cudaEventRecord(ctx->scratch_ev, ctx->compute));
        PB_CUDA(cudaStreamWaitEvent(ctx->copy, ctx->scratch_ev, 0));
    }
    if(b != nullptr)
        b->touched = true;
    else if(b == nullptr)
    {
        // foreign destination: order after everything queued on the compute stream
        PB_CUDA(cudaEventRecord(ctx->scratch_ev, ctx->compute));
        PB_CUDA(cudaStreamWaitEvent(ctx->copy, ctx->scratch_ev, 0));
    }
    PB_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->copy));
    PB_CUDA(cudaEventRecord(ctx->h2d_done, ctx->copy));
    ctx->h2d_any = true;
    PB_CUDA(cudaStreamWaitEvent(ctx->compute, ctx->h2d_done, 0));
    return PARIS_B200_OK;
}

extern "C" int paris_b200_h2d_done(paris_b200_ctx* ctx, int* done)
{
    PB_CHECK_ARG(ctx != nullptr && done != nullptr);
    PB_TRY(bind(ctx));
    *done = 1;
    if(ctx->h2d_any)
    {
        const cudaError_t e = cudaEventQuery(ctx->h2d_done);
        if(e == cudaErrorNotReady)
        {
            (void)cudaGetLastError();
            *done = 0;
        }
        else
            PB_CUDA(e);
    }
    return PARIS_B200_OK;
}

extern "C" int paris_b200_h2d_wait(paris_b200_ctx* ctx)
{
    PB_CHECK_ARG(ctx != nullptr);
    PB_TRY(bind(ctx));
    PB_CUDA(cudaStreamSynchronize(ctx->copy));
    return PARIS_B200_OK;
}

extern "C" int paris_b200_proj_d2h(paris_b200_ctx* ctx, const float* d_src, float* h_dst, uint32_t dim_x,
                                   uint32_t dim_y)
{
    PB_CHECK_ARG(ctx != nullptr && d_src != nullptr && h_dst != nullptr);
    PB_TRY(bind(ctx));
    PB_TRY(flush_if_held(ctx, d_src));
    const size_t bytes = static_cast<size_t>(dim_x) * dim_y * sizeof(float);
    PB_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->compute));
    PB_CUDA(cudaStreamSynchronize(ctx->compute));
    return PARIS_B200_OK;
}

extern "C" int paris_b200_vol_h2d(paris_b200_ctx* ctx, const float* h_src, float* d_dst, size_t n_voxels)
{
    PB_CHECK_ARG(ctx != nullptr && h_src != nullptr && d_dst != nullptr);
    PB_TRY(bind(ctx));
    PB_TRY(paris_b200_flush(ctx));
    PB_CUDA(cudaMemcpyAsync(d_dst, h_src, n_voxels * sizeof(float), cudaMemcpyHostToDevice, ctx->compute));
    PB_CUDA(cudaStreamSynchronize(ctx->compute));
    return PARIS_B200_OK;
}

static int run_pending_filter(paris_b200_ctx* ctx);

// Backproject `count` stack slots into the target and bring the volume to the host, z-chunk by z-chunk: the
// download of chunk c (copy stream) runs behind the backprojection of chunk c+1 (compute stream).  Chunks end
// at multiples of 64 slices in GLOBAL slice indices -- the kernel's tile anchors -- so no tile is computed twice
// and the result is bit-identical to the one-piece launch.  Returns with the host copy complete.
// wait == false: returns with the last chunk's copy still in flight on the copy stream (group.cu overlaps it with the
// next slab's backprojection).
// h_row_floats: floats from one row of the host destination to the next (0 or v_dim_x: contiguous; larger: the slab is
// a box inside a wider host volume).
int pb::backproject_and_download(paris_b200_ctx* ctx, const float* d_stack, size_t slot_floats, uint32_t pitch,
                                 uint32_t first, uint32_t count, const float* sn, const float* cs, const bp_target& t,
                                 uint32_t layout, float* h_dst, bool wait, uint32_t h_row_floats)
{
    const size_t h_row = h_row_floats > t.v_dim_x ? h_row_floats : t.v_dim_x;
    const size_t h_slice = h_row * t.v_dim_y;
    const uint32_t off0 = (t.enable_roi ? t.roi.z1 : 0u) + t.v_offset;
    const size_t slice = static_cast<size_t>(t.v_dim_x) * t.v_dim_y;
    uint32_t step = 64u;
    while((t.v_dim_z + step - 1u) / step > 32u)
        step *= 2u;
    for(uint32_t z = 0; z < t.v_dim_z;)
    {
        const uint32_t end = std::min<uint32_t>(((off0 + z) / step + 1u) * step - off0, t.v_dim_z);
        bp_target c = t;
        c.d_vol = t.d_vol + static_cast<size_t>(z) * slice;
        c.v_dim_z = end - z;
        c.v_offset = t.v_offset + z;
        const uint32_t launches = (count + static_cast<uint32_t>(ctx->bp_batch) - 1u) / static_cast<uint32_t>(ctx->bp_batch);
        const uint32_t per_launch = launches > 0u ? (count + launches - 1u) / launches : 0u;
        for(uint32_t done = 0; done < count;)
        {
            const uint32_t n = std::min<uint32_t>(count - done, per_launch);
            PB_TRY(launch_backproject(ctx, d_stack, slot_floats, pitch, first + done, n, sn + done, cs + done, c, layout));
            done += n;
        }
        PB_CUDA(cudaEventRecord(ctx->scratch_ev, ctx->compute));
        PB_CUDA(cudaStreamWaitEvent(ctx->copy, ctx->scratch_ev, 0));
        if(h_row == t.v_dim_x)
            PB_CUDA(cudaMemcpyAsync(h_dst + static_cast<size_t>(z) * slice, c.d_vol, static_cast<size_t>(end - z) * slice * sizeof(float),
                                    cudaMemcpyDeviceToHost, ctx->copy));
        else
            PB_CUDA(cudaMemcpy2DAsync(h_dst + static_cast<size_t>(z) * h_slice, h_row * sizeof(float), c.d_vol,
                                      static_cast<size_t>(t.v_dim_x) * sizeof(float), static_cast<size_t>(t.v_dim_x) * sizeof(float),
                                      static_cast<size_t>(end - z) * t.v_dim_y, cudaMemcpyDeviceToHost, ctx->copy));
        z = end;
    }
    if(wait)
        PB_CUDA(cudaStreamSynchronize(ctx->copy));
    return PARIS_B200_OK;
}

extern "C" int paris_b200_vol_d2h(paris_b200_ctx* ctx, const float* d_src, float* h_dst, size_t n_voxels)
{
    PB_CHECK_ARG(ctx != nullptr && d_src != nullptr && h_dst != nullptr);
    PB_TRY(bind(ctx));
    const bp_target& t = ctx->target;
    if(ctx->pending > 0 && t.d_vol == d_src
       && n_voxels == static_cast<size_t>(t.v_dim_x) * t.v_dim_y * t.v_dim_z)
    {
        // the pending batch goes into exactly this volume: overlap its backprojection with the download
        PB_TRY(run_pending_filter(ctx));
        ++ctx->stat_flush;
        ++ctx->flushes_for_target;
        const uint32_t n = static_cast<uint32_t>(ctx->pending);
        ctx->pending = 0;
        return backproject_and_download(ctx, ctx->stack, ctx->stack_slot_floats, ctx->stack_pitch, 0u, n, ctx->pend_sin,
                                        ctx->pend_cos, t, ctx->stack_layout, h_dst, true, 0u);
    }
    PB_TRY(paris_b200_flush(ctx));
    PB_CUDA(cudaMemcpyAsync(h_dst, d_src, n_voxels * sizeof(float), cudaMemcpyDeviceToHost, ctx->compute));
    PB_CUDA(cudaStreamSynchronize(ctx->compute));
    return PARIS_B200_OK;
}

// ---- filter table ------------------------------------------------------------------------------------------

namespace
{
    // radix-2 decimation-in-time FFT in double, host side, used once per filter
    void host_fft(std::vector<std::complex<double>>& a)
    {
        const size_t n = a.size();
        for(size_t i = 1, j = 0; i < n; ++i)
        {
            size_t bit = n >> 1;
            for(; j & bit; bit >>= 1)
                j ^= bit;
            j ^= bit;
            if(i < j)
                std::swap(a[i], a[j]);
        }
        for(size_t len = 2; len <= n; len <<= 1)
        {
            for(size_t i = 0; i < n; i += len)
                for(size_t k = 0; k < len / 2; ++k)
                {
                    const double ang = -2.0 * M_PI * static_cast<double>(k) / static_cast<double>(len);
                    const std::complex<double> w(std::cos(ang), std::sin(ang));
                    const auto u = a[i + k];
                    const auto v = a[i + k + len / 2] * w;
                    a[i + k] = u + v;
                    a[i + k + len / 2] = u - v;
                }
        }
    }
}

extern "C" int paris_b200_filter_create(paris_b200_ctx* ctx, uint32_t size, float tau, paris_b200_filter** out)
{
    PB_CHECK_ARG(ctx != nullptr && out != nullptr);
    PB_CHECK_ARG(size >= 32 && size <= 8192 && (size & (size - 1)) == 0);
    PB_CHECK_ARG(tau > 0.f);
    PB_TRY(bind(ctx));
    *out = nullptr;

    // spatial taps in float, exactly as src/openmp/filtering.cpp:52-73 writes them
    std::vector<std::complex<double>> r(size);
    const int32_t j0 = -(static_cast<int32_t>(size) - 2) / 2;
    const float pi_f = static_cast<float>(M_PI);
    for(uint32_t x = 0; x < size; ++x)
    {
        const int32_t j = j0 + static_cast<int32_t>(x);
        float tap;
        if(j == 0)
            tap = (1.f / 8.f) * (1.f / std::pow(tau, 2.f));
        else if(j % 2 == 0)
            tap = 0.f;
        else
            tap = -(1.f / (2.f * static_cast<float>(j * j) * (pi_f * pi_f) * (tau * tau)));
        r[x] = std::complex<double>(static_cast<double>(tap), 0.0);
    }
    host_fft(r);

    // K[x] = tau * |R[x]| on the float-rounded transform (:155-162); K/N is exact (N = 2^k)
    const uint32_t n_trans = size / 2 + 1;
    std::vector<float> k(n_trans), kn(n_trans);
    for(uint32_t x = 0; x < n_trans; ++x)
    {
        const float re = static_cast<float>(r[x].real());
        const float im = static_cast<float>(r[x].imag());
        k[x] = tau * std::abs(std::sqrt(std::pow(re, 2.f) + std::pow(im, 2.f)));
        kn[x] = k[x] / static_cast<float>(size);
    }
    // the same table in the order the forward transform leaves its output in (filter.cu)
    int log2n = 0;
    while((1u << log2n) < size)
        ++log2n;
    std::vector<float> knp(size);
    for(uint32_t pos = 0; pos < size; ++pos)
    {
        const uint32_t fq = static_cast<uint32_t>(frequency_of_position(log2n, static_cast<int>(pos)));
        knp[pos] = kn[fq <= size / 2 ? fq : size - fq];
    }
    std::vector<float2> tw(size);
    for(uint32_t i = 0; i < size; ++i)
    {
        const double ang = -2.0 * M_PI * static_cast<double>(i) / static_cast<double>(size);
        tw[i] = make_float2(static_cast<float>(std::cos(ang)), static_cast<float>(std::sin(ang)));
    }

    std::vector<float2> twc(3 * (2 * static_cast<size_t>(size) - 8) / 4);
    for(uint32_t span = 8; span <= size; span <<= 1)
    {
        float2* t = twc.data() + 3 * (span - 8) / 4;
        for(uint32_t i = 0; i < 3 * span / 4; ++i)
        {
            const double ang = -2.0 * M_PI * static_cast<double>(i) / static_cast<double>(span);
            t[i] = make_float2(static_cast<float>(std::cos(ang)), static_cast<float>(std::sin(ang)));
        }
    }

    auto* f = new paris_b200_filter{};
    f->device = ctx->device;
    f->size = size;
    f->tau = tau;
    PB_CUDA(cudaMalloc(reinterpret_cast<void**>(&f->d_k), n_trans * sizeof(float)));
    PB_CUDA(cudaMalloc(reinterpret_cast<void**>(&f->d_kn), n_trans * sizeof(float)));
    PB_CUDA(cudaMalloc(reinterpret_cast<void**>(&f->d_knp), size * sizeof(float)));
    PB_CUDA(cudaMemcpy(f->d_knp, knp.data(), size * sizeof(float), cudaMemcpyHostToDevice));
    PB_CUDA(cudaMalloc(reinterpret_cast<void**>(&f->d_tw), size * sizeof(float2)));
    PB_CUDA(cudaMemcpy(f->d_k, k.data(), n_trans * sizeof(float), cudaMemcpyHostToDevice));
    PB_CUDA(cudaMemcpy(f->d_kn, kn.data(), n_trans * sizeof(float), cudaMemcpyHostToDevice));
    PB_CUDA(cudaMemcpy(f->d_tw, tw.data(), size * sizeof(float2), cudaMemcpyHostToDevice));
    PB_CUDA(cudaMalloc(reinterpret_cast<void**>(&f->d_twc), twc.size() * sizeof(float2)));
    PB_CUDA(cudaMemcpy(f->d_twc, twc.data(), twc.size() * sizeof(float2), cudaMemcpyHostToDevice));
    *out = f;
    return PARIS_B200_OK;
}

extern "C" int paris_b200_filter_destroy(paris_b200_filter* f)
{
    if(f == nullptr)
        return PARIS_B200_OK;
    PB_CUDA(cudaSetDevice(f->device));
    cudaFree(f->d_k);
    cudaFree(f->d_kn);
    cudaFree(f->d_knp);
    cudaFree(f->d_tw);
    cudaFree(f->d_twc);
    delete f;
    return PARIS_B200_OK;
}

extern "C" int paris_b200_filter_read(paris_b200_ctx* ctx, const paris_b200_filter* f, float* h_k)
{
    PB_CHECK_ARG(ctx != nullptr && f != nullptr && h_k != nullptr);
    PB_TRY(bind(ctx));
    PB_CUDA(cudaMemcpy(h_k, f->d_k, (f->size / 2 + 1) * sizeof(float), cudaMemcpyDeviceToHost));
    return PARIS_B200_OK;
}

// ---- stages ------------------------------------------------------------------------------------------------

extern "C" int paris_b200_weight(paris_b200_ctx* ctx, float* d_proj, uint32_t dim_x, uint32_t dim_y, float h_min,
                                 float v_min, float d_sd, float l_px_row, float l_px_col)
{
    PB_CHECK_ARG(ctx != nullptr && d_proj != nullptr && dim_x > 0 && dim_y > 0);
    PB_TRY(bind(ctx));
    PB_TRY(flush_if_held(ctx, d_proj));
    touch(ctx, d_proj);
    return launch_weight(ctx, d_proj, dim_x, dim_y, h_min, v_min, d_sd, l_px_row, l_px_col);
}

extern "C" int paris_b200_apply_filter(paris_b200_ctx* ctx, float* d_proj, uint32_t dim_x, uint32_t dim_y,
                                       const paris_b200_filter* filter, uint32_t filter_size, uint32_t n_col)
{
    PB_CHECK_ARG(ctx != nullptr && d_proj != nullptr && filter != nullptr);
    PB_CHECK_ARG(filter_size == filter->size && n_col == dim_y && dim_x <= filter_size);
    PB_TRY(bind(ctx));
    PB_TRY(flush_if_held(ctx, d_proj));
    touch(ctx, d_proj);
    return launch_filter(ctx, d_proj, d_proj, dim_x, dim_y, filter, weight_params{}, false, 0);
}

extern "C" int paris_b200_weight_filter(paris_b200_ctx* ctx, float* d_proj, uint32_t dim_x, uint32_t dim_y,
                                        float h_min, float v_min, float d_sd, float l_px_row, float l_px_col,
                                        const paris_b200_filter* filter, uint32_t filter_size)
{
    PB_CHECK_ARG(ctx != nullptr && d_proj != nullptr && filter != nullptr);
    PB_CHECK_ARG(filter_size == filter->size && dim_x <= filter_size);
    PB_TRY(bind(ctx));
    PB_TRY(flush_if_held(ctx, d_proj));
    touch(ctx, d_proj);
    weight_params w{1, h_min, v_min, d_sd, l_px_row, l_px_col};
    return launch_filter(ctx, d_proj, d_proj, dim_x, dim_y, filter, w, false, 0);
}

static weight_params weighting_from_detector(const paris_b200_detector_geometry* det)
{
    // src/weighting.cpp:37-42
    const float n_row_f = static_cast<float>(det->n_row);
    const float n_col_f = static_cast<float>(det->n_col);
    weight_params w{};
    w.enable = 1;
    w.h_min = (det->delta_s * det->l_px_row) - ((n_row_f * det->l_px_row) / 2);
    w.v_min = (det->delta_t * det->l_px_col) - ((n_col_f * det->l_px_col) / 2);
    w.d_sd = std::fabs(det->d_so) + std::fabs(det->d_od);
    w.l_px_row = det->l_px_row;
    w.l_px_col = det->l_px_col;
    return w;
}

// ---- deferred backprojection ------------------------------------------------------------------------------

extern "C" int paris_b200_choose_stack_layout(const paris_b200_detector_geometry* det,
                                              const paris_b200_volume_geometry* vol_full, uint32_t* layout)
{
    PB_CHECK_ARG(det != nullptr && vol_full != nullptr && layout != nullptr);
    *layout = choose_stack_layout(*det, *vol_full);
    return PARIS_B200_OK;
}

extern "C" int paris_b200_stack_slot_bytes(uint32_t n_row, uint32_t n_col, size_t* bytes, uint32_t* pitch)
{
    PB_CHECK_ARG(n_row > 0 && n_col > 0);
    const uint32_t p = stack_pitch_for(n_col);
    if(bytes) *bytes = static_cast<size_t>(p) * n_row * sizeof(float);
    if(pitch) *pitch = p;
    return PARIS_B200_OK;
}

static bool same_target(const bp_target& a, const bp_target& b)
{
    return a.d_vol == b.d_vol && a.v_dim_x == b.v_dim_x && a.v_dim_y == b.v_dim_y && a.v_dim_z == b.v_dim_z
        && a.v_offset == b.v_offset && std::memcmp(&a.det, &b.det, sizeof(a.det)) == 0
        && std::memcmp(&a.vol_full, &b.vol_full, sizeof(a.vol_full)) == 0 && a.enable_roi == b.enable_roi
        && (!a.enable_roi || std::memcmp(&a.roi, &b.roi, sizeof(a.roi)) == 0) && a.delta_s_mm == b.delta_s_mm
        && a.delta_t_mm == b.delta_t_mm;
}

static int ensure_stack(paris_b200_ctx* ctx, uint32_t n_row, uint32_t n_col, uint32_t layout)
{
    ctx->stack_layout = layout; // (callers flush before changing the layout of a non-empty batch)
    const uint32_t pitch = stack_pitch_for(n_col);
    const uint32_t slots = static_cast<uint32_t>(ctx->bp_batch);
    if(ctx->stack != nullptr && ctx->stack_n_row == n_row && ctx->stack_n_col == n_col && ctx->stack_slots >= slots)
        return PARIS_B200_OK;
    if(ctx->stack != nullptr)
    {
        PB_CUDA(cudaStreamSynchronize(ctx->compute));
        PB_CUDA(cudaFree(ctx->stack));
        ctx->stack = nullptr;
    }
    ctx->stack_slot_floats = static_cast<size_t>(pitch) * n_row;
    PB_CUDA(cudaMalloc(reinterpret_cast<void**>(&ctx->stack), ctx->stack_slot_floats * slots * sizeof(float)));
    // pad columns (pitch > n_col) are never written by the filter kernel; keep them zero
    PB_CUDA(cudaMemsetAsync(ctx->stack, 0, ctx->stack_slot_floats * slots * sizeof(float), ctx->compute));
    ctx->stack_n_row = n_row;
    ctx->stack_n_col = n_col;
    ctx->stack_pitch = pitch;
    ctx->stack_slots = slots;
    ctx->tma.valid = false;
    return PARIS_B200_OK;
}

// enqueue the deferred fused weight+filter launch of the pending batch and let go of its raw buffers
static int run_pending_filter(paris_b200_ctx* ctx)
{
    if(ctx->pend_raw_count == 0)
        return PARIS_B200_OK;
    const int n = ctx->pend_raw_count;
    ctx->pend_raw_count = 0;
    const int rc = launch_filter_batch(ctx, ctx->pend_raw, nullptr, static_cast<uint32_t>(n), ctx->stack,
                                       static_cast<uint32_t>(ctx->pend_raw_first), ctx->stack_slot_floats,
                                       ctx->stack_n_row, ctx->stack_n_col, ctx->pend_filter, ctx->pend_w, true,
                                       ctx->stack_pitch, ctx->stack_layout);
    for(auto& b : ctx->pool)
    {
        if(!b.held)
            continue;
        b.held = false;
        if(b.free_pending)
        {
            b.free_pending = false;
            PB_CUDA(cudaEventRecord(b.freed, ctx->compute));
            b.freed_valid = true;
            b.in_use = false;
            ctx->free_fifo.push_back(static_cast<size_t>(&b - ctx->pool.data()));
        }
    }
    return rc;
}

extern "C" int paris_b200_flush(paris_b200_ctx* ctx)
{
    PB_CHECK_ARG(ctx != nullptr);
    if(ctx->pending == 0)
        return PARIS_B200_OK;
    PB_TRY(bind(ctx));
    PB_TRY(run_pending_filter(ctx));
    ++ctx->stat_flush;
    ++ctx->flushes_for_target;
    const int n = ctx->pending;
    ctx->pending = 0;
    return launch_backproject(ctx, ctx->stack, ctx->stack_slot_floats, ctx->stack_pitch, 0u, static_cast<uint32_t>(n),
                              ctx->pend_sin, ctx->pend_cos, ctx->target, ctx->stack_layout);
}

// a buffer a deferred launch still has to read must not be touched before that launch is enqueued
static int flush_if_held(paris_b200_ctx* ctx, const void* d_ptr)
{
    for(const auto& b : ctx->pool)
        if(b.ptr == d_ptr && b.held)
            return paris_b200_flush(ctx);
    return PARIS_B200_OK;
}

extern "C" int paris_b200_backproject(paris_b200_ctx* ctx, const float* d_proj, uint32_t dim_x, uint32_t dim_y,
                                      float* d_vol, uint32_t v_dim_x, uint32_t v_dim_y, uint32_t v_dim_z,
                                      uint32_t v_offset, const paris_b200_detector_geometry* det,
                                      const paris_b200_volume_geometry* vol_full, int enable_roi,
                                      const paris_b200_roi* roi, float sin_phi, float cos_phi, float delta_s_mm,
                                      float delta_t_mm, uint32_t flags, const paris_b200_filter* filter,
                                      const paris_b200_weighting* weighting)
{
    PB_CHECK_ARG(ctx != nullptr && d_proj != nullptr && d_vol != nullptr && det != nullptr && vol_full != nullptr);
    PB_CHECK_ARG(dim_x == det->n_row && dim_y == det->n_col);
    PB_CHECK_ARG(v_dim_x > 0 && v_dim_y > 0 && v_dim_z > 0);
    PB_CHECK_ARG(!enable_roi || roi != nullptr);
    const bool fuse = (flags & PARIS_B200_BP_FUSE_WEIGHT_FILTER) != 0u;
    PB_CHECK_ARG(!fuse || (filter != nullptr && dim_x <= filter->size));
    PB_TRY(bind(ctx));

    bp_target t{};
    t.d_vol = d_vol;
    t.v_dim_x = v_dim_x;
    t.v_dim_y = v_dim_y;
    t.v_dim_z = v_dim_z;
    t.v_offset = v_offset;
    t.det = *det;
    t.vol_full = *vol_full;
    t.enable_roi = enable_roi ? 1 : 0;
    if(enable_roi)
        t.roi = *roi;
    t.delta_s_mm = delta_s_mm;
    t.delta_t_mm = delta_t_mm;

    if(!same_target(ctx->target, t))
    {
        if(ctx->pending > 0)
            PB_TRY(paris_b200_flush(ctx));
        ctx->flushes_for_target = 0;
    }
    else
    {
        // A full batch is launched when the NEXT projection arrives, not when it fills: the last batch of a
        // scan is then still pending when the volume is read back, and vol_d2h() can backproject it slab by
        // slab with the download of each finished slab running behind the next one.
        // The first batches into a volume are short -- 16 projections, then half as many again each time -- so that
        // the backprojection starts while most of the scan is still being uploaded and never waits long for the
        // next batch to fill; full batches then amortise the volume traffic and the kernel prologue.
        int threshold = 16;
        for(int i = 0; i < ctx->flushes_for_target && threshold < ctx->bp_batch; ++i)
            threshold += threshold / 2;
        threshold = std::min(threshold, ctx->bp_batch);
        bool launch = ctx->pending >= threshold;
        if(!launch && ctx->pending >= std::min(16, ctx->bp_batch))
        {
            // upload-bound scans: if the GPU has nothing left to do, give it what has arrived (batches then stay
            // small, and so does the work left when the last projection comes in); while the GPU is busy the batch
            // keeps growing
            launch = cudaStreamQuery(ctx->compute) == cudaSuccess;
            (void)cudaGetLastError();   // cudaErrorNotReady is not an error
        }
        if(launch)
            PB_TRY(paris_b200_flush(ctx));
    }
    // (the layout is a function of the target geometry, so it is constant within a batch)
    PB_TRY(ensure_stack(ctx, dim_x, dim_y, choose_stack_layout(*det, *vol_full)));
    ctx->target = t;

    float* slot = ctx->stack + ctx->stack_slot_floats * static_cast<size_t>(ctx->pending);
    touch(ctx, d_proj);
    if(fuse)
    {
        weight_params w = weighting_from_detector(det);
        if(weighting != nullptr)
        {
            w.h_min = weighting->h_min;
            w.v_min = weighting->v_min;
            w.d_sd = weighting->d_sd;
            w.l_px_row = weighting->l_px_row;
            w.l_px_col = weighting->l_px_col;
        }
        pb::raw_buffer* buf = find_buffer(ctx, d_proj);
        if(buf != nullptr && buf->in_use)
        {
            // Pooled buffer: defer the fused weight+filter kernel so that the whole batch is ONE launch
            // (filter.cu).  The deferred run must be a consecutive range of slots with one filter and one
            // set of weighting scalars.
            const bool extends = ctx->pend_raw_count > 0 && ctx->pend_filter == filter
                              && std::memcmp(&ctx->pend_w, &w, sizeof(w)) == 0
                              && ctx->pend_raw_first + ctx->pend_raw_count == ctx->pending;
            if(ctx->pend_raw_count > 0 && !extends)
                PB_TRY(run_pending_filter(ctx));
            if(ctx->pend_raw_count == 0)
            {
                ctx->pend_raw_first = ctx->pending;
                ctx->pend_filter = filter;
                ctx->pend_w = w;
            }
            ctx->pend_raw[ctx->pend_raw_count++] = d_proj;
            buf->held = true;
        }
        else
            PB_TRY(launch_filter(ctx, d_proj, slot, dim_x, dim_y, filter, w, true, ctx->stack_pitch, ctx->stack_layout));
    }
    else
    {
        PB_TRY(flush_if_held(ctx, d_proj));
        PB_TRY(launch_transpose_to_slot(ctx, d_proj, slot, dim_x, dim_y, ctx->stack_pitch, ctx->stack_layout));
    }

    ctx->pend_sin[ctx->pending] = sin_phi;
    ctx->pend_cos[ctx->pending] = cos_phi;
    ++ctx->pending;
    return PARIS_B200_OK;
}

// ---- stack-level entry points -----------------------------------------------------------------------------

extern "C" int paris_b200_stack_alloc(paris_b200_ctx* ctx, uint32_t n_row, uint32_t n_col, uint32_t slots, float** d_stack)
{
    PB_CHECK_ARG(ctx != nullptr && d_stack != nullptr && n_row > 0 && n_col > 0 && slots > 0);
    PB_TRY(bind(ctx));
    const size_t bytes = static_cast<size_t>(stack_pitch_for(n_col)) * n_row * slots * sizeof(float);
    *d_stack = nullptr;
    PB_CUDA(cudaMalloc(reinterpret_cast<void**>(d_stack), bytes));
    PB_CUDA(cudaMemsetAsync(*d_stack, 0, bytes, ctx->compute));
    return PARIS_B200_OK;
}

extern "C" int paris_b200_stack_free(paris_b200_ctx* ctx, float* d_stack)
{
    PB_CHECK_ARG(ctx != nullptr);
    if(d_stack == nullptr)
        return PARIS_B200_OK;
    PB_TRY(bind(ctx));
    PB_CUDA(cudaStreamSynchronize(ctx->compute));
    PB_CUDA(cudaFree(d_stack));
    return PARIS_B200_OK;
}

extern "C" int paris_b200_filter_to_stack_batch(paris_b200_ctx* ctx, const float* d_raw, size_t raw_stride,
                                                uint32_t count, const paris_b200_detector_geometry* det,
                                                const paris_b200_filter* filter, float* d_stack, uint32_t first_slot,
                                                uint32_t layout)
{
    PB_CHECK_ARG(layout == kLayoutPlain || layout == kLayoutSplit2);
    PB_CHECK_ARG(ctx != nullptr && d_raw != nullptr && det != nullptr && filter != nullptr && d_stack != nullptr);
    PB_CHECK_ARG(det->n_row <= filter->size);
    PB_CHECK_ARG(count <= 1u || raw_stride >= static_cast<size_t>(det->n_row) * det->n_col);
    PB_TRY(bind(ctx));
    const uint32_t pitch = stack_pitch_for(det->n_col);
    const size_t slot_floats = static_cast<size_t>(pitch) * det->n_row;
    const weight_params w = weighting_from_detector(det);
    const float* src[kMaxBatch];
    for(uint32_t done = 0; done < count;)
    {
        const uint32_t n = std::min<uint32_t>(count - done, static_cast<uint32_t>(kMaxBatch));
        for(uint32_t i = 0; i < n; ++i)
            src[i] = d_raw + raw_stride * (done + i);
        PB_TRY(launch_filter_batch(ctx, src, nullptr, n, d_stack, first_slot + done, slot_floats, det->n_row, det->n_col,
                                   filter, w, true, pitch, layout));
        done += n;
    }
    return PARIS_B200_OK;
}

// the same for detector-native 16-bit samples (raw_stride in samples): widened to float by the kernel's first load
extern "C" int paris_b200_filter_to_stack_batch_u16(paris_b200_ctx* ctx, const uint16_t* d_raw, size_t raw_stride,
                                                    uint32_t count, const paris_b200_detector_geometry* det,
                                                    const paris_b200_filter* filter, float* d_stack, uint32_t first_slot,
                                                    uint32_t layout)
{
    PB_CHECK_ARG(layout == kLayoutPlain || layout == kLayoutSplit2);
    PB_CHECK_ARG(ctx != nullptr && d_raw != nullptr && det != nullptr && filter != nullptr && d_stack != nullptr);
    PB_CHECK_ARG(det->n_row <= filter->size);
    PB_CHECK_ARG(count <= 1u || raw_stride >= static_cast<size_t>(det->n_row) * det->n_col);
    PB_TRY(bind(ctx));
    const uint32_t pitch = stack_pitch_for(det->n_col);
    const size_t slot_floats = static_cast<size_t>(pitch) * det->n_row;
    const weight_params w = weighting_from_detector(det);
    const float* src[kMaxBatch];
    for(uint32_t done = 0; done < count;)
    {
        const uint32_t n = std::min<uint32_t>(count - done, static_cast<uint32_t>(kMaxBatch));
        for(uint32_t i = 0; i < n; ++i)
            src[i] = reinterpret_cast<const float*>(d_raw + raw_stride * (done + i));
        PB_TRY(launch_filter_batch(ctx, src, nullptr, n, d_stack, first_slot + done, slot_floats, det->n_row, det->n_col,
                                   filter, w, true, pitch, layout, true));
        done += n;
    }
    return PARIS_B200_OK;
}

extern "C" int paris_b200_filter_to_stack(paris_b200_ctx* ctx, const float* d_raw,
                                          const paris_b200_detector_geometry* det, const paris_b200_filter* filter,
                                          float* d_stack, uint32_t slot, uint32_t layout)
{
    PB_CHECK_ARG(ctx != nullptr);
    PB_TRY(flush_if_held(ctx, d_raw));
    touch(ctx, d_raw);
    return paris_b200_filter_to_stack_batch(ctx, d_raw, 0, 1u, det, filter, d_stack, slot, layout);
}

extern "C" int paris_b200_backproject_stack(paris_b200_ctx* ctx, const float* d_stack, uint32_t first, uint32_t count,
                                            const float* sin_phi, const float* cos_phi, float* d_vol,
                                            uint32_t v_dim_x, uint32_t v_dim_y, uint32_t v_dim_z, uint32_t v_offset,
                                            const paris_b200_detector_geometry* det,
                                            const paris_b200_volume_geometry* vol_full, int enable_roi,
                                            const paris_b200_roi* roi, uint32_t layout)
{
    PB_CHECK_ARG(ctx != nullptr && d_stack != nullptr && d_vol != nullptr && det != nullptr && vol_full != nullptr);
    PB_CHECK_ARG(layout == kLayoutPlain || layout == kLayoutSplit2);
    PB_CHECK_ARG(sin_phi != nullptr && cos_phi != nullptr);
    PB_CHECK_ARG(!enable_roi || roi != nullptr);
    PB_TRY(bind(ctx));
    PB_TRY(paris_b200_flush(ctx));
    bp_target t{};
    t.d_vol = d_vol;
    t.v_dim_x = v_dim_x;
    t.v_dim_y = v_dim_y;
    t.v_dim_z = v_dim_z;
    t.v_offset = v_offset;
    t.det = *det;
    t.vol_full = *vol_full;
    t.enable_roi = enable_roi ? 1 : 0;
    if(enable_roi)
        t.roi = *roi;
    // src/backprojection.cpp:49-50
    t.delta_s_mm = det->delta_s * det->l_px_row;
    t.delta_t_mm = det->delta_t * det->l_px_col;
    const uint32_t pitch = stack_pitch_for(det->n_col);
    const size_t slot_floats = static_cast<size_t>(pitch) * det->n_row;
    // launches of (nearly) equal size: 272 projections go as 136 + 136, not 256 + 16
    const uint32_t launches = (count + static_cast<uint32_t>(ctx->bp_batch) - 1u) / static_cast<uint32_t>(ctx->bp_batch);
    const uint32_t per_launch = launches > 0u ? (count + launches - 1u) / launches : 0u;
    for(uint32_t done = 0; done < count;)
    {
        const uint32_t n = std::min<uint32_t>(count - done, per_launch);
        PB_TRY(launch_backproject(ctx, d_stack, slot_floats, pitch, first + done, n, sin_phi + done, cos_phi + done, t,
                                  layout));
        done += n;
    }
    return PARIS_B200_OK;
}

extern "C" int paris_b200_backproject_stack_d2h(paris_b200_ctx* ctx, const float* d_stack, uint32_t first, uint32_t count,
                                                const float* sin_phi, const float* cos_phi, float* d_vol,
                                                uint32_t v_dim_x, uint32_t v_dim_y, uint32_t v_dim_z, uint32_t v_offset,
                                                const paris_b200_detector_geometry* det,
                                                const paris_b200_volume_geometry* vol_full, int enable_roi,
                                                const paris_b200_roi* roi, uint32_t layout, float* h_dst)
{
    PB_CHECK_ARG(ctx != nullptr && d_stack != nullptr && d_vol != nullptr && det != nullptr && vol_full != nullptr);
    PB_CHECK_ARG(layout == kLayoutPlain || layout == kLayoutSplit2);
    PB_CHECK_ARG(sin_phi != nullptr && cos_phi != nullptr && h_dst != nullptr);
    PB_CHECK_ARG(!enable_roi || roi != nullptr);
    PB_TRY(bind(ctx));
    PB_TRY(paris_b200_flush(ctx));
    bp_target t{};
    t.d_vol = d_vol;
    t.v_dim_x = v_dim_x;
    t.v_dim_y = v_dim_y;
    t.v_dim_z = v_dim_z;
    t.v_offset = v_offset;
    t.det = *det;
    t.vol_full = *vol_full;
    t.enable_roi = enable_roi ? 1 : 0;
    if(enable_roi)
        t.roi = *roi;
    t.delta_s_mm = det->delta_s * det->l_px_row;   // src/backprojection.cpp:49-50
    t.delta_t_mm = det->delta_t * det->l_px_col;
    const uint32_t pitch = stack_pitch_for(det->n_col);
    const size_t slot_floats = static_cast<size_t>(pitch) * det->n_row;
    return backproject_and_download(ctx, d_stack, slot_floats, pitch, first, count, sin_phi, cos_phi, t, layout, h_dst, true, 0u);
}

extern "C" int paris_b200_phantom_project(paris_b200_ctx* ctx, const double* ellipsoids, uint32_t n_ellipsoids,
                                          const paris_b200_detector_geometry* det, uint32_t first_idx,
                                          uint32_t n_proj, float* d_stack_raw)
{
    PB_CHECK_ARG(ctx != nullptr && ellipsoids != nullptr && det != nullptr && d_stack_raw != nullptr);
    PB_CHECK_ARG(n_ellipsoids >= 1 && n_ellipsoids <= 64);
    PB_TRY(bind(ctx));
    return launch_phantom(ctx, ellipsoids, n_ellipsoids, det, first_idx, n_proj, d_stack_raw);
}
