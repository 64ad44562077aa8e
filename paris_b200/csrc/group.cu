// group.cu -- one scan reconstructed by the GPUs of one box: the multi-GPU step of SURVEY 8(e) in C++ behind the C ABI.
//
// The reference's multi-device scheme is task parallelism with nothing exchanged: every device re-reads, re-weights
// and re-filters EVERY projection for each slab it owns (/root/reference/src/main.cpp:93-105, slab arithmetic
// src/cuda/subvolume_information.cpp:112-116, src/make_volume.cpp:32-34, src/main.cpp:96).  Here a GROUP of `world`
// members (one per GPU; one process each, or several host threads of one process) shares the work:
//
//   member r   uploads and filters 1/world of the projections (fused weight+filter kernel, filter.cu), cut into
//              ROUNDS of consecutive projections so that the backprojection starts early;
//   exchange   after every round each member copies what it filtered straight into its peers' stacks over NVLink --
//              peer memory, copy engines (cudaMemcpy2DAsync on a stream of its own: no SM-resident collective next
//              to the backprojection) -- and only the BAND of detector rows the peer's slabs can ever read (the
//              scheme sketched in doc "Geometrie - Definitionen fuer Subvolumen", never implemented in the reference):
//              ~1/world of the bytes of an all-gather.  Arrival is announced by a stream memory operation
//              (cuStreamWriteValue32) on a flag word in the peer's memory; the consumer's backprojection stream waits
//              on its own flags with a one-thread polling kernel that gives up after a timeout (wait_flag_kernel).
//              No host synchronisation anywhere inside a step;
//   member r   backprojects ALL projections into its z-slabs (no reduction) and streams them to the host: several
//              slabs per member loop over the ONE gathered stack, the download of slab k running behind the
//              backprojection of slab k + 1; the host volume is assembled by writing every slab at its offset.
//
// Slabs are bit-identical crops of the one-piece result (tile anchors are global, backproject_tma.cu), so the
// decomposition changes no voxel.
#include "common.cuh"
#include "backproject.cuh"

#include <cudaTypedefs.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace pb
{
    struct round_t
    {
        uint32_t first, count;   // projections [first, first + count) of the scan
    };

    // member r's share of a round: consecutive projections, as even as possible
    static void share_of(const round_t& rd, uint32_t world, uint32_t r, uint32_t* first, uint32_t* count)
    {
        const uint32_t base = rd.count / world, extra = rd.count % world;
        *count = base + (r < extra ? 1u : 0u);
        *first = rd.first + r * base + std::min(r, extra);
    }

    struct slab_t
    {
        uint32_t z_first, dz;    // in slices of the region
    };

    struct band_t
    {
        uint32_t lo, hi;         // detector rows [lo, hi), multiples of 8
    };

    // everything a peer needs to reach a member's memory (fixed size, exchanged by the host side)
    struct group_handle
    {
        uint32_t magic, rank;
        int32_t pid, device;
        uint64_t stack_ptr, flags_ptr;              // valid inside process `pid`
        uint64_t stack_off, flags_off;              // offsets of those pointers inside their allocations
        cudaIpcMemHandle_t stack_ipc, flags_ipc;    // for every other process
    };
    static_assert(sizeof(group_handle) <= PARIS_B200_GROUP_HANDLE_BYTES, "handle blob too small");
    constexpr uint32_t kHandleMagic = 0x50423247u;  // "PB2G"
    constexpr uint32_t kErrorWord = 1024u;          // flags[kErrorWord]: 0, or 1 + the member a wait gave up on

    using stream_value32_fn = CUresult (*)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);

    static stream_value32_fn driver_entry(const char* name)
    {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q{};
        if(cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<stream_value32_fn>(p);
    }

    // one word written into (peer) memory from a kernel: the fallback where stream memory operations refuse the address
    __global__ void flag_store_kernel(uint32_t* flag, uint32_t value)
    {
        __threadfence_system();
        *reinterpret_cast<volatile uint32_t*>(flag) = value;
        __threadfence_system();
    }

    // The consumer's side of an arrival flag: ONE thread polls the word (flags only grow: cyclic comparison) and the
    // stream continues when it has reached `value`.  A kernel rather than cuStreamWaitValue32 for two reasons: it
    // cannot wait for ever -- after `timeout_ns` it records the failure in *error and lets the stream go on, so a
    // member that died takes the step down with an error instead of hanging the GPU -- and it holds nothing but one
    // warp slot, where a memory-operation wait occupies the hardware queue its stream is mapped to together with
    // whatever other stream shares that queue (with members sharing one GPU that can be the exchange stream whose
    // signal a peer is waiting for).  PARIS_B200_GROUP_WAIT=memop selects the memory operation for comparison.
    __global__ void wait_flag_kernel(const uint32_t* flag, uint32_t value, unsigned long long timeout_ns, uint32_t* error,
                                     uint32_t code)
    {
        const volatile uint32_t* f = flag;
        unsigned long long t0 = 0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while(static_cast<int32_t>(*f - value) < 0)
        {
            __nanosleep(256);
            unsigned long long t = 0;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if(t - t0 > timeout_ns)
            {
                atomicMax(error, code);
                break;
            }
        }
        __threadfence_system();
    }

    // rows [lo, hi) of `lines` stack lines, 16 bytes per thread and step (the SM-driven form of the exchange;
    // the default is the copy engine)
    __global__ void push_band_kernel(const float4* __restrict__ src, float4* __restrict__ dst, uint32_t pitch4, uint32_t lo4,
                                     uint32_t width4, uint32_t lines)
    {
        const uint64_t total = static_cast<uint64_t>(width4) * lines;
        for(uint64_t e = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; e < total;
            e += static_cast<uint64_t>(gridDim.x) * blockDim.x)
        {
            const uint64_t line = e / width4, c = e % width4;
            const uint64_t idx = line * pitch4 + lo4 + c;
            dst[idx] = src[idx];
        }
    }
}

using namespace pb;

struct paris_b200_group
{
    paris_b200_group_config cfg{};
    paris_b200_ctx* ctx = nullptr;    // backprojection (compute stream) and downloads (copy stream)
    paris_b200_ctx* fctx = nullptr;   // uploads (copy stream) and the fused weight+filter kernel (compute stream)
    cudaStream_t push = nullptr;      // the exchange: copies into peer memory + arrival signals
    paris_b200_filter* filter = nullptr;

    uint32_t region_x = 0, region_y = 0, region_z = 0, region_z0 = 0;   // region dims; first slice in the full volume
    uint32_t x_first = 0, x_count = 0;          // this member's columns of the region
    uint32_t host_row = 0;                      // floats per row of the host destination
    uint32_t layout = 0, pitch = 0;
    size_t slot_floats = 0, px = 0;
    size_t sample_bytes = 4;                    // of a raw sample: 4 (float) or 2 (detector-native counts)
    std::vector<round_t> rounds;
    std::vector<uint32_t> local_first;          // index of a round's first projection among this member's own
    uint32_t my_count = 0;
    std::vector<slab_t> slabs;                  // this member's slabs, consecutive in z
    std::vector<band_t> bands;                  // per member: the rows its slabs can read
    std::vector<float> sn, cs;                  // per projection of the scan

    float* stack = nullptr;                     // n_proj slots
    uint32_t* flags = nullptr;                  // [0, world): arrived[src]; [world, 2 world): consumed[dst]
    std::vector<float*> peer_stack;
    std::vector<uint32_t*> peer_flags;
    std::vector<bool> peer_ipc;
    bool connected = false;

    float* raw[2] = {nullptr, nullptr};         // upload buffers, one round each (floats or 16-bit counts)
    size_t raw_bytes = 0;
    cudaEvent_t raw_free[2] = {nullptr, nullptr};      // the filter launch that read raw[i] has run
    bool raw_free_valid[2] = {false, false};
    std::vector<cudaEvent_t> filtered;          // per round: this member's share is in the stack
    std::vector<cudaEvent_t> uploaded;          // per round: this member's uploads have left their host buffers
    cudaEvent_t step_done = nullptr, pushed = nullptr;
    // PARIS_B200_GROUP_TRACE=1: per round, when the member's share was filtered, when it had been pushed to every
    // peer, when the round's backprojection finished -- milliseconds from the step's start, printed by group_end
    bool trace = false;
    cudaEvent_t trace_t0 = nullptr;
    std::vector<cudaEvent_t> trace_filtered, trace_pushed, trace_bp;
    bool step_open = false;                     // between step_open and step_finish
    uint32_t next_round = 0;
    float* step_h_slabs = nullptr;
    std::vector<float*> vol;                    // slab buffers on the device
    std::vector<cudaEvent_t> slab_down;         // per slab buffer: its slab has reached the host
    std::vector<bool> slab_down_valid;
    uint32_t steps = 0;                         // completed + begun steps (sequence numbers derive from it)
    bool in_step = false;
    cudaStream_t side = nullptr;                // diagnostics only
    uint32_t* h_debug = nullptr;
    bool memops_ok = true;
    bool wait_memops = false;                   // waits as stream memory operations instead of polling kernels (A/B only)
    unsigned long long wait_timeout_ns = 30ull * 1000ull * 1000ull * 1000ull;
    uint64_t bytes_pushed = 0;
};

namespace
{
    int bind(const paris_b200_group* g) { return paris_b200_ctx_bind(g->ctx); }

    // src/backprojection.cpp:53-63 (float arithmetic, libm sin/cos)
    void angle_sin_cos(float phi_deg, float* s, float* c)
    {
        const float phi = phi_deg * (static_cast<float>(M_PI) / 180.f);
        *s = std::sin(phi);
        *c = std::cos(phi);
    }

    // Detector rows the voxels of slices [z_first, z_first + dz) of the region can read at any angle:
    // v = (z_m * factor - min_v) / l_px - 0.5 with factor = d_sd / (s + d_so) and |s| <= the region's largest distance
    // from the rotation axis (src/openmp/backprojection.cpp:120-133).  Conservative: tiles stick out of the region by
    // less than one tile in x and y (their extra voxels are computed and dropped), four rows of slack cover the
    // reference's own float rounding of v and the + 1 neighbour; rounded outwards to multiples of 8 rows (16-byte
    // runs in either parity plane).
    band_t band_of(const paris_b200_group_config& c, uint32_t rx, uint32_t ry, uint32_t z_global_first, uint32_t dz,
                   uint32_t pitch)
    {
        const auto& v = c.vol_full;
        const uint32_t x1 = c.enable_roi ? c.roi.x1 : 0u, y1 = c.enable_roi ? c.roi.y1 : 0u;
        auto centred = [](double i, double dim, double size) { return -(dim * size / 2.0) + size / 2.0 + i * size; };
        constexpr double kTile = 16.0;   // the widest tile: voxels computed beyond the region's border
        double r = 0.0;
        for(int cx = 0; cx < 2; ++cx)
            for(int cy = 0; cy < 2; ++cy)
            {
                const double x = centred(cx ? x1 + rx - 1.0 + kTile : x1 - kTile, v.dim_x, v.l_vx_x);
                const double y = centred(cy ? y1 + ry - 1.0 + kTile : y1 - kTile, v.dim_y, v.l_vx_y);
                r = std::max(r, std::hypot(x, y));
            }
        const double d_so = c.det.d_so, d_sd = std::fabs(static_cast<double>(c.det.d_so)) + std::fabs(static_cast<double>(c.det.d_od));
        band_t b{0u, pitch};
        if(!(d_so - r > 0.02 * d_so))
            return b;   // (source almost inside the region: no useful bound)
        const double f_lo = d_sd / (d_so + r), f_hi = d_sd / (d_so - r);
        const double l_px = c.det.l_px_col;
        const double min_v = -(c.det.n_col * l_px / 2.0) - static_cast<double>(c.det.delta_t) * l_px;
        double lo = 1e300, hi = -1e300;
        for(int cz = 0; cz < 2; ++cz)
        {
            const double z = centred(cz ? z_global_first + dz - 1.0 : z_global_first, v.dim_z, v.l_vx_z);
            for(const double f : {f_lo, f_hi})
            {
                const double row = (z * f - min_v) / l_px - 0.5;
                lo = std::min(lo, row);
                hi = std::max(hi, row);
            }
        }
        const double last = static_cast<double>(c.det.n_col) - 1.0;
        const double lo_c = std::min(std::max(std::floor(lo) - 4.0, 0.0), last);
        const double hi_c = std::min(std::max(std::floor(hi) + 1.0 + 4.0, 0.0), last);
        b.lo = (static_cast<uint32_t>(lo_c) / 8u) * 8u;
        b.hi = std::min<uint32_t>(((static_cast<uint32_t>(hi_c) + 1u + 7u) / 8u) * 8u, pitch);
        return b;
    }

    int signal32(paris_b200_group* g, cudaStream_t s, uint32_t* d_flag, uint32_t value)
    {
        static stream_value32_fn write32 = driver_entry("cuStreamWriteValue32");
        if(g->memops_ok && write32 != nullptr)
        {
            const CUresult r = write32(reinterpret_cast<CUstream>(s), reinterpret_cast<CUdeviceptr>(d_flag), value,
                                       CU_STREAM_WRITE_VALUE_DEFAULT);
            if(r == CUDA_SUCCESS)
                return PARIS_B200_OK;
            g->memops_ok = false;   // this driver / address space refuses: one-word kernels from here on
        }
        flag_store_kernel<<<1, 1, 0, s>>>(d_flag, value);
        PB_CUDA(cudaGetLastError());
        return PARIS_B200_OK;
    }

    // the stream continues once *d_flag >= value (see wait_flag_kernel); who = the member waited for (error report)
    int wait32(paris_b200_group* g, cudaStream_t s, const uint32_t* d_flag, uint32_t value, uint32_t who)
    {
        if(!g->wait_memops)
        {
            wait_flag_kernel<<<1, 1, 0, s>>>(d_flag, value, g->wait_timeout_ns, g->flags + kErrorWord, who + 1u);
            PB_CUDA(cudaGetLastError());
            return PARIS_B200_OK;
        }
        static stream_value32_fn wait32_fn = driver_entry("cuStreamWaitValue32");
        if(wait32_fn == nullptr)
        {
            set_error("cuStreamWaitValue32 is not available from the driver");
            return PARIS_B200_ECUDA;
        }
        const CUresult r = wait32_fn(reinterpret_cast<CUstream>(s), reinterpret_cast<CUdeviceptr>(const_cast<uint32_t*>(d_flag)),
                                     value, CU_STREAM_WAIT_VALUE_GEQ);
        if(r != CUDA_SUCCESS)
        {
            set_error("cuStreamWaitValue32 failed with CUresult %d", static_cast<int>(r));
            return PARIS_B200_ECUDA;
        }
        return PARIS_B200_OK;
    }

    bp_target target_of(const paris_b200_group* g, const slab_t& s, float* d_vol)
    {
        bp_target t{};
        t.d_vol = d_vol;
        t.v_dim_x = g->x_count;
        t.v_dim_y = g->region_y;
        t.v_dim_z = s.dz;
        t.v_offset = s.z_first;
        t.det = g->cfg.det;
        t.vol_full = g->cfg.vol_full;
        t.enable_roi = g->cfg.enable_roi ? 1 : 0;
        if(t.enable_roi)
            t.roi = g->cfg.roi;
        if(g->x_first != 0u)
        {
            // a member that owns only part of the columns: the kernel shifts voxel indices by (roi.x1, roi.y1, roi.z1)
            // (src/openmp/backprojection.cpp:105-109), so the part is one more shift of x
            if(!t.enable_roi)
                t.roi = paris_b200_roi{0u, 0u, 0u, 0u, 0u, 0u};
            t.enable_roi = 1;
            t.roi.x1 += g->x_first;
        }
        t.delta_s_mm = g->cfg.det.delta_s * g->cfg.det.l_px_row;   // src/backprojection.cpp:49-50
        t.delta_t_mm = g->cfg.det.delta_t * g->cfg.det.l_px_col;
        return t;
    }

    // projections [first, first + count) of the stack into one slab, in launches of (nearly) equal size
    int backproject_range(paris_b200_group* g, uint32_t first, uint32_t count, const slab_t& s, float* d_vol)
    {
        const bp_target t = target_of(g, s, d_vol);
        const uint32_t batch = static_cast<uint32_t>(g->ctx->bp_batch);
        const uint32_t launches = (count + batch - 1u) / batch;
        const uint32_t per_launch = launches > 0u ? (count + launches - 1u) / launches : 0u;
        for(uint32_t done = 0; done < count;)
        {
            const uint32_t n = std::min(count - done, per_launch);
            PB_TRY(launch_backproject(g->ctx, g->stack, g->slot_floats, g->pitch, first + done, n, g->sn.data() + first + done,
                                      g->cs.data() + first + done, t, g->layout));
            done += n;
        }
        return PARIS_B200_OK;
    }

    // this member's share of round rd, filtered into its own stack, goes to every peer (band rows only)
    int push_round(paris_b200_group* g, uint32_t rd, uint32_t seq)
    {
        const uint32_t world = static_cast<uint32_t>(g->cfg.world), me = static_cast<uint32_t>(g->cfg.rank);
        uint32_t first = 0, count = 0;
        share_of(g->rounds[rd], world, me, &first, &count);
        PB_CUDA(cudaStreamWaitEvent(g->push, g->filtered[rd], 0));
        const size_t line_bytes = static_cast<size_t>(g->pitch) * sizeof(float);
        const size_t lines = static_cast<size_t>(g->cfg.det.n_row) * count;
        for(uint32_t step = 1; step < world; ++step)
        {
            const uint32_t k = (me + step) % world;   // (members start with different peers)
            const band_t b = g->bands[k];
            if(count > 0 && b.hi > b.lo)
            {
                const float* src = g->stack + g->slot_floats * first;
                float* dst = g->peer_stack[k] + g->slot_floats * first;
                const uint32_t planes = g->layout == kLayoutSplit2 ? 2u : 1u;
                for(uint32_t p = 0; p < planes; ++p)
                {
                    // split layout: even rows, then odd rows, half a line each; the band is [lo/2, hi/2) in both
                    const size_t off = planes == 2u ? p * (g->pitch / 2u) + b.lo / 2u : b.lo;
                    const size_t width = (planes == 2u ? (b.hi - b.lo) / 2u : b.hi - b.lo) * sizeof(float);
                    if(g->cfg.exchange == PARIS_B200_EXCHANGE_KERNEL)
                    {
                        const uint32_t blocks = static_cast<uint32_t>(std::min<uint64_t>((width / 16u * lines + 255u) / 256u, 16u));
                        push_band_kernel<<<blocks, 256, 0, g->push>>>(reinterpret_cast<const float4*>(src),
                                                                     reinterpret_cast<float4*>(dst), g->pitch / 4u,
                                                                     static_cast<uint32_t>(off / 4u),
                                                                     static_cast<uint32_t>(width / 16u),
                                                                     static_cast<uint32_t>(lines));
                        PB_CUDA(cudaGetLastError());
                    }
                    else
                        PB_CUDA(cudaMemcpy2DAsync(dst + off, line_bytes, src + off, line_bytes, width, lines,
                                                  cudaMemcpyDeviceToDevice, g->push));
                    g->bytes_pushed += width * lines;
                }
            }
            PB_TRY(signal32(g, g->push, g->peer_flags[k] + me, seq));
        }
        return PARIS_B200_OK;
    }
}

extern "C" size_t paris_b200_group_handle_bytes(void) { return PARIS_B200_GROUP_HANDLE_BYTES; }

// Who filters what, who owns which slices, which detector rows travel to whom: host arithmetic only (no device is
// touched), so the decomposition can be checked on a machine without GPUs.
extern "C" int paris_b200_group_plan(const paris_b200_group_config* cfg, paris_b200_group_plan_t* plan)
{
    PB_CHECK_ARG(cfg != nullptr && plan != nullptr);
    PB_CHECK_ARG(cfg->world >= 1 && cfg->world <= PARIS_B200_GROUP_MAX_MEMBERS && cfg->rank >= 0 && cfg->rank < cfg->world);
    PB_CHECK_ARG(cfg->n_proj >= 1 && cfg->det.n_row > 0 && cfg->det.n_col > 0);
    std::memset(plan, 0, sizeof(*plan));
    // region and its slabs (src/cuda/subvolume_information.cpp:112-116: dim_z / num, remainder on the last)
    paris_b200_volume_geometry region = cfg->vol_full;
    if(cfg->enable_roi)
        PB_TRY(paris_b200_apply_roi(&cfg->vol_full, &cfg->roi, &region));
    plan->region_x = region.dim_x;
    plan->region_y = region.dim_y;
    plan->region_z = region.dim_z;
    plan->region_z0 = cfg->enable_roi ? cfg->roi.z1 : 0u;
    const uint32_t world = static_cast<uint32_t>(cfg->world);
    const uint32_t spr = std::max(1u, cfg->slabs_per_member);
    const uint32_t xp = std::max(1u, cfg->x_parts);
    if(world % xp != 0u || xp > region.dim_x)
    {
        set_error("x_parts = %u must divide the %u members and not exceed the region's %u columns", xp, world, region.dim_x);
        return PARIS_B200_EINVAL;
    }
    const uint32_t world_z = world / xp;           // members along z
    const uint32_t total = world_z * spr;
    plan->x_parts = xp;
    plan->x_dx = region.dim_x / xp;
    plan->x_remainder = region.dim_x % xp;
    if(total > region.dim_z)
    {
        set_error("%u slabs for %u slices: every slab needs at least one slice", total, region.dim_z);
        return PARIS_B200_EINVAL;
    }
    plan->slabs_total = total;
    plan->slab_dz = region.dim_z / total;
    plan->slab_remainder = region.dim_z % total;
    plan->layout = choose_stack_layout(cfg->det, cfg->vol_full);
    plan->pitch = stack_pitch_for(cfg->det.n_col);

    // rounds: short ones first so that the backprojection starts early, long ones afterwards (few large exchanges
    // overlap with the backprojection far better than many small ones)
    const uint32_t first_round = cfg->first_round ? cfg->first_round : 64u;
    const uint32_t max_round = cfg->max_round ? cfg->max_round : static_cast<uint32_t>(kMaxBatch);
    uint32_t remaining = cfg->n_proj, cap = std::max(first_round, world), at = 0;
    while(remaining > 0)
    {
        uint32_t c = std::min(remaining, cap);
        if(remaining - c < std::max(1u, c / 4u))
            c = remaining;   // do not leave a sliver for the last round
        if(plan->rounds == PARIS_B200_GROUP_MAX_ROUNDS)
        {
            set_error("more than %d exchange rounds: raise first_round / max_round", PARIS_B200_GROUP_MAX_ROUNDS);
            return PARIS_B200_EINVAL;
        }
        plan->round_first[plan->rounds] = at;
        plan->round_count[plan->rounds] = c;
        ++plan->rounds;
        at += c;
        remaining -= c;
        cap = std::min(std::max(max_round, world), cap * 2u);
    }

    // bands: the detector rows each member's slabs can read
    for(uint32_t k = 0; k < world; ++k)
    {
        const uint32_t kz = k / xp;                // (members that share a z-run receive the same band)
        const uint32_t zf = kz * spr * plan->slab_dz;
        const uint32_t zn = (kz == world_z - 1u) ? region.dim_z - zf : spr * plan->slab_dz;
        const band_t b = cfg->whole_projections ? band_t{0u, plan->pitch}
                                                : band_of(*cfg, region.dim_x, region.dim_y, plan->region_z0 + zf, zn, plan->pitch);
        plan->band_lo[k] = b.lo;
        plan->band_hi[k] = b.hi;
    }
    return PARIS_B200_OK;
}

// member `member`'s share of round `round`: projections [first, first + count) of the scan
extern "C" int paris_b200_group_share(const paris_b200_group_plan_t* plan, uint32_t world, uint32_t round, uint32_t member,
                                      uint32_t* first, uint32_t* count)
{
    PB_CHECK_ARG(plan != nullptr && first != nullptr && count != nullptr && world >= 1 && member < world && round < plan->rounds);
    share_of(round_t{plan->round_first[round], plan->round_count[round]}, world, member, first, count);
    return PARIS_B200_OK;
}

extern "C" int paris_b200_group_create(int device, const paris_b200_group_config* cfg, paris_b200_group** out)
{
    PB_CHECK_ARG(cfg != nullptr && out != nullptr);
    *out = nullptr;
    PB_CHECK_ARG(cfg->world >= 1 && cfg->world <= PARIS_B200_GROUP_MAX_MEMBERS && cfg->rank >= 0 && cfg->rank < cfg->world);
    PB_CHECK_ARG(cfg->n_proj >= 1 && cfg->det.n_row > 0 && cfg->det.n_col > 0);
    PB_CHECK_ARG(cfg->sample_type == PARIS_B200_SAMPLES_F32 || cfg->sample_type == PARIS_B200_SAMPLES_U16);
    PB_CHECK_ARG(cfg->exchange == PARIS_B200_EXCHANGE_COPY_ENGINE || cfg->exchange == PARIS_B200_EXCHANGE_KERNEL);

    auto* g = new paris_b200_group{};
    g->cfg = *cfg;
    g->cfg.angles_deg = nullptr;   // (copied below; the caller's array need not outlive this call)
    const auto fail = [&](int rc) {
        paris_b200_group_destroy(g);
        return rc;
    };
#define PB_GTRY(expr)                        \
    do                                       \
    {                                        \
        const int rc_ = (expr);              \
        if(rc_ != PARIS_B200_OK)             \
            return fail(rc_);                \
    } while(0)
#define PB_GCUDA(expr)                                                                                       \
    do                                                                                                       \
    {                                                                                                        \
        const cudaError_t e_ = (expr);                                                                       \
        if(e_ != cudaSuccess)                                                                                \
        {                                                                                                    \
            set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__);           \
            return fail(e_ == cudaErrorMemoryAllocation ? PARIS_B200_ENOMEM : PARIS_B200_ECUDA);             \
        }                                                                                                    \
    } while(0)

    if(const char* e = std::getenv("PARIS_B200_GROUP_WAIT"))
        g->wait_memops = std::strcmp(e, "memop") == 0;
    if(const char* e = std::getenv("PARIS_B200_GROUP_TRACE"))
        g->trace = std::atoi(e) != 0;
    if(const char* e = std::getenv("PARIS_B200_GROUP_TIMEOUT_S"))
        g->wait_timeout_ns = static_cast<unsigned long long>(std::max(1.0, std::atof(e)) * 1e9);
    PB_GTRY(paris_b200_ctx_create(device, &g->ctx));
    PB_GTRY(paris_b200_ctx_create(device, &g->fctx));
    // the exchange outranks everything else on the device (its SM-driven form must not queue behind a backprojection)
    int prio_lo = 0, prio_hi = 0;
    PB_GCUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    PB_GCUDA(cudaStreamCreateWithPriority(&g->push, cudaStreamNonBlocking, prio_hi));
    PB_GCUDA(cudaStreamCreateWithFlags(&g->side, cudaStreamNonBlocking));
    PB_GCUDA(cudaHostAlloc(reinterpret_cast<void**>(&g->h_debug), 2u * PARIS_B200_GROUP_MAX_MEMBERS * sizeof(uint32_t),
                           cudaHostAllocPortable));

    // region, slabs, rounds, bands: pure host arithmetic, shared with paris_b200_group_plan
    paris_b200_group_plan_t plan{};
    PB_GTRY(paris_b200_group_plan(cfg, &plan));
    g->region_x = plan.region_x;
    g->region_y = plan.region_y;
    g->region_z = plan.region_z;
    g->region_z0 = plan.region_z0;
    const uint32_t world = static_cast<uint32_t>(cfg->world), me = static_cast<uint32_t>(cfg->rank);
    const uint32_t spr = std::max(1u, cfg->slabs_per_member);
    const uint32_t total = plan.slabs_total;
    const uint32_t xp = plan.x_parts, mz = me / xp, mx = me % xp;
    g->x_first = mx * plan.x_dx;
    g->x_count = plan.x_dx + (mx == xp - 1u ? plan.x_remainder : 0u);
    g->host_row = cfg->host_row_floats ? cfg->host_row_floats : g->x_count;
    if(g->host_row < g->x_count)
    {
        set_error("host_row_floats = %u is less than the member's %u columns", cfg->host_row_floats, g->x_count);
        return fail(PARIS_B200_EINVAL);
    }
    for(uint32_t s = 0; s < spr; ++s)
    {
        const uint32_t id = mz * spr + s;
        // src/main.cpp:96, src/make_volume.cpp:32-34
        g->slabs.push_back(slab_t{id * plan.slab_dz, plan.slab_dz + (id == total - 1u ? plan.slab_remainder : 0u)});
    }

    // stack geometry, filter, angles
    g->layout = plan.layout;
    g->pitch = plan.pitch;
    g->slot_floats = static_cast<size_t>(g->pitch) * cfg->det.n_row;
    g->px = static_cast<size_t>(cfg->det.n_row) * cfg->det.n_col;
    g->sample_bytes = cfg->sample_type == PARIS_B200_SAMPLES_U16 ? 2u : 4u;
    PB_GTRY(paris_b200_filter_create(g->fctx, paris_b200_filter_size(cfg->det.n_row), cfg->det.l_px_row, &g->filter));
    g->sn.resize(cfg->n_proj);
    g->cs.resize(cfg->n_proj);
    for(uint32_t i = 0; i < cfg->n_proj; ++i)
        angle_sin_cos(cfg->angles_deg != nullptr ? cfg->angles_deg[i] : static_cast<float>(i) * cfg->det.delta_phi, &g->sn[i],
                      &g->cs[i]);

    for(uint32_t rd = 0; rd < plan.rounds; ++rd)
        g->rounds.push_back(round_t{plan.round_first[rd], plan.round_count[rd]});
    size_t max_share = 0;
    for(const auto& rd : g->rounds)
    {
        uint32_t f = 0, c = 0;
        share_of(rd, world, me, &f, &c);
        g->local_first.push_back(g->my_count);
        g->my_count += c;
        max_share = std::max<size_t>(max_share, c);
    }
    for(uint32_t k = 0; k < world; ++k)
        g->bands.push_back(band_t{plan.band_lo[k], plan.band_hi[k]});

    // device memory: the stack (every slot, same offsets on every member), flags, upload buffers, slab buffers
    PB_GCUDA(cudaMalloc(reinterpret_cast<void**>(&g->stack), g->slot_floats * cfg->n_proj * sizeof(float)));
    PB_GCUDA(cudaMemsetAsync(g->stack, 0, g->slot_floats * cfg->n_proj * sizeof(float), g->ctx->compute));
    PB_GCUDA(cudaMalloc(reinterpret_cast<void**>(&g->flags), 2u << 20));   // (an allocation of its own: exportable)
    PB_GCUDA(cudaMemsetAsync(g->flags, 0, 2u << 20, g->ctx->compute));
    g->raw_bytes = std::max<size_t>(max_share, 1u) * g->px * g->sample_bytes;
    for(int i = 0; i < 2; ++i)
    {
        PB_GCUDA(cudaMalloc(reinterpret_cast<void**>(&g->raw[i]), g->raw_bytes));
        PB_GCUDA(cudaEventCreateWithFlags(&g->raw_free[i], cudaEventDisableTiming));
    }
    PB_GCUDA(cudaEventCreateWithFlags(&g->pushed, cudaEventDisableTiming));
    PB_GCUDA(cudaEventCreateWithFlags(&g->step_done, cudaEventDisableTiming));
    g->filtered.resize(g->rounds.size());
    for(auto& e : g->filtered)
        PB_GCUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    g->uploaded.resize(g->rounds.size());
    for(auto& e : g->uploaded)
        PB_GCUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    if(g->trace)
    {
        PB_GCUDA(cudaEventCreate(&g->trace_t0));
        for(auto* v : {&g->trace_filtered, &g->trace_pushed, &g->trace_bp})
        {
            v->resize(g->rounds.size());
            for(auto& e : *v)
                PB_GCUDA(cudaEventCreate(&e));
        }
    }
    const uint32_t n_buf = cfg->stream_slabs ? std::min(spr, 2u) : spr;
    for(uint32_t b = 0; b < n_buf; ++b)
    {
        // (a streamed buffer must hold the largest of this member's slabs: the last one carries the remainder)
        uint32_t need = 0;
        for(uint32_t s = b; s < spr; s += n_buf)
            need = std::max(need, g->slabs[s].dz);
        float* v = nullptr;
        PB_GCUDA(cudaMalloc(reinterpret_cast<void**>(&v), static_cast<size_t>(g->x_count) * g->region_y * need * sizeof(float)));
        g->vol.push_back(v);
        cudaEvent_t e = nullptr;
        PB_GCUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        g->slab_down.push_back(e);
        g->slab_down_valid.push_back(false);
    }
    PB_GCUDA(cudaStreamSynchronize(g->ctx->compute));

    // every kernel a step can launch is loaded NOW (see common.cuh: a lazy load inside a step deadlocks the group)
    preload_backprojection_kernels();
    preload_filter_kernels(g->filter->size);
    {
        cudaFuncAttributes a{};
        (void)cudaFuncGetAttributes(&a, flag_store_kernel);
        (void)cudaFuncGetAttributes(&a, wait_flag_kernel);
        (void)cudaFuncGetAttributes(&a, push_band_kernel);
        (void)cudaGetLastError();
    }

    g->peer_stack.assign(world, nullptr);
    g->peer_flags.assign(world, nullptr);
    g->peer_ipc.assign(world, false);
    g->peer_stack[me] = g->stack;
    g->peer_flags[me] = g->flags;
    g->connected = world == 1u;
#undef PB_GTRY
#undef PB_GCUDA
    *out = g;
    return PARIS_B200_OK;
}

extern "C" int paris_b200_group_destroy(paris_b200_group* g)
{
    if(g == nullptr)
        return PARIS_B200_OK;
    if(g->ctx != nullptr)
    {
        paris_b200_ctx_bind(g->ctx);
        cudaDeviceSynchronize();
    }
    for(size_t k = 0; k < g->peer_stack.size(); ++k)
        if(g->peer_ipc[k])
        {
            // (the mapped base pointers were stored unshifted in the ipc slots below)
            cudaIpcCloseMemHandle(g->peer_stack[k]);
            cudaIpcCloseMemHandle(g->peer_flags[k]);
        }
    for(float* v : g->vol)
        cudaFree(v);
    for(int i = 0; i < 2; ++i)
    {
        if(g->raw[i]) cudaFree(g->raw[i]);
        if(g->raw_free[i]) cudaEventDestroy(g->raw_free[i]);
    }
    for(auto e : g->slab_down)
        if(e) cudaEventDestroy(e);
    if(g->pushed) cudaEventDestroy(g->pushed);
    for(auto e : g->filtered)
        if(e) cudaEventDestroy(e);
    for(auto e : g->uploaded)
        if(e) cudaEventDestroy(e);
    if(g->trace_t0) cudaEventDestroy(g->trace_t0);
    for(auto* v : {&g->trace_filtered, &g->trace_pushed, &g->trace_bp})
        for(auto e : *v)
            if(e) cudaEventDestroy(e);
    if(g->step_done) cudaEventDestroy(g->step_done);
    if(g->stack) cudaFree(g->stack);
    if(g->flags) cudaFree(g->flags);
    if(g->filter) paris_b200_filter_destroy(g->filter);
    if(g->push) cudaStreamDestroy(g->push);
    if(g->side) cudaStreamDestroy(g->side);
    if(g->h_debug) cudaFreeHost(g->h_debug);
    if(g->fctx) paris_b200_ctx_destroy(g->fctx);
    if(g->ctx) paris_b200_ctx_destroy(g->ctx);
    delete g;
    return PARIS_B200_OK;
}

extern "C" int paris_b200_group_export(paris_b200_group* g, unsigned char* handle, size_t handle_bytes)
{
    PB_CHECK_ARG(g != nullptr && handle != nullptr && handle_bytes >= PARIS_B200_GROUP_HANDLE_BYTES);
    PB_TRY(bind(g));
    group_handle h{};
    h.magic = kHandleMagic;
    h.rank = static_cast<uint32_t>(g->cfg.rank);
    h.pid = static_cast<int32_t>(getpid());
    h.device = g->ctx->device;
    h.stack_ptr = reinterpret_cast<uint64_t>(g->stack);
    h.flags_ptr = reinterpret_cast<uint64_t>(g->flags);
    // an IPC handle names the ALLOCATION a pointer lies in: keep the pointer's offset inside it
    CUdeviceptr base = 0;
    size_t size = 0;
    using range_fn = CUresult (*)(CUdeviceptr*, size_t*, CUdeviceptr);
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q{};
    range_fn get_range = nullptr;
    if(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fp, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
        get_range = reinterpret_cast<range_fn>(fp);
    if(get_range != nullptr && get_range(&base, &size, reinterpret_cast<CUdeviceptr>(g->stack)) == CUDA_SUCCESS)
        h.stack_off = h.stack_ptr - static_cast<uint64_t>(base);
    if(get_range != nullptr && get_range(&base, &size, reinterpret_cast<CUdeviceptr>(g->flags)) == CUDA_SUCCESS)
        h.flags_off = h.flags_ptr - static_cast<uint64_t>(base);
    PB_CUDA(cudaIpcGetMemHandle(&h.stack_ipc, g->stack));
    PB_CUDA(cudaIpcGetMemHandle(&h.flags_ipc, g->flags));
    std::memset(handle, 0, PARIS_B200_GROUP_HANDLE_BYTES);
    std::memcpy(handle, &h, sizeof(h));
    return PARIS_B200_OK;
}

extern "C" int paris_b200_group_connect(paris_b200_group* g, const unsigned char* handles, size_t handle_bytes)
{
    PB_CHECK_ARG(g != nullptr && handles != nullptr && handle_bytes >= PARIS_B200_GROUP_HANDLE_BYTES);
    PB_TRY(bind(g));
    const uint32_t world = static_cast<uint32_t>(g->cfg.world), me = static_cast<uint32_t>(g->cfg.rank);
    for(uint32_t k = 0; k < world; ++k)
    {
        if(k == me)
            continue;
        group_handle h{};
        std::memcpy(&h, handles + static_cast<size_t>(k) * handle_bytes, sizeof(h));
        if(h.magic != kHandleMagic || h.rank != k)
        {
            set_error("handle %u is not the export of member %u", k, k);
            return PARIS_B200_EINVAL;
        }
        if(h.pid == static_cast<int32_t>(getpid()))
        {
            // same process (one host thread per device, src/main.cpp:157-169): the pointers are valid as they are
            if(h.device != g->ctx->device)
            {
                int can = 0;
                PB_CUDA(cudaDeviceCanAccessPeer(&can, g->ctx->device, h.device));
                if(!can)
                {
                    set_error("device %d cannot access the memory of device %d", g->ctx->device, h.device);
                    return PARIS_B200_ECUDA;
                }
                const cudaError_t e = cudaDeviceEnablePeerAccess(h.device, 0);
                if(e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                    PB_CUDA(e);
                (void)cudaGetLastError();
            }
            g->peer_stack[k] = reinterpret_cast<float*>(h.stack_ptr);
            g->peer_flags[k] = reinterpret_cast<uint32_t*>(h.flags_ptr);
        }
        else
        {
            void* ps = nullptr;
            void* pf = nullptr;
            PB_CUDA(cudaIpcOpenMemHandle(&ps, h.stack_ipc, cudaIpcMemLazyEnablePeerAccess));
            PB_CUDA(cudaIpcOpenMemHandle(&pf, h.flags_ipc, cudaIpcMemLazyEnablePeerAccess));
            g->peer_ipc[k] = true;
            g->peer_stack[k] = static_cast<float*>(ps);   // (recorded first: group_destroy unmaps whatever was mapped)
            g->peer_flags[k] = static_cast<uint32_t*>(pf);
            if(h.stack_off != 0 || h.flags_off != 0)
            {
                set_error("peer %u exported pointers inside larger allocations (offsets %llu, %llu): unsupported", k,
                          static_cast<unsigned long long>(h.stack_off), static_cast<unsigned long long>(h.flags_off));
                return PARIS_B200_ESTATE;
            }
        }
    }
    g->connected = true;
    return PARIS_B200_OK;
}

extern "C" int paris_b200_group_info(const paris_b200_group* g, paris_b200_group_info_t* info)
{
    PB_CHECK_ARG(g != nullptr && info != nullptr);
    std::memset(info, 0, sizeof(*info));
    const uint32_t me = static_cast<uint32_t>(g->cfg.rank);
    info->my_projections = g->my_count;
    info->rounds = static_cast<uint32_t>(g->rounds.size());
    info->slabs = static_cast<uint32_t>(g->slabs.size());
    info->z_first = g->slabs.front().z_first;
    info->z_count = g->slabs.back().z_first + g->slabs.back().dz - g->slabs.front().z_first;
    info->x_first = g->x_first;
    info->x_count = g->x_count;
    info->region_x = g->region_x;
    info->region_y = g->region_y;
    info->region_z = g->region_z;
    info->band_lo = g->bands[me].lo;
    info->band_hi = g->bands[me].hi;
    info->layout = g->layout;
    info->pitch = g->pitch;
    info->d_stack = g->stack;
    info->slab_buffers = static_cast<uint32_t>(g->vol.size());
    info->d_first_slab = g->vol.empty() ? nullptr : g->vol[0];
    info->bytes_pushed = g->bytes_pushed;
    info->ctx = g->ctx;
    info->filter_ctx = g->fctx;
    info->memops = g->memops_ok ? 1u : 0u;
    return PARIS_B200_OK;
}

// floats from one row to the next in the host memory handed to the following steps (0: the member's own box)
extern "C" int paris_b200_group_set_host_row(paris_b200_group* g, uint32_t host_row_floats)
{
    PB_CHECK_ARG(g != nullptr);
    if(g->in_step || g->step_open)
    {
        set_error("group_set_host_row while a step is in flight");
        return PARIS_B200_ESTATE;
    }
    PB_CHECK_ARG(host_row_floats == 0u || host_row_floats >= g->x_count);
    g->host_row = host_row_floats ? host_row_floats : g->x_count;
    return PARIS_B200_OK;
}

// scan index of this member's local projection `local` (the order its raw projections are handed over in)
extern "C" int paris_b200_group_projection_index(const paris_b200_group* g, uint32_t local, uint32_t* index)
{
    PB_CHECK_ARG(g != nullptr && index != nullptr && local < g->my_count);
    const uint32_t world = static_cast<uint32_t>(g->cfg.world), me = static_cast<uint32_t>(g->cfg.rank);
    for(size_t rd = g->rounds.size(); rd-- > 0;)
        if(g->local_first[rd] <= local)
        {
            uint32_t f = 0, c = 0;
            share_of(g->rounds[rd], world, me, &f, &c);
            *index = f + (local - g->local_first[rd]);
            return PARIS_B200_OK;
        }
    return PARIS_B200_EINVAL;
}

// ---- one step, piece by piece: open -> round 0 .. rounds-1 -> finish -> end -----------------------------------
// (callers that have all their projections at hand use group_begin; the command-line driver reads the next round's
// frames from disk while the device works on the previous one)

extern "C" int paris_b200_group_step_open(paris_b200_group* g, float* h_slabs)
{
    PB_CHECK_ARG(g != nullptr);
    if(!g->connected)
    {
        set_error("group step before group_connect");
        return PARIS_B200_ESTATE;
    }
    if(g->in_step || g->step_open)
    {
        set_error("group step opened while another one is in flight (call group_end first)");
        return PARIS_B200_ESTATE;
    }
    if(g->slabs.size() > g->vol.size() && h_slabs == nullptr)
    {
        set_error("slabs are streamed through %zu buffers: a host destination is required", g->vol.size());
        return PARIS_B200_EINVAL;
    }
    PB_TRY(bind(g));
    paris_b200_ctx* ctx = g->ctx;
    paris_b200_ctx* fctx = g->fctx;
    const uint32_t world = static_cast<uint32_t>(g->cfg.world), me = static_cast<uint32_t>(g->cfg.rank);
    const uint32_t step = g->steps;            // steps completed before this one
    const size_t slice = static_cast<size_t>(g->x_count) * g->region_y;

    // ---- write-after-read guards: the stack slots are rewritten ------------------------------------------------
    // my own previous backprojections have read my slots; every peer's must have read what I am about to push
    if(step > 0)
    {
        PB_CUDA(cudaStreamWaitEvent(fctx->compute, g->step_done, 0));
        PB_CUDA(cudaStreamWaitEvent(fctx->compute, g->pushed, 0));   // (the exchange reads my slots as well)
        for(uint32_t k = 0; k < world; ++k)
            if(k != me)
                PB_TRY(wait32(g, g->push, g->flags + world + k, step, k));
    }
    // (first use of buffer 0 in this step: the previous step's download of it finished in group_end)
    if(g->trace)
        PB_CUDA(cudaEventRecord(g->trace_t0, ctx->compute));
    PB_CUDA(cudaMemsetAsync(g->vol[0], 0, slice * g->slabs[0].dz * sizeof(float), ctx->compute));
    g->step_h_slabs = h_slabs;
    g->next_round = 0;
    g->step_open = true;
    return PARIS_B200_OK;
}

// Round `rd` of the open step (rounds go in order).  Exactly one of h_raw / d_raw unless this member has no share of
// the round: h_raw[j] = pinned host address of the j-th projection of this member's share, d_raw = the share on the
// device, contiguous.  Everything is only enqueued: the host buffers must stay untouched until group_end (or until
// paris_b200_group_uploaded reports the round).
extern "C" int paris_b200_group_step_round(paris_b200_group* g, uint32_t rd, const float* const* h_raw, const float* d_raw)
{
    PB_CHECK_ARG(g != nullptr);
    if(!g->step_open || rd != g->next_round || rd >= g->rounds.size())
    {
        set_error("group_step_round(%u): rounds of an open step go in order (next: %u of %zu)", rd, g->next_round, g->rounds.size());
        return PARIS_B200_ESTATE;
    }
    PB_TRY(bind(g));
    paris_b200_ctx* ctx = g->ctx;
    paris_b200_ctx* fctx = g->fctx;
    const uint32_t world = static_cast<uint32_t>(g->cfg.world), me = static_cast<uint32_t>(g->cfg.rank);
    const uint32_t n_rounds = static_cast<uint32_t>(g->rounds.size());
    const uint32_t step = g->steps;
    uint32_t first = 0, count = 0;
    share_of(g->rounds[rd], world, me, &first, &count);
    PB_CHECK_ARG(count == 0 || (h_raw != nullptr) != (d_raw != nullptr));
    const uint32_t seq = step * n_rounds + rd + 1u;
    const weight_params w = [&] {
        // src/weighting.cpp:37-42
        const auto& det = g->cfg.det;
        const float n_row_f = static_cast<float>(det.n_row), n_col_f = static_cast<float>(det.n_col);
        weight_params p{};
        p.enable = 1;
        p.h_min = (det.delta_s * det.l_px_row) - ((n_row_f * det.l_px_row) / 2);
        p.v_min = (det.delta_t * det.l_px_col) - ((n_col_f * det.l_px_col) / 2);
        p.d_sd = std::fabs(det.d_so) + std::fabs(det.d_od);
        p.l_px_row = det.l_px_row;
        p.l_px_col = det.l_px_col;
        return p;
    }();

    if(count > 0)
    {
        const float* src = d_raw;
        if(h_raw != nullptr)
        {
            // upload this round's share into one of the two upload buffers (copy stream of the filter context)
            const int b = static_cast<int>(rd & 1u);
            if(g->raw_free_valid[b])
                PB_CUDA(cudaStreamWaitEvent(fctx->copy, g->raw_free[b], 0));
            for(uint32_t j = 0; j < count; ++j)
                PB_CUDA(cudaMemcpyAsync(reinterpret_cast<unsigned char*>(g->raw[b]) + g->px * g->sample_bytes * j, h_raw[j],
                                        g->px * g->sample_bytes, cudaMemcpyHostToDevice, fctx->copy));
            PB_CUDA(cudaEventRecord(g->uploaded[rd], fctx->copy));
            PB_CUDA(cudaStreamWaitEvent(fctx->compute, g->uploaded[rd], 0));
            src = g->raw[b];
        }
        const float* ptrs[kMaxBatch];
        for(uint32_t done = 0; done < count;)
        {
            const uint32_t n = std::min<uint32_t>(count - done, 64u);   // (persistent CTAs: 64 projections fill the GPU)
            for(uint32_t i = 0; i < n; ++i)
                ptrs[i] = reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(src)
                                                         + g->px * g->sample_bytes * (done + i));
            PB_TRY(launch_filter_batch(fctx, ptrs, nullptr, n, g->stack, first + done, g->slot_floats, g->cfg.det.n_row,
                                       g->cfg.det.n_col, g->filter, w, true, g->pitch, g->layout, g->sample_bytes == 2u));
            done += n;
        }
        if(h_raw != nullptr)
        {
            PB_CUDA(cudaEventRecord(g->raw_free[rd & 1u], fctx->compute));
            g->raw_free_valid[rd & 1u] = true;
        }
    }
    if(count == 0 || h_raw == nullptr)
        PB_CUDA(cudaEventRecord(g->uploaded[rd], fctx->copy));   // (nothing to wait for)
    PB_CUDA(cudaEventRecord(g->filtered[rd], fctx->compute));
    if(g->trace)
        PB_CUDA(cudaEventRecord(g->trace_filtered[rd], fctx->compute));
    if(world > 1u)
        PB_TRY(push_round(g, rd, seq));
    if(g->trace)
        PB_CUDA(cudaEventRecord(g->trace_pushed[rd], g->push));
    // the round is complete here once my own share is filtered and every peer's has arrived
    PB_CUDA(cudaStreamWaitEvent(ctx->compute, g->filtered[rd], 0));
    for(uint32_t k = 0; k < world; ++k)
        if(k != me)
            PB_TRY(wait32(g, ctx->compute, g->flags + k, seq, k));
    const slab_t s0 = g->slabs[0];
    const bool last_round = rd + 1u == n_rounds;
    if(last_round && g->step_h_slabs != nullptr)
    {
        // the last launch into the slab is cut into z-chunks whose download runs behind the next chunk's kernel
        const bp_target t = target_of(g, s0, g->vol[0]);
        PB_TRY(backproject_and_download(ctx, g->stack, g->slot_floats, g->pitch, g->rounds[rd].first, g->rounds[rd].count,
                                        g->sn.data() + g->rounds[rd].first, g->cs.data() + g->rounds[rd].first, t,
                                        g->layout, g->step_h_slabs, false, g->host_row));
        PB_CUDA(cudaEventRecord(g->slab_down[0], ctx->copy));
        g->slab_down_valid[0] = true;
    }
    else
        PB_TRY(backproject_range(g, g->rounds[rd].first, g->rounds[rd].count, s0, g->vol[0]));
    if(g->trace)
        PB_CUDA(cudaEventRecord(g->trace_bp[rd], ctx->compute));
    g->next_round = rd + 1u;
    return PARIS_B200_OK;
}

// 1 once the uploads of round `rd` have left their host buffers (those may then be reused)
extern "C" int paris_b200_group_uploaded(paris_b200_group* g, uint32_t rd, int* done)
{
    PB_CHECK_ARG(g != nullptr && done != nullptr && rd < g->rounds.size());
    PB_TRY(bind(g));
    const cudaError_t e = cudaEventQuery(g->uploaded[rd]);
    (void)cudaGetLastError();
    *done = e == cudaSuccess ? 1 : 0;
    return PARIS_B200_OK;
}

// After the last round: the member's further slabs loop over the ONE gathered stack, the download of slab k behind
// the backprojection of slab k + 1; then every peer learns that its pushes into this member's stack have been consumed.
extern "C" int paris_b200_group_step_finish(paris_b200_group* g)
{
    PB_CHECK_ARG(g != nullptr);
    if(!g->step_open || g->next_round != g->rounds.size())
    {
        set_error("group_step_finish before every round of the step was given");
        return PARIS_B200_ESTATE;
    }
    PB_TRY(bind(g));
    paris_b200_ctx* ctx = g->ctx;
    const uint32_t world = static_cast<uint32_t>(g->cfg.world), me = static_cast<uint32_t>(g->cfg.rank);
    const uint32_t step = g->steps;
    const size_t slice = static_cast<size_t>(g->x_count) * g->region_y;
    const size_t h_slice = static_cast<size_t>(g->host_row) * g->region_y;
    float* h_slabs = g->step_h_slabs;
    for(uint32_t s = 1; s < g->slabs.size(); ++s)
    {
        const slab_t sl = g->slabs[s];
        const uint32_t b = s % static_cast<uint32_t>(g->vol.size());
        if(g->slab_down_valid[b])
            PB_CUDA(cudaStreamWaitEvent(ctx->compute, g->slab_down[b], 0));   // the buffer's previous slab is on the host
        PB_CUDA(cudaMemsetAsync(g->vol[b], 0, slice * sl.dz * sizeof(float), ctx->compute));
        const uint32_t n_proj = g->cfg.n_proj;
        if(h_slabs != nullptr)
        {
            const uint32_t batch = static_cast<uint32_t>(ctx->bp_batch);
            const uint32_t head = n_proj > batch ? n_proj - batch : 0u;
            if(head > 0u)
                PB_TRY(backproject_range(g, 0u, head, sl, g->vol[b]));
            const bp_target t = target_of(g, sl, g->vol[b]);
            PB_TRY(backproject_and_download(ctx, g->stack, g->slot_floats, g->pitch, head, n_proj - head, g->sn.data() + head,
                                            g->cs.data() + head, t, g->layout,
                                            h_slabs + h_slice * (sl.z_first - g->slabs[0].z_first), false, g->host_row));
            PB_CUDA(cudaEventRecord(g->slab_down[b], ctx->copy));
            g->slab_down_valid[b] = true;
        }
        else
            PB_TRY(backproject_range(g, 0u, n_proj, sl, g->vol[b]));
    }
    PB_CUDA(cudaEventRecord(g->step_done, ctx->compute));
    PB_CUDA(cudaEventRecord(g->pushed, g->push));
    for(uint32_t k = 0; k < world; ++k)
        if(k != me)
            PB_TRY(signal32(g, ctx->compute, g->peer_flags[k] + world + me, step + 1u));
    g->steps = step + 1u;
    g->step_open = false;
    g->in_step = true;
    return PARIS_B200_OK;
}

extern "C" int paris_b200_group_begin(paris_b200_group* g, const float* const* h_raw, const float* d_raw, float* h_slabs)
{
    PB_CHECK_ARG(g != nullptr);
    PB_CHECK_ARG((h_raw != nullptr) != (d_raw != nullptr) || g->my_count == 0);
    PB_TRY(paris_b200_group_step_open(g, h_slabs));
    for(uint32_t rd = 0; rd < g->rounds.size(); ++rd)
    {
        const uint32_t local = g->local_first[rd];
        PB_TRY(paris_b200_group_step_round(g, rd, h_raw != nullptr ? h_raw + local : nullptr,
                                           d_raw != nullptr ? reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(d_raw)
                                                                                              + g->px * g->sample_bytes * local)
                                                            : nullptr));
    }
    return paris_b200_group_step_finish(g);
}

extern "C" int paris_b200_group_end(paris_b200_group* g)
{
    PB_CHECK_ARG(g != nullptr);
    if(!g->in_step)
        return PARIS_B200_OK;
    PB_TRY(bind(g));
    PB_CUDA(cudaStreamSynchronize(g->ctx->compute));
    PB_CUDA(cudaStreamSynchronize(g->ctx->copy));
    PB_CUDA(cudaStreamSynchronize(g->push));
    std::fill(g->slab_down_valid.begin(), g->slab_down_valid.end(), false);
    g->in_step = false;
    if(g->trace)
    {
        std::string line = "[group trace] member " + std::to_string(g->cfg.rank) + " step " + std::to_string(g->steps)
                         + ": round(projections) filtered / pushed / backprojected at ms:";
        for(size_t rd = 0; rd < g->rounds.size(); ++rd)
        {
            float f = 0.f, p = 0.f, b = 0.f;
            cudaEventElapsedTime(&f, g->trace_t0, g->trace_filtered[rd]);
            cudaEventElapsedTime(&p, g->trace_t0, g->trace_pushed[rd]);
            cudaEventElapsedTime(&b, g->trace_t0, g->trace_bp[rd]);
            char buf[96];
            std::snprintf(buf, sizeof(buf), " %zu(%u) %.1f/%.1f/%.1f", rd, g->rounds[rd].count, f, p, b);
            line += buf;
        }
        (void)cudaGetLastError();
        std::fprintf(stderr, "%s\n", line.c_str());
    }
    uint32_t gave_up = 0;
    PB_CUDA(cudaMemcpy(&gave_up, g->flags + kErrorWord, sizeof(gave_up), cudaMemcpyDeviceToHost));
    if(gave_up != 0u)
    {
        set_error("member %d gave up waiting for member %u after %.0f s: the step's result is incomplete", g->cfg.rank,
                  gave_up - 1u, static_cast<double>(g->wait_timeout_ns) * 1e-9);
        PB_CUDA(cudaMemset(g->flags + kErrorWord, 0, sizeof(uint32_t)));
        return PARIS_B200_ESTATE;
    }
    return PARIS_B200_OK;
}

// diagnostics: which of the member's streams still have work (1) and the current flag words, without touching any
// of those streams.  out: [0..4] compute, download, filter, upload, exchange; [5 .. 5 + 2 world) arrived[], consumed[]
extern "C" int paris_b200_group_debug_state(paris_b200_group* g, uint32_t* out, uint32_t n)
{
    PB_CHECK_ARG(g != nullptr && out != nullptr);
    const uint32_t world = static_cast<uint32_t>(g->cfg.world);
    PB_CHECK_ARG(n >= 5u + 2u * world);
    PB_TRY(bind(g));
    cudaStream_t streams[5] = {g->ctx->compute, g->ctx->copy, g->fctx->compute, g->fctx->copy, g->push};
    for(int i = 0; i < 5; ++i)
    {
        out[i] = cudaStreamQuery(streams[i]) == cudaSuccess ? 0u : 1u;
        (void)cudaGetLastError();
    }
    // (a stream and a pinned landing zone made at create time: nothing here can queue behind the member's own work)
    PB_CUDA(cudaMemcpyAsync(g->h_debug, g->flags, 2u * world * sizeof(uint32_t), cudaMemcpyDeviceToHost, g->side));
    for(int spin = 0; spin < 2000 && cudaStreamQuery(g->side) != cudaSuccess; ++spin)
        usleep(1000);
    (void)cudaGetLastError();
    const bool landed = cudaStreamQuery(g->side) == cudaSuccess;
    (void)cudaGetLastError();
    for(uint32_t i = 0; i < 2u * world; ++i)
        out[5 + i] = landed ? g->h_debug[i] : 0xffffffffu;
    return PARIS_B200_OK;
}

extern "C" int paris_b200_group_reconstruct(paris_b200_group* g, const float* const* h_raw, const float* d_raw, float* h_slabs)
{
    PB_TRY(paris_b200_group_begin(g, h_raw, d_raw, h_slabs));
    return paris_b200_group_end(g);
}
