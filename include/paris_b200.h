/*
 * paris_b200.h -- C ABI of the B200-native FDK backend (libparis_b200.so).
 *
 * This is the drop-in boundary for the hot path of hzdr/PARIS: everything the
 * reference's compile-time backend contract (src/backend.h:26-46, canonical text
 * src/generic/backend.h:42-88) asks of a backend, as plain C entry points -- no C++
 * types, no torch types, plain pointers and sizes.  The C++ namespace paris::b200
 * (paris_b200/cpp/b200/backend.h) is a set of <=10-line forwarders onto these, and
 * INTEGRATION.md shows the three-line change that selects it inside the reference.
 *
 * Conventions
 *  - every call returns 0 on success, else a PARIS_B200_E* code; paris_b200_last_error()
 *    gives the text for the calling thread.  Nothing throws, nothing calls exit().
 *  - a context (paris_b200_ctx) belongs to one device and one host thread at a time,
 *    like the reference's per-device thread (src/main.cpp:157-169).  All work of a
 *    context is ordered on its compute stream; H2D/D2H copies use a second stream and
 *    are ordered against compute with events.
 *  - "d_" pointers are device memory of the context's device, "h_" pointers are host.
 *  - projections are dim_x = n_row samples per detector row (fastest) times dim_y = n_col
 *    rows, contiguous (src/projection.h:31-46); volumes are x fastest, z slowest,
 *    contiguous (src/volume.h:31-45).  All samples float32.
 *  - there is NO CPU fallback: every entry point that computes launches sm_100a kernels
 *    and fails with PARIS_B200_ECUDA when no such device is present.
 */
#ifndef PARIS_B200_H_
#define PARIS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PARIS_B200_OK 0
#define PARIS_B200_EINVAL 1   /* bad argument */
#define PARIS_B200_ECUDA 2    /* CUDA runtime / driver error (text in last_error) */
#define PARIS_B200_ENOMEM 3   /* allocation failed */
#define PARIS_B200_ESTATE 4   /* call not valid in the context's current state */

/* src/geometry.h:30-47 */
typedef struct paris_b200_detector_geometry
{
    uint32_t n_row;   /* pixels per detector row */
    uint32_t n_col;   /* number of detector rows */
    float l_px_row;   /* pixel size along a row [mm] */
    float l_px_col;   /* pixel size across rows [mm] */
    float delta_s;    /* horizontal detector offset [px] */
    float delta_t;    /* vertical detector offset [px] */
    float d_so;       /* source -> object [mm] */
    float d_od;       /* object -> detector [mm] */
    float delta_phi;  /* angle step [deg] */
} paris_b200_detector_geometry;

/* src/geometry.h:49-58 */
typedef struct paris_b200_volume_geometry
{
    uint32_t dim_x, dim_y, dim_z;
    float l_vx_x, l_vx_y, l_vx_z;
} paris_b200_volume_geometry;

/* src/region_of_interest.h:30-38 */
typedef struct paris_b200_roi
{
    uint32_t x1, x2, y1, y2, z1, z2;
} paris_b200_roi;

/* src/subvolume_information.h:30-34 with src/geometry.h:60-69 flattened */
typedef struct paris_b200_subvolume_info
{
    uint32_t dim_x, dim_y, dim_z;
    uint32_t remainder; /* extra slices carried by the LAST slab (src/make_volume.cpp:32-34) */
    int32_t num;        /* number of slabs */
} paris_b200_subvolume_info;

typedef struct paris_b200_ctx paris_b200_ctx;       /* per-device state */
typedef struct paris_b200_filter paris_b200_filter; /* device-resident ramp-filter table */

/* ---- library / devices ------------------------------------------------------------------ */

const char* paris_b200_last_error(void);
const char* paris_b200_version(void);

/* get_devices / set_device: src/generic/backend.h:84-86, src/cuda/device.cpp:38-47 */
int paris_b200_device_count(int* count);
int paris_b200_ctx_create(int device, paris_b200_ctx** ctx);
int paris_b200_ctx_destroy(paris_b200_ctx* ctx);
int paris_b200_ctx_device(const paris_b200_ctx* ctx, int* device);
/* make the context's device current for the calling thread (set_device) */
int paris_b200_ctx_bind(paris_b200_ctx* ctx);
/* block until all work of the context (both streams) is complete */
int paris_b200_ctx_sync(paris_b200_ctx* ctx);
/* the context's compute stream as a cudaStream_t, for callers that time with CUDA events */
int paris_b200_ctx_stream(paris_b200_ctx* ctx, void** stream);
/* counters since ctx_create: kernels launched by this library on this context */
int paris_b200_ctx_launch_count(const paris_b200_ctx* ctx, uint64_t* launches);

/* Which backprojection kernel ran: the instantiation of the most recent backprojection launch as text (tile, box,
 * stages, stack layout; "bp_exact_kernel" for the reference-order fallback a geometry takes when its footprint does
 * not fit the TMA kernel's tiles), and how many launches went to either since ctx_create.  Any pointer may be NULL. */
int paris_b200_ctx_bp_kernel_info(const paris_b200_ctx* ctx, char* name, size_t name_len, uint64_t* tma_launches,
                                  uint64_t* exact_launches);

/* diagnostics: stats[0..5] = kernel launches, pool cudaMallocs, pool reuses of an idle buffer, pool reuses of a
 * buffer still being read (the upload waits), backprojection flushes, pool size; n >= 6 */
int paris_b200_ctx_stats(const paris_b200_ctx* ctx, uint64_t* stats, int n);

/* CUDA-event timing on the context's compute stream (bench / profiling plumbing) */
typedef struct paris_b200_event paris_b200_event;
int paris_b200_event_create(paris_b200_ctx* ctx, paris_b200_event** ev);
int paris_b200_event_record(paris_b200_ctx* ctx, paris_b200_event* ev);
/* waits for `stop`, then returns the milliseconds between the two records */
int paris_b200_event_elapsed_ms(paris_b200_event* start, paris_b200_event* stop, float* ms);
int paris_b200_event_destroy(paris_b200_event* ev);

/* Tunables.  "bp_batch": projections accumulated per backprojection launch (default 256, max 256);
 * "bp_kernel": 0 = auto, 1 = generic L1-gather kernel, 2 = TMA-staged kernel.
 * "bp_tile": 0 = auto (16x8x64 voxel tiles, two CTAs per SM, when the footprint fits), 1 = 16x16x64 tiles only.
 * "bp_swizzle": backprojection CTAs are numbered in n x n super-blocks of (x, y) tiles (default 16; 0 = row-major).
 * "filter_wide": 4096-point transforms with two row pairs per 512-thread CTA (1, default) or one per 256 (0). */
int paris_b200_ctx_set_option(paris_b200_ctx* ctx, const char* name, int64_t value);

/* ---- geometry (host-side, pure arithmetic) ------------------------------------------------ */

/* calculate_volume_geometry: src/geometry.cpp:36-84 */
int paris_b200_calculate_volume_geometry(const paris_b200_detector_geometry* det, paris_b200_volume_geometry* vol);
/* apply_roi: src/geometry.cpp:86-130 (an invalid ROI leaves the geometry unchanged, as the reference does) */
int paris_b200_apply_roi(const paris_b200_volume_geometry* vol, const paris_b200_roi* roi,
                         paris_b200_volume_geometry* out);
/* filter_size = 2 * 2^ceil(log2(n_row)): src/filtering.cpp:38 */
uint32_t paris_b200_filter_size(uint32_t n_row);
/* make_subvolume_information: src/cuda/subvolume_information.cpp:63-118.  num_slabs > 0 forces that many
 * equal z-slabs (remainder on the last); num_slabs == 0 sizes slabs by the free memory of the
 * context's device like the reference (volume + 10 projections must fit, halve until it does). */
int paris_b200_make_subvolume_information(paris_b200_ctx* ctx, const paris_b200_volume_geometry* vol,
                                          const paris_b200_detector_geometry* det, int num_slabs,
                                          paris_b200_subvolume_info* out);

/* ---- memory: make_* / copy_* of the backend contract (src/openmp/memory.cpp:33-79,
 *      src/cuda/memory.cpp:34-102) ----------------------------------------------------------- */

/* pinned host memory (make_projection_host / make_volume_host); zero == nonzero clears it */
int paris_b200_host_alloc(size_t bytes, int zero, void** h_ptr);
int paris_b200_host_free(void* h_ptr);
/* page-lock / release host memory the caller owns (e.g. a shared-memory mapping that several member processes of a
 * group download their slabs into, so that the host volume is assembled by the downloads themselves) */
int paris_b200_host_register(void* h_ptr, size_t bytes);
int paris_b200_host_unregister(void* h_ptr);
/* stream-ordered device memory (make_projection_device); freeing is ordered after queued work */
int paris_b200_dev_alloc(paris_b200_ctx* ctx, size_t bytes, void** d_ptr);
int paris_b200_dev_free(paris_b200_ctx* ctx, void* d_ptr);
/* make_volume_device: zero-initialised dim_x*dim_y*dim_z floats */
int paris_b200_volume_alloc(paris_b200_ctx* ctx, uint32_t dim_x, uint32_t dim_y, uint32_t dim_z, float** d_vol);
int paris_b200_volume_free(paris_b200_ctx* ctx, float* d_vol);
/* zero a device volume again (asynchronous on the compute stream); pending batches targeting it are dropped */
int paris_b200_volume_clear(paris_b200_ctx* ctx, float* d_vol, uint32_t dim_x, uint32_t dim_y, uint32_t dim_z);

/* copy_h2d(projection): asynchronous when h_src is pinned; later work of the context waits for it.
 * The host buffer must stay valid until paris_b200_h2d_done() reports completion (or ctx_sync). */
int paris_b200_proj_h2d(paris_b200_ctx* ctx, const float* h_src, float* d_dst, uint32_t dim_x, uint32_t dim_y);
/* nonzero *done when every H2D copy issued so far has finished */
int paris_b200_h2d_done(paris_b200_ctx* ctx, int* done);
/* copy_d2h(projection): returns when the data is in h_dst */
int paris_b200_proj_d2h(paris_b200_ctx* ctx, const float* d_src, float* h_dst, uint32_t dim_x, uint32_t dim_y);
/* copy_h2d / copy_d2h(volume).  vol_d2h first flushes pending backprojections into d_src
 * (it is the only observer of the volume, sink.cpp:77) and returns when the data is in h_dst. */
int paris_b200_vol_h2d(paris_b200_ctx* ctx, const float* h_src, float* d_dst, size_t n_voxels);
int paris_b200_vol_d2h(paris_b200_ctx* ctx, const float* d_src, float* h_dst, size_t n_voxels);

/* ---- pipeline stages --------------------------------------------------------------------- */

/* backend::weight (src/openmp/weighting.cpp:32-57): in place,
 * p[s + t*dim_x] *= d_sd / sqrt(d_sd^2 + h_s^2 + v_t^2) */
int paris_b200_weight(paris_b200_ctx* ctx, float* d_proj, uint32_t dim_x, uint32_t dim_y,
                      float h_min, float v_min, float d_sd, float l_px_row, float l_px_col);

/* backend::make_filter (src/openmp/filtering.cpp:139-165): K[x] = tau*|DFT(r)[x]|, x = 0..size/2,
 * r the spatial Ram-Lak taps of :52-73.  size must be a power of two, 32..8192. */
int paris_b200_filter_create(paris_b200_ctx* ctx, uint32_t size, float tau, paris_b200_filter** filter);
int paris_b200_filter_destroy(paris_b200_filter* filter);
/* copy the size/2+1 table entries to the host (tests) */
int paris_b200_filter_read(paris_b200_ctx* ctx, const paris_b200_filter* filter, float* h_k);

/* backend::apply_filter (src/openmp/filtering.cpp:167-219): per detector row zero-pad to
 * filter_size, DFT, scale by K, inverse DFT, keep the first dim_x samples, divide by filter_size.
 * In place.  n_col must equal dim_y. */
int paris_b200_apply_filter(paris_b200_ctx* ctx, float* d_proj, uint32_t dim_x, uint32_t dim_y,
                            const paris_b200_filter* filter, uint32_t filter_size, uint32_t n_col);

/* weight + apply_filter fused in one kernel (the weighted row never reaches HBM).  In place. */
int paris_b200_weight_filter(paris_b200_ctx* ctx, float* d_proj, uint32_t dim_x, uint32_t dim_y,
                             float h_min, float v_min, float d_sd, float l_px_row, float l_px_col,
                             const paris_b200_filter* filter, uint32_t filter_size);

/* backend::backproject (src/openmp/backprojection.cpp:156-199): accumulate one filtered projection
 * into the slab d_vol (v_dim_x * v_dim_y * v_dim_z, x fastest).  vol_full is the FULL volume geometry;
 * with enable_roi the voxel indices are shifted by (roi.x1, roi.y1, roi.z1); v_offset is the slab's z
 * offset; sin/cos of the projection angle; delta_s/delta_t in millimetres (src/backprojection.cpp:49-50).
 *
 * DEFERRED: the projection is copied (transposed) into the context's filtered stack and the call
 * returns; the volume is updated when the batch fills, on paris_b200_flush(), or when the volume is
 * observed by paris_b200_vol_d2h().  d_proj may be freed (paris_b200_dev_free) right after the call.
 * If flags has PARIS_B200_BP_FUSE_WEIGHT_FILTER, d_proj holds the RAW projection and weight + filter
 * (with `filter`; the weighting scalars from `weighting`, or derived from det as src/weighting.cpp:37-42
 * does when that is NULL) are applied on the way into the stack by the fused kernel; d_proj itself is
 * left untouched. */
#define PARIS_B200_BP_FUSE_WEIGHT_FILTER 1u
/* the five scalars backend::weight receives (src/weighting.cpp:37-44) */
typedef struct paris_b200_weighting
{
    float h_min, v_min, d_sd, l_px_row, l_px_col;
} paris_b200_weighting;
int paris_b200_backproject(paris_b200_ctx* ctx, const float* d_proj, uint32_t dim_x, uint32_t dim_y,
                           float* d_vol, uint32_t v_dim_x, uint32_t v_dim_y, uint32_t v_dim_z, uint32_t v_offset,
                           const paris_b200_detector_geometry* det, const paris_b200_volume_geometry* vol_full,
                           int enable_roi, const paris_b200_roi* roi,
                           float sin_phi, float cos_phi, float delta_s_mm, float delta_t_mm,
                           uint32_t flags, const paris_b200_filter* filter, const paris_b200_weighting* weighting);
/* block the calling thread until every H2D copy issued so far has left its host buffer */
int paris_b200_h2d_wait(paris_b200_ctx* ctx);
/* run every pending backprojection batch of the context (asynchronous on the compute stream) */
int paris_b200_flush(paris_b200_ctx* ctx);

/* ---- stack-level entry points (multi-GPU path: filter 1/N of the projections, all-gather the
 *      stack, backproject all of them into the local slab) ---------------------------------- */

/* A stack slot holds one filtered projection TRANSPOSED: n_row lines (one per detector column) of `pitch`
 * floats, detector-row index v fastest.  Two line layouts exist:
 *   PARIS_B200_LAYOUT_PLAIN   line[v]
 *   PARIS_B200_LAYOUT_SPLIT2  even rows, then odd rows: line[(v & 1) * pitch/2 + (v >> 1)]
 * The backprojection kernel prefers SPLIT2 when one voxel step in z spans about two detector rows (volumes of
 * K^3 voxels from a (2K)^2 detector); paris_b200_choose_stack_layout picks it from the geometry.  Filter and
 * backprojection calls on the same stack must be given the same layout. */
#define PARIS_B200_LAYOUT_PLAIN 0u
#define PARIS_B200_LAYOUT_SPLIT2 1u
int paris_b200_choose_stack_layout(const paris_b200_detector_geometry* det, const paris_b200_volume_geometry* vol_full,
                                   uint32_t* layout);
/* Size in bytes of one stack slot and its pitch in floats. */
int paris_b200_stack_slot_bytes(uint32_t n_row, uint32_t n_col, size_t* bytes, uint32_t* pitch);
/* a zero-initialised stack of `slots` slots (the padding columns pitch > n_col must be zero; a caller that brings
 * its own buffer, e.g. one NCCL can see, zeroes it itself) */
int paris_b200_stack_alloc(paris_b200_ctx* ctx, uint32_t n_row, uint32_t n_col, uint32_t slots, float** d_stack);
int paris_b200_stack_free(paris_b200_ctx* ctx, float* d_stack);
/* weight + filter a RAW device projection into slot `slot` of an external stack buffer */
int paris_b200_filter_to_stack(paris_b200_ctx* ctx, const float* d_raw, const paris_b200_detector_geometry* det,
                               const paris_b200_filter* filter, float* d_stack, uint32_t slot, uint32_t layout);
/* the same for `count` raw projections laid out raw_stride floats apart, into slots first_slot.. ; one
 * launch per 64 projections */
int paris_b200_filter_to_stack_batch(paris_b200_ctx* ctx, const float* d_raw, size_t raw_stride, uint32_t count,
                                     const paris_b200_detector_geometry* det, const paris_b200_filter* filter,
                                     float* d_stack, uint32_t first_slot, uint32_t layout);
/* the same for detector-native 16-bit samples (raw_stride in samples): the widening the HIS reader does on the host
 * (src/his.cpp:169-185) becomes the kernel's first load, so that half the bytes cross PCIe and HBM */
int paris_b200_filter_to_stack_batch_u16(paris_b200_ctx* ctx, const uint16_t* d_raw, size_t raw_stride, uint32_t count,
                                         const paris_b200_detector_geometry* det, const paris_b200_filter* filter,
                                         float* d_stack, uint32_t first_slot, uint32_t layout);
/* backproject slots [first, first+count) of an external stack into d_vol; sin_phi/cos_phi are host
 * arrays of `count` entries (slot first+i uses entry i). */
int paris_b200_backproject_stack(paris_b200_ctx* ctx, const float* d_stack, uint32_t first, uint32_t count,
                                 const float* sin_phi, const float* cos_phi,
                                 float* d_vol, uint32_t v_dim_x, uint32_t v_dim_y, uint32_t v_dim_z,
                                 uint32_t v_offset, const paris_b200_detector_geometry* det,
                                 const paris_b200_volume_geometry* vol_full, int enable_roi,
                                 const paris_b200_roi* roi, uint32_t layout);
/* the same, followed by the download of the volume into h_dst (pinned host memory, v_dim_x*v_dim_y*v_dim_z
 * floats): the work is cut into z-chunks at the kernel's tile anchors and the copy of each finished chunk runs
 * behind the backprojection of the next one (what sink::save's copy_d2h does after the last projection,
 * /root/reference/src/sink.cpp:76-77, without serialising it behind the whole backprojection).  Returns with the
 * host copy complete. */
int paris_b200_backproject_stack_d2h(paris_b200_ctx* ctx, const float* d_stack, uint32_t first, uint32_t count,
                                     const float* sin_phi, const float* cos_phi,
                                     float* d_vol, uint32_t v_dim_x, uint32_t v_dim_y, uint32_t v_dim_z,
                                     uint32_t v_offset, const paris_b200_detector_geometry* det,
                                     const paris_b200_volume_geometry* vol_full, int enable_roi,
                                     const paris_b200_roi* roi, uint32_t layout, float* h_dst);

/* ---- one scan across the GPUs of a box (csrc/group.cu) ------------------------------------------
 *
 * The reference's multi-device scheme is task parallelism: every device re-reads and re-filters every projection for
 * each slab it owns and nothing is exchanged (src/main.cpp:93-105; slabs src/cuda/subvolume_information.cpp:112-116,
 * src/make_volume.cpp:32-34, offset src/main.cpp:96).  A GROUP replaces that loop: `world` members, one per GPU --
 * one process each, or several host threads of one process like the reference's std::async per device
 * (src/main.cpp:157-169).  Member r uploads and filters 1/world of the projections; after every round of projections
 * each member copies its filtered share straight into its peers' stacks over NVLink (peer memory, copy engines, only
 * the band of detector rows a peer's slabs can read) and announces it with a flag word the peer's stream waits on;
 * every member backprojects all projections into its own z-slabs -- slabs_per_member of them, looping over the one
 * gathered stack, each downloaded behind the next one's backprojection.  No reduction, no host synchronisation inside
 * a step; the host volume is assembled by writing every slab at its z offset.
 *
 *   create (every member)  ->  export (64..128-byte handle)  ->  [the host side hands every member all handles]  ->
 *   connect  ->  reconstruct / begin + end, any number of times  ->  destroy
 */
#define PARIS_B200_SAMPLES_F32 0u
#define PARIS_B200_SAMPLES_U16 1u           /* detector-native counts: h_raw / d_raw address uint16_t samples */
#define PARIS_B200_EXCHANGE_COPY_ENGINE 0u   /* cudaMemcpy2DAsync into peer memory on a stream of its own (default) */
#define PARIS_B200_EXCHANGE_KERNEL 1u        /* a small copy kernel on a high-priority stream */
#define PARIS_B200_GROUP_HANDLE_BYTES 256

typedef struct paris_b200_group paris_b200_group;

typedef struct paris_b200_group_config
{
    int32_t rank, world;
    paris_b200_detector_geometry det;
    paris_b200_volume_geometry vol_full;  /* the FULL volume geometry */
    int32_t enable_roi;                   /* the reconstructed region: the ROI box if set, else the full volume */
    paris_b200_roi roi;
    uint32_t n_proj;                      /* projections of the scan; projection i is slot i of every member's stack */
    const float* angles_deg;              /* n_proj angles, or NULL: angle i = float(i) * det.delta_phi (src/backprojection.cpp:53-57) */
    uint32_t slabs_per_member;            /* >= 1: the region is cut into world * slabs_per_member equal z-slabs (remainder on
                                             the last), member r owns slabs [r * spm, (r + 1) * spm) */
    uint32_t stream_slabs;                /* 1: at most two slab buffers on the device, slabs go to the host one by one (a host
                                             destination is then required); 0: every slab stays resident */
    uint32_t sample_type;                 /* PARIS_B200_SAMPLES_F32 or _U16 (raw projections as 16-bit counts: half the upload) */
    uint32_t first_round, max_round;      /* projections per exchanged round: first one, upper bound (0 = 64 / 256) */
    uint32_t whole_projections;           /* 1: exchange every detector row (an all-gather); 0: only the band of rows the
                                             receiving member's slabs can read */
    uint32_t exchange;                    /* PARIS_B200_EXCHANGE_* */
    uint32_t x_parts;                     /* 0 or 1: members own whole slices (z-slabs).  n > 1 (n divides world): the region is
                                             also cut into n equal parts along x (remainder on the last); member r owns x-part
                                             r % n of z-run r / n.  Two uses: world / n = 2 mirror z-runs give every member
                                             the same mix of cheap inner and costly outermost slices (bench.py from 4 GPUs
                                             on), and volumes with few slices per GPU keep z-runs of >= 128 slices */
    uint32_t host_row_floats;             /* floats from one row to the next in the host memory handed to group_begin: 0 = the
                                             member's own box, contiguous; region_x = a box inside one region-wide volume */
} paris_b200_group_config;

typedef struct paris_b200_group_info_t
{
    uint32_t my_projections;              /* how many projections this member uploads and filters */
    uint32_t rounds, slabs;
    uint32_t z_first, z_count;            /* this member's slices of the region */
    uint32_t x_first, x_count;            /* this member's columns of the region (everything unless x_parts > 1) */
    uint32_t region_x, region_y, region_z;
    uint32_t band_lo, band_hi;            /* detector rows [lo, hi) this member receives from its peers */
    uint32_t layout, pitch;               /* stack layout (PARIS_B200_LAYOUT_*) and line pitch in floats */
    float* d_stack;                       /* the member's filtered stack: n_proj slots (rows outside the band are only
                                             valid for the member's own projections) */
    uint32_t slab_buffers;                /* device slab buffers (== slabs unless stream_slabs) */
    float* d_first_slab;                  /* device buffer of the member's first slab */
    uint64_t bytes_pushed;                /* bytes this member has copied into peer memory since create */
    paris_b200_ctx* ctx;                  /* the context whose compute stream runs the backprojection (events, options) */
    paris_b200_ctx* filter_ctx;           /* the context that uploads and filters */
    uint32_t memops;                      /* 1: arrival flags are written by stream memory operations, 0: by one-word kernels */
} paris_b200_group_info_t;

/* The decomposition a configuration leads to -- who filters what, who owns which slices, which detector rows travel
 * to whom.  Pure host arithmetic: no device is touched (checked on CPU-only machines, tests/test_multi_cpu.py). */
#define PARIS_B200_GROUP_MAX_MEMBERS 64
#define PARIS_B200_GROUP_MAX_ROUNDS 256
typedef struct paris_b200_group_plan_t
{
    uint32_t region_x, region_y, region_z;   /* the reconstructed region */
    uint32_t region_z0;                       /* its first slice in the full volume */
    uint32_t layout, pitch;                   /* stack layout and line pitch (floats) */
    uint32_t slabs_total, slab_dz, slab_remainder;   /* (world / x_parts) * slabs_per_member slabs of slab_dz slices, remainder on the last */
    uint32_t x_parts, x_dx, x_remainder;              /* x_parts parts of x_dx columns, remainder on the last */
    uint32_t rounds;
    uint32_t round_first[PARIS_B200_GROUP_MAX_ROUNDS], round_count[PARIS_B200_GROUP_MAX_ROUNDS];
    uint32_t band_lo[PARIS_B200_GROUP_MAX_MEMBERS], band_hi[PARIS_B200_GROUP_MAX_MEMBERS];   /* rows [lo, hi) member k receives */
} paris_b200_group_plan_t;
int paris_b200_group_plan(const paris_b200_group_config* cfg, paris_b200_group_plan_t* plan);
/* member `member`'s share of round `round`: projections [first, first + count) of the scan */
int paris_b200_group_share(const paris_b200_group_plan_t* plan, uint32_t world, uint32_t round, uint32_t member,
                           uint32_t* first, uint32_t* count);

size_t paris_b200_group_handle_bytes(void);
int paris_b200_group_create(int device, const paris_b200_group_config* cfg, paris_b200_group** group);
int paris_b200_group_destroy(paris_b200_group* group);
/* what a peer needs to reach this member's stack and flags (handle_bytes >= PARIS_B200_GROUP_HANDLE_BYTES) */
int paris_b200_group_export(paris_b200_group* group, unsigned char* handle, size_t handle_bytes);
/* handles: world blobs of handle_bytes each, blob k = member k's export.  Members of the same process are reached through
 * their pointers (peer access is enabled), members of other processes through CUDA IPC. */
int paris_b200_group_connect(paris_b200_group* group, const unsigned char* handles, size_t handle_bytes);
int paris_b200_group_info(const paris_b200_group* group, paris_b200_group_info_t* info);
/* change host_row_floats for the following steps (0: the member's own box, contiguous) */
int paris_b200_group_set_host_row(paris_b200_group* group, uint32_t host_row_floats);
/* scan index of this member's local projection `local` (0 <= local < my_projections): the order in which its raw
 * projections are handed to group_begin */
int paris_b200_group_projection_index(const paris_b200_group* group, uint32_t local, uint32_t* index);
/* One reconstruction.  Exactly one of h_raw / d_raw: h_raw[j] = (pinned) host address of this member's local raw
 * projection j (n_col x n_row floats), uploaded inside the step; d_raw = the same projections already on the device,
 * contiguous in local order.  h_slabs: where this member's first slab goes on the host (z_count slices of region_x x
 * region_y floats, pinned), or NULL to leave the slabs on the device.  begin() only enqueues -- members that share a
 * host thread begin one after the other and then end; end() returns when this member's slabs are complete. */
int paris_b200_group_begin(paris_b200_group* group, const float* const* h_raw, const float* d_raw, float* h_slabs);
/* (with PARIS_B200_SAMPLES_U16 the h_raw[j] / d_raw pointers address uint16_t samples, n_col x n_row per projection) */
/* The same step piece by piece, for callers that produce their projections while the device works (the command-line
 * driver reads the next round's frames from disk meanwhile): open, then every round in order -- h_raw[j] / d_raw now
 * address only the member's share of THAT round (paris_b200_group_share tells which projections those are) -- then
 * finish, then group_end.  group_uploaded reports when a round's host buffers may be reused. */
int paris_b200_group_step_open(paris_b200_group* group, float* h_slabs);
int paris_b200_group_step_round(paris_b200_group* group, uint32_t round, const float* const* h_raw, const float* d_raw);
int paris_b200_group_uploaded(paris_b200_group* group, uint32_t round, int* done);
int paris_b200_group_step_finish(paris_b200_group* group);
int paris_b200_group_end(paris_b200_group* group);
/* diagnostics: out[0..4] = 1 while the member's backprojection / download / filter / upload / exchange stream still
 * has work; out[5 .. 5 + world) = arrival flags (last round each member delivered), the next `world` words = how
 * many steps each member has consumed.  n >= 5 + 2 * world.  Touches none of those streams. */
int paris_b200_group_debug_state(paris_b200_group* group, uint32_t* out, uint32_t n);
int paris_b200_group_reconstruct(paris_b200_group* group, const float* const* h_raw, const float* d_raw, float* h_slabs);

/* ---- synthetic input (bench / tests): analytic cone-beam line integrals of ellipsoids ------ */

/* ellipsoids: n x 8 doubles {density, a, b, c, x0, y0, z0, theta_deg} in millimetres (host memory).
 * Writes projections first_idx .. first_idx+n_proj-1 (angle = float(idx)*delta_phi) to d_stack_raw,
 * n_proj x n_col x n_row floats, row-major. */
int paris_b200_phantom_project(paris_b200_ctx* ctx, const double* ellipsoids, uint32_t n_ellipsoids,
                               const paris_b200_detector_geometry* det, uint32_t first_idx, uint32_t n_proj,
                               float* d_stack_raw);

#ifdef __cplusplus
}
#endif

#endif /* PARIS_B200_H_ */
