// backproject_tma.cu -- K2, production kernel: TMA-staged, register-accumulating backprojection.
//
// Work decomposition (one CTA = one voxel tile, all projections of the batch):
//   * tile = TX x TY voxel columns x TZ = 32*NZ slices.  A warp's 32 LANES RUN ALONG z for one column
//     group (lane l owns slices l, l+32, ...): for a fixed column (x, y) and projection, the detector
//     column position h, the bilinear x-weights and the magnification are the same for every z, and the
//     detector row position is affine in z:  v(z) = v_base + z * dv.  All per-(column, projection)
//     terms -- the reference's rotate / perspective-divide arithmetic of
//     /root/reference/src/openmp/backprojection.cpp:120-133, including the one IEEE division -- are
//     therefore computed ONCE per column and projection by one thread, put in a shared-memory table and
//     read back as a warp-wide broadcast, once for the NZ slices of a lane.
//   * v(z) is evaluated with the REFERENCE'S OWN FLOAT OPERATIONS (:130-133 -> :45-50: multiply by the
//     magnification, subtract the detector's lower edge, divide by the pixel size, subtract 1/2), two slices
//     at a time in packed f32x2 arithmetic, the IEEE division as a reciprocal multiply plus one exact-residual
//     correction (bit-identical to the division on every value checked, scripts/check_division.py).  floor and
//     fraction come from one round-down addition of 1.5 * 2^23: the sum's low mantissa bits ARE the row index
//     (no conversion instruction), the y-weight is v minus the rounded value.  The kernel therefore interpolates at
//     exactly the reference's coordinates -- round 1 carried rows in fixed point, more accurately than the
//     reference, and was up to 1.05e-4 of the phantom contrast away from it on single voxels for that very reason
//     (the reference rounds v to float32 after every operation: ulp 1.2e-4 rows at |v| ~ 2000).
//   * the filtered stack is stored transposed (detector-row index v fastest), so the 32 lanes of an
//     update read 32 nearly consecutive floats of one stack line; the tile's footprint on projection p,
//     a BH x BV box, is fetched by ONE 3-D TMA load (cp.async.bulk.tensor) into a ring of shared-memory
//     stages, out-of-detector parts zero-filled by the TMA unit.  mbarrier transaction counts signal
//     arrival; loads run S-1 projections ahead of the math.
//   * each thread keeps CPW accumulators (its columns at its z) in registers for the whole batch, seeded
//     from the volume and stored once: the volume is read and written once per batch, and the per-voxel
//     summation order is the reference's (projection order), so no re-association error is introduced.
//   * exact FP32 bilinear interpolation on the four point-fetched samples with the reference's
//     "all four neighbours inside, else 0" rule (:65-71) -- no hardware texture filtering (8-bit weights
//     break the tolerance).
//
// Tensor cores are not used: this is a gather plus interpolation, not a contraction.
#include "common.cuh"
#include "backproject.cuh"

#include <cuda.h>
#include <cudaTypedefs.h>

namespace pb
{
    // ---- tile configuration ---------------------------------------------------------------------------
    template <int TX_, int TY_, int NZ_, int CPW_, int BH_, int BV_, int STAGES_, bool SPLIT_ = false>
    struct tile_cfg
    {
        // SPLIT: the stack (and the staged box) keeps even detector rows and odd detector rows in two planes
        static constexpr bool SPLIT = SPLIT_;
        static constexpr int BVH = BV_ / 2;                 // rows per parity plane of a staged column
        static constexpr int TX = TX_, TY = TY_, NZ = NZ_, CPW = CPW_, BH = BH_, BV = BV_, STAGES = STAGES_;
        static constexpr int TZ = 32 * NZ;                // slices per tile; lane l owns l, l+32, ...
        static constexpr int COLS = TX * TY;
        static constexpr int WARPS = COLS / CPW;          // one warp per group of CPW columns
        static constexpr int THREADS = 32 * WARPS;
        static constexpr int STAGE_BYTES = BH * BV * 4;
        static constexpr int TAB_BYTES = COLS * (16 + 4);      // float4 + the valid-slice interval per column
        // table sets: two (the table of projection p + 1 is built while projection p is consumed)
        static constexpr int TABLE_SLOTS = 2;
        // shared memory of an instantiation: stages, two table sets, box origins, barriers
        static constexpr size_t smem(bool straddle)
        {
            (void)straddle;   // (tiles anchored at the slab's first slice need nothing extra)
            return size_t(STAGES) * STAGE_BYTES + TABLE_SLOTS * TAB_BYTES + kMaxBatch * 8 + STAGES * 8 + 128;
        }
        static constexpr size_t SMEM = smem(true);   // (scratch-size checks use the larger one)
        // 256-thread tiles rely on TWO resident CTAs per SM (228 KB of shared memory, 1 KB reserved per CTA)
        static_assert(THREADS > 256 || 2 * (smem(true) + 1024) <= 233472, "two CTAs of this tile must fit one SM");
        static_assert(COLS % CPW == 0, "columns per warp must divide the tile");
        static_assert(CPW % TX == 0 || TX % CPW == 0, "a warp's columns must be whole or partial x-runs");
        static_assert(BV % 4 == 0, "TMA inner box extent must be a multiple of 16 bytes");
        static_assert(BH <= 256 && (SPLIT ? BV / 2 : BV) <= 256, "TMA box extents are limited to 256");
        static_assert(STAGE_BYTES % 128 == 0, "stages must keep the 128-byte TMA destination alignment");
        static_assert(!SPLIT || BV % 8 == 0, "parity planes: 16-byte plane rows");
    };

    struct box_origin   // 8 bytes per projection of the batch
    {
        short h0, v0;             // detector coordinates of the box's first element
        unsigned char all_valid;  // 1: every bilinear cell the tile touches lies inside the detector and the box;
                                  // 2: the tile's shadow misses the detector altogether (nothing to add); 0: mixed
        unsigned char fits;       // the box covers the tile's footprint (guaranteed by the host-side check; else rows are clamped)
        unsigned short pad;
    };
    static_assert(sizeof(box_origin) == 8, "box origins are budgeted at 8 bytes per projection");

    // ---- small PTX wrappers -----------------------------------------------------------------------------
    __device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

    __device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
    {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
    }

    __device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
    {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    }

    __device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
    {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "WAIT_LOOP:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
            "@p bra WAIT_DONE;\n"
            "bra WAIT_LOOP;\n"
            "WAIT_DONE:\n"
            "}\n" ::"r"(bar), "r"(parity) : "memory");
    }

    __device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2)
    {
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
            ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
    }

    // (a & b) | c in one LOP3
    __device__ __forceinline__ uint32_t and_or(uint32_t a, uint32_t b, uint32_t c)
    {
        uint32_t d;
        asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
        return d;
    }

    __device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                                int c3)
    {
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
            ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
    }

    // one staged box: rows v0.. of columns h0.. of slot `slot` (v0 a multiple of 4, of 8 for the split layout)
    template <class CFG>
    __device__ __forceinline__ void load_box(uint32_t dst, const CUtensorMap* map, uint32_t bar, int v0, int h0, int slot)
    {
        if(CFG::SPLIT)
            tma_load_4d(dst, map, bar, v0 >> 1, 0, h0, slot);   // {row pair, parity, column, slot}
        else
            tma_load_3d(dst, map, bar, v0, h0, slot);
    }

    // Blackwell packed FP32: one instruction works on two floats held in a 64-bit register pair
    __device__ __forceinline__ uint64_t pack2(float lo, float hi)
    {
        uint64_t r;
        asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
        return r;
    }

    __device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi)
    {
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
    }

    __device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c)
    {
        uint64_t d;
        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
        return d;
    }

    __device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b)
    {
        uint64_t d;
        asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
        return d;
    }

    __device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b)
    {
        uint64_t d;
        asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
        return d;
    }

    // round towards minus infinity (FADD2.RM)
    __device__ __forceinline__ uint64_t add2_rm(uint64_t a, uint64_t b)
    {
        uint64_t d;
        asm("add.rm.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
        return d;
    }

    __device__ __forceinline__ uint64_t fma2_rm(uint64_t a, uint64_t b, uint64_t c)
    {
        uint64_t d;
        asm("fma.rm.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
        return d;
    }

#ifdef PB_BOUNDS_CHECK
    // Checked build (make check-lib; compute-sanitizer is closed on the GPU pool): every sample load of the
    // backprojection must fall inside the CTA's ring of staged boxes.  [0] loads checked, [1] loads outside (skipped,
    // they return 0), [2] / [3] lowest / highest offending offset relative to the ring.
    __device__ unsigned long long g_bounds[4] = {0ull, 0ull, ~0ull, 0ull};
    __device__ uint32_t g_ring_lo_of_cta, g_ring_bytes;   // (written by every CTA with the same values)

    __device__ __forceinline__ float lds_f32(uint32_t addr)
    {
        float v = 0.f;
        const uint32_t off = addr - g_ring_lo_of_cta;
        atomicAdd(&g_bounds[0], 1ull);
        if(off + 4u > g_ring_bytes || (addr & 3u) != 0u)
        {
            atomicAdd(&g_bounds[1], 1ull);
            atomicMin(&g_bounds[2], static_cast<unsigned long long>(static_cast<long long>(static_cast<int>(off))));
            atomicMax(&g_bounds[3], static_cast<unsigned long long>(off));
            return v;
        }
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
        return v;
    }
#else
    __device__ __forceinline__ float lds_f32(uint32_t addr)
    {
        float v;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
        return v;
    }
#endif

    // ---- per-projection geometry helpers ------------------------------------------------------------------

    __device__ __forceinline__ float centered(uint32_t coord, uint32_t dim, float size)
    {
        // src/openmp/backprojection.cpp:39-43, same float operations
        const float size2 = size / 2.f;
        return __fadd_rn(__fadd_rn(-__fmul_rn(static_cast<float>(dim), size2), size2),
                         __fmul_rn(static_cast<float>(coord), size));
    }

    struct column_terms
    {
        float h;        // fractional detector column, exactly the reference's float arithmetic (:121-129)
        float factor;   // d_sd / (s + d_so), IEEE
        float u;        // d_so / (s + d_so)
    };

    __device__ __forceinline__ column_terms project_column(float x_k, float y_l, float sn, float cs, const bp_geometry& g)
    {
        column_terms c;
        const float s = __fadd_rn(__fmul_rn(x_k, cs), __fmul_rn(y_l, sn));
        const float t = __fadd_rn(__fmul_rn(-x_k, sn), __fmul_rn(y_l, cs));
        const float denom = __fadd_rn(s, g.d_so);
        c.factor = __fdiv_rn(g.d_sd, denom);
        const float size2 = g.l_px_x / 2.f;
        const float min_h = __fsub_rn(-__fmul_rn(static_cast<float>(g.p_dim_x), size2), g.delta_s);
        c.h = __fsub_rn(__fdiv_rn(__fsub_rn(__fmul_rn(t, c.factor), min_h), g.l_px_x), 0.5f);
        c.u = c.factor * g.so_over_sd;   // d_so / (s + d_so); scales the sample only, no decision hangs on it
        return c;
    }

    // fractional detector row of slice coordinate z_m (double): (z_m*factor - min_v)/l_px_y - 0.5 (:130-133)
    __device__ __forceinline__ double row_of(double z_m, double factor, const bp_geometry& g)
    {
        return (z_m * factor - g.min_v_d) * g.inv_l_px_y_d - 0.5;
    }

    __device__ __forceinline__ double centered_d(uint32_t coord, uint32_t dim, float size)
    {
        const double sz = static_cast<double>(size);
        return -(static_cast<double>(dim) * (sz / 2.0)) + sz / 2.0 + static_cast<double>(coord) * sz;
    }

    // ---- the kernel ------------------------------------------------------------------------------------------

    // Boundary handling.  The reference adds a sample only if all four bilinear neighbours lie on the detector
    // (src/openmp/backprojection.cpp:65-71); at rows 0 and dim_y - 1 that is a DISCONTINUITY, so a row that
    // rounds to the other side of the border would add or drop a whole sample.  The reference's row
    //     v(z) = (z_m * factor - min_v) / l_px_y - 0.5        (:130-133, :45-50, z_m from :39-43)
    // is a chain of monotone float operations, hence monotone in z: per (column, projection) the valid slices
    // of a tile form ONE interval.  For columns that come within a cell of the border the table builder finds
    // that interval exactly -- it locates the two crossings with the affine model and settles each by
    // evaluating v(z) with the reference's own float operations at the three candidate slices -- and the inner
    // loop only compares the slice index with it.  Interior tiles (no column near the border) skip even that.
    __device__ __forceinline__ float reference_row(uint32_t z_global, float factor, const bp_geometry& g)
    {
        const float size2 = g.l_vx_z / 2.f;
        const float z_m = __fadd_rn(__fadd_rn(-__fmul_rn(static_cast<float>(g.full_z), size2), size2),
                                    __fmul_rn(static_cast<float>(z_global), g.l_vx_z));
        const float min_v = __fsub_rn(-__fmul_rn(static_cast<float>(g.p_dim_y), g.l_px_y / 2.f), g.delta_t);
        return __fsub_rn(__fdiv_rn(__fsub_rn(__fmul_rn(z_m, factor), min_v), g.l_px_y), 0.5f);
    }

    // First and count of the tile-local slices (0 .. TZ-1) whose rows r = floor(v) satisfy r >= 0 and
    // r + 1 < dim_y, packed as first | count << 8.  `first_row`/`dv` are the affine model in detector rows.
    template <int TZ>
    __device__ __forceinline__ uint32_t valid_slices(uint32_t z0, float factor, double first_row, double dv,
                                                     const bp_geometry& g)
    {
        const float dim_y_f = static_cast<float>(g.p_dim_y);
        // lower crossing: first slice with floor(v) >= 0
        int lo = 0;
        {
            const double zc = ceil((0.0 - first_row) / dv);
            const int c = static_cast<int>(fmin(fmax(zc, -2.0), static_cast<double>(TZ + 2)));
            lo = c + 2;   // if none of the candidates passes (cannot happen while the model is within a slice)
            #pragma unroll
            for(int t = 1; t >= -1; --t)
            {
                const int z = c + t;
                if(z < 0 || (z < TZ && floorf(reference_row(z0 + static_cast<uint32_t>(z), factor, g)) >= 0.f))
                    lo = z;   // monotone: ends at the smallest passing candidate
            }
            lo = max(lo, 0);
        }
        // upper crossing: last slice with floor(v) + 1 < dim_y
        int hi = TZ - 1;
        {
            const double zc = ceil((static_cast<double>(g.p_dim_y) - 1.0 - first_row) / dv) - 1.0;
            const int c = static_cast<int>(fmin(fmax(zc, -3.0), static_cast<double>(TZ + 1)));
            hi = c - 2;
            #pragma unroll
            for(int t = -1; t <= 1; ++t)
            {
                const int z = c + t;
                if(z >= TZ
                   || (z >= 0 && __fadd_rn(floorf(reference_row(z0 + static_cast<uint32_t>(z), factor, g)), 1.f) < dim_y_f))
                    hi = z;   // monotone: ends at the largest passing candidate
            }
            hi = min(hi, TZ - 1);
        }
        const int count = max(hi - lo + 1, 0);
        return static_cast<uint32_t>(min(lo, 255)) | (static_cast<uint32_t>(count) << 8);
    }

    // 1.5 * 2^23: adding it (rounding down) to a row -2^22 < v < 2^22 leaves floor(v) in the low mantissa bits
    constexpr uint32_t kMagicBits = 0x4B400000u;

    // per-launch constants of the row arithmetic, packed for the two slices of a pair
    struct row_consts
    {
        uint64_t neg_min_v2;   // -(min_v): (coord - min) of src/openmp/backprojection.cpp:49 as an addition
        uint64_t inv_px2;      // RN(1 / l_px_y)
        uint64_t neg_px2;      // -l_px_y
        uint64_t neg_half2, half2, magic2, neg_magic2, neg_one2, neg_two2;
        uint64_t zero2;        // a zero the assembler cannot see: z_m * factor + zero stays a ROUNDED product (ptxas fuses
                               // mul.rn.f32x2 + add.rn.f32x2 into one FFMA2, which would skip the reference's rounding)
    };

    // The four samples of one voxel update.  `tb` = bits of magic + floor(v); `base` = the staged column's address
    // minus the box origin's share (see the table builder), so that the row index needs no subtraction here.
    struct update_samples
    {
        float q11, q12, q21, q22;   // q?1: first row of the pair the layout fetches first, q?2: the other one
    };

    // plain layout: rows r, r + 1 are neighbours in the staged column; q?1 = row r, q?2 = row r + 1
    template <class CFG>
    __device__ __forceinline__ update_samples fetch_plain(uint32_t base, uint32_t tb)
    {
        update_samples u;
        const uint32_t addr = base + (tb << 2);
        u.q11 = lds_f32(addr);
        u.q12 = lds_f32(addr + 4);
        u.q21 = lds_f32(addr + 4 * CFG::BV);
        u.q22 = lds_f32(addr + 4 * CFG::BV + 4);
        return u;
    }

    // split layout: rows r = 2k + p and r + 1 are one even and one odd row; the odd one sits in the odd plane at pair
    // index k = floor(v / 2), the even one in the even plane at k + p = round(v / 2).  q?1 = EVEN row, q?2 = ODD row
    // whatever p is: every load instruction then reads ONE plane for all 32 lanes (lanes whose rows differ in parity
    // would otherwise hop between planes inside one instruction and collide).  tb_odd / tb_even = magic bits + those
    // two pair indices, straight from two additions of the magic number (round down / round to nearest): no shift,
    // no mask, no parity test.
    template <class CFG>
    __device__ __forceinline__ update_samples fetch_split(uint32_t base_even, uint32_t tb_even, uint32_t tb_odd)
    {
        update_samples u;
        const uint32_t a_even = base_even + (tb_even << 2);
        const uint32_t a_odd = base_even + (tb_odd << 2);          // (+ the odd plane's offset, an immediate below)
        u.q11 = lds_f32(a_even);
        u.q21 = lds_f32(a_even + 4 * CFG::BV);
        u.q12 = lds_f32(a_odd + 4 * CFG::BVH);
        u.q22 = lds_f32(a_odd + 4 * CFG::BVH + 4 * CFG::BV);
        return u;
    }

    // MIXED: the tile has voxels on both sides of the detector border: per-slice validity from the table, and row
    // indices are kept inside the staged box (a slice whose row lies outside the detector may lie outside the box as
    // well; its value is discarded by the validity select).
    template <class CFG, bool MIXED>
    __device__ __forceinline__ void consume(uint64_t (&acc)[CFG::CPW][CFG::NZ / 2], const float4* __restrict__ tab_a,
                                            const uint32_t* __restrict__ tab_c, const uint64_t (&zm)[CFG::NZ / 2],
                                            const row_consts& rc, uint32_t ct, int col0, uint32_t lane)
    {
        static_assert(CFG::NZ % 2 == 0, "the slices of a lane are processed as packed f32x2 pairs (l + 64j, l + 64j + 32)");
        const uint64_t minus_one = pack2(-1.f, -1.f);
        #pragma unroll
        for(int i = 0; i < CFG::CPW; ++i)
        {
            // {address of the staged column x1, magnification d_sd / (s + d_so), w*(1-fx), w*fx}, valid slices (first |
            // count << 8; boundary tiles only): read ONCE for the NZ slices of the lane
            const float4 ea = tab_a[col0 + i];
            const uint32_t colbase = __float_as_uint(ea.x);
            // (an index arrives as magic bits + its value; the box origin and the magic are taken off the base once:
            // ct = magic bits + the box's first row, or + its first row PAIR for the split layout)
            const uint32_t base = colbase - 4u * ct;
            const uint64_t f2 = pack2(ea.y, ea.y), wa2 = pack2(ea.z, ea.z), wb2 = pack2(ea.w, ea.w);
            uint32_t rel = 0, count = 0;
            if(MIXED)
            {
                // the reference's "all four neighbours inside" as a slice interval (see valid_slices)
                const uint32_t vs = tab_c[col0 + i];
                rel = lane - (vs & 0xffu);
                count = vs >> 8;
            }
            #pragma unroll
            for(int j = 0; j < CFG::NZ / 2; ++j)
            {
                // the reference's float arithmetic of the detector row (src/openmp/backprojection.cpp:130-133, :45-50),
                // both slices of the pair at once:  v = (z_m * factor - min_v) / l_px - 1/2
                const uint64_t num = add2(fma2(zm[j], f2, rc.zero2), rc.neg_min_v2);
                const uint64_t q0 = mul2(num, rc.inv_px2);
                const uint64_t q = fma2(fma2(q0, rc.neg_px2, num), rc.inv_px2, q0);   // == num / l_px (IEEE), see header
                const uint64_t v = add2(q, rc.neg_half2);
                uint64_t wy;                 // weight of q?2 against q?1
                update_samples s0, s1;
                uint32_t b = base;
                if(CFG::SPLIT)
                {
                    // pair indices: k = floor(v/2) holds the odd row, round(v/2) = k + p the even one (a tie -- v an odd
                    // integer -- may round either way: the even row's weight is then exactly 0).  With u = v - 2k in
                    // [0, 2) the EVEN row's weight is |u - 1| whichever of the two rows comes first, the odd row's 1 - |u - 1|.
                    // (v/2 is exact, so the fused multiply-adds round exactly like an addition of v/2 would)
                    const uint64_t t_odd = fma2_rm(v, rc.half2, rc.magic2);            // magic + floor(v/2)
                    const uint64_t t_even = fma2(v, rc.half2, rc.magic2);              // magic + round(v/2)
                    const uint64_t um1 = add2(fma2(add2(t_odd, rc.neg_magic2), rc.neg_two2, v), rc.neg_one2);   // u - 1, exactly
                    float o0, o1, e0, e1, w0, w1;
                    unpack2(t_odd, o0, o1);
                    unpack2(t_even, e0, e1);
                    unpack2(um1, w0, w1);
                    uint32_t to0 = __float_as_uint(o0), to1 = __float_as_uint(o1);
                    uint32_t te0 = __float_as_uint(e0), te1 = __float_as_uint(e1);
                    if(MIXED)
                    {
                        // box-relative pair indices, kept inside the box
                        to0 = static_cast<uint32_t>(min(max(static_cast<int>(to0 - ct), 0), CFG::BVH - 1));
                        to1 = static_cast<uint32_t>(min(max(static_cast<int>(to1 - ct), 0), CFG::BVH - 1));
                        te0 = static_cast<uint32_t>(min(max(static_cast<int>(te0 - ct), 0), CFG::BVH - 1));
                        te1 = static_cast<uint32_t>(min(max(static_cast<int>(te1 - ct), 0), CFG::BVH - 1));
                        b = colbase;
                    }
                    s0 = fetch_split<CFG>(b, te0, to0);
                    s1 = fetch_split<CFG>(b, te1, to1);
                    // d = g_odd + |u - 1| * (g_even - g_odd): swap the roles so that the common lerp below applies
                    wy = pack2(fabsf(w0), fabsf(w1));
                    float tmp;
                    tmp = s0.q11; s0.q11 = s0.q12; s0.q12 = tmp;
                    tmp = s0.q21; s0.q21 = s0.q22; s0.q22 = tmp;
                    tmp = s1.q11; s1.q11 = s1.q12; s1.q12 = tmp;
                    tmp = s1.q21; s1.q21 = s1.q22; s1.q22 = tmp;
                }
                else
                {
                    const uint64_t t = add2_rm(v, rc.magic2);                          // magic + floor(v), exactly
                    wy = fma2(add2(t, rc.neg_magic2), minus_one, v);                    // v - floor(v), exactly
                    float t0f, t1f;
                    unpack2(t, t0f, t1f);
                    uint32_t tb0 = __float_as_uint(t0f), tb1 = __float_as_uint(t1f);
                    if(MIXED)
                    {
                        // box-relative row, kept inside the box
                        tb0 = static_cast<uint32_t>(min(max(static_cast<int>(tb0 - ct), 0), CFG::BV - 2));
                        tb1 = static_cast<uint32_t>(min(max(static_cast<int>(tb1 - ct), 0), CFG::BV - 2));
                        b = colbase;
                    }
                    s0 = fetch_plain<CFG>(b, tb0);
                    s1 = fetch_plain<CFG>(b, tb1);
                }
                // both slices at once: g1 = wa*q11 + wb*q21, g2 = wa*q12 + wb*q22, d = g1 + wy*(g2 - g1)
                // (src/openmp/backprojection.cpp:73-83 with the weight 0.5*u^2 of :147 folded into wa, wb)
                const uint64_t g1 = fma2(wb2, pack2(s0.q21, s1.q21), mul2(wa2, pack2(s0.q11, s1.q11)));
                const uint64_t g2 = fma2(wb2, pack2(s0.q22, s1.q22), mul2(wa2, pack2(s0.q12, s1.q12)));
                uint64_t d = fma2(wy, fma2(g1, minus_one, g2), g1);
                if(MIXED)
                {
                    float d0, d1;
                    unpack2(d, d0, d1);
                    d = pack2((rel + 64u * j) < count ? d0 : 0.f, (rel + 64u * j + 32u) < count ? d1 : 0.f);
                }
                acc[i][j] = add2(acc[i][j], d);
            }
        }
    }

    template <class CFG, bool STRADDLE = false>
    __global__ void __launch_bounds__(CFG::THREADS, CFG::THREADS <= 256 ? 2 : 1)
    bp_tma_kernel(const __grid_constant__ CUtensorMap tmap, float* __restrict__ vol, const bp_geometry g,
                  const bp_angles ang, const uint32_t first_slot, const uint32_t tiles_x, const uint32_t tiles_y,
                  const uint32_t super)
    {
        extern __shared__ __align__(128) unsigned char smem[];
        unsigned char* stage_mem = smem;                                                   // STAGES x BH x BV floats
        float4* tab_a = reinterpret_cast<float4*>(smem + size_t(CFG::STAGES) * CFG::STAGE_BYTES);
        constexpr int TS = CFG::TABLE_SLOTS;
        uint32_t* tab_c = reinterpret_cast<uint32_t*>(tab_a + TS * CFG::COLS);
        box_origin* origin = reinterpret_cast<box_origin*>(tab_c + TS * CFG::COLS);
        uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(origin) + kMaxBatch * 8);

        const int tid = threadIdx.x;
        const uint32_t lane = tid & 31;
        const int warp = tid >> 5;          // which group of CPW columns
        const int count = ang.count;

        // Tiles are anchored at multiples of the tile size in GLOBAL voxel indices (ROI and slab offsets
        // included), and every per-tile quantity below is derived from the full, unclipped tile.  A voxel
        // therefore sees the same arithmetic whichever ROI / z-slab decomposition it is reconstructed in:
        // slabs and ROI blocks are bit-identical to the corresponding crop of the one-piece reconstruction.
        // CTAs are numbered along grid.x in SUPER-BLOCKS of super x super tiles, so that the ~300 CTAs resident at any
        // time cover a compact square of the slab instead of a few full-width rows of tiles.  Their boxes on a
        // projection then cover a narrow detector strip and the boxes of a whole batch stay in L2 from one wave of
        // tiles to the next (with row-major numbering the TMA loads of a 256-projection batch re-read ~57 GB from
        // HBM per launch, profiles/r1_ncu_summary.md).  Which CTA computes a tile changes nothing about the tile.
        const uint32_t sb = blockIdx.x / (super * super), within = blockIdx.x % (super * super);
        const uint32_t sb_per_row = (tiles_x + super - 1u) / super;
        const uint32_t bx = (sb % sb_per_row) * super + within % super;
        const uint32_t by = (sb / sb_per_row) * super + within / super;
        if(bx >= tiles_x || by >= tiles_y)
            return;   // (padding of the last super-blocks; uniform for the CTA, before any barrier is initialised)
        const uint32_t x0 = (g.off_x / CFG::TX + bx) * CFG::TX;   // global index of the tile's first voxel
        const uint32_t y0 = (g.off_y / CFG::TY + by) * CFG::TY;
        // (z: the STRADDLE instantiation anchors its tiles at the slab's first slice instead -- no partly empty tile
        // layers for regions whose z offset is not a multiple of the tile height; the ROW anchors below stay global,
        // so the voxels are bit-identical either way)
        const uint32_t z0 = STRADDLE ? g.off_z + blockIdx.z * CFG::TZ : (g.off_z / CFG::TZ + blockIdx.z) * CFG::TZ;
        // ---- prologue 1: barriers, box origins for every projection of the batch ---------------------------
        if(tid == 0)
        {
            #pragma unroll
            for(int s = 0; s < CFG::STAGES; ++s)
            {
                mbar_init(smem_u32(&bars[s]), 1);                                  // the stage's box has landed
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        if(tid < count)
        {
            const float sn = ang.sn[tid], cs = ang.cs[tid];
            // extremes over the tile's corners (h and factor are projective/monotone over the convex tile)
            const uint32_t xe = x0 + CFG::TX - 1u, ye = y0 + CFG::TY - 1u, ze = z0 + CFG::TZ - 1u;
            const double zlo = centered_d(z0, g.full_z, g.l_vx_z);
            const double zhi = centered_d(ze, g.full_z, g.l_vx_z);
            float hmin = 3.0e38f, hmax = -3.0e38f;
            double vmin = 1.0e300, vmax = -1.0e300;
            #pragma unroll
            for(int c = 0; c < 4; ++c)
            {
                const uint32_t k = (c & 1) ? xe : x0, l = (c & 2) ? ye : y0;
                const column_terms ct = project_column(centered(k, g.full_x, g.l_vx_x),
                                                       centered(l, g.full_y, g.l_vx_y), sn, cs, g);
                hmin = fminf(hmin, ct.h);
                hmax = fmaxf(hmax, ct.h);
                const double va = row_of(zlo, static_cast<double>(ct.factor), g);
                const double vb = row_of(zhi, static_cast<double>(ct.factor), g);
                vmin = fmin(vmin, fmin(va, vb));
                vmax = fmax(vmax, fmax(va, vb));
            }
            const bool finite = hmin == hmin && hmax == hmax && vmin == vmin && vmax == vmax;
            // clamp so the integer conversions are safe for columns far off the detector
            const float lim_h = static_cast<float>(g.p_dim_x) + 8.f, lim_v = static_cast<float>(g.p_dim_y) + 8.f;
            const float hlo_f = fminf(fmaxf(floorf(hmin), -8.f), lim_h), hhi_f = fminf(fmaxf(floorf(hmax), -8.f), lim_h);
            const float vlo_f = fminf(fmaxf(floorf(static_cast<float>(vmin)), -8.f), lim_v);
            const float vhi_f = fminf(fmaxf(floorf(static_cast<float>(vmax)), -8.f), lim_v);
            const int hlo = static_cast<int>(hlo_f), hhi = static_cast<int>(hhi_f);
            const int vlo = static_cast<int>(vlo_f), vhi = static_cast<int>(vhi_f);
            box_origin o;
            o.h0 = static_cast<short>(hlo - 1);
            // the TMA unit wants the innermost start coordinate on a 16-byte boundary (measured on B200:
            // anything else raises an illegal-instruction fault), so round down to a multiple of 4 floats
            // (a multiple of 8 rows = 4 row pairs for the split layout)
            o.v0 = static_cast<short>(CFG::SPLIT ? ((vlo - 1) >> 3) << 3 : ((vlo - 1) >> 2) << 2);
            // cells used: columns hlo-1 .. hhi+2, rows vlo-1 .. vhi+2 (one cell of slack for float rounding)
            const bool fits = (hhi + 2 - o.h0) < CFG::BH && (vhi + 2 - o.v0) < CFG::BV;
            const bool inside = hlo - 1 >= 0 && hhi + 2 <= static_cast<int>(g.p_dim_x) - 1
                             && vlo - 1 >= 0 && vhi + 2 <= static_cast<int>(g.p_dim_y) - 1;
            // no voxel of the tile has all four neighbours on the detector: x1 = floor(h) < 0 or x1+1 >= dim_x
            // for every column, or the same for the rows (two cells of slack on the safe side)
            const bool outside = hmax < -2.f || hmin > static_cast<float>(g.p_dim_x) + 1.f
                              || vmax < -2.0 || vmin > static_cast<double>(g.p_dim_y) + 1.0;
            o.all_valid = !finite ? 0 : outside ? 2 : (fits && inside) ? 1 : 0;
            o.fits = (fits && finite) ? 1 : 0;
            origin[tid] = o;
#ifdef PB_BP_STATS
            atomicAdd(&g_bp_stats[o.all_valid], 1ull);                                 // [0] mixed, [1] interior, [2] skipped
#endif
        }
        __syncthreads();

        // ---- prologue 2: seed the accumulators, start the TMA pipeline, build the first table -------------------
        // The tile of the volume travels through shared memory (the still unused stage buffers) so that global
        // memory sees runs of TX consecutive voxels; a lane's own voxels are 32 slices apart.
        // (local indices wrap to huge values for voxels before the region's origin and fail the range test)
        const int col0 = warp * CFG::CPW;
        const size_t slice = static_cast<size_t>(g.v_dim_x) * g.v_dim_y;
        float* scratch = reinterpret_cast<float*>(stage_mem);          // [TZ][COLS + 1]
        constexpr int kPitch = CFG::COLS + 1;
        static_assert(size_t(CFG::TZ) * kPitch * 4 <= size_t(CFG::STAGES) * CFG::STAGE_BYTES, "scratch must fit the stages");
        for(int e = tid; e < CFG::TZ * CFG::COLS; e += CFG::THREADS)
        {
            const int c = e % CFG::COLS, zl = e / CFG::COLS;
            const uint32_t x = x0 + c % CFG::TX - g.off_x, y = y0 + c / CFG::TX - g.off_y, z = z0 + zl - g.off_z;
            const bool ok = z < g.v_dim_z && x < g.v_dim_x && y < g.v_dim_y;
            scratch[zl * kPitch + c] = ok ? vol[static_cast<size_t>(z) * slice + static_cast<size_t>(y) * g.v_dim_x + x] : 0.f;
        }
        __syncthreads();
        uint64_t acc[CFG::CPW][CFG::NZ / 2];   // (slice lane + 64j, slice lane + 64j + 32) of column col0 + i, packed
        #pragma unroll
        for(int i = 0; i < CFG::CPW; ++i)
        {
            #pragma unroll
            for(int j = 0; j < CFG::NZ / 2; ++j)
                acc[i][j] = pack2(scratch[(lane + 64u * j) * kPitch + col0 + i],
                                  scratch[(lane + 64u * j + 32u) * kPitch + col0 + i]);
        }
        __syncthreads(); // scratch is dead: the stages may be filled

        if(tid == 0)
        {
            const int pre = count < CFG::STAGES ? count : CFG::STAGES;
            for(int p = 0; p < pre; ++p)
            {
                const uint32_t bar = smem_u32(&bars[p]);
                mbar_expect_tx(bar, CFG::STAGE_BYTES);
                load_box<CFG>(smem_u32(stage_mem + size_t(p) * CFG::STAGE_BYTES), &tmap, bar, origin[p].v0, origin[p].h0,
                              static_cast<int>(first_slot) + p);
            }
        }

        // ---- the lane's slices: z_m exactly as the reference computes it (:39-43), packed in pairs ------------------
        uint64_t zm[CFG::NZ / 2];
        #pragma unroll
        for(int j = 0; j < CFG::NZ / 2; ++j)
            zm[j] = pack2(centered(z0 + lane + 64u * j, g.full_z, g.l_vx_z), centered(z0 + lane + 64u * j + 32u, g.full_z, g.l_vx_z));
        row_consts rc;
        {
            // proj_real_coordinate (:45-50): min = -(dim * size/2) - offset, in the reference's float operations
            const float min_v = __fsub_rn(-__fmul_rn(static_cast<float>(g.p_dim_y), g.l_px_y / 2.f), g.delta_t);
            const float inv_px = __frcp_rn(g.l_px_y);
            rc.neg_min_v2 = pack2(-min_v, -min_v);
            rc.inv_px2 = pack2(inv_px, inv_px);
            rc.neg_px2 = pack2(-g.l_px_y, -g.l_px_y);
            rc.neg_half2 = pack2(-0.5f, -0.5f);
            rc.magic2 = pack2(__uint_as_float(kMagicBits), __uint_as_float(kMagicBits));
            rc.neg_magic2 = pack2(-__uint_as_float(kMagicBits), -__uint_as_float(kMagicBits));
            rc.half2 = pack2(0.5f, 0.5f);
            rc.neg_one2 = pack2(-1.f, -1.f);
            rc.neg_two2 = pack2(-2.f, -2.f);
            rc.zero2 = pack2(g.zero, g.zero);
        }

        // table builder state: thread c < COLS owns tile column c
        const bool builder = tid < CFG::COLS;
        float bx_k = 0.f, by_l = 0.f;
        double z_m0 = 0.0;
        if(builder)
        {
            bx_k = centered(x0 + tid % CFG::TX, g.full_x, g.l_vx_x);
            by_l = centered(y0 + tid / CFG::TX, g.full_y, g.l_vx_y);
            z_m0 = centered_d(z0, g.full_z, g.l_vx_z);
        }
        const uint32_t stage_base0 = smem_u32(stage_mem);
#ifdef PB_BOUNDS_CHECK
        if(tid == 0)
        {
            g_ring_lo_of_cta = stage_base0;   // (dynamic shared memory starts at the same address in every CTA of a launch)
            g_ring_bytes = CFG::STAGES * CFG::STAGE_BYTES;
        }
        __syncthreads();
#endif

        // entry of tile column `col` for projection p, into table set `slot`
        auto build = [&](int p, int col, int slot, float bx_k, float by_l) {
            const box_origin o = origin[p];
            if(o.all_valid == 2)
                return; // the projection is skipped for this tile
            const column_terms ct = project_column(bx_k, by_l, ang.sn[p], ang.cs[p], g);
            const float x1 = floorf(ct.h);
            const bool valid_x = x1 >= 0.f && x1 + 1.f < static_cast<float>(g.p_dim_x);
            // A dead entry (detector column off the detector) reads column 0 of the box with zero weights AND an empty
            // interval of valid slices: dead entries only occur in tiles that take the MIXED path, whose select then
            // yields an exact 0 whatever was staged there (0 * NaN from a bad pixel must not reach voxels the reference
            // leaves untouched, src/openmp/backprojection.cpp:65-71).
            float4 ea = make_float4(0.f, ct.factor, 0.f, 0.f);
            uint32_t ec = 0u;                                    // no slice valid
            int x1rel = 0;
            if(valid_x)
            {
                const double fd = static_cast<double>(ct.factor);
                const double dv = fd * g.dv_scale_d;   // l_vx_z * factor / l_px_y
                const float fx = ct.h - x1;
                const float w = 0.5f * ct.u * ct.u;
                ec = static_cast<uint32_t>(CFG::TZ) << 8;        // every slice valid, unless found otherwise below
                x1rel = min(max(static_cast<int>(x1) - o.h0, 0), CFG::BH - 2);
                ea.z = w * (1.f - fx);
                ea.w = w * fx;
                // detector rows of the tile's first and last slice (affine model in double); one cell of slack
                const double first = row_of(z_m0, fd, g), last = first + dv * (CFG::TZ - 1);
                const bool safe = fmin(first, last) >= 1.0 && fmax(first, last) + 2.0 <= static_cast<double>(g.p_dim_y) - 1.0;
                if(!safe && o.all_valid == 0)
                    ec = valid_slices<CFG::TZ>(z0, ct.factor, first, dv, g);
            }
            // address of row 0 (of the box) of the staged column x1; split layout: in the even plane
            const uint32_t base = stage_base0 + static_cast<uint32_t>(p % CFG::STAGES) * CFG::STAGE_BYTES
                                + 4u * static_cast<uint32_t>(x1rel * CFG::BV);
            ea.x = __uint_as_float(base);
#ifdef PB_BP_STATS
            atomicAdd(&g_bp_stats[3], 1ull);                                           // table entries built
            if(ec != (static_cast<uint32_t>(CFG::TZ) << 8)) atomicAdd(&g_bp_stats[4], 1ull);   // columns near the border
            if(ea.z == 0.f && ea.w == 0.f) atomicAdd(&g_bp_stats[5], 1ull);            // dead columns
            if(o.all_valid == 0) atomicAdd(&g_bp_stats[6], 1ull);                      // entries in mixed tiles
#endif
            tab_a[slot * CFG::COLS + col] = ea;
            tab_c[slot * CFG::COLS + col] = ec;
        };

        if(builder && count > 0)
            build(0, tid, 0, bx_k, by_l);
        __syncthreads();

        // ---- main loop over the projections of the batch -----------------------------------------------------------
        #pragma unroll 1
        for(int p = 0; p < count; ++p)
        {
            if(builder && p + 1 < count)
                build(p + 1, tid, (p + 1) & 1, bx_k, by_l);

            const int stage = p % CFG::STAGES;
            mbar_wait(smem_u32(&bars[stage]), static_cast<uint32_t>((p / CFG::STAGES) & 1));

            const box_origin o = origin[p];
            const float4* ta = tab_a + (p & 1) * CFG::COLS;
            const uint32_t* tc = tab_c + (p & 1) * CFG::COLS;
            // magic bits + the box's first row (split layout: first row PAIR; v0 is a multiple of 8): what an index
            // carries beyond its place inside the box
            const uint32_t ct = kMagicBits + static_cast<uint32_t>(CFG::SPLIT ? static_cast<int>(o.v0) / 2 : static_cast<int>(o.v0));
            if(o.all_valid == 1)
                consume<CFG, false>(acc, ta, tc, zm, rc, ct, col0, lane);
            else if(o.all_valid == 0)
                consume<CFG, true>(acc, ta, tc, zm, rc, ct, col0, lane);

            __syncthreads(); // stage and table[p&1] are free again; table[(p+1)&1] is complete
            if(tid == 0 && p + CFG::STAGES < count)
            {
                const int q = p + CFG::STAGES;
                const uint32_t bar = smem_u32(&bars[stage]);
                mbar_expect_tx(bar, CFG::STAGE_BYTES);
                load_box<CFG>(smem_u32(stage_mem + size_t(stage) * CFG::STAGE_BYTES), &tmap, bar, origin[q].v0,
                              origin[q].h0, static_cast<int>(first_slot) + q);
            }
        }

        // ---- epilogue: back through shared memory, one coalesced store per voxel -------------------------------------
        // (the loop's last __syncthreads guarantees nobody reads the stages any more and no TMA load is in flight)
        #pragma unroll
        for(int i = 0; i < CFG::CPW; ++i)
        {
            #pragma unroll
            for(int j = 0; j < CFG::NZ / 2; ++j)
            {
                float lo, hi;
                unpack2(acc[i][j], lo, hi);
                scratch[(lane + 64u * j) * kPitch + col0 + i] = lo;
                scratch[(lane + 64u * j + 32u) * kPitch + col0 + i] = hi;
            }
        }
        __syncthreads();
        for(int e = tid; e < CFG::TZ * CFG::COLS; e += CFG::THREADS)
        {
            const int c = e % CFG::COLS, zl = e / CFG::COLS;
            const uint32_t x = x0 + c % CFG::TX - g.off_x, y = y0 + c / CFG::TX - g.off_y, z = z0 + zl - g.off_z;
            if(z < g.v_dim_z && x < g.v_dim_x && y < g.v_dim_y)
                vol[static_cast<size_t>(z) * slice + static_cast<size_t>(y) * g.v_dim_x + x] = scratch[zl * kPitch + c];
        }
    }

    // ---- host side -------------------------------------------------------------------------------------------------

    static PFN_cuTensorMapEncodeTiled_v12000 get_encode()
    {
        static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
        if(fn == nullptr)
        {
            void* p = nullptr;
            cudaDriverEntryPointQueryResult q{};
            if(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess
               && q == cudaDriverEntryPointSuccess)
                fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
        }
        return fn;
    }

    static int make_tensor_map(paris_b200_ctx* ctx, const float* d_stack, uint32_t n_row, uint32_t pitch, uint32_t slots,
                               size_t slot_floats, uint32_t box_v, uint32_t box_h, bool split)
    {
        auto& c = ctx->tma;
        const uint32_t layout = split ? kLayoutSplit2 : kLayoutPlain;
        if(c.valid && c.base == d_stack && c.n_row == n_row && c.pitch == pitch && c.slots == slots && c.box_v == box_v
           && c.box_h == box_h && c.layout == layout)
            return PARIS_B200_OK;
        auto encode = get_encode();
        if(encode == nullptr)
        {
            set_error("cuTensorMapEncodeTiled is not available from the driver");
            return PARIS_B200_ECUDA;
        }
        CUresult r;
        if(!split)
        {
            // {detector row, detector column, slot}
            const cuuint64_t dims[3] = {pitch, n_row, slots};
            const cuuint64_t strides[2] = {static_cast<cuuint64_t>(pitch) * 4u, static_cast<cuuint64_t>(slot_floats) * 4u};
            const cuuint32_t box[3] = {box_v, box_h, 1u};
            const cuuint32_t elem[3] = {1u, 1u, 1u};
            r = encode(&c.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(d_stack), dims, strides, box, elem,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        }
        else
        {
            // {row pair, parity plane, detector column, slot}: the box lands as [column][parity][row pair]
            const cuuint64_t dims[4] = {pitch / 2u, 2u, n_row, slots};
            const cuuint64_t strides[3] = {static_cast<cuuint64_t>(pitch / 2u) * 4u, static_cast<cuuint64_t>(pitch) * 4u,
                                           static_cast<cuuint64_t>(slot_floats) * 4u};
            const cuuint32_t box[4] = {box_v / 2u, 2u, box_h, 1u};
            const cuuint32_t elem[4] = {1u, 1u, 1u, 1u};
            r = encode(&c.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(d_stack), dims, strides, box, elem,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        }
        if(r != CUDA_SUCCESS)
        {
            set_error("cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
            return PARIS_B200_ECUDA;
        }
        c.base = d_stack;
        c.n_row = n_row;
        c.pitch = pitch;
        c.slots = slots;
        c.box_v = box_v;
        c.box_h = box_h;
        c.layout = layout;
        c.valid = true;
        return PARIS_B200_OK;
    }

    // Conservative extent of a tile's footprint on the detector, over every angle and every tile of the slab.
    struct footprint
    {
        bool ok;
        int need_h, need_v;
    };

    static footprint tile_footprint(const bp_geometry& g, int tx, int ty, int tz)
    {
        footprint f{false, 0, 0};
        // farthest voxel column from the rotation axis, and farthest slice from the mid-plane, in millimetres
        // (tiles are anchored at global multiples of the tile size and may stick out of the region)
        auto extent = [](uint32_t off, uint32_t n, uint32_t full, float size, int tile) {
            // (covers both anchorings: global multiples of the tile size, and tiles starting at `off`)
            const uint32_t first = off / tile * tile;
            const uint32_t last = std::max((off + n - 1u) / tile * tile + tile - 1u, off + (n + tile - 1u) / tile * tile - 1u);
            const double lo = (static_cast<double>(first) + 0.5 - full / 2.0) * size;
            const double hi = (static_cast<double>(last) + 0.5 - full / 2.0) * size;
            return std::max(std::fabs(lo), std::fabs(hi));
        };
        const double rx = extent(g.off_x, g.v_dim_x, g.full_x, g.l_vx_x, tx);
        const double ry = extent(g.off_y, g.v_dim_y, g.full_y, g.l_vx_y, ty);
        const double rz = extent(g.off_z, g.v_dim_z, g.full_z, g.l_vx_z, tz);
        const double r = std::sqrt(rx * rx + ry * ry);
        const double d_so = g.d_so;
        if(!(d_so > 0.0) || !(d_so - r > 0.05 * d_so))
            return f; // source (nearly) inside the slab: magnification unbounded
        const double fmax = g.d_sd / (d_so - r);
        const double diag = std::sqrt(std::pow((tx - 1) * static_cast<double>(g.l_vx_x), 2)
                                    + std::pow((ty - 1) * static_cast<double>(g.l_vx_y), 2));
        const double dfac = fmax * fmax / g.d_sd * diag;                 // change of the magnification across a tile
        const double span_h = (diag * fmax + r * dfac) / g.l_px_x;       // |d(t*factor)| <= |dt|*f + |t|*|df|
        const double span_v = ((tz - 1) * static_cast<double>(g.l_vx_z) * fmax + rz * dfac) / g.l_px_y;
        // kernel needs (floor(max) + 2) - (floor(min) - 1) < B  <=  span + 1 + 3 < B
        f.need_h = static_cast<int>(std::ceil(span_h)) + 5;
        f.need_v = static_cast<int>(std::ceil(span_v)) + 5; // + 3 (plain) or 7 (split) for the box start alignment
        f.ok = true;
        return f;
    }

    template <class CFG>
    static int launch_cfg(paris_b200_ctx* ctx, const float* d_stack, size_t slot_floats, uint32_t first,
                          const bp_geometry& g, const bp_angles& a, float* d_vol)
    {
        // the tensor map spans the slots this launch can touch: [0, first + count)
        const uint32_t slots = first + static_cast<uint32_t>(a.count);
        PB_TRY(make_tensor_map(ctx, d_stack, g.p_dim_x, g.pitch, slots, slot_floats, CFG::BV, CFG::BH, CFG::SPLIT));
        // x, y: tiles anchored at global multiples of the tile size (first tile holds off, last holds off + dim - 1);
        // z: the same when the slab starts on a tile boundary, else tiles anchored at the slab's first slice (the
        // STRADDLE instantiation: no partly empty tile layers, row anchors stay global)
        const bool straddle = (g.off_z % static_cast<uint32_t>(CFG::TZ)) != 0u;
        auto kern = straddle ? bp_tma_kernel<CFG, true> : bp_tma_kernel<CFG, false>;
        const size_t smem = CFG::smem(straddle);
        PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        auto tiles = [](uint32_t off, uint32_t dim, uint32_t t) { return (off + dim - 1u) / t - off / t + 1u; };
        const uint32_t tiles_x = tiles(g.off_x, g.v_dim_x, CFG::TX), tiles_y = tiles(g.off_y, g.v_dim_y, CFG::TY);
        // (x, y) tiles are numbered in super x super blocks along grid.x (see the kernel); super = tiles_x gives the
        // plain row-major order ("bp_swizzle" = 0)
        const uint32_t super = ctx->bp_swizzle > 0 ? static_cast<uint32_t>(ctx->bp_swizzle) : std::max(tiles_x, tiles_y);
        const uint32_t super_blocks = ((tiles_x + super - 1u) / super) * ((tiles_y + super - 1u) / super);
        const dim3 grid(super_blocks * super * super, 1u,
                        straddle ? (g.v_dim_z + CFG::TZ - 1u) / CFG::TZ : tiles(g.off_z, g.v_dim_z, CFG::TZ));
        if(grid.y > 65535u || grid.z > 65535u)
        {
            set_error("slab too large for the backprojection grid");
            return PARIS_B200_EINVAL;
        }
        kern<<<grid, CFG::THREADS, smem, ctx->compute>>>(ctx->tma.map, d_vol, g, a, first, tiles_x, tiles_y, super);
        PB_CUDA(cudaGetLastError());
        ++ctx->launches;
        ++ctx->bp_launches_tma;
        std::snprintf(ctx->bp_last_kernel, sizeof(ctx->bp_last_kernel), "bp_tma_kernel<%dx%dx%d tile, box %dx%d, %d stages, %s%s>",
                      CFG::TX, CFG::TY, CFG::TZ, CFG::BH, CFG::BV, CFG::STAGES, CFG::SPLIT ? "split2" : "plain",
                      straddle ? ", straddle" : "");
        return PARIS_B200_OK;
    }

    //                            TX  TY  NZ CPW  BH   BV  STAGES SPLIT
    using cfg_fine         = tile_cfg<16, 16, 2, 16, 40, 96, 6>;          // ~1 detector row per voxel (PARIS-derived volumes)
    using cfg_coarse       = tile_cfg<16, 16, 2, 16, 64, 164, 4>;         // ~2 rows per voxel, plain stack layout
    using cfg_coarse_split = tile_cfg<16, 16, 2, 16, 64, 168, 4, true>;   // ~2 rows per voxel, parity-split stack layout
    // half tiles: 256 threads, two CTAs per SM, so that one CTA's barriers, table building and volume tile
    // traffic are covered by the other CTA's interpolation work
    using cfg_fine_half         = tile_cfg<16, 8, 2, 16, 32, 96, 5>;
    using cfg_coarse_split_half = tile_cfg<16, 8, 2, 16, 52, 168, 3, true>;
    // tall tiles: 8 x 8 columns x 128 slices, four slices per lane -- the per-(column, projection) table entry is
    // read once per FOUR updates of a lane and built for half as many columns per voxel; used when the slab is
    // at least one such tile thick
    using cfg_fine_tall         = tile_cfg<8, 8, 4, 8, 24, 176, 4>;
    using cfg_coarse_split_tall = tile_cfg<8, 8, 4, 8, 32, 312, 2, true>;

    template <class CFG>
    static bool fits_cfg(const bp_geometry& g)
    {
        const footprint f = tile_footprint(g, CFG::TX, CFG::TY, CFG::TZ);
        return f.ok && f.need_h <= CFG::BH && f.need_v + (CFG::SPLIT ? 7 : 3) <= CFG::BV;
    }

#ifdef PB_BOUNDS_CHECK
    // out[0..3]: sample loads checked, loads outside the staged ring, lowest / highest offending offset; resets
    extern "C" int paris_b200_debug_bounds(unsigned long long* out)
    {
        cudaDeviceSynchronize();
        cudaMemcpyFromSymbol(out, g_bounds, sizeof(unsigned long long) * 4);
        const unsigned long long fresh[4] = {0ull, 0ull, ~0ull, 0ull};
        cudaMemcpyToSymbol(g_bounds, fresh, sizeof(fresh));
        return 0;
    }
#endif

#ifdef PB_BP_STATS
    extern "C" int paris_b200_debug_bp_stats(unsigned long long* out)
    {
        cudaDeviceSynchronize();
        cudaMemcpyFromSymbol(out, g_bp_stats, sizeof(unsigned long long) * 8);
        unsigned long long zero[8] = {0};
        cudaMemcpyToSymbol(g_bp_stats, zero, sizeof(zero));
        return 0;
    }
#endif

    // Load every instantiation now (CUDA loads kernels lazily, on first launch, and that load synchronises the
    // context: inside a group step, where streams wait for flags other work has yet to set, it deadlocks).
    template <class CFG>
    static void preload_cfg()
    {
        cudaFuncAttributes a{};
        (void)cudaFuncGetAttributes(&a, bp_tma_kernel<CFG, false>);
        (void)cudaFuncGetAttributes(&a, bp_tma_kernel<CFG, true>);
    }

    void preload_bp_tma_kernels()
    {
        preload_cfg<cfg_fine>();
        preload_cfg<cfg_coarse>();
        preload_cfg<cfg_coarse_split>();
        preload_cfg<cfg_fine_half>();
        preload_cfg<cfg_coarse_split_half>();
        preload_cfg<cfg_fine_tall>();
        preload_cfg<cfg_coarse_split_tall>();
        (void)cudaGetLastError();
    }

    int launch_bp_tma(paris_b200_ctx* ctx, const float* d_stack, size_t slot_floats, uint32_t first,
                      const bp_geometry& g, const bp_angles& a, float* d_vol, bool required, bool* handled)
    {
        *handled = false;
        // (box origins are kept as 16-bit detector coordinates)
        const bool aligned = (reinterpret_cast<uintptr_t>(d_stack) % 16u) == 0u && (g.pitch % 8u) == 0u
                          && g.p_dim_x <= 32000u && g.p_dim_y <= 32000u;
        const bool half = ctx->bp_tile != 1;   // 0 = automatic (tall, then half tiles when they fit), 1 = full tiles only, 2 = half or full
#define PB_TRY_CFG(CFG)                                                                            \
        if(fits_cfg<CFG>(g))                                                                       \
        {                                                                                          \
            PB_TRY(launch_cfg<CFG>(ctx, d_stack, slot_floats, first, g, a, d_vol));                \
            *handled = true;                                                                       \
            return PARIS_B200_OK;                                                                  \
        }
        // Tiles are anchored at multiples of the tile size in GLOBAL voxel indices in x and y, so a region whose offsets
        // are not multiples of the tile size pays for partly empty tiles; the tall tile (fastest per voxel) is used only
        // when the voxels its tiles cover, times its relative cost per voxel, are fewer than the half tile's.
        auto covered = [&](uint32_t tx, uint32_t ty, uint32_t tz) {
            auto span = [](uint32_t off, uint32_t dim, uint32_t t) { return ((off + dim - 1u) / t - off / t + 1u) * static_cast<double>(t); };
            // (z: a slab that does not start on a tile boundary gets tiles anchored at its first slice, at ~10 % more
            // instructions per voxel)
            const double z = (g.off_z % tz) ? 1.1 * ((g.v_dim_z + tz - 1u) / tz) * static_cast<double>(tz) : span(g.off_z, g.v_dim_z, tz);
            return span(g.off_x, g.v_dim_x, tx) * span(g.off_y, g.v_dim_y, ty) * z;
        };
        const bool tall = ctx->bp_tile == 0 && covered(8, 8, 128) <= 1.035 * covered(16, 8, 64);
        if(aligned)
        {
            if(g.layout == kLayoutSplit2)
            {
                if(tall)
                    PB_TRY_CFG(cfg_coarse_split_tall)
                if(half)
                    PB_TRY_CFG(cfg_coarse_split_half)
                PB_TRY_CFG(cfg_coarse_split)
            }
            else
            {
                if(tall)
                    PB_TRY_CFG(cfg_fine_tall)
                if(half)
                    PB_TRY_CFG(cfg_fine_half)
                PB_TRY_CFG(cfg_fine)
                PB_TRY_CFG(cfg_coarse)
            }
        }
#undef PB_TRY_CFG
        const footprint f = tile_footprint(g, 16, 16, 64);
        if(required)
        {
            set_error("geometry does not fit the TMA kernel's tiles (footprint %d x %d detector cells per tile, layout %u)",
                      f.need_h, f.need_v, g.layout);
            return PARIS_B200_ESTATE;
        }
        return PARIS_B200_OK;
    }
}
