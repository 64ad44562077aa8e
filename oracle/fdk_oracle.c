/*
 * oracle/fdk_oracle.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Plain-C restatement of the reference's FDK hot path (weight -> ramp filter ->
 * voxel-driven backprojection) as implemented by its OpenMP backend and the
 * backend-agnostic wrappers.  Every function cites the reference file:line it
 * follows; the float expression ORDER is kept so that, built with the same
 * flags as oracle/_ref (no FMA contraction), the two agree bit for bit
 * (tests/test_oracle.py pins that, and tests/golden/ holds outputs of the
 * reference itself).
 *
 * Parity status: the reference has no tests or golden vectors of its own
 * (SURVEY F2), and its FFT lives in FFTW3f, an un-vendored, un-pinned
 * dependency.  This oracle is pinned against the reference's own compiled code
 * (oracle/_ref, built from /root/reference/src by oracle/Makefile) with
 * oracle/fft_shim.c standing in for FFTW in both; at the FFT boundary itself
 * parity is "unpinned" in the sense of the task statement.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference
 * legs may use this file.  The product (paris_b200/) never does.
 *
 * Unlike the reference, nothing here is frozen in function-local statics
 * (SURVEY F8): geometry is passed per call.  The values computed are the same.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <omp.h>

#include "shim/fftw3.h"

typedef struct
{
    uint32_t n_row, n_col;
    float l_px_row, l_px_col, delta_s, delta_t, d_so, d_od, delta_phi;
} oracle_detector_geometry; /* src/geometry.h:30-47 */

typedef struct
{
    uint32_t dim_x, dim_y, dim_z;
    float l_vx_x, l_vx_y, l_vx_z;
} oracle_volume_geometry; /* src/geometry.h:49-58 */

typedef struct
{
    uint32_t x1, x2, y1, y2, z1, z2;
} oracle_roi; /* src/region_of_interest.h:30-38 */

typedef struct
{
    uint32_t dim_x, dim_y, dim_z, remainder;
    int num;
} oracle_subvolume_info; /* src/subvolume_information.h:30-34 + src/geometry.h:60-69 */

int oracle_num_threads(void) { return omp_get_max_threads(); }

/* ---- geometry ------------------------------------------------------------------------- */

/* src/geometry.cpp:36-67 (make_volume_geometry) */
void oracle_calculate_volume_geometry(const oracle_detector_geometry* det, oracle_volume_geometry* out)
{
    const float n_row = (float)det->n_row;
    const float l_px_row = det->l_px_row;
    const float delta_s = fabsf(det->delta_s * l_px_row);

    const float n_col = (float)det->n_col;
    const float l_px_col = det->l_px_col;
    const float delta_t = fabsf(det->delta_t * l_px_col);

    const float d_so = fabsf(det->d_so);
    const float d_sd = fabsf(det->d_od) + d_so;

    const float alpha = atanf((((n_row * l_px_row) / 2.f) + delta_s) / d_sd);
    const float r = d_so * sinf(alpha);

    out->l_vx_x = r / ((((n_row * l_px_row) / 2.f) + delta_s) / l_px_row);
    out->l_vx_y = out->l_vx_x;

    out->dim_x = (uint32_t)((2.f * r) / out->l_vx_x);
    out->dim_y = out->dim_x;

    out->l_vx_z = out->l_vx_x;
    out->dim_z = (uint32_t)(((n_col * l_px_col / 2.f) + delta_t) * (d_so / d_sd) * (2.f / out->l_vx_z));
}

/* src/geometry.cpp:86-130 (apply_roi): dim = x2 - x1, +1 iff x1 == 0; rejected ROIs leave the geometry unchanged */
void oracle_apply_roi(const oracle_volume_geometry* vol, const oracle_roi* r, oracle_volume_geometry* out)
{
    *out = *vol;
    if(r->x1 < r->x2 && r->y1 < r->y2 && r->z1 < r->z2)
    {
        uint32_t dim_x = r->x2 - r->x1;
        uint32_t dim_y = r->y2 - r->y1;
        uint32_t dim_z = r->z2 - r->z1;
        if(r->x1 == 0) ++dim_x;
        if(r->y1 == 0) ++dim_y;
        if(r->z1 == 0) ++dim_z;
        if(dim_x <= vol->dim_x && dim_y <= vol->dim_y && dim_z <= vol->dim_z)
        {
            out->dim_x = dim_x;
            out->dim_y = dim_y;
            out->dim_z = dim_z;
        }
    }
}

/* z-slab split: src/cuda/subvolume_information.cpp:112-116 (dim_z / num, remainder on the last slab,
 * src/make_volume.cpp:32-34); slab offset = id * dim_z (src/main.cpp:96).  num is given, not memory-derived. */
void oracle_make_subvolume_information(const oracle_volume_geometry* vol, int num, oracle_subvolume_info* out)
{
    out->dim_x = vol->dim_x;
    out->dim_y = vol->dim_y;
    out->dim_z = vol->dim_z / (uint32_t)num;
    out->remainder = vol->dim_z % (uint32_t)num;
    out->num = num;
}

/* ---- weighting ------------------------------------------------------------------------ */

/* src/weighting.cpp:32-45 (constants) + src/openmp/weighting.cpp:32-57 (loop); in place */
void oracle_weight(float* p, const oracle_detector_geometry* det)
{
    const float n_row_f = (float)det->n_row;
    const float n_col_f = (float)det->n_col;
    const float h_min = (det->delta_s * det->l_px_row) - ((n_row_f * det->l_px_row) / 2);
    const float v_min = (det->delta_t * det->l_px_col) - ((n_col_f * det->l_px_col) / 2);
    const float d_sd = fabsf(det->d_so) + fabsf(det->d_od);
    const float l_px_row = det->l_px_row;
    const float l_px_col = det->l_px_col;
    const uint32_t dim_x = det->n_row, dim_y = det->n_col;

    #pragma omp parallel for collapse(2)
    for(uint32_t t = 0u; t < dim_y; ++t)
    {
        for(uint32_t s = 0u; s < dim_x; ++s)
        {
            const uint32_t coord = s + t * dim_x;
            const float s_f = (float)s;
            const float t_f = (float)t;
            const float h_s = (l_px_row / 2) + s_f * l_px_row + h_min;
            const float v_t = (l_px_col / 2) + t_f * l_px_col + v_min;
            const float w_st = d_sd / sqrtf(d_sd * d_sd + h_s * h_s + v_t * v_t);
            p[coord] *= w_st;
        }
    }
}

/* ---- filtering ------------------------------------------------------------------------ */

/* src/filtering.cpp:38 */
uint32_t oracle_filter_size(uint32_t n_row)
{
    return (uint32_t)(2 * pow(2.0, ceil(log2((double)n_row))));
}

/* src/openmp/filtering.cpp:52-73 (make_filter_real) + :139-165 (make_filter); k_out[x], x = 0..size/2
 * (the reference stores the same value in re and im) */
void oracle_make_filter(uint32_t size, float tau, float* k_out)
{
    const uint32_t size_trans = size / 2 + 1;
    float* r = (float*)fftwf_malloc(size * sizeof(float));
    fftwf_complex* k = (fftwf_complex*)fftwf_malloc(size_trans * sizeof(fftwf_complex));
    fftwf_plan plan = fftwf_plan_dft_r2c_1d((int)size, r, k, FFTW_MEASURE | FFTW_PRESERVE_INPUT);

    const int32_t j0 = -((int32_t)size - 2) / 2;
    const float pi_f = (float)M_PI;
    for(uint32_t x = 0u; x < size; ++x)
    {
        const int32_t j = j0 + (int32_t)x;
        if(j == 0)
            r[x] = (1.f / 8.f) * (1.f / powf(tau, 2.f));
        else if(j % 2 == 0)
            r[x] = 0.f;
        else
            r[x] = -(1.f / (2.f * (float)(j * j) * (pi_f * pi_f) * (tau * tau)));
    }

    fftwf_execute(plan);

    for(uint32_t x = 0u; x < size_trans; ++x)
        k_out[x] = tau * fabsf(sqrtf(powf(k[x][0], 2.f) + powf(k[x][1], 2.f)));

    fftwf_destroy_plan(plan);
    fftwf_free(r);
    fftwf_free(k);
}

/* src/openmp/filtering.cpp:167-219 (apply_filter: expand :75, r2c, do_filtering :92, c2r, shrink :107,
 * normalize :120); p is dim_x x dim_y (dim_x = n_row fastest), in place; k has filter_size/2+1 entries */
void oracle_apply_filter(float* p, uint32_t dim_x, uint32_t dim_y, const float* k, uint32_t filter_size)
{
    const uint32_t size_trans = filter_size / 2 + 1;
    const int n = (int)filter_size;
    float* p_exp = (float*)fftwf_malloc((size_t)filter_size * dim_y * sizeof(float));
    fftwf_complex* p_trans = (fftwf_complex*)fftwf_malloc((size_t)size_trans * dim_y * sizeof(fftwf_complex));
    const int nembed_exp = (int)filter_size, nembed_trans = (int)size_trans;

    fftwf_plan forward = fftwf_plan_many_dft_r2c(1, &n, (int)dim_y, p_exp, &nembed_exp, 1, (int)filter_size,
                                                 p_trans, &nembed_trans, 1, (int)size_trans,
                                                 FFTW_MEASURE | FFTW_PRESERVE_INPUT);
    fftwf_plan inverse = fftwf_plan_many_dft_c2r(1, &n, (int)dim_y, p_trans, &nembed_trans, 1, (int)size_trans,
                                                 p_exp, &nembed_exp, 1, (int)filter_size,
                                                 FFTW_MEASURE | FFTW_DESTROY_INPUT);

    /* expand: zero-pad each row on the right */
    #pragma omp parallel for collapse(2)
    for(uint32_t y = 0u; y < dim_y; ++y)
        for(uint32_t x = 0u; x < filter_size; ++x)
            p_exp[x + (size_t)y * filter_size] = x < dim_x ? p[x + (size_t)y * dim_x] : 0.f;

    fftwf_execute(forward);

    #pragma omp parallel for collapse(2)
    for(uint32_t y = 0u; y < dim_y; ++y)
        for(uint32_t x = 0u; x < size_trans; ++x)
        {
            const size_t coord = x + (size_t)y * size_trans;
            p_trans[coord][0] *= k[x];
            p_trans[coord][1] *= k[x];
        }

    fftwf_execute(inverse);

    /* shrink + normalize */
    #pragma omp parallel for collapse(2)
    for(uint32_t y = 0u; y < dim_y; ++y)
        for(uint32_t x = 0u; x < dim_x; ++x)
        {
            float val = p_exp[x + (size_t)y * filter_size];
            val /= (float)filter_size;
            p[x + (size_t)y * dim_x] = val;
        }

    fftwf_destroy_plan(forward);
    fftwf_destroy_plan(inverse);
    fftwf_free(p_exp);
    fftwf_free(p_trans);
}

/* src/filtering.cpp:32-45: filter_size from n_row, tau = l_px_row, K built once, then apply */
void oracle_filter(float* p, const oracle_detector_geometry* det)
{
    const uint32_t filter_size = oracle_filter_size(det->n_row);
    float* k = (float*)malloc((filter_size / 2 + 1) * sizeof(float));
    oracle_make_filter(filter_size, det->l_px_row, k);
    oracle_apply_filter(p, det->n_row, det->n_col, k, filter_size);
    free(k);
}

/* ---- backprojection -------------------------------------------------------------------- */

/* src/openmp/backprojection.cpp:39-43 */
static inline float vol_centered_coordinate(uint32_t coord, uint32_t dim, float size)
{
    const float size2 = size / 2.f;
    return -((float)dim * size2) + size2 + (float)coord * size;
}

/* src/openmp/backprojection.cpp:45-50 */
static inline float proj_real_coordinate(float coord, uint32_t dim, float size, float offset)
{
    const float size2 = size / 2.f;
    const float min = -((float)dim * size2) - offset;
    return (coord - min) / size - (1.f / 2.f);
}

/* src/openmp/backprojection.cpp:52-84: bilinear, zero unless all four neighbours are inside */
static inline float interpolate(const float* p, float x, float y, uint32_t dim_x, uint32_t dim_y)
{
    const float x1 = floorf(x);
    const float x2 = x1 + 1.f;
    const float y1 = floorf(y);
    const float y2 = y1 + 1.f;

    float interp = 0.f;
    if(x1 >= 0.f && x2 < (float)dim_x && y1 >= 0.f && y2 < (float)dim_y)
    {
        const uint32_t x1u = (uint32_t)x1, x2u = (uint32_t)x2, y1u = (uint32_t)y1, y2u = (uint32_t)y2;
        const float q11 = p[x1u + y1u * dim_x];
        const float q12 = p[x1u + y2u * dim_x];
        const float q21 = p[x2u + y1u * dim_x];
        const float q22 = p[x2u + y2u * dim_x];
        const float interp_y1 = (x2 - x) / (x2 - x1) * q11 + (x - x1) / (x2 - x1) * q21;
        const float interp_y2 = (x2 - x) / (x2 - x1) * q12 + (x - x1) / (x2 - x1) * q22;
        interp = (y2 - y) / (y2 - y1) * interp_y1 + (y - y1) / (y2 - y1) * interp_y2;
    }
    return interp;
}

/*
 * src/backprojection.cpp:37-69 (angle -> sin/cos, delta_s/t in mm) +
 * src/openmp/backprojection.cpp:86-153 (do_backprojection) and :156-199 (constants).
 * vol is v_dim_x x v_dim_y x v_dim_z (x fastest), accumulated into.  vol_full is the FULL volume
 * geometry; with enable_roi the voxel index is shifted by (roi.x1, roi.y1, roi.z1); v_offset is the
 * z offset of the slab.
 */
void oracle_backproject(const float* proj, uint32_t idx, float phi_deg, int enable_angles,
                        float* vol, uint32_t v_dim_x, uint32_t v_dim_y, uint32_t v_dim_z, uint32_t v_offset,
                        const oracle_detector_geometry* det, const oracle_volume_geometry* vol_full,
                        int enable_roi, const oracle_roi* roi)
{
    const float delta_s = det->delta_s * det->l_px_row;
    const float delta_t = det->delta_t * det->l_px_col;

    float phi = enable_angles ? phi_deg : (float)idx * det->delta_phi;
    phi *= (float)M_PI / 180.f;
    const float sn = sinf(phi);
    const float cs = cosf(phi);

    const uint32_t fx = vol_full->dim_x, fy = vol_full->dim_y, fz = vol_full->dim_z;
    const float l_vx_x = vol_full->l_vx_x, l_vx_y = vol_full->l_vx_y, l_vx_z = vol_full->l_vx_z;
    const float l_px_x = det->l_px_row, l_px_y = det->l_px_col;
    const float d_so = det->d_so;
    const float d_sd = fabsf(det->d_so) + fabsf(det->d_od);
    const uint32_t p_dim_x = det->n_row, p_dim_y = det->n_col;
    const uint32_t rx = enable_roi ? roi->x1 : 0u, ry = enable_roi ? roi->y1 : 0u, rz = enable_roi ? roi->z1 : 0u;

    #pragma omp parallel for collapse(3)
    for(uint32_t m = 0u; m < v_dim_z; ++m)
    {
        for(uint32_t l = 0u; l < v_dim_y; ++l)
        {
            for(uint32_t k = 0u; k < v_dim_x; ++k)
            {
                const uint32_t coord = k + l * v_dim_x + m * v_dim_x * v_dim_y;

                const float x_k = vol_centered_coordinate(k + rx, fx, l_vx_x);
                const float y_l = vol_centered_coordinate(l + ry, fy, l_vx_y);
                const float z_m = vol_centered_coordinate(m + rz + v_offset, fz, l_vx_z);

                const float s = x_k * cs + y_l * sn;
                const float t = -x_k * sn + y_l * cs;

                const float factor = d_sd / (s + d_so);
                const float h = proj_real_coordinate(t * factor, p_dim_x, l_px_x, delta_s);
                const float v = proj_real_coordinate(z_m * factor, p_dim_y, l_px_y, delta_t);

                const float det_val = interpolate(proj, h, v, p_dim_x, p_dim_y);

                const float u = -(d_so / (s + d_so));
                vol[coord] += 0.5f * det_val * u * u;
            }
        }
    }
}

/* The hot loop of src/main.cpp:98-105 over an in-memory stack (same contract as oracle/ref_api.cpp) */
void oracle_reconstruct(const float* stack, uint32_t n_proj, uint32_t first_idx, uint32_t idx_stride,
                        float* vol, uint32_t v_dim_x, uint32_t v_dim_y, uint32_t v_dim_z,
                        const oracle_detector_geometry* det, const oracle_volume_geometry* vol_full,
                        int enable_roi, const oracle_roi* roi, double* times)
{
    const size_t px = (size_t)det->n_row * det->n_col;
    const uint32_t filter_size = oracle_filter_size(det->n_row);
    float* k = (float*)malloc((filter_size / 2 + 1) * sizeof(float));
    float* p = (float*)malloc(px * sizeof(float));
    double t_w = 0.0, t_f = 0.0, t_b = 0.0;

    oracle_make_filter(filter_size, det->l_px_row, k);
    for(uint32_t i = 0u; i < n_proj; ++i)
    {
        memcpy(p, stack + (size_t)i * px, px * sizeof(float));
        const double t0 = omp_get_wtime();
        oracle_weight(p, det);
        const double t1 = omp_get_wtime();
        oracle_apply_filter(p, det->n_row, det->n_col, k, filter_size);
        const double t2 = omp_get_wtime();
        oracle_backproject(p, first_idx + i * idx_stride, 0.f, 0, vol, v_dim_x, v_dim_y, v_dim_z, 0u,
                           det, vol_full, enable_roi, roi);
        const double t3 = omp_get_wtime();
        if(i > 0 || n_proj == 1)
        {
            t_w += t1 - t0;
            t_f += t2 - t1;
            t_b += t3 - t2;
        }
    }
    if(times)
    {
        times[0] = t_w;
        times[1] = t_f;
        times[2] = t_b;
    }
    free(p);
    free(k);
}

/* ---- full-size checks block by block ---------------------------------------------------- */

/*
 * Rows [row0, row0 + n_rows) of weight -> filter, in place in a FULL-SIZE projection buffer.  Not a different
 * algorithm: the weight of a pixel depends only on its own (s, t) (src/openmp/weighting.cpp:44-52, t absolute), and
 * the reference filters detector rows independently of each other (batched 1-D plans, howmany = n_col,
 * src/openmp/filtering.cpp:199-204), so a band of rows goes through exactly the operations it would see inside the
 * whole projection.  tests/test_oracle.py checks bit-identity with the crop of the full-projection result.
 * This is what makes oracle checks at 2048^2 x 1440 affordable: a 32^3 block of voxels only ever reads a band of
 * detector rows.
 */
void oracle_weight_filter_rows(float* p, const oracle_detector_geometry* det, const float* k, uint32_t filter_size,
                               uint32_t row0, uint32_t n_rows)
{
    const float n_row_f = (float)det->n_row;
    const float n_col_f = (float)det->n_col;
    const float h_min = (det->delta_s * det->l_px_row) - ((n_row_f * det->l_px_row) / 2);
    const float v_min = (det->delta_t * det->l_px_col) - ((n_col_f * det->l_px_col) / 2);
    const float d_sd = fabsf(det->d_so) + fabsf(det->d_od);
    const float l_px_row = det->l_px_row;
    const float l_px_col = det->l_px_col;
    const uint32_t dim_x = det->n_row;

    #pragma omp parallel for collapse(2)
    for(uint32_t t = row0; t < row0 + n_rows; ++t)
    {
        for(uint32_t s = 0u; s < dim_x; ++s)
        {
            const size_t coord = s + (size_t)t * dim_x;
            const float s_f = (float)s;
            const float t_f = (float)t;
            const float h_s = (l_px_row / 2) + s_f * l_px_row + h_min;
            const float v_t = (l_px_col / 2) + t_f * l_px_col + v_min;
            const float w_st = d_sd / sqrtf(d_sd * d_sd + h_s * h_s + v_t * v_t);
            p[coord] *= w_st;
        }
    }
    oracle_apply_filter(p + (size_t)row0 * dim_x, dim_x, n_rows, k, filter_size);
}

/*
 * The hot loop of src/main.cpp:98-105 for ONE box of voxels through the ROI path (src/openmp/backprojection.cpp:
 * 105-118), fed with only the band of detector rows the box can touch.  band_stack: n_proj x n_rows x n_row raw
 * samples (rows [row0, row0 + n_rows) of every projection).  Every other row of the scratch projection is NaN, so
 * a band chosen too small poisons the result instead of going unnoticed.
 */
void oracle_reconstruct_block(const float* band_stack, uint32_t n_proj, uint32_t first_idx, uint32_t idx_stride,
                              uint32_t row0, uint32_t n_rows, float* vol, uint32_t v_dim_x, uint32_t v_dim_y,
                              uint32_t v_dim_z, const oracle_detector_geometry* det,
                              const oracle_volume_geometry* vol_full, const oracle_roi* roi)
{
    const size_t px = (size_t)det->n_row * det->n_col;
    const size_t band_px = (size_t)det->n_row * n_rows;
    const uint32_t filter_size = oracle_filter_size(det->n_row);
    float* k = (float*)malloc((filter_size / 2 + 1) * sizeof(float));
    float* p = (float*)malloc(px * sizeof(float));
    for(size_t i = 0; i < px; ++i)
        p[i] = NAN;
    oracle_make_filter(filter_size, det->l_px_row, k);
    for(uint32_t i = 0u; i < n_proj; ++i)
    {
        memcpy(p + (size_t)row0 * det->n_row, band_stack + (size_t)i * band_px, band_px * sizeof(float));
        oracle_weight_filter_rows(p, det, k, filter_size, row0, n_rows);
        oracle_backproject(p, first_idx + i * idx_stride, 0.f, 0, vol, v_dim_x, v_dim_y, v_dim_z, 0u, det, vol_full, 1,
                           roi);
    }
    free(p);
    free(k);
}

void oracle_set_num_threads(int n) { omp_set_num_threads(n > 0 ? n : 1); }
