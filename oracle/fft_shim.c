/*
 * oracle/fft_shim.c -- TEST INFRASTRUCTURE, not product code.
 *
 * The six FFTW3f entry points the reference's OpenMP backend calls
 * (/root/reference/src/openmp/filtering.cpp:41,48,136,149,153,199,201-204,208,214),
 * implemented from the published definition of the DFT:
 *
 *   r2c:  X[k] = sum_n x[n] exp(-2 pi i n k / N),  k = 0..N/2
 *   c2r:  x[n] = sum_k X[k] exp(+2 pi i n k / N),  Hermitian input, unnormalised,
 *         imaginary parts of X[0] and X[N/2] ignored (FFTW manual, "The 1d Real-data DFT")
 *
 * FFTW3 is an un-vendored and un-pinned dependency of the reference
 * (CMakeLists.txt:50, FIND_PACKAGE(FFTW) without a version), so its rounding is
 * not part of any contract; this stand-in computes in DOUBLE and rounds once to
 * float on output -- the most defensible answer for "any correct FFTW".
 * Power-of-two N only (the reference only ever asks for N = 2*2^k,
 * src/filtering.cpp:38).
 *
 * Batched plans are executed with OpenMP over the batch.  The reference's FFTW
 * use is single-threaded; threading here only makes the CPU baseline faster.
 */
#include "shim/fftw3.h"

#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

struct oracle_fft_plan_s
{
    int kind; /* 0 = r2c, 1 = c2r */
    int n;
    int howmany;
    void* in;
    void* out;
    int idist;
    int odist;
    /* tables for the half-size complex transform */
    int m;          /* n / 2 */
    int log2m;
    double* tw_re;  /* exp(-2 pi i k / n), k = 0..n/2 */
    double* tw_im;
    uint32_t* rev;  /* bit reversal for size m */
};

void* fftwf_malloc(size_t n)
{
    void* p = NULL;
    if(posix_memalign(&p, 64, n ? n : 64) != 0)
        return NULL;
    return p;
}

void fftwf_free(void* p) { free(p); }

static fftwf_plan make_plan(int kind, int n, int howmany, void* in, int idist, void* out, int odist)
{
    if(n < 2 || (n & (n - 1)) != 0)
        return NULL;
    struct oracle_fft_plan_s* p = (struct oracle_fft_plan_s*)calloc(1, sizeof(*p));
    p->kind = kind;
    p->n = n;
    p->howmany = howmany;
    p->in = in;
    p->out = out;
    p->idist = idist;
    p->odist = odist;
    p->m = n / 2;
    p->log2m = 0;
    while((1 << p->log2m) < p->m)
        ++p->log2m;
    p->tw_re = (double*)malloc(sizeof(double) * (size_t)(n / 2 + 1));
    p->tw_im = (double*)malloc(sizeof(double) * (size_t)(n / 2 + 1));
    for(int k = 0; k <= n / 2; ++k)
    {
        const double a = -2.0 * M_PI * (double)k / (double)n;
        p->tw_re[k] = cos(a);
        p->tw_im[k] = sin(a);
    }
    p->rev = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)p->m);
    for(int i = 0; i < p->m; ++i)
    {
        uint32_t r = 0;
        for(int b = 0; b < p->log2m; ++b)
            if(i & (1 << b))
                r |= 1u << (p->log2m - 1 - b);
        p->rev[i] = r;
    }
    return p;
}

fftwf_plan fftwf_plan_dft_r2c_1d(int n, float* in, fftwf_complex* out, unsigned flags)
{
    (void)flags;
    return make_plan(0, n, 1, in, n, out, n / 2 + 1);
}

fftwf_plan fftwf_plan_many_dft_r2c(int rank, const int* n, int howmany,
                                   float* in, const int* inembed, int istride, int idist,
                                   fftwf_complex* out, const int* onembed, int ostride, int odist,
                                   unsigned flags)
{
    (void)inembed; (void)onembed; (void)flags;
    if(rank != 1 || istride != 1 || ostride != 1)
        return NULL;
    return make_plan(0, n[0], howmany, in, idist, out, odist);
}

fftwf_plan fftwf_plan_many_dft_c2r(int rank, const int* n, int howmany,
                                   fftwf_complex* in, const int* inembed, int istride, int idist,
                                   float* out, const int* onembed, int ostride, int odist,
                                   unsigned flags)
{
    (void)inembed; (void)onembed; (void)flags;
    if(rank != 1 || istride != 1 || ostride != 1)
        return NULL;
    return make_plan(1, n[0], howmany, in, idist, out, odist);
}

void fftwf_destroy_plan(fftwf_plan p)
{
    if(!p)
        return;
    free(p->tw_re);
    free(p->tw_im);
    free(p->rev);
    free(p);
}

/* In-place complex FFT of size m on (re, im); sign = -1 forward, +1 inverse (unnormalised).
 * Twiddles exp(-2 pi i k / n) with n = 2m are read at even k. */
static void cfft(const struct oracle_fft_plan_s* p, double* re, double* im, int sign)
{
    const int m = p->m;
    for(int i = 0; i < m; ++i)
    {
        const int j = (int)p->rev[i];
        if(j > i)
        {
            double t = re[i]; re[i] = re[j]; re[j] = t;
            t = im[i]; im[i] = im[j]; im[j] = t;
        }
    }
    for(int len = 2; len <= m; len <<= 1)
    {
        const int half = len >> 1;
        const int step = p->n / len; /* index stride into the size-n twiddle table */
        for(int base = 0; base < m; base += len)
        {
            for(int j = 0; j < half; ++j)
            {
                const double wr = p->tw_re[j * step];
                const double wi = sign < 0 ? p->tw_im[j * step] : -p->tw_im[j * step];
                const int a = base + j;
                const int b = a + half;
                const double xr = re[b] * wr - im[b] * wi;
                const double xi = re[b] * wi + im[b] * wr;
                re[b] = re[a] - xr;
                im[b] = im[a] - xi;
                re[a] += xr;
                im[a] += xi;
            }
        }
    }
}

static void r2c_one(const struct oracle_fft_plan_s* p, const float* x, fftwf_complex* X, double* re, double* im)
{
    const int n = p->n, m = p->m;
    if(m == 1)
    {
        X[0][0] = (float)((double)x[0] + (double)x[1]); X[0][1] = 0.f;
        X[1][0] = (float)((double)x[0] - (double)x[1]); X[1][1] = 0.f;
        return;
    }
    for(int i = 0; i < m; ++i)
    {
        re[i] = (double)x[2 * i];
        im[i] = (double)x[2 * i + 1];
    }
    cfft(p, re, im, -1);
    for(int k = 0; k <= m; ++k)
    {
        const int k1 = k % m;
        const int k2 = (m - k) % m;
        /* E = (Z[k] + conj Z[m-k]) / 2, O = (Z[k] - conj Z[m-k]) / (2i) */
        const double er = 0.5 * (re[k1] + re[k2]);
        const double ei = 0.5 * (im[k1] - im[k2]);
        const double dr = 0.5 * (re[k1] - re[k2]);
        const double di = 0.5 * (im[k1] + im[k2]);
        const double or_ = di;
        const double oi = -dr;
        const double wr = p->tw_re[k], wi = p->tw_im[k];
        X[k][0] = (float)(er + (or_ * wr - oi * wi));
        X[k][1] = (float)(ei + (or_ * wi + oi * wr));
    }
    X[0][1] = 0.f;
    X[m][1] = 0.f;
    (void)n;
}

static void c2r_one(const struct oracle_fft_plan_s* p, const fftwf_complex* X, float* x, double* re, double* im)
{
    const int m = p->m;
    if(m == 1)
    {
        x[0] = (float)((double)X[0][0] + (double)X[1][0]);
        x[1] = (float)((double)X[0][0] - (double)X[1][0]);
        return;
    }
    for(int k = 0; k < m; ++k)
    {
        double ar = (double)X[k][0], ai = (double)X[k][1];
        double br = (double)X[m - k][0], bi = -(double)X[m - k][1]; /* conj X[m-k] */
        if(k == 0)
        {
            ai = 0.0; /* imaginary parts of X[0] and X[N/2] are ignored */
            bi = 0.0;
        }
        const double er = ar + br, ei = ai + bi;       /* 2E */
        const double dr = ar - br, di = ai - bi;
        const double wr = p->tw_re[k], wi = -p->tw_im[k]; /* conj W^k */
        const double or_ = dr * wr - di * wi;          /* 2O */
        const double oi = dr * wi + di * wr;
        re[k] = er - oi;                               /* 2E + i 2O */
        im[k] = ei + or_;
    }
    cfft(p, re, im, +1);
    for(int i = 0; i < m; ++i)
    {
        x[2 * i] = (float)re[i];
        x[2 * i + 1] = (float)im[i];
    }
}

void fftwf_execute(const fftwf_plan p)
{
    if(!p)
        return;
    const int m = p->m;
    #pragma omp parallel
    {
        double* re = (double*)malloc(sizeof(double) * (size_t)(m > 0 ? m : 1));
        double* im = (double*)malloc(sizeof(double) * (size_t)(m > 0 ? m : 1));
        #pragma omp for schedule(static)
        for(int b = 0; b < p->howmany; ++b)
        {
            if(p->kind == 0)
                r2c_one(p, (const float*)p->in + (size_t)b * (size_t)p->idist,
                        (fftwf_complex*)p->out + (size_t)b * (size_t)p->odist, re, im);
            else
                c2r_one(p, (const fftwf_complex*)p->in + (size_t)b * (size_t)p->idist,
                        (float*)p->out + (size_t)b * (size_t)p->odist, re, im);
        }
        free(re);
        free(im);
    }
}
