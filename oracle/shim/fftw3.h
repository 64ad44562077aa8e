/*
 * oracle/shim/fftw3.h -- TEST INFRASTRUCTURE, not product code.
 *
 * Minimal stand-in for <fftw3.h> so that the reference's OpenMP backend
 * (/root/reference/src/openmp/filtering.cpp, which calls FFTW3f at lines
 * 41,48,136,149,153,199,201-204,208,214) compiles unmodified in an image that
 * has no FFTW.  Only the symbols the reference uses are declared; they are
 * implemented by oracle/fft_shim.cpp (double-precision radix-2 DFT, results
 * rounded to float).  FFTW itself is an un-vendored, un-pinned third-party
 * dependency of the reference (CMakeLists.txt:50), see DESIGN.md "oracle".
 */
#ifndef PARIS_B200_ORACLE_FFTW3_SHIM_H_
#define PARIS_B200_ORACLE_FFTW3_SHIM_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef float fftwf_complex[2];
typedef struct oracle_fft_plan_s* fftwf_plan;

#define FFTW_MEASURE (0U)
#define FFTW_DESTROY_INPUT (1U << 0)
#define FFTW_PRESERVE_INPUT (1U << 4)
#define FFTW_ESTIMATE (1U << 6)

void* fftwf_malloc(size_t n);
void fftwf_free(void* p);

fftwf_plan fftwf_plan_dft_r2c_1d(int n, float* in, fftwf_complex* out, unsigned flags);

fftwf_plan fftwf_plan_many_dft_r2c(int rank, const int* n, int howmany,
                                   float* in, const int* inembed, int istride, int idist,
                                   fftwf_complex* out, const int* onembed, int ostride, int odist,
                                   unsigned flags);

fftwf_plan fftwf_plan_many_dft_c2r(int rank, const int* n, int howmany,
                                   fftwf_complex* in, const int* inembed, int istride, int idist,
                                   float* out, const int* onembed, int ostride, int odist,
                                   unsigned flags);

void fftwf_execute(const fftwf_plan p);
void fftwf_destroy_plan(fftwf_plan p);

#ifdef __cplusplus
}
#endif

#endif
