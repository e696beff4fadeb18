/*
 * fftw3.h -- SHIM, test infrastructure only (see oracle/README.md).
 *
 * The reference's hot-path translation unit (src/datastore.cxx) includes
 * <fftw3.h> (src/datastore.h:28) and libfftw3f is not installed in this image.
 * This header declares exactly the seven FFTW entry points that file uses
 * (src/datastore.cxx:30,31,32-33,43,44,45,82) so that the UNMODIFIED reference
 * source can be compiled from /root/reference into oracle/_ref/.  The
 * definitions live in shim_fft.cpp: a stand-in single-precision Stockham FFT,
 * NOT FFTW.  Everything else (loop, carry-over, truncation, bin order, double
 * accumulation, queue protocol) is then the reference's own code.
 */
#ifndef RPF_ORACLE_FFTW3_SHIM_H
#define RPF_ORACLE_FFTW3_SHIM_H

#ifdef __cplusplus
extern "C" {
#endif

typedef float fftwf_complex[2];
typedef struct rpf_shim_plan_s *fftwf_plan;

#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)

fftwf_complex *fftwf_alloc_complex(unsigned long n);
void fftwf_free(void *p);
fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex *in, fftwf_complex *out, int sign, unsigned flags);
void fftwf_execute(const fftwf_plan plan);
void fftwf_destroy_plan(fftwf_plan plan);

#ifdef __cplusplus
}
#endif
#endif
