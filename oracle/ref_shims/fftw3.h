/* Stand-in for the six FFTW r2r kinds the reference's FftwPatchSolver plans
 * (FFTW 3 is a third-party dependency that is NOT vendored in the reference tree and
 * is not installed here).  Transforms are the unnormalised definitions of the FFTW
 * manual section "1d Real-even DFTs (DCTs)" / "1d Real-odd DFTs (DSTs)", evaluated as naive
 * separable O(n^2) sums in shims.cpp.  Parity use only - not representative of FFTW speed.
 * TEST INFRASTRUCTURE (oracle/_ref build only). */
#ifndef ORACLE_SHIM_FFTW3_H
#define ORACLE_SHIM_FFTW3_H
enum fftw_r2r_kind { FFTW_R2HC, FFTW_HC2R, FFTW_DHT, FFTW_REDFT00, FFTW_REDFT01, FFTW_REDFT10,
                     FFTW_REDFT11, FFTW_RODFT00, FFTW_RODFT01, FFTW_RODFT10, FFTW_RODFT11 };
#define FFTW_MEASURE 0u
#define FFTW_DESTROY_INPUT 1u
#define FFTW_ESTIMATE 64u
struct fftw_plan_s;
typedef fftw_plan_s *fftw_plan;
fftw_plan fftw_plan_r2r(int rank, const int *n, double *in, double *out, const fftw_r2r_kind *kind, unsigned flags);
void      fftw_execute(const fftw_plan p);
void      fftw_destroy_plan(fftw_plan p);
#endif
