/* TEST INFRASTRUCTURE ONLY — CPU restatement of fastACE's Economy::time_step hot path.
 * See fastace_oracle.c.  Nothing in the product path may include or link this. */
#ifndef FASTACE_ORACLE_H
#define FASTACE_ORACLE_H
#include "../include/fastace_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* One Economy::time_step() for every economy, in place on HOST arrays.
 * Economies e with first_econ <= e < first_econ + count are stepped (so callers can
 * spread economies over threads); returns 0, or -1 on invalid dims. */
/* function families used by the next step calls (process-wide; default CES / CES) */
void fastace_oracle_set_function_kinds(int util_kind, int prod_kind);
/* VecToScalar::f of the given family; share/theta read with `stride` (theta may be NULL) */
double fastace_oracle_function_f(int kind, double tfp, const double* share, const double* theta, double rho,
                                 const double* x, int n, int stride);

int fastace_oracle_step(const fastace_dims_t* dims, fastace_state_t* state,
                        const fastace_actions_t* actions, const fastace_step_out_t* out,
                        uint32_t flags, uint32_t time_before, int first_econ, int count);

/* same, spreading economies over `nthreads` pthreads (CPU-baseline timing) */
int fastace_oracle_step_mt(const fastace_dims_t* dims, fastace_state_t* state,
                           const fastace_actions_t* actions, const fastace_step_out_t* out,
                           uint32_t flags, uint32_t time_before, int nthreads);

/* function plugins */
double fastace_oracle_ces_f(double tfp, const double* share_norm, double rho, const double* x, int n, int stride);
void   fastace_oracle_ces_params(const double* share_raw, double elasticity, int n, double* share_out, double* rho_out);
double fastace_oracle_cobb_douglas_f(double tfp, const double* elast, const double* x, int n);
/* (int) conversion as x86-64 g++ performs it (cvttsd2si) */
int32_t fastace_oracle_double_to_int(double x);

#ifdef __cplusplus
}
#endif
#endif
