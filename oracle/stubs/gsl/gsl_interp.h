/* Type-only stand-in for GSL (absent offline).  TEST INFRASTRUCTURE ONLY. */
#ifndef STUB_GSL_INTERP_H
#define STUB_GSL_INTERP_H
typedef struct gsl_interp gsl_interp;
typedef struct gsl_interp_accel gsl_interp_accel;
#endif
