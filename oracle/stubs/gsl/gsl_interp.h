/* Stand-in for GSL (absent offline).  TEST INFRASTRUCTURE ONLY.  Only gravpm.c's massive-neutrino
 * branch (gravpm.c:304-326,415-437) touches these; the fixtures never enable it, so the
 * functions abort if they are ever reached. */
#ifndef STUB_GSL_INTERP_H
#define STUB_GSL_INTERP_H
#include <stddef.h>
#include <stdlib.h>
typedef struct gsl_interp gsl_interp;
typedef struct gsl_interp_accel gsl_interp_accel;
typedef struct gsl_interp_type gsl_interp_type;
static const gsl_interp_type *const gsl_interp_linear = NULL;
static inline gsl_interp *gsl_interp_alloc(const gsl_interp_type *t, size_t n) { abort(); return NULL; }
static inline gsl_interp_accel *gsl_interp_accel_alloc(void) { abort(); return NULL; }
static inline int gsl_interp_init(gsl_interp *s, const double *x, const double *y, size_t n) { abort(); return 0; }
static inline double gsl_interp_eval(const gsl_interp *s, const double *x, const double *y, double v, gsl_interp_accel *a) { abort(); return 0; }
static inline void gsl_interp_free(gsl_interp *s) { abort(); }
static inline void gsl_interp_accel_free(gsl_interp_accel *a) { abort(); }
#endif
