/* Stand-in for GSL's integration header (absent offline).  TEST INFRASTRUCTURE ONLY.  The one
 * reference file built against it, timebinmgr.c, integrates only inside time_to_present
 * (timebinmgr.c:120-147, excursion-set sync points); the fixtures never enable those, so the
 * functions abort if they are ever reached. */
#ifndef STUB_GSL_INTEGRATION_H
#define STUB_GSL_INTEGRATION_H
#include <stddef.h>
#include <stdlib.h>
typedef struct { double (*function)(double x, void *params); void *params; } gsl_function;
typedef struct gsl_integration_workspace gsl_integration_workspace;
#define GSL_INTEG_GAUSS21 2
#define GSL_INTEG_GAUSS61 6
static inline gsl_integration_workspace *gsl_integration_workspace_alloc(size_t n) { abort(); return NULL; }
static inline void gsl_integration_workspace_free(gsl_integration_workspace *w) { abort(); }
static inline int gsl_integration_qag(const gsl_function *f, double a, double b, double epsabs, double epsrel, size_t limit,
                                      int key, gsl_integration_workspace *w, double *result, double *abserr) { abort(); return 0; }
#endif
