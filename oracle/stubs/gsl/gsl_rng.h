/* Stand-in for gsl_rng (absent offline): utils/system.c only uses it to fill
 * the random-number table, which the force path never touches.  TEST ONLY. */
#ifndef STUB_GSL_RNG_H
#define STUB_GSL_RNG_H
#include <stdlib.h>
typedef struct { unsigned long long s; } gsl_rng;
typedef int gsl_rng_type;
static const gsl_rng_type *gsl_rng_ranlxd2 = 0;
static inline gsl_rng *gsl_rng_alloc(const gsl_rng_type *t) { return (gsl_rng *) calloc(1, sizeof(gsl_rng)); }
static inline void gsl_rng_set(gsl_rng *r, unsigned long seed) { r->s = seed * 2862933555777941757ULL + 3037000493ULL; }
static inline double gsl_rng_uniform(gsl_rng *r) { r->s = r->s * 6364136223846793005ULL + 1442695040888963407ULL; return (double) (r->s >> 11) / 9007199254740992.0; }
static inline void gsl_rng_free(gsl_rng *r) { free(r); }
#endif
