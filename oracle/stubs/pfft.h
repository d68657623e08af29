/* Type-only stand-in for PFFT (absent offline; the tree/SPH files only see it
 * through petapm.h's struct definitions).  TEST INFRASTRUCTURE ONLY. */
#ifndef STUB_PFFT_H
#define STUB_PFFT_H
#include <stddef.h>
typedef double pfft_complex[2];
typedef void *pfft_plan;
#endif
