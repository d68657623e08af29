/* stand-in, see cuda_runtime.h in this directory: engine.h only names the handle type */
#ifndef EMUL_CUFFT_H
#define EMUL_CUFFT_H
typedef int cufftHandle;
#endif
