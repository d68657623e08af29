/* stand-in, see ../cuda_runtime.h: the one CUB call steploop.cu makes, as a serial loop */
#ifndef EMUL_CUB_H
#define EMUL_CUB_H
#include <cuda_runtime.h>
namespace cub {
struct DeviceSelect {
    template <class In, class Fl, class Out>
    static cudaError_t Flagged(void *temp, size_t &bytes, const In *in, const Fl *flags, Out *out, int *num, int n, cudaStream_t)
    {
        if(!temp) { bytes = 1; return 0; }
        int k = 0;
        for(int i = 0; i < n; i++) if(flags[i]) out[k++] = in[i];
        *num = k;
        return 0;
    }
};
}
#endif
