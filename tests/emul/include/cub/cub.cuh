/* stand-in, see ../cuda_runtime.h: the one CUB call steploop.cu makes, as a serial loop */
#ifndef EMUL_CUB_H
#define EMUL_CUB_H
#include <cuda_runtime.h>
#include <algorithm>
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
struct DeviceRadixSort {
    /* stable sort of (key, value) pairs by key */
    template <class K, class V>
    static cudaError_t SortPairs(void *temp, size_t &bytes, const K *kin, K *kout, const V *vin, V *vout, int n, int, int, cudaStream_t)
    {
        if(!temp) { bytes = 1; return 0; }
        int *order = (int *) malloc(sizeof(int) * (n > 0 ? n : 1));
        for(int i = 0; i < n; i++) order[i] = i;
        std::stable_sort(order, order + n, [&](int a, int b) { return kin[a] < kin[b]; });
        for(int i = 0; i < n; i++) { kout[i] = kin[order[i]]; vout[i] = vin[order[i]]; }
        free(order);
        return 0;
    }
};
}
#endif
