// pm_slab.h -- state of the x-slab PM of one rank (pm_slab.cu), shared with the multi-GPU driver (sharded.cu).
#pragma once
#include "engine.h"

namespace b200 {

struct SlabPM {
    double Box = 0, Asmth = 0, G = 0;
    int N = 0, Nz = 0, rank = 0, nranks = 1, halo = 0;
    int x0 = 0, nx = 0, y0 = 0, ny = 0;
    DevBuf<double> real, cplx, cplxT, ktab;
    DevBuf<uint8_t> work;       // shared cuFFT work area
    DevBuf<int> err;            // [0] particles outside slab + halo (deposit / readout)
    cufftHandle p2f = 0, p2i = 0, p1 = 0;
    bool plans = false;
};

int pmslab_set_stream(Engine *E, cudaStream_t st);

} // namespace b200
