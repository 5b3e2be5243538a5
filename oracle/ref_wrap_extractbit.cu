// extern "C" entry to the REFERENCE's own CUDA kernel (compiled from
// /root/reference/src/cuda/extractbit.cu in place; see oracle/Makefile).  GPU-side oracle for
// stage 2, used by tests only.
#include <stdint.h>
#include "cuda/extractbit.h"
extern "C" int nsb_ref_extract_bits(float* d_dest, const uint64_t* d_src, int batch, int channels,
                                    int channels_first) {
    if (channels_first)
        nshogi::engine::cuda::extractBits<true>(d_dest, d_src, batch, channels, 0);
    else
        nshogi::engine::cuda::extractBits<false>(d_dest, d_src, batch, channels, 0);
    return (int)cudaDeviceSynchronize();
}
