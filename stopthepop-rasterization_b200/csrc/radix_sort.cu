// radix_sort.cu -- device radix sort used by the binning stage.
// Round-1 bring-up implementation: CUB's DeviceRadixSort from the CUDA toolkit (the same library
// call the reference makes, rasterizer_impl.cu:347-352).  The hand-written onesweep replacement
// lives behind the same two functions.
#include "radix_sort.cuh"
#include <cub/device/device_radix_sort.cuh>

namespace stp {

size_t radix_sort_temp_bytes(size_t n) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                    (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n);
    return bytes;
}

int radix_sort_kernel_launches(size_t, int) { return 0; }  // CUB = library code, not counted

cudaError_t radix_sort_pairs(void* temp, size_t temp_bytes, const uint64_t* keys_in, uint64_t* keys_out,
                             const uint32_t* vals_in, uint32_t* vals_out, size_t n, int end_bit, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    return cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, vals_in, vals_out, (int)n, 0, end_bit,
                                           stream);
}

}  // namespace stp
