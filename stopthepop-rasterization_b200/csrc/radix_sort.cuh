// radix_sort.cuh -- stable LSD radix sort of (u64 key, u32 value) pairs on key bits [0,end_bit).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace stp {
size_t radix_sort_temp_bytes(size_t n);
int radix_sort_kernel_launches(size_t n, int end_bit);
cudaError_t radix_sort_pairs(void* temp, size_t temp_bytes, const uint64_t* keys_in, uint64_t* keys_out,
                             const uint32_t* vals_in, uint32_t* vals_out, size_t n, int end_bit, cudaStream_t stream);
}  // namespace stp
