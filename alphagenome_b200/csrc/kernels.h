// Internal host-side declarations shared by the .cu files of libagb200.so (not part of the C ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/agb200.h"

namespace agb {

// thread-local error text; always returns non-zero so callers can `return set_error(...)`
int set_error(const char* fmt, ...);
// cudaGetLastError() after a launch; bumps the launch counter; formats watchdog info if present
int check_launch(const char* what);
// host-mapped pinned word block [4] the device watchdog writes before trapping (device pointer)
unsigned int* device_error_word(int device);
int sm_count(int device);
// tiled, 128B-swizzled bf16 tensor map (rank 2..4); dims[0] is the contiguous dimension
int make_tensor_map_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box);

int launch_gemm(const AgGemmArgs& a, int device, cudaStream_t stream);

}  // namespace agb
