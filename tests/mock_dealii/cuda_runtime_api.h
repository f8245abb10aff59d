// mock of the two CUDA runtime calls the shim uses (the real header is found first when the CUDA toolkit is on the include path)
#pragma once
#include <cstddef>
#include <cstdlib>
enum cudaError_t { cudaSuccess = 0 };
inline cudaError_t cudaMallocHost(void **p, std::size_t n) { *p = std::malloc(n); return cudaSuccess; }
inline cudaError_t cudaFreeHost(void *p) { std::free(p); return cudaSuccess; }
