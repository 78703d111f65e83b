// libpulser_diff_b200.so: the C ABI of include/pulser_diff_b200.h over the CUDA backend.
#include "cuda_backend.cuh"
#define PD_BACKEND pd::CudaBackend
#include "cabi_impl.hpp"
