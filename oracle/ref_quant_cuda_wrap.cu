// Compiles the REFERENCE's own CUDA kernel from where it lies (/root/reference/ant_quantization/quant/),
// unmodified and uncopied, as a torch extension named `ref_quant_cuda` under oracle/_ref/.
// TEST INFRASTRUCTURE ONLY: the GPU-side second oracle and the "kernel to beat" (tools/ref_gpu_bench.py).
//
// The reference does not build against torch >= 2.x as shipped: quant_kernel.cu:51 passes `x.type()`
// (DeprecatedTypeProperties) where AT_DISPATCH_FLOATING_TYPES now wants a ScalarType.  Every header the
// two reference files include is pulled in FIRST (include guards make their own #includes no-ops), then
// the token `type` is mapped to `scalar_type` for the reference's 61 + 28 lines only.
#include <torch/extension.h>
#include <cuda.h>
#include <cuda_runtime.h>
#include <thrust/device_vector.h>
#include <iostream>
#include <assert.h>
#include <stdio.h>

#define type scalar_type
#include ANTQ_REF_KERNEL_CU
#undef TORCH_EXTENSION_NAME
#define TORCH_EXTENSION_NAME ref_quant_cuda
#include ANTQ_REF_BINDING_CPP
#undef type
