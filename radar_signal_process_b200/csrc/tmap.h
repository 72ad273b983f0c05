// tmap.h -- cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda), shared by the kernels that fetch
// their tiles with tensor-map TMA loads (pcw_kernel.cu, onepass_kernel.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace rb {

inline cudaError_t tensor_map_encode_tiled(CUtensorMap* map, CUtensorMapDataType type, cuuint32_t rank, const void* base, const cuuint64_t* dims,
                                           const cuuint64_t* strides, const cuuint32_t* box, CUtensorMapSwizzle swizzle) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
        if (e != cudaSuccess) return e;
        if (qres != cudaDriverEntryPointSuccess || !ptr) return cudaErrorNotSupported;
        fn = reinterpret_cast<EncodeFn>(ptr);
    }
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const CUresult r = fn(map, type, rank, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

}  // namespace rb
