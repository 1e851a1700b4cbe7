// Host-side TMA tensor-map construction.  cuTensorMapEncodeTiled is looked up through
// cudaGetDriverEntryPoint so libomnissm.so links only against the (static) CUDA runtime.
#include <mutex>

#include "umma.cuh"

namespace omni {

PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

int make_tmap_16bit(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, bool is_bf16, int swizzle) {
  return make_tmap(out, base, rank, dims, strides_bytes, box, is_bf16 ? OMNI_BF16 : OMNI_F16, swizzle);
}

int make_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
              const uint32_t* box, int dtype, int swizzle) {
  PFN_encodeTiled enc = get_encode_tiled();
  OMNI_CHECK(enc != nullptr, OMNI_CUDA_ERROR, "cuTensorMapEncodeTiled is not available from this driver");
  OMNI_CHECK(aligned16(base), OMNI_BAD_STRIDE, "TMA: base pointer must be 16-byte aligned");
  // The encode is a DRIVER call and needs a current context on the calling thread.  PyTorch's autograd worker threads only
  // record their device (no context is bound until the first runtime call), so bind the primary context of the thread's
  // current device once per (thread, device) - otherwise a backward that starts with a tensor-map encode fails with 201.
  {
    static thread_local int bound_dev = -1;
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && dev != bound_dev) {
      cudaSetDevice(dev);
      bound_dev = dev;
    }
  }
  cuuint64_t gdim[5], gstr[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) {
      OMNI_CHECK(strides_bytes[i - 1] % 16 == 0, OMNI_BAD_STRIDE, "TMA: stride %d (%llu bytes) is not a multiple of 16", i,
                 (unsigned long long)strides_bytes[i - 1]);
      gstr[i - 1] = strides_bytes[i - 1];
    }
  }
  const CUtensorMapDataType ty = dtype == OMNI_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                 : dtype == OMNI_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = enc(out, ty, (cuuint32_t)rank,
                   const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (swizzle == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  OMNI_CHECK(r == CUDA_SUCCESS, OMNI_CUDA_ERROR, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return OMNI_OK;
}

}  // namespace omni
