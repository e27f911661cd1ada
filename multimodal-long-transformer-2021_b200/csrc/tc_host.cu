// Host helpers of the tcgen05 path: TMA tensor-map encoding via the driver entry point.
#include "tc_ptx.cuh"

#include <mutex>

namespace mlt {

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

int make_qkv_tensor_map(CUtensorMap* out, const void* ptr, int64_t sb, int64_t sl, int64_t sh, int B,
                        int len, int H, int box_rows) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return (int)CUDA_ERROR_NOT_SUPPORTED;
  // dims fastest-first: d, len, H, B.  A zero stride (broadcast) is not representable: use 16 B.
  cuuint64_t dims[4] = {64, (cuuint64_t)len, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)sl * 2, (cuuint64_t)sh * 2, (cuuint64_t)sb * 2};
  for (int i = 0; i < 3; ++i)
    if (strides[i] == 0) strides[i] = 16;
  cuuint32_t box[4] = {64, (cuuint32_t)box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return (int)r;
}

}  // namespace mlt
