// Host helpers of the tcgen05 path: TMA tensor-map encoding via the driver entry point, with a
// small cache (a training step re-encodes the same ~40 views every call).
#include "tc_ptx.cuh"

#include <cstring>
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

namespace {
// A tensor map is a pure function of (address, strides, extents, box): it can be cached by value.
// (Unified addressing makes the pointer unique across devices; the map never depends on the data.)
struct MapKey {
  const void* ptr;
  int64_t sb, sl, sh;
  int32_t B, len, H, box;
};
struct MapSlot {
  MapKey key;
  bool valid;
  CUtensorMap map;
};
constexpr int kSlots = 256;   // direct-mapped
MapSlot g_slots[kSlots];
std::mutex g_slots_mu;

inline uint32_t key_hash(const MapKey& k) {
  uint64_t h = reinterpret_cast<uint64_t>(k.ptr) * 0x9E3779B97F4A7C15ull;
  h ^= (uint64_t)k.sb * 0xC2B2AE3D27D4EB4Full + (uint64_t)k.sl * 0x165667B19E3779F9ull + (uint64_t)k.sh;
  h ^= ((uint64_t)(uint32_t)k.len << 32) ^ ((uint64_t)(uint32_t)k.B << 20) ^ ((uint64_t)(uint32_t)k.H << 8) ^ (uint32_t)k.box;
  h ^= h >> 29;
  h *= 0xBF58476D1CE4E5B9ull;
  h ^= h >> 32;
  return (uint32_t)h;
}
}  // namespace

int make_qkv_tensor_map(CUtensorMap* out, const void* ptr, int64_t sb, int64_t sl, int64_t sh, int B,
                        int len, int H, int box_rows) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return (int)CUDA_ERROR_NOT_SUPPORTED;
  // A zero (broadcast) stride cannot be encoded; over an extent of 1 the stride is never used, so any
  // legal value does.  Over a larger extent the view is not addressable by TMA: refuse (the callers
  // check tc_t4_ok() first and route such views to the SIMT kernels).
  if ((sb == 0 && B > 1) || (sl == 0 && len > 1) || (sh == 0 && H > 1)) return (int)CUDA_ERROR_INVALID_VALUE;
  MapKey key;
  std::memset(&key, 0, sizeof(key));
  key.ptr = ptr; key.sb = sb; key.sl = sl; key.sh = sh;
  key.B = B; key.len = len; key.H = H; key.box = box_rows;
  MapSlot& slot = g_slots[key_hash(key) % kSlots];
  {
    std::lock_guard<std::mutex> lk(g_slots_mu);
    if (slot.valid && std::memcmp(&slot.key, &key, sizeof(key)) == 0) {
      *out = slot.map;
      return 0;
    }
  }
  // dims fastest-first: d, len, H, B
  cuuint64_t dims[4] = {64, (cuuint64_t)len, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)sl * 2, (cuuint64_t)sh * 2, (cuuint64_t)sb * 2};
  for (int i = 0; i < 3; ++i)
    if (strides[i] == 0) strides[i] = 16;   // extent 1 (checked above): never dereferenced
  cuuint32_t box[4] = {64, (cuuint32_t)box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r == CUDA_SUCCESS) {
    std::lock_guard<std::mutex> lk(g_slots_mu);
    slot.key = key;
    slot.map = *out;
    slot.valid = true;
  }
  return (int)r;
}

}  // namespace mlt
