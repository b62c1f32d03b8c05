// pm_tmap.cu — host-side TMA tensor-map encoding through the driver entry point.
#include "pm_common.cuh"

namespace pm {

static PFN_encodeTiled g_encode = nullptr;

int pm_get_encode_fn(PFN_encodeTiled* out) {
  if (g_encode == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) return PM_ERR_NO_DRIVER;
    g_encode = reinterpret_cast<PFN_encodeTiled>(fn);
  }
  *out = g_encode;
  return PM_OK;
}

static CUtensorMapDataType dtype_of(int elt_bytes) {
  return elt_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
}

int pm_make_tmap_2d(CUtensorMap* map, const void* base, int elt_bytes, uint64_t rows, uint64_t cols,
                    uint64_t ld, uint32_t box_rows, uint32_t box_cols) {
  PFN_encodeTiled enc;
  int rc = pm_get_encode_fn(&enc);
  if (rc != PM_OK) return rc;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || ((ld * elt_bytes) & 15) != 0) return PM_ERR_INVALID;
  if (box_cols * elt_bytes != 128 || box_rows > 256) return PM_ERR_INVALID;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * elt_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, dtype_of(elt_bytes), 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? PM_OK : PM_ERR_TENSORMAP;
}

// 64-byte rows (box_cols * elt_bytes == 64) with the 64B swizzle: K = 32 bf16 operand tiles (pm_vq.cu)
int pm_make_tmap_2d_sw64(CUtensorMap* map, const void* base, int elt_bytes, uint64_t rows, uint64_t cols,
                         uint64_t ld, uint32_t box_rows, uint32_t box_cols) {
  PFN_encodeTiled enc;
  int rc = pm_get_encode_fn(&enc);
  if (rc != PM_OK) return rc;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || ((ld * elt_bytes) & 15) != 0) return PM_ERR_INVALID;
  if (box_cols * elt_bytes != 64 || box_rows > 256) return PM_ERR_INVALID;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * elt_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, dtype_of(elt_bytes), 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? PM_OK : PM_ERR_TENSORMAP;
}

int pm_make_tmap_3d(CUtensorMap* map, const void* base, int elt_bytes, uint64_t batch, uint64_t rows,
                    uint64_t cols, uint64_t ld_row, uint64_t ld_batch, uint32_t box_rows,
                    uint32_t box_cols) {
  return pm_make_tmap_3d_box(map, base, elt_bytes, batch, rows, cols, ld_row, ld_batch, 1, box_rows, box_cols);
}

int pm_make_tmap_3d_box(CUtensorMap* map, const void* base, int elt_bytes, uint64_t batch, uint64_t rows,
                        uint64_t cols, uint64_t ld_row, uint64_t ld_batch, uint32_t box_batch, uint32_t box_rows,
                        uint32_t box_cols) {
  PFN_encodeTiled enc;
  int rc = pm_get_encode_fn(&enc);
  if (rc != PM_OK) return rc;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || ((ld_row * elt_bytes) & 15) != 0 ||
      ((ld_batch * elt_bytes) & 15) != 0)
    return PM_ERR_INVALID;
  if (box_cols * elt_bytes != 128 || box_rows > 256) return PM_ERR_INVALID;
  cuuint64_t gdim[3] = {cols, rows, batch};
  cuuint64_t gstride[2] = {ld_row * elt_bytes, ld_batch * elt_bytes};
  cuuint32_t box[3] = {box_cols, box_rows, box_batch};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, dtype_of(elt_bytes), 3, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? PM_OK : PM_ERR_TENSORMAP;
}

}  // namespace pm
