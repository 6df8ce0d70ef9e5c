// tma_gather_test.cu — probe of cp.async.bulk.tensor tile::gather4 (sm_100a): which tensor-map box shape it wants and
// how the four gathered rows land in a SWIZZLE_64B shared-memory tile.  Developer tool, not part of the library.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d: %s\n", #x, __LINE__, cudaGetErrorString(e_)); return 1; } } while (0)

typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k_probe(const __grid_constant__ CUtensorMap map, int col, int r0, int r1, int r2, int r3, uint32_t bytes,
                        uint16_t *out, int *status) {
  __shared__ __align__(1024) uint16_t tile[2048];     // 4 KB
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2048; ++i) tile[i] = 0xFFFF;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(smem_u32(tile)), "l"(&map), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0, spins = 0;
    while (!ok && spins < (1u << 22)) {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
      ++spins;
    }
    *status = ok ? 1 : -1;
    for (int i = 0; i < 2048; ++i) out[i] = tile[i];
  }
}

int main(int argc, char **argv) {
  const int rows = 64, cols = 512;
  std::vector<uint16_t> h((size_t)rows * cols);
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < cols; ++c) h[(size_t)r * cols + c] = (uint16_t)(r * 512 + c);    // element id
  uint16_t *d, *dout; int *dst;
  CK(cudaMalloc(&d, h.size() * 2)); CK(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&dout, 4096)); CK(cudaMalloc(&dst, 4));
  void *fn = nullptr; cudaDriverEntryPointQueryResult qr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
  EncodeTiled enc = (EncodeTiled)fn;
  const int box_rows_opts[2] = {1, 4};
  const CUtensorMapSwizzle sw_opts[2] = {CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_SWIZZLE_64B};
  for (int bi = 0; bi < 2; ++bi)
    for (int si = 0; si < 2; ++si) {
      CUtensorMap map;
      cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
      cuuint64_t gstr[1] = {(cuuint64_t)cols * 2};
      cuuint32_t box[2] = {32, (cuuint32_t)box_rows_opts[bi]};
      cuuint32_t estr[2] = {1, 1};
      CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw_opts[si],
                       CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      printf("== box {32,%d} swizzle %s: encode -> %d\n", box_rows_opts[bi], si ? "64B" : "none", (int)r);
      if (r != CUDA_SUCCESS) continue;
      CK(cudaMemset(dout, 0, 4096)); CK(cudaMemset(dst, 0, 4));
      k_probe<<<1, 32>>>(map, 32, 5, 17, 3, 40, 4 * 64, dout, dst);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("   kernel error: %s\n", cudaGetErrorString(e)); return 1; }
      int st; std::vector<uint16_t> o(2048);
      CK(cudaMemcpy(&st, dst, 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(o.data(), dout, 4096, cudaMemcpyDeviceToHost));
      printf("   status %d; 16-byte pieces of the first 512 B (row:col of first element, -- = untouched):\n", st);
      for (int p = 0; p < 32; ++p) {
        uint16_t v = o[p * 8];
        if (v == 0xFFFF) printf(" --"); else printf(" %d:%d", v / 512, v % 512);
        if (p % 4 == 3) printf(" |");
      }
      printf("\n");
    }
  return 0;
}
