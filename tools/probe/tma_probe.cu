// Bisects the TMA box load used by k_wt53_inv_level_tma.  Usage: tma_probe <variant>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
struct alignas(64) TmaDesc { unsigned char bytes[128]; };
constexpr int BW = 36, BH = 34;
template <int RANK>
__global__ void probe(const __grid_constant__ TmaDesc tmap, int* out, int cx, int cy, int cz, int use_proxy_fence) {
  struct alignas(128) Box { int v[BH][BW]; };
  __shared__ Box box;
  __shared__ __align__(8) unsigned long long s_bar;
  const unsigned bar = (unsigned)__cvta_generic_to_shared(&s_bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (use_proxy_fence) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((unsigned)(BW * BH * 4)) : "memory");
    const unsigned dst = (unsigned)__cvta_generic_to_shared(&box.v[0][0]);
    if (RANK == 3)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                   ::"r"(dst), "l"(&tmap), "r"(cx), "r"(cy), "r"(cz), "r"(bar) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(dst), "l"(&tmap), "r"(cx), "r"(cy), "r"(bar) : "memory");
  }
  unsigned done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0, 0x989680;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar) : "memory");
  for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = box.v[i / BW][i % BW];
}
int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  const unsigned rows = 64, cols = 96, nimg = 2;
  std::vector<int> h(rows * cols * nimg);
  for (size_t i = 0; i < h.size(); i++) h[i] = (int)i;
  int *d, *o;
  cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, BW * BH * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeFn fn = (EncodeFn)p;
  TmaDesc m;
  const int rank = (variant & 1) ? 2 : 3;
  const cuuint64_t dims[3] = {cols, rows * (rank == 2 ? nimg : 1), nimg};
  const cuuint64_t strides[2] = {(cuuint64_t)cols * 4, (cuuint64_t)rows * cols * 4};
  const cuuint32_t box[3] = {BW, BH, 1}, es[3] = {1, 1, 1};
  CUresult r = fn((CUtensorMap*)&m, (variant == 21 || variant == 22) ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : ((variant == 23 || variant == 24) ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_INT32), rank, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                  (variant & 4) ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("variant %d rank %d encode rc=%d\n", variant, rank, (int)r);
  int cx = (variant & 2) ? -1 : 4, cy = (variant & 2) ? -1 : 2;
  if (variant == 16) { cx = (int)cols - 10; cy = 2; }
  if (variant == 17) { cx = 4; cy = (int)rows - 5; }
  if (variant == 18) { cx = -1; cy = 2; }
  if (variant == 19) { cx = 4; cy = -1; }
  if (variant == 20) { cx = -4; cy = 2; }
  if (variant == 22 || variant == 24) { cx = -1; cy = -1; }
  if (rank == 3) probe<3><<<1, 128>>>(m, o, cx, cy, 1, (variant & 8) ? 0 : 1);
  else probe<2><<<1, 128>>>(m, o, cx, cy, 0, (variant & 8) ? 0 : 1);
  cudaError_t e = cudaDeviceSynchronize();
  printf("  kernel: %s\n", cudaGetErrorString(e));
  if (e == cudaSuccess) {
    std::vector<int> ho(BW * BH);
    cudaMemcpy(ho.data(), o, BW * BH * 4, cudaMemcpyDeviceToHost);
    printf("  box[0][0..3]=%d %d %d %d box[1][1]=%d (expect plane[%d][%d])\n", ho[0], ho[1], ho[2], ho[3], ho[BW + 1], cy + 1, cx + 1);
  }
  return 0;
}
