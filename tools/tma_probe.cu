// Standalone probe: 3-D TMA tile load of a float field into shared memory, descriptor passed (a) as __grid_constant__
// kernel parameter and (b) through a pointer to global memory.  Prints what arrives.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <vector>
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int B>
__global__ void probe(const __grid_constant__ CUtensorMap pmap, const CUtensorMap* gmap, int use_global, int c0, int c1, int c2, float* out) {
  extern __shared__ __align__(128) unsigned char sm[];
  float* brick = (float*)sm;
  uint64_t* bar = (uint64_t*)(sm + B * B * B * 4);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(B * B * B * 4) : "memory");
    const CUtensorMap* m = use_global ? gmap : &pmap;
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(s32(brick)), "l"((unsigned long long)m), "r"(c0), "r"(c1), "r"(c2), "r"(s32(bar)) : "memory");
  }
  uint32_t ok = 0, spins = 0;
  while (!ok && spins < (1u << 22)) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(bar)), "r"(0) : "memory");
    ++spins;
  }
  if (threadIdx.x == 0) out[0] = ok ? 1.f : -1.f, out[1] = (float)spins;
  for (int i = threadIdx.x; i < B * B * B; i += blockDim.x) out[2 + i] = ok ? brick[i] : -7.f;
}
template <int B>
int run(PFN_encodeTiled enc, float* d, int nx, int ny, int nz, int nzp, int c0, int c1, int c2, const std::vector<float>& h) {
  CUtensorMap map;
  cuuint64_t gdim[3] = {(cuuint64_t)nz, (cuuint64_t)ny, (cuuint64_t)nx};
  cuuint64_t gstr[2] = {(cuuint64_t)nzp * 4, (cuuint64_t)ny * nzp * 4};
  cuuint32_t box[3] = {B, B, B}, es[3] = {1, 1, 1};
  CUresult rc = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("B=%d encode rc=%d\n", B, (int)rc);
  if (rc) return 1;
  CUtensorMap* gm;
  cudaMalloc(&gm, sizeof(map));
  cudaMemcpy(gm, &map, sizeof(map), cudaMemcpyHostToDevice);
  float* out;
  cudaMalloc(&out, (2 + B * B * B) * 4);
  size_t smem = B * B * B * 4 + 64;
  cudaFuncSetAttribute(probe<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int ug = 0; ug < 2; ++ug) {
    cudaMemset(out, 0, (2 + B * B * B) * 4);
    probe<B><<<1, 128, smem>>>(map, gm, ug, c0, c1, c2, out);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> o(2 + B * B * B);
    cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int lx = 0; lx < B; ++lx) for (int ly = 0; ly < B; ++ly) for (int lz = 0; lz < B; ++lz) {
      int x = c2 + lx, y = c1 + ly, z = c0 + lz;
      float want = (x >= 0 && y >= 0 && z >= 0 && x < nx && y < ny && z < nz) ? h[((size_t)x * ny + y) * nzp + z] : 0.f;
      if (o[2 + (lx * B + ly) * B + lz] != want) ++bad;
    }
    printf("  B=%d desc=%s coords(z,y,x)=(%d,%d,%d): err=%s ok=%g spins=%g mismatches=%d\n", B, ug ? "global-ptr" : "grid_constant", c0, c1, c2,
           cudaGetErrorString(e), o[0], o[1], bad);
    if (e != cudaSuccess) return 2;
  }
  return 0;
}
int main(int argc, char** argv) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (!fn) { printf("no entry point %s\n", cudaGetErrorString(e)); return 1; }
  int B = argc > 1 ? atoi(argv[1]) : 8, c0 = argc > 2 ? atoi(argv[2]) : 4, c1 = argc > 3 ? atoi(argv[3]) : 5, c2 = argc > 4 ? atoi(argv[4]) : 6;
  int nx = 64, ny = 64, nz = 62, nzp = 64;
  std::vector<float> h((size_t)nx * ny * nzp);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 9973) * 0.5f;
  float* d;
  cudaMalloc(&d, h.size() * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  PFN_encodeTiled enc = (PFN_encodeTiled)fn;
  if (B == 8) return run<8>(enc, d, nx, ny, nz, nzp, c0, c1, c2, h);
  if (B == 16) return run<16>(enc, d, nx, ny, nz, nzp, c0, c1, c2, h);
  if (B == 24) return run<24>(enc, d, nx, ny, nz, nzp, c0, c1, c2, h);
  if (B == 12) return run<12>(enc, d, nx, ny, nz, nzp, c0, c1, c2, h);
  if (B == 20) return run<20>(enc, d, nx, ny, nz, nzp, c0, c1, c2, h);
  return 0;
}
