// tma_probe.cu -- standalone bisect of the TMA u8 tile load used by l0_fused_kernel.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int BW, int BH>
__global__ void probe(const __grid_constant__ CUtensorMap map, int cx, int cy, unsigned char* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem + ((BW * BH + 127) / 128) * 128);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(BW * BH) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(smem)),
        "l"(reinterpret_cast<unsigned long long>(&map)), "r"(cx), "r"(cy), "r"(smem_u32(bar))
        : "memory");
  }
  asm volatile(
      "{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(
          smem_u32(bar)),
      "r"(0)
      : "memory");
  for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = smem[i];
}

typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int BW, int BH>
int run(Enc enc, unsigned char* dimg, const std::vector<unsigned char>& img, int W, int H, int cx, int cy) {
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H};
  cuuint64_t strides[1] = {(cuuint64_t)W};
  cuuint32_t box[2] = {BW, BH};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, dimg, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("box %dx%d encode rc=%d\n", BW, BH, (int)r);
  if (r != CUDA_SUCCESS) return 1;
  unsigned char* dout;
  cudaMalloc(&dout, BW * BH);
  const int smem = ((BW * BH + 127) / 128) * 128 + 16;
  cudaFuncSetAttribute(probe<BW, BH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<BW, BH><<<1, 256, smem>>>(m, cx, cy, dout);
  cudaError_t e = cudaDeviceSynchronize();
  printf("  launch: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 2;
  std::vector<unsigned char> out(BW * BH);
  cudaMemcpy(out.data(), dout, BW * BH, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int y = 0; y < BH; ++y)
    for (int x = 0; x < BW; ++x) {
      int gx = cx + x, gy = cy + y;
      unsigned char want = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? img[gy * W + gx] : 0;
      if (out[y * BW + x] != want) ++bad;
    }
  printf("  mismatches: %d\n", bad);
  cudaFree(dout);
  return bad != 0;
}

int main(int argc, char** argv) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  printf("entry point: %s q=%d p=%p\n", cudaGetErrorString(e), (int)q, p);
  Enc enc = (Enc)p;
  const int W = 320, H = 240;
  std::vector<unsigned char> img(W * H);
  for (int i = 0; i < W * H; ++i) img[i] = (unsigned char)((i * 7 + i / W) & 255);
  unsigned char* dimg;
  cudaMalloc(&dimg, W * H);
  cudaMemcpy(dimg, img.data(), W * H, cudaMemcpyHostToDevice);
  int rc = 0;
  int cx = argc > 1 ? atoi(argv[1]) : 0, cy = argc > 2 ? atoi(argv[2]) : 0;
  printf("coords (%d,%d): ", cx, cy);
  rc |= run<96, 74>(enc, dimg, img, W, H, cx, cy);
  printf("probe rc=%d\n", rc);
  return rc;
}
