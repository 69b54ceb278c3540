// Probe (hardware question, not product code): may a tcgen05 shared-memory descriptor start at an arbitrary ROW of a
// TMA-written swizzled tile?  The halo form of the 3x3 convolution loads 170+ consecutive board positions once and
// feeds the nine taps as row-shifted views of the same tile (shift = dy*(W+1)+dx rows), so the operand start address is
// 128-byte (SWIZZLE_128B, fp16) or 64-byte (SWIZZLE_64B, e4m3) aligned but not aligned to the swizzle pattern repeat.
// For every shift 0..47 and both settings of the descriptor's base-offset field (0, or (addr >> 7) & 7 as the PTX ISA's
// matrix-descriptor table words it) the probe compares D = A[shift : shift+128] x B^T with the host result.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I agogo_b200/csrc tools/probe_rowshift.cu -o tools/probe_rowshift -lcudart -ldl
#include <cuda_fp8.h>

#include <cstdio>
#include <vector>

#include "tc_common.cuh"

namespace {

constexpr int HALO = 176, N = 64, KB = 64;

__device__ __forceinline__ void umma_f8_1(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}

// fp8 = 0: fp16 tiles, 128-byte rows, SWIZZLE_128B; fp8 = 1: e4m3 tiles, 64-byte rows, SWIZZLE_64B
__global__ void __launch_bounds__(128, 1)
k_probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int fp8, int row0, int shift, int bo_mode,
        float* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t row_bytes = fp8 ? 64 : 128;
  const uint32_t sa = base, sb = base + 32768, bar = base + 49152, mbar2 = bar + 8;
  uint32_t* tmem_slot = (uint32_t*)(smem_raw + (base - smem_u32(smem_raw)) + 49152 + 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1); mbar_init(mbar2, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, HALO * row_bytes + N * row_bytes);
    tma_load_2d(sa, &tmA, bar, 0, row0);
    tma_load_2d(sb, &tmB, bar, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t a_start = sa + shift * row_bytes;
    uint64_t dA = fp8 ? make_desc_sw<32>(a_start) : make_desc_sw<64>(a_start);
    const uint64_t dB = fp8 ? make_desc_sw<32>(sb) : make_desc_sw<64>(sb);
    if (bo_mode == 1) dA |= (uint64_t)((a_start >> 7) & 7) << 49;
    const uint32_t idesc = make_idesc(128, N);
    if (fp8) {
      for (int ks = 0; ks < 2; ks++) umma_f8_1(tmem, dA + (uint64_t)(ks * 2), dB + (uint64_t)(ks * 2), idesc, ks ? 1u : 0u);
    } else {
      for (int ks = 0; ks < 4; ks++) umma_f16(tmem, dA + (uint64_t)(ks * 2), dB + (uint64_t)(ks * 2), idesc, ks ? 1u : 0u);
    }
    umma_commit(mbar2);
  }
  __syncthreads();
  mbar_wait(mbar2, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t r[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int i = 0; i < 32; i++) out[(size_t)(warp * 32 + lane) * N + c0 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64));
}

CUtensorMap map2d(void* basep, CUtensorMapDataType dt, int esz, uint64_t rows, uint32_t box_rows, CUtensorMapSwizzle sw) {
  CUtensorMap m;
  cuuint64_t dims[2] = {KB, rows};
  cuuint64_t strides[1] = {(cuuint64_t)KB * esz};
  cuuint32_t box[2] = {KB, box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = get_encode()(&m, dt, 2, basep, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(2); }
  return m;
}

}  // namespace

int main() {
  const int R = 256;
  std::vector<float> a((size_t)R * KB), b((size_t)N * KB);
  uint64_t s = 12345;
  auto rnd = [&]() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (int)((s >> 33) % 9) - 4; };
  for (auto& v : a) v = (float)rnd();
  for (auto& v : b) v = (float)rnd();
  std::vector<__half> a16(a.size()), b16(b.size());
  std::vector<uint8_t> a8(a.size()), b8(b.size());
  for (size_t i = 0; i < a.size(); i++) { a16[i] = __float2half(a[i]); a8[i] = (uint8_t)__nv_cvt_float_to_fp8(a[i], __NV_SATFINITE, __NV_E4M3); }
  for (size_t i = 0; i < b.size(); i++) { b16[i] = __float2half(b[i]); b8[i] = (uint8_t)__nv_cvt_float_to_fp8(b[i], __NV_SATFINITE, __NV_E4M3); }
  void *dA16, *dB16, *dA8, *dB8;
  float* dOut;
  CUDA_CHECK(cudaMalloc(&dA16, a16.size() * 2)); CUDA_CHECK(cudaMalloc(&dB16, b16.size() * 2));
  CUDA_CHECK(cudaMalloc(&dA8, a8.size())); CUDA_CHECK(cudaMalloc(&dB8, b8.size()));
  CUDA_CHECK(cudaMalloc(&dOut, 128 * N * 4));
  CUDA_CHECK(cudaMemcpy(dA16, a16.data(), a16.size() * 2, cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMemcpy(dB16, b16.data(), b16.size() * 2, cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMemcpy(dA8, a8.data(), a8.size(), cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMemcpy(dB8, b8.data(), b8.size(), cudaMemcpyHostToDevice));
  CUtensorMap mA16 = map2d(dA16, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, R, HALO, CU_TENSOR_MAP_SWIZZLE_128B);
  CUtensorMap mB16 = map2d(dB16, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, N, N, CU_TENSOR_MAP_SWIZZLE_128B);
  CUtensorMap mA8 = map2d(dA8, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, R, HALO, CU_TENSOR_MAP_SWIZZLE_64B);
  CUtensorMap mB8 = map2d(dB8, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, N, N, CU_TENSOR_MAP_SWIZZLE_64B);
  CUDA_CHECK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 52 * 1024));
  const int row0 = 3;
  std::vector<float> out(128 * N);
  for (int fp8 = 0; fp8 < 2; fp8++)
    for (int bo = 0; bo < 2; bo++) {
      printf("%s base_offset=%s :", fp8 ? "e4m3/SW64 " : "fp16/SW128", bo ? "(addr>>7)&7" : "0");
      int ok_count = 0;
      for (int shift = 0; shift <= 47; shift++) {
        k_probe<<<1, 128, 52 * 1024>>>(fp8 ? mA8 : mA16, fp8 ? mB8 : mB16, fp8, row0, shift, bo, dOut);
        CUDA_CHECK(cudaGetLastError());
        CUDA_CHECK(cudaDeviceSynchronize());
        CUDA_CHECK(cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int i = 0; i < 128; i++)
          for (int j = 0; j < N; j++) {
            float ref = 0;
            for (int k = 0; k < KB; k++) ref += a[(size_t)(row0 + shift + i) * KB + k] * b[(size_t)j * KB + k];
            if (out[(size_t)i * N + j] != ref) bad++;
          }
        printf(" %d:%s", shift, bad ? "X" : "ok");
        ok_count += bad == 0;
      }
      printf("  => %d/48 shifts exact\n", ok_count);
    }
  return 0;
}
