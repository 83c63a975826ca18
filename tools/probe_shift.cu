// Dev probe (not part of the product library): does a tcgen05 shared-memory descriptor whose start address is
// shifted by a whole number of 128-byte rows inside a 128B-swizzled region read the rows the shift implies?
//   kernel 1 (K-major A, conv forward style):  D[128][N] = sum_t  Apix[r + shift_t][0:64] . B_t[n][0:64]
//   kernel 2 (MN-major A, wgrad style):        D[m][n]   = sum_p  Apix[p + s(m)][m % 64] * dY[p][n],  s(m) = m < 64 ? s0 : s1
// mode 0: descriptor base_offset field = 0;  mode 1: base_offset = (shift & 7).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -shared -I accel_rl_b200/csrc \
//             -o tools/libprobe_shift.so tools/probe_shift.cu
#include "common.cuh"

using namespace arl;

__global__ void __launch_bounds__(128) probe_fwd_kernel(const __nv_bfloat16* __restrict__ apix, int P,
                                                        const __nv_bfloat16* __restrict__ B, const int* __restrict__ shifts,
                                                        int T, int N, float* __restrict__ D, int mode) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = base;
  const uint32_t b_base = base + ((P * 128 + 1023) & ~1023);
  const uint32_t bar = b_base + T * N * 128;
  const uint32_t tptr = bar + 8;
  const int tid = threadIdx.x;
  for (int i = tid; i < P * 8; i += 128) {
    int r = i >> 3, c = i & 7;
    uint4 v = *reinterpret_cast<const uint4*>(apix + (long)r * 64 + c * 8);
    st_shared_v4(a_base + swz_off<128>(r, c), v);
  }
  for (int i = tid; i < T * N * 8; i += 128) {
    int t = i / (N * 8), rem = i % (N * 8), r = rem >> 3, c = rem & 7;
    uint4 v = *reinterpret_cast<const uint4*>(B + ((long)t * N + r) * 64 + c * 8);
    st_shared_v4(b_base + t * N * 128 + swz_off<128>(r, c), v);
  }
  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (tid < 32) tmem_alloc(tptr, 64);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(tptr));
  if (tid == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
    for (int t = 0; t < T; ++t) {
      const int s = shifts[t];
      for (int k = 0; k < 4; ++k) {
        uint64_t ad = make_smem_desc(a_base + s * 128 + k * 32, 16, 1024, 2);
        if (mode == 1) ad |= (uint64_t)(s & 7) << 49;
        uint64_t bd = make_smem_desc(b_base + t * N * 128 + k * 32, 16, 1024, 2);
        umma_bf16(tmem, ad, bd, idesc, (t > 0 || k > 0) ? 1u : 0u);
      }
    }
    umma_commit(bar);
  }
  mbar_wait(bar, 0, 1);
  tc_fence_after();
  const int warp = tid >> 5;
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t r[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) D[(long)tid * N + c0 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem, 64);
}

__global__ void __launch_bounds__(128) probe_wgrad_kernel(const __nv_bfloat16* __restrict__ apix, int P,
                                                          const __nv_bfloat16* __restrict__ dy /*[128][64]*/, int s0, int s1,
                                                          float* __restrict__ D /*[128][64]*/, int mode) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = base;
  const uint32_t b_base = base + ((P * 128 + 1023) & ~1023);
  const uint32_t bar = b_base + 128 * 128;
  const uint32_t tptr = bar + 8;
  const int tid = threadIdx.x;
  for (int i = tid; i < P * 8; i += 128) {
    int r = i >> 3, c = i & 7;
    uint4 v = *reinterpret_cast<const uint4*>(apix + (long)r * 64 + c * 8);
    st_shared_v4(a_base + swz_off<128>(r, c), v);
  }
  for (int i = tid; i < 128 * 8; i += 128) {
    int r = i >> 3, c = i & 7;
    uint4 v = *reinterpret_cast<const uint4*>(dy + (long)r * 64 + c * 8);
    st_shared_v4(b_base + swz_off<128>(r, c), v);
  }
  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (tid < 32) tmem_alloc(tptr, 64);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(tptr));
  if (tid == 0) {
    const uint32_t idesc = make_idesc_bf16(128, 64, 1, 1);
    for (int k = 0; k < 8; ++k) {   // 128 pixels, 16 per MMA
      uint64_t ad = make_smem_desc(a_base + s0 * 128 + k * 2048, (uint32_t)(s1 - s0) * 128, 1024, 2);
      if (mode == 1) ad |= (uint64_t)(s0 & 7) << 49;
      uint64_t bd = make_smem_desc(b_base + k * 2048, 8192, 1024, 2);
      umma_bf16(tmem, ad, bd, idesc, k > 0 ? 1u : 0u);
    }
    umma_commit(bar);
  }
  mbar_wait(bar, 0, 1);
  tc_fence_after();
  const int warp = tid >> 5;
  for (int c0 = 0; c0 < 64; c0 += 16) {
    uint32_t r[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) D[(long)tid * 64 + c0 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem, 64);
}

extern "C" int probe_fwd(const void* apix, int P, const void* B, const int* shifts_dev, int T, int N, float* D, int mode,
                         void* stream) {
  int smem = ((P * 128 + 1023) & ~1023) + T * N * 128 + 1024 + 64;
  cudaFuncSetAttribute(probe_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe_fwd_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)apix, P, (const __nv_bfloat16*)B, shifts_dev,
                                                           T, N, D, mode);
  return (int)cudaGetLastError();
}

extern "C" int probe_wgrad(const void* apix, int P, const void* dy, int s0, int s1, float* D, int mode, void* stream) {
  int smem = ((P * 128 + 1023) & ~1023) + 128 * 128 + 1024 + 64;
  cudaFuncSetAttribute(probe_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe_wgrad_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)apix, P, (const __nv_bfloat16*)dy, s0, s1, D,
                                                             mode);
  return (int)cudaGetLastError();
}
