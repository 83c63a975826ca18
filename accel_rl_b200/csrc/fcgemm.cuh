// TMA-fed tcgen05 tiles for the three dense contractions of the hidden FC layer (sm_100a).
//
// Every operand tile is ONE contiguous, pre-swizzled block in HBM, so a stage is filled by 2-5 plain 1-D bulk copies
// (cp.async.bulk, SASS UBLKCP) issued by one elected lane — no per-thread cp.async gather, no tensor maps:
//   act_fc [HW][rows][64]            the last conv layer's output, one 64-channel plane per pixel (written by the conv
//                                    epilogue), 16-byte chunks XOR-swizzled by (row & 7)
//   wfc_t  [HW][H/64][64 c][64 j]    the FC weights W[(c,hw)][j] as 8 KB tiles, chunks swizzled by (c & 7) — the SAME
//                                    tile is the N-major B operand of the forward GEMM and the K-major B operand of the
//                                    data-gradient GEMM
//   dh_t   [H/64][rows][64]          dL/dh, one 64-column plane per block of hidden units, swizzled by (row & 7)
// GEMMs (M = 128 per tile):
//   KIND 0  forward : part[split][n][H]   = act[n][K] * W[K][H]          A K-major, B N-major, split-K over pixels
//   KIND 1  dgrad   : dY[n][(hw,c)]       = dh[n][H] * W^T, masked by act > 0, scattered into the conv gradient grid
//   KIND 2  wgrad   : dW[(c,hw)][H] (fp32) = act^T[K][n] * dh[n][H]      both MN-major, two pixel planes per tile
//   KIND 3  wgrad, FOUR pixel planes per tile: two M = 128 accumulators (all 512 TMEM columns) fed by the same dh stage —
//           the kernel is bound by L2 -> shared-memory operand traffic (128 x 256 tiles re-read dh_t 54 times), so
//           a 256 x 256 tile moves 2/3 of the bytes per FLOP and occupies half the SMs
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM), warps 2..9 = epilogue.
#pragma once
#include "common.cuh"

namespace arl {

constexpr int kFcThreads = 320;
constexpr int kFcMaxCopies = 8;

struct FcCopy {
  const __nv_bfloat16* base;
  long mul_x, mul_y, mul_z, mul_it;   // element offsets per blockIdx.x / .y / .z / k-iteration
  uint32_t bytes, smem_off;
};

struct FcParams {
  FcCopy cp[kFcMaxCopies];
  int ncopies;
  int a_bytes;                 // B tile starts here inside a stage
  int stage_bytes, stages;
  int niter;                   // k-iterations per CTA (forward: per split, clipped to niter_total)
  int niter_total;
  int M;                       // valid rows of the GEMM (epilogue guard)
  // epilogue
  float* out_f32;              // KIND 0: partial [split][M][ldo];  KIND 2: flat gradient (fp32)
  int ldo;
  // KIND 1
  __nv_bfloat16* dy;           // conv gradient grid (chunk-swizzled 128-byte rows)
  const __nv_bfloat16* act;    // act_fc planes (mask source), plane stride act_plane
  long act_plane;
  int sc_Wo, sc_S, sc_Wp, sc_pad;
  // KIND 2 / 3
  // KIND 0, cluster > 1: the `cluster` CTAs of consecutive blockIdx.z (one thread-block cluster) add their split-K partials
  // through distributed shared memory — CTA rank r sums rows [r*128/cluster, ...) of all of them in rank order — and only
  // the sums go to out_f32[blockIdx.z / cluster]: 1/cluster of the partial traffic on the FC-forward -> head edge
  int cluster;
  int fc_HW;                   // D row (hw, c) -> gradient row c*HW + hw
  double* ss_out;              // optional: per-epilogue-warp sums of squares of the gradient rows written,
                               // [(blockIdx.y * gridDim.x + blockIdx.x) * 8 + warp - 2] (global-norm partials)
};

template <int KIND, int BN>
__global__ void __launch_bounds__(kFcThreads, 1) fc_gemm_kernel(const __grid_constant__ FcParams p) {
  constexpr bool A_MN = (KIND >= 2), B_MN = (KIND != 1);
  constexpr int NACC = (KIND == 3) ? 2 : 1;                // accumulators (M = 128 each) sharing one B stage
  constexpr int TMEM_COLS = NACC * BN <= 128 ? 128 : (NACC * BN <= 256 ? 256 : 512);
  static_assert(NACC * BN <= 512, "TMEM columns");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + p.stages * p.stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (8 + s); };
  const uint32_t done_bar = bar_base + 8u * 16;
  const uint32_t tmem_ptr_addr = bar_base + 8u * 17;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int niter = p.niter;
  if (KIND == 0) niter = max(0, min(p.niter, p.niter_total - (int)blockIdx.z * p.niter));

  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_addr, TMEM_COLS);
  pdl_wait();
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));

  if (warp == 0) {
    // ===================== TMA producer =====================
    uint32_t total = 0;
    for (int i = 0; i < p.ncopies; ++i) total += p.cp[i].bytes;
    for (int it = 0; it < niter; ++it) {
      const int s = it % p.stages;
      const uint32_t ph = (it / p.stages) & 1;
      mbar_wait(empty_bar(s), ph ^ 1, 41);
      if (elect_one()) {
        mbar_arrive_expect_tx(full_bar(s), total);
        const uint32_t dst = smem_base + s * p.stage_bytes;
        for (int i = 0; i < p.ncopies; ++i) {
          const FcCopy& c = p.cp[i];
          const __nv_bfloat16* src = c.base + blockIdx.x * c.mul_x + blockIdx.y * c.mul_y + blockIdx.z * c.mul_z + it * c.mul_it;
          bulk_g2s(dst + c.smem_off, src, c.bytes, full_bar(s));
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t tmem_u = make_uniform(tmem_base);
    constexpr uint32_t idesc = make_idesc_bf16(128, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
    for (int it = 0; it < niter; ++it) {
      const int s = it % p.stages;
      const uint32_t ph = (it / p.stages) & 1;
      mbar_wait(full_bar(s), ph, 42);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_tile = smem_base + s * p.stage_bytes;
        const uint32_t b_tile = a_tile + p.a_bytes;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t bd = B_MN ? make_smem_desc(b_tile + k * 2048, 8192, 1024, 2) : make_smem_desc(b_tile + k * 32, 16, 1024, 2);
#pragma unroll
          for (int a = 0; a < NACC; ++a) {    // accumulator a: pixel planes 2a, 2a+1 of the stage (16 KB apart)
            const uint64_t ad = A_MN ? make_smem_desc(a_tile + a * 16384 + k * 2048, 8192, 1024, 2)
                                     : make_smem_desc(a_tile + k * 32, 16, 1024, 2);
            umma_bf16(tmem_u + a * BN, ad, bd, idesc, (it > 0 || k > 0) ? 1u : 0u);
          }
        }
        umma_commit(empty_bar(s));
        if (it == niter - 1) umma_commit(done_bar);
      }
      __syncwarp();
    }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3, h = (warp - 2) >> 2;
    constexpr int HC = BN / 2;
    const int r = q * 32 + lane;
    if (niter > 0) {
      mbar_wait(done_bar, 0, 43);
      tc_fence_after();
    }
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + h * HC;
    // all MMAs are complete (done_bar): the operand stages are free -> per-warp transpose scratch for the row stores
    const uint32_t scratch = smem_base + (uint32_t)(warp - 2) * kRowStoreScratch;
    double ssq = 0.0;
#pragma unroll 1
    for (int cc0 = 0; cc0 < NACC * HC; cc0 += 32) {
      const int acc = cc0 / HC;               // KIND 3: second accumulator = planes 2, 3 of the tile
      const int c0 = cc0 - acc * HC;
      uint32_t v[32];
      if (niter > 0) { tmem_ld32(taddr + acc * BN + c0, v); tmem_ld_wait(); }
      else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0;
      }
      const int col = h * HC + c0;            // first of 32 tile columns
      if (KIND == 0 && p.cluster > 1) {
        // own partial -> shared memory [128 rows][32 float4], float4 slot j of row r at (j ^ (r & 31)): conflict-free for the
        // row-per-lane writes here and for the row-major reads of the reduction below
        const uint32_t red = smem_base + (uint32_t)r * 512u;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t slot = (uint32_t)((col >> 2) + j) ^ (uint32_t)(r & 31);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(red + slot * 16u), "r"(v[4 * j]), "r"(v[4 * j + 1]),
                       "r"(v[4 * j + 2]), "r"(v[4 * j + 3]) : "memory");
        }
      } else if (KIND == 0) {
        const int row = blockIdx.x * 128 + r;
        float* dst = p.out_f32 + ((long)blockIdx.z * p.M + row) * p.ldo + blockIdx.y * BN + col;
        store_rows32_coalesced(scratch, v, dst, row < p.M, lane);
      } else if (KIND == 1) {
        // (a shared-memory-transposed variant of these 64-byte row pieces — load_rows16_coalesced / store_rows16_coalesced
        // with the swizzle as a chunk permutation — was measured: 10.1 vs 9.6 us, the extra shuffles and shared-memory
        // round trips cost more than the LSU wavefronts they save; the direct form stays)
        const int row = blockIdx.x * 128 + r;   // image
        if (row < p.M) {
          const int n0 = blockIdx.y * BN + col; // GEMM column = hw*64 + c
          const int hw = n0 >> 6, cc = (n0 & 63) >> 3;
          const __nv_bfloat16* arow = p.act + hw * p.act_plane + (long)row * 64;
          const int a7 = row & 7;
          const int i_ = hw / p.sc_Wo, j_ = hw - i_ * p.sc_Wo;
          const long pos = (long)row * p.sc_S + (i_ + p.sc_pad) * p.sc_Wp + (j_ + p.sc_pad);
          __nv_bfloat16* prow = p.dy + pos * 64;
          const int x7 = (int)(pos & 7);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 m = __ldg(reinterpret_cast<const uint4*>(arow + ((cc + j) ^ a7) * 8));
            uint32_t mw[4] = {m.x, m.y, m.z, m.w};
            uint32_t pk[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              float lo = bf16_lo(mw[k]) > 0.f ? __uint_as_float(v[8 * j + 2 * k]) : 0.f;
              float hi = bf16_hi(mw[k]) > 0.f ? __uint_as_float(v[8 * j + 2 * k + 1]) : 0.f;
              pk[k] = pack_bf16x2(lo, hi);
            }
            *reinterpret_cast<uint4*>(prow + ((cc + j) ^ x7) * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
      } else {
        // D row r of accumulator acc = (plane blockIdx.x*2*NACC + 2*acc + r/64, channel r%64) -> gradient row c*HW + hw
        const int hw = blockIdx.x * 2 * NACC + 2 * acc + (r >> 6), c = r & 63;
        float* dst = p.out_f32 + ((long)c * p.fc_HW + hw) * p.ldo + blockIdx.y * BN + col;
        store_rows32_coalesced(scratch, v, dst, hw < p.fc_HW, lane);
        if (p.ss_out && hw < p.fc_HW) {
          float s32 = 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i) s32 = fmaf(__uint_as_float(v[i]), __uint_as_float(v[i]), s32);
          ssq += (double)s32;
        }
      }
    }
    if (KIND >= 2 && p.ss_out) {
      ssq = warp_sum_d(ssq);
      if (lane == 0) p.ss_out[((long)blockIdx.y * gridDim.x + blockIdx.x) * 8 + (warp - 2)] = ssq;
    }
    tc_fence_before();
  }
  if (KIND == 0 && p.cluster > 1) {
    // every thread of every CTA of the cluster: partials are in shared memory -> reduce my row slice -> peers may exit
    cluster_sync_all();
    if (warp >= 2) {
      const int S = p.cluster;
      const uint32_t crank = cluster_ctarank();
      const int rows_per = 128 / S;                       // (cluster in {2, 4, 8})
      const int t = tid - 64;                             // 0 .. 255
      for (int e = t; e < rows_per * 32; e += 256) {
        const int rr = (int)crank * rows_per + (e >> 5), f = e & 31;
        const uint32_t local = smem_base + (uint32_t)rr * 512u + (uint32_t)((f ^ (rr & 31)) * 16);
        float4 acc4 = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int z = 0; z < S; ++z) {                     // fixed rank order: bit-reproducible
          const float4 x = ld_dsmem_f4(local, (uint32_t)z);
          acc4.x += x.x; acc4.y += x.y; acc4.z += x.z; acc4.w += x.w;
        }
        const int row = blockIdx.x * 128 + rr;
        if (row < p.M)
          *reinterpret_cast<float4*>(p.out_f32 + ((long)(blockIdx.z / S) * p.M + row) * p.ldo + blockIdx.y * BN + f * 4) = acc4;
      }
    }
    cluster_sync_all();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace arl
