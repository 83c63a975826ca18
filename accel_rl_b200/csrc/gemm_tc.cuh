// tcgen05 tensor-core tiles for the policy/value network (sm_100a).
//
// Two kernels cover every dense contraction on the path:
//
//   rowgemm_kernel : D[128 rows x BN] = A[128 x K] * B[BN x K]^T        (A K-major)
//       conv forward (implicit GEMM, rows = output pixels), conv dgrad (rows = input
//       pixels of one stride-parity class), FC forward (split-K) and FC dgrad.
//   wgrad_kernel   : D[K' x BN] = sum_rows A[row, K']^T * dY[row, BN]   (both MN-major)
//       conv / FC weight gradients, split over row ranges.
//
// Operand tiles are staged in shared memory in the UMMA canonical 128-byte-swizzled
// layout by 8 producer warps (the im2col gather is fused into that load: the patch is
// never materialised in HBM); one elected thread issues tcgen05.mma with the fp32
// accumulator in TMEM; the same 8 warps read it back with tcgen05.ld for the fused
// epilogue (bias + ReLU + bf16 pack, dReLU mask, or fp32 split partials).
// Producer <-> MMA hand-off is an mbarrier ring (full/empty per stage).
#pragma once
#include "common.cuh"

namespace arl {

constexpr int kProducerWarps = 8;
constexpr int kWgradProducerWarps = 16;                    // wgrad: two groups of 8 alternate stages
constexpr int kWgradThreads = kWgradProducerWarps * 32 + 32;
constexpr int kProducerThreads = kProducerWarps * 32;  // 256
constexpr int kGemmThreads = kProducerThreads + 32;    // + MMA warp
constexpr int kBK = 64;                                // K elements per stage (128 B of bf16)

// ---------------------------------------------------------------------------
// operand loaders (run by the 256 producer threads; tid in [0,256))
// ---------------------------------------------------------------------------

// Row permutation of a dense source: logical row q (in (hw, c) order) lives at source row
// (q % C) * HW + q / C  — the FC weight matrix is kept in the reference's (c, h, w) row order while
// the activations are NHWC.  C == 0: identity.
struct RowPerm {
  int C, HW;
};
ARL_DEVINL int perm_row(const RowPerm& p, int q) {
  if (p.C == 0) return q;
  int hw = q / p.C;
  return (q - hw * p.C) * p.HW + hw;
}

// Dense row-major bf16 source: copies R rows x ROWB bytes into a swizzled tile.
//   tile row r  <- src[perm(row0 + r) * ld + col0 .. + ROWB/2)   (zero when row0+r >= nrows)
// NT = number of threads cooperating (tid in [0, NT)).
template <int R, int ROWB, int NT = 256>
ARL_DEVINL void fill_dense(uint32_t tile, const __nv_bfloat16* __restrict__ src, long ld, int row0, int nrows,
                           int col0, int tid, RowPerm perm = RowPerm{0, 0}) {
  constexpr int CH = ROWB / 16;        // 16-byte chunks per row
  constexpr int TOTAL = R * CH;        // chunks in the tile
#pragma unroll
  for (int i = tid; i < TOTAL; i += NT) {
    int r = i / CH, c = i % CH;
    uint4 v = make_uint4(0, 0, 0, 0);
    int row = row0 + r;
    if (row < nrows) v = __ldg(reinterpret_cast<const uint4*>(src + (long)perm_row(perm, row) * ld + col0 + c * 8));
    st_shared_v4(tile + swz_off<ROWB>(r, c), v);
  }
}

// Asynchronous variant: cp.async (LDGSTS) straight into the swizzled tile, no register staging; rows past
// nrows are zero-filled (src-size 0).  Completion is tracked by cp_async_mbar_arrive on the stage barrier.
template <int R, int ROWB, int NT = 256>
ARL_DEVINL void fill_dense_async(uint32_t tile, const __nv_bfloat16* __restrict__ src, long ld, int row0, int nrows,
                                 int col0, int tid, RowPerm perm = RowPerm{0, 0}) {
  constexpr int CH = ROWB / 16;
  constexpr int TOTAL = R * CH;
#pragma unroll
  for (int i = tid; i < TOTAL; i += NT) {
    int r = i / CH, c = i % CH;
    int row = row0 + r;
    bool ok = row < nrows;
    const __nv_bfloat16* p = ok ? src + (long)perm_row(perm, row) * ld + col0 + c * 8 : src;
    cp_async16(tile + swz_off<ROWB>(r, c), p, ok ? 16u : 0u);
  }
}

// Gather geometry for implicit-GEMM tiles whose rows are spatial positions.
// Rows enumerate (b, qy, qx); K index k' = (ty * Tx + tx) * C + c over an NHWC bf16 source.
//   src_y = qy * sy + y0 + ty * dty,  src_x = qx * sx + x0 + tx * dtx   (out of range -> 0)
struct ConvGeom {
  const __nv_bfloat16* src;
  const int* idx;        // optional image gather: image b reads source image idx[off*nb + b]
  const int* idx_off;    // optional device scalar `off` (minibatch index, graph-replayed training)
  int Qh, Qw;            // row grid per image
  int Hs, Ws, C;         // source dims
  int sy, y0, dty;
  int sx, x0, dtx;
  int Tx;                // taps per tap-row
  int nrows;             // nb * Qh * Qw
};

// Decoded position of one GEMM row: source image, and the (y, x) of tap (0,0) in it
struct RowPos {
  int img;     // source image index (after the optional gather), -1 when the row is out of range
  int ysxs;    // (qy*sy + y0) in the low 16 bits, (qx*sx + x0) in the high 16 bits (both signed)
};

ARL_DEVINL RowPos conv_row_pos(const ConvGeom& g, int row) {
  RowPos rp;
  if (row >= g.nrows) {
    rp.img = -1; rp.ysxs = 0;
    return rp;
  }
  int per = g.Qh * g.Qw;
  int b = row / per;
  int rem = row - b * per;
  int qy = rem / g.Qw;
  int qx = rem - qy * g.Qw;
  int img = b;
  if (g.idx) img = g.idx[(g.idx_off ? (long)g.idx_off[0] * (g.nrows / per) : 0) + b];
  rp.img = img;
  int ys = qy * g.sy + g.y0, xs = qx * g.sx + g.x0;
  rp.ysxs = (ys & 0xffff) | (xs << 16);
  return rp;
}

// One [R rows x 64 k'] tile (k-block kb).  Thread owns chunk column j = tid & 7 and rows (tid>>3) + 32*i.
// Everything that does not change from k-block to k-block is hoisted: the swizzled shared-memory offsets,
// the per-row source pointer / origin, and a per-CTA table (shared memory) that decodes the K chunk index
// into its tap offset — so the gather costs a handful of integer ops and one cp.async per 16 bytes.
template <int R>
struct ConvLoader {
  static constexpr bool kNeedsTable = true;
  ConvGeom g;
  const int4* ktab;                        // [K/8] {dy, dx, element offset of (dy,dx,c0), 0} (shared memory)
  const __nv_bfloat16* rptr[R / 32];       // image base pointer of each owned row (nullptr: row out of range)
  int roff[R / 32];                        // element offset of the row's tap-(0,0) pixel inside its image
  int ysxs[R / 32];
  uint32_t soff[R / 32];                   // swizzled byte offset of (row, chunk column) inside a tile

  // called by all threads of the CTA once (followed by a __syncthreads in the kernel)
  ARL_DEVINL void build_table(int4* tab, int nchunks, int tid, int nthreads, int kc0 = 0) const {
    const int cpt = g.C >> 3;
    for (int i = tid; i < nchunks; i += nthreads) {
      int kc = kc0 + i;                      // global 16-byte chunk index along K; tab is indexed locally
      int tap = kc / cpt;
      int cc = kc - tap * cpt;
      int ty = tap / g.Tx;
      int tx = tap - ty * g.Tx;
      int dy = ty * g.dty, dx = tx * g.dtx;
      tab[i] = make_int4(dy, dx, (dy * g.Ws + dx) * g.C + cc * 8, 0);
    }
  }
  ARL_DEVINL void init_thread(const int4* tab, int tid) {
    ktab = tab;
#pragma unroll
    for (int i = 0; i < R / 32; ++i) soff[i] = swz_off<128>((tid >> 3) + 32 * i, tid & 7);
  }
  ARL_DEVINL void set_row(int i, RowPos rp) {
    int ys = (int)(short)(rp.ysxs & 0xffff), xs = rp.ysxs >> 16;
    rptr[i] = rp.img >= 0 ? g.src + (long)rp.img * ((long)g.Hs * g.Ws * g.C) : nullptr;
    roff[i] = (ys * g.Ws + xs) * g.C;
    ysxs[i] = rp.ysxs;
  }
  ARL_DEVINL void prepare(int row0, int tid) {
#pragma unroll
    for (int i = 0; i < R / 32; ++i) set_row(i, conv_row_pos(g, row0 + (tid >> 3) + 32 * i));
  }
  // rows decoded earlier into a shared-memory table (wgrad: one decode per row per CTA)
  ARL_DEVINL void prepare_tab(const RowPos* rowtab, int local_row0, int tid) {
#pragma unroll
    for (int i = 0; i < R / 32; ++i) set_row(i, rowtab[local_row0 + (tid >> 3) + 32 * i]);
  }
  ARL_DEVINL void fill_async(uint32_t tile, int kb, int tid) const {
    const int4 e = ktab[kb * 8 + (tid & 7)];
#pragma unroll
    for (int i = 0; i < R / 32; ++i) {
      int y = (int)(short)(ysxs[i] & 0xffff) + e.x, x = (ysxs[i] >> 16) + e.y;
      bool ok = rptr[i] != nullptr && (unsigned)y < (unsigned)g.Hs && (unsigned)x < (unsigned)g.Ws;
      const __nv_bfloat16* p = ok ? rptr[i] + (roff[i] + e.z) : g.src;
      cp_async16(tile + soff[i], p, ok ? 16u : 0u);
    }
  }
};

// Dense K-major A operand (FC forward / FC dgrad: rows = batch rows of a row-major matrix)
template <int R>
struct DenseLoader {
  static constexpr bool kNeedsTable = false;
  const __nv_bfloat16* src;
  long ld;
  int nrows;
  int row0;
  ARL_DEVINL void build_table(int4*, int, int, int, int = 0) const {}
  ARL_DEVINL void init_thread(const int4*, int) {}
  ARL_DEVINL void prepare(int r0, int) { row0 = r0; }
  ARL_DEVINL void prepare_tab(const RowPos*, int, int) {}
  ARL_DEVINL void fill_async(uint32_t tile, int kb, int tid) const {
    fill_dense_async<R, 128>(tile, src, ld, row0, nrows, kb * kBK, tid);
  }
};

// ---------------------------------------------------------------------------
// epilogue description (runtime-switched; the branch is uniform per launch)
// ---------------------------------------------------------------------------
enum EpiMode { EPI_BIAS_RELU_BF16 = 0, EPI_MASK_BF16 = 1, EPI_PARTIAL_F32 = 2, EPI_BIAS_BF16 = 3 };

struct RowEpi {
  int mode;
  float scale;                 // applied to the accumulator before bias (conv1: 1/255)
  const float* bias;           // [N] (modes 0,3)
  __nv_bfloat16* out;          // bf16 destination (modes 0,1,3), row pitch ldo
  const __nv_bfloat16* act;    // forward activation at the destination (mode 1: dReLU mask)
  float* partial;              // fp32 [split][M][ldo] (mode 2)
  int ldo;
  int M;                       // valid rows
  // destination-row map: identity when map_s == 0, else rows enumerate (b,qy,qx) of a
  // stride-parity class and land at (b, map_s*qy + map_y0, map_s*qx + map_x0) of an H x W grid
  int map_s, map_y0, map_x0, map_Qh, map_Qw, map_H, map_W;
  // EPI_MASK_BF16 scatter (FC data gradient -> the last conv layer's gradient grid of pconv.cuh): GEMM row = image,
  // column n = (i*sc_Wo + j)*64 + c  ->  position row*sc_S + (i+sc_pad)*sc_Wp + (j+sc_pad), 128-byte rows whose
  // 16-byte chunks are XOR-swizzled by (position & 7).  The mask is still read from the dense `act`.
  int sc_on, sc_Wo, sc_S, sc_Wp, sc_pad;
};

ARL_DEVINL long epi_dest_row(const RowEpi& e, int row) {
  if (e.map_s == 0) return row;
  int per = e.map_Qh * e.map_Qw;
  int b = row / per;
  int rem = row - b * per;
  int qy = rem / e.map_Qw;
  int qx = rem - qy * e.map_Qw;
  return ((long)b * e.map_H + (e.map_s * qy + e.map_y0)) * e.map_W + (e.map_s * qx + e.map_x0);
}

// 32 consecutive accumulator columns of one row -> destination
ARL_DEVINL void epi_store32(const RowEpi& e, int row, int n0, const uint32_t (&r)[32], int split) {
  if (row >= e.M) return;
  if (e.mode == EPI_PARTIAL_F32) {
    float4* dst = reinterpret_cast<float4*>(e.partial + ((long)split * e.M + row) * e.ldo + n0);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      dst[i] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]),
                           __uint_as_float(r[4 * i + 3]));
    return;
  }
  long drow = epi_dest_row(e, row);
  uint32_t packed[16];
  if (e.mode == EPI_MASK_BF16) {
    const uint4* a = reinterpret_cast<const uint4*>(e.act + drow * e.ldo + n0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 m = __ldg(a + i);
      uint32_t mw[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float lo = __uint_as_float(r[8 * i + 2 * k]), hi = __uint_as_float(r[8 * i + 2 * k + 1]);
        lo = bf16_lo(mw[k]) > 0.f ? lo : 0.f;
        hi = bf16_hi(mw[k]) > 0.f ? hi : 0.f;
        packed[4 * i + k] = pack_bf16x2(lo, hi);
      }
    }
  } else {
    const bool relu = (e.mode == EPI_BIAS_RELU_BF16);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float lo = __uint_as_float(r[2 * i]) * e.scale + __ldg(e.bias + n0 + 2 * i);
      float hi = __uint_as_float(r[2 * i + 1]) * e.scale + __ldg(e.bias + n0 + 2 * i + 1);
      if (relu) { lo = fmaxf(lo, 0.f); hi = fmaxf(hi, 0.f); }
      packed[i] = pack_bf16x2(lo, hi);
    }
  }
  if (e.sc_on) {
    const int hw = n0 >> 6, c0 = (n0 & 63) >> 3;
    const int i = hw / e.sc_Wo, j = hw - i * e.sc_Wo;
    const long pos = (long)row * e.sc_S + (i + e.sc_pad) * e.sc_Wp + (j + e.sc_pad);
    __nv_bfloat16* prow = e.out + pos * 64;
    const int x7 = (int)(pos & 7);
#pragma unroll
    for (int q = 0; q < 4; ++q)
      *reinterpret_cast<uint4*>(prow + ((c0 + q) ^ x7) * 8) =
          make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
    return;
  }
  uint4* dst = reinterpret_cast<uint4*>(e.out + drow * e.ldo + n0);
#pragma unroll
  for (int i = 0; i < 4; ++i) dst[i] = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
}

ARL_DEVINL void epi_store16(const RowEpi& e, int row, int n0, const uint32_t (&r16)[16], int split) {
  // 16-column variant (BN == 32 tiles): widen into two halves of the 32-column path is not possible,
  // so handle it directly.
  if (row >= e.M) return;
  if (e.mode == EPI_PARTIAL_F32) {
    float4* dst = reinterpret_cast<float4*>(e.partial + ((long)split * e.M + row) * e.ldo + n0);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      dst[i] = make_float4(__uint_as_float(r16[4 * i]), __uint_as_float(r16[4 * i + 1]),
                           __uint_as_float(r16[4 * i + 2]), __uint_as_float(r16[4 * i + 3]));
    return;
  }
  long drow = epi_dest_row(e, row);
  uint32_t packed[8];
  if (e.mode == EPI_MASK_BF16) {
    const uint4* a = reinterpret_cast<const uint4*>(e.act + drow * e.ldo + n0);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      uint4 m = __ldg(a + i);
      uint32_t mw[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float lo = __uint_as_float(r16[8 * i + 2 * k]), hi = __uint_as_float(r16[8 * i + 2 * k + 1]);
        lo = bf16_lo(mw[k]) > 0.f ? lo : 0.f;
        hi = bf16_hi(mw[k]) > 0.f ? hi : 0.f;
        packed[4 * i + k] = pack_bf16x2(lo, hi);
      }
    }
  } else {
    const bool relu = (e.mode == EPI_BIAS_RELU_BF16);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float lo = __uint_as_float(r16[2 * i]) * e.scale + __ldg(e.bias + n0 + 2 * i);
      float hi = __uint_as_float(r16[2 * i + 1]) * e.scale + __ldg(e.bias + n0 + 2 * i + 1);
      if (relu) { lo = fmaxf(lo, 0.f); hi = fmaxf(hi, 0.f); }
      packed[i] = pack_bf16x2(lo, hi);
    }
  }
  uint4* dst = reinterpret_cast<uint4*>(e.out + drow * e.ldo + n0);
#pragma unroll
  for (int i = 0; i < 2; ++i) dst[i] = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
}

// ---------------------------------------------------------------------------
// rowgemm: D[128 x BN] = A[128 x K] * B^T
//   B K-major : weights [Ntot][ldb] row-major (ldb = K), tile rows n0..n0+BN
//   B N-major : matrix  [K][ldb]   row-major (ldb = Ntot), tile cols n0..n0+BN (BN multiple of 64)
// grid = (ceil(M/128), Ntot/BN, splits); each split covers kb_per_split k-blocks.
// ---------------------------------------------------------------------------
template <int BN>
struct RowGemmCfg {
  static constexpr int STAGES = 3;
  static constexpr int A_BYTES = 128 * 128;      // 128 rows x 64 bf16
  static constexpr int B_BYTES = BN * 128;       // BN rows x 64 bf16 (either major)
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
};

struct WeightSrc {
  const __nv_bfloat16* w;
  long ldb;
  int kdim;      // rows available along K (N-major) / unused (K-major)
  RowPerm perm;  // permutation of the source rows (tile rows for K-major, k rows for N-major)
};

template <class ALoad, bool B_NMAJOR, int BN>
ARL_DEVINL void rowgemm_body(ALoad& aload, const WeightSrc& bsrc, const RowEpi& epi, int num_kb, int kb_per_split,
                             int m_tile, int n_tile, int split) {
  using Cfg = RowGemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES;
  // barriers: full[s] @ +8*s, empty[s] @ +8*(STAGES+s), tmem_full @ +8*2*STAGES, tmem ptr @ +8*2*STAGES+8
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * Cfg::STAGES);
  const uint32_t tmem_ptr_addr = tmem_full_bar + 8u;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int m0 = m_tile * 128;
  const int n0 = n_tile * BN;
  const int kb0 = split * kb_per_split;
  const int kb1 = min(num_kb, kb0 + kb_per_split);
  const int niter = kb1 - kb0;

  if (tid == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(full_bar(s), kProducerThreads);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == kProducerWarps) tmem_alloc(tmem_ptr_addr, Cfg::TMEM_COLS);
  pdl_wait();
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));

  if (warp < kProducerWarps) {
    // ===================== producers =====================
    aload.prepare(m0, tid);
    for (int it = 0; it < niter; ++it) {
      const int s = it % Cfg::STAGES;
      const uint32_t ph = (it / Cfg::STAGES) & 1;
      mbar_wait(empty_bar(s), ph ^ 1, 1);
      const uint32_t a_tile = smem_base + s * Cfg::STAGE_BYTES;
      const uint32_t b_tile = a_tile + Cfg::A_BYTES;
      const int kb = kb0 + it;
      aload.fill_async(a_tile, kb, tid);
      if (!B_NMAJOR) {
        fill_dense_async<BN, 128>(b_tile, bsrc.w, bsrc.ldb, n0, 1 << 30, kb * kBK, tid, bsrc.perm);
      } else {
#pragma unroll
        for (int at = 0; at < BN / 64; ++at)
          fill_dense_async<64, 128>(b_tile + at * 8192, bsrc.w, bsrc.ldb, kb * kBK, bsrc.kdim, n0 + at * 64, tid,
                                    bsrc.perm);
      }
      cp_async_mbar_arrive(full_bar(s));   // phase completes only once this thread's copies have landed
      mbar_arrive(full_bar(s));            // ... and every producer thread has issued its share
    }
    // ===================== epilogue =====================
    if (niter > 0) {
      mbar_wait(tmem_full_bar, 0, 2);
      tc_fence_after();
      const int q = warp & 3;
      const int half = warp >> 2;
      const int row = m0 + q * 32 + (tid & 31);
      const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
      if constexpr (BN >= 64) {
        constexpr int COLS_PER_WARP = BN / 2;
#pragma unroll
        for (int c = 0; c < COLS_PER_WARP; c += 32) {
          uint32_t r[32];
          tmem_ld32(lane_addr + half * COLS_PER_WARP + c, r);
          tmem_ld_wait();
          epi_store32(epi, row, n0 + half * COLS_PER_WARP + c, r, split);
        }
      } else if (BN == 32 || half == 0) {
        // BN == 32: two column halves of 16; BN == 16: warps 0-3 only (warp-uniform branch)
        const int coff = (BN == 32) ? half * 16 : 0;
        uint32_t r16[16];
        tmem_ld16(lane_addr + coff, r16);
        tmem_ld_wait();
        epi_store16(epi, row, n0 + coff, r16, split);
      }
      tc_fence_before();
    }
  } else {
    // ===================== MMA issuer (converged warp, one elected lane issues: see common.cuh elect_one) ==========
    constexpr uint32_t idesc = make_idesc_bf16(128, BN, 0, B_NMAJOR ? 1 : 0);
    const uint32_t tmem_u = make_uniform(tmem_base);
    for (int it = 0; it < niter; ++it) {
      const int s = it % Cfg::STAGES;
      const uint32_t ph = (it / Cfg::STAGES) & 1;
      mbar_wait(full_bar(s), ph, 3);
      fence_proxy_async();                 // cp.async (generic-proxy) writes -> tcgen05 (async-proxy) reads
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_tile = smem_base + s * Cfg::STAGE_BYTES;
        const uint32_t b_tile = a_tile + Cfg::A_BYTES;
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k) {
          uint64_t adesc = make_smem_desc(a_tile + k * 32, 16, 1024, 2);
          uint64_t bdesc = B_NMAJOR ? make_smem_desc(b_tile + k * 2048, 8192, 1024, 2)
                                    : make_smem_desc(b_tile + k * 32, 16, 1024, 2);
          umma_bf16(tmem_u, adesc, bdesc, idesc, (it > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(empty_bar(s));
        if (it == niter - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
    }
  }
  __syncthreads();
  if (warp == kProducerWarps) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <class ALoad, bool B_NMAJOR, int BN>
__global__ void __launch_bounds__(kGemmThreads) rowgemm_kernel(ALoad aload, WeightSrc bsrc, RowEpi epi,
                                                               int num_kb, int kb_per_split) {
  rowgemm_body<ALoad, B_NMAJOR, BN>(aload, bsrc, epi, num_kb, kb_per_split, blockIdx.x, blockIdx.y, blockIdx.z);
}

// Several independent GEMMs of the same shape class in ONE launch (blockIdx.y selects the problem):
// the stride-parity classes of a conv dgrad.
constexpr int kMaxMulti = 4;
template <class ALoad>
struct RowGemmMulti {
  ALoad a[kMaxMulti];
  WeightSrc b[kMaxMulti];
  RowEpi e[kMaxMulti];
  int mtiles[kMaxMulti];
  int num_kb[kMaxMulti];
};

// ---------------------------------------------------------------------------
// conv_gemm_persist: persistent, warp-specialised implicit-GEMM conv tiles.
//   D[128 rows x BN] = im2col(A)[128 x K] * W[BN x K]^T  for every 128-row tile of up to kMaxMulti
//   independent problems ("classes": 1 for a forward conv, the stride-parity classes for a dgrad).
// grid = (CTAs per class, classes); each CTA loops over the tiles of its class (tile += gridDim.x):
//   warps 0-7  : producers — cp.async im2col gather of the A tile into a 4-stage ring
//   warp  8    : one thread issues tcgen05.mma; accumulators ping-pong between two TMEM buffers
//   warps 9-12 : epilogue — tcgen05.ld of the finished buffer, bias/ReLU/mask, bf16 stores, while the
//                next tile's MMAs run
// The whole weight matrix W (BN x K bf16, <= 72 KB) is loaded ONCE per CTA and stays resident in shared
// memory; barriers/TMEM are set up once per CTA instead of once per tile.
// ---------------------------------------------------------------------------
constexpr int kPersistThreads = kProducerThreads + 32 + 128;   // 416
#ifdef ARL_TRACE
__device__ long long g_trace[4096];
#define ARL_T(slot) do { if (blockIdx.x == 0 && blockIdx.y == 0 && (slot) < 4096) g_trace[(slot)] = clock64(); } while (0)
#else
#define ARL_T(slot) do {} while (0)
#endif
constexpr int kPersistStages = 4;

template <int BN>
__host__ __device__ constexpr int conv_persist_smem(int K) {
  return BN * K * 2 + kPersistStages * 16384 + 1024 + 256 + (K / 8) * 16;   // B, A ring, align, barriers, k-table
}

template <int BN>
__global__ void __launch_bounds__(kPersistThreads, 2) conv_gemm_persist_kernel(
    const __grid_constant__ RowGemmMulti<ConvLoader<128>> p) {
  const int cls = blockIdx.y;
  const int ntiles = p.mtiles[cls];
  if ((int)blockIdx.x >= ntiles) return;          // whole CTA exits before any barrier / TMEM use
  const int num_kb = p.num_kb[cls];
  constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_base = smem_base;                                   // num_kb tiles of [BN x 128 B]
  const uint32_t a_base = smem_base + num_kb * (BN * 128);             // (BN*128 is a multiple of 1024 for BN >= 8)
  const uint32_t bar_base = a_base + kPersistStages * 16384;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kPersistStages + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * kPersistStages + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * kPersistStages + 2 + b); };
  const uint32_t bfull_bar = bar_base + 8u * (2 * kPersistStages + 4);
  const uint32_t tmem_ptr_addr = bfull_bar + 8u;
  int4* ktab = reinterpret_cast<int4*>(smem_raw + (bar_base + 256u - smem_u32(smem_raw)));

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < kPersistStages; ++s) {
      mbar_init(full_bar(s), kProducerThreads);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), 4);
    }
    mbar_init(bfull_bar, kProducerThreads);
    fence_mbar_init();
  }
  p.a[cls].build_table(ktab, num_kb * 8, tid, kPersistThreads);
  if (warp == kProducerWarps) tmem_alloc(tmem_ptr_addr, TMEM_COLS);
  pdl_wait();
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));

  if (warp < kProducerWarps) {
    // ===================== producers =====================
    const WeightSrc& bsrc = p.b[cls];
    for (int kb = 0; kb < num_kb; ++kb)
      fill_dense_async<BN, 128>(b_base + kb * (BN * 128), bsrc.w, bsrc.ldb, 0, 1 << 30, kb * kBK, tid, bsrc.perm);
    cp_async_mbar_arrive(bfull_bar);
    mbar_arrive(bfull_bar);
    ConvLoader<128> aload = p.a[cls];
    aload.init_thread(ktab, tid);
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      aload.prepare(tile * 128, tid);
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = it % kPersistStages;
        const uint32_t ph = (it / kPersistStages) & 1;
        mbar_wait(empty_bar(s), ph ^ 1, 11);
        if (tid == 0) ARL_T(it * 6 + 0);
        aload.fill_async(a_base + s * 16384, kb, tid);
        cp_async_mbar_arrive(full_bar(s));
        mbar_arrive(full_bar(s));
        if (tid == 0) ARL_T(it * 6 + 1);
      }
    }
  } else if (warp == kProducerWarps) {
    // ===================== MMA issuer =====================
    if (tid == kProducerThreads) {
      constexpr uint32_t idesc = make_idesc_bf16(128, BN, 0, 0);
      mbar_wait(bfull_bar, 0, 12);
      int it = 0, tcount = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
        const int acc = tcount & 1;
        const uint32_t aph = (tcount >> 1) & 1;
        mbar_wait(tempty_bar(acc), aph ^ 1, 13);       // epilogue has drained this accumulator
        tc_fence_after();
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % kPersistStages;
          const uint32_t ph = (it / kPersistStages) & 1;
          mbar_wait(full_bar(s), ph, 14);
          ARL_T(it * 6 + 2);
          fence_proxy_async();
          tc_fence_after();
          const uint32_t a_tile = a_base + s * 16384;
          const uint32_t b_tile = b_base + kb * (BN * 128);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            uint64_t adesc = make_smem_desc(a_tile + k * 32, 16, 1024, 2);
            uint64_t bdesc = make_smem_desc(b_tile + k * 32, 16, 1024, 2);
            umma_bf16(tmem_base + acc * BN, adesc, bdesc, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(empty_bar(s));
          ARL_T(it * 6 + 3);
        }
        umma_commit(tfull_bar(acc));
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const RowEpi& epi = p.e[cls];
    const int q = warp & 3;                         // TMEM lane quadrant this warp may access
    int tcount = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
      const int acc = tcount & 1;
      const uint32_t aph = (tcount >> 1) & 1;
      mbar_wait(tfull_bar(acc), aph, 15);
      if (warp == 9 && (tid & 31) == 0) ARL_T(tcount * 6 + 4);
      tc_fence_after();
      const int row = tile * 128 + q * 32 + (tid & 31);
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
      if constexpr (BN >= 32) {
        uint32_t r[BN / 32][32];
#pragma unroll
        for (int c = 0; c < BN / 32; ++c) tmem_ld32(taddr + c * 32, r[c]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(tempty_bar(acc));   // accumulator is in registers: release it
#pragma unroll
        for (int c = 0; c < BN / 32; ++c) epi_store32(epi, row, c * 32, r[c], 0);
        if (warp == 9 && (tid & 31) == 0) ARL_T(tcount * 6 + 5);
      } else {
        uint32_t r16[16];
        tmem_ld16(taddr, r16);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(tempty_bar(acc));
        epi_store16(epi, row, 0, r16, 0);
      }
    }
  }
  __syncthreads();
  if (warp == kProducerWarps) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------
// wgrad: D[K' x BN] = sum over rows of A[row, K']^T * dY[row, BN]
//   A tiles: [64 rows x 64 k'] atoms filled by the same loaders as the forward pass
//            (MN-major for the MMA: k' is the contiguous dimension)
//   B tile : [64 rows x BN] from the row-major dY (N-major); BN*2 bytes per row
//            (32/64/128-byte rows -> SW32/64/128; wider rows split into 64-column atoms)
// One CTA owns MT consecutive 128-row M-tiles of K' (starting at blockIdx.x*MT), BN columns
// starting at blockIdx.y*BN, and the row range of split blockIdx.z.  Sixteen producer warps in
// two groups fill alternate stages (two stages' gathers in flight per SM); while copying dY the
// producers also accumulate its column sums = the layer's bias gradient (per-split partial).
// ---------------------------------------------------------------------------
struct WgradEpi {
  float* out;        // fp32 destination
  int mode;          // 0: partial[split][Kp][ldo]  1: FC direct (row map k'=(hw*C+c) -> c*HW+hw)
  int Kvalid;        // valid k' rows
  int Kp;            // padded rows of the partial buffer
  int ldo;           // row pitch (floats)
  int fc_C, fc_HW;   // mode 1 row permutation
  float* bias_out;   // optional [splits][BN] column sums of dY over this split's rows
};

template <int MT, int BN>
struct WgradCfg {
  static constexpr int ROWB = (BN * 2 >= 128) ? 128 : BN * 2;  // bytes per B smem row
  static constexpr int B_ATOMS = (BN * 2 + 127) / 128;
  static constexpr int A_BYTES = MT * 2 * 8192;                // MT*2 atoms of [64 x 128B]
  static constexpr int B_BYTES = 64 * BN * 2;
  static constexpr int STAGE_BYTES = ((A_BYTES + B_BYTES + 1023) / 1024) * 1024;
  static constexpr int STAGES = (STAGE_BYTES * 4 <= 200 * 1024) ? 4 : (STAGE_BYTES * 3 <= 200 * 1024) ? 3 : 2;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 256;
  static constexpr int ACC_COLS = MT * BN;
  static constexpr int TMEM_COLS = ACC_COLS <= 32 ? 32 : ACC_COLS <= 64 ? 64 : ACC_COLS <= 128 ? 128 : ACC_COLS <= 256 ? 256 : 512;
  static_assert(ACC_COLS <= 512, "accumulator does not fit TMEM");
  static_assert(STAGES * STAGE_BYTES >= 512 * 8 * 4, "stage memory is reused as the column-sum scratch");
  static constexpr int KTAB_BYTES = MT * 2 * 8 * 16;          // k-chunk decode table of this CTA's atoms
  static constexpr int ROWTAB_ROWS = 2048;                     // decoded row positions (rows_per_split <= this)
  static constexpr int SMEM_TOTAL = SMEM + KTAB_BYTES + ROWTAB_ROWS * 8;
};

template <class ALoad64, int MT, int BN>
__global__ void __launch_bounds__(kWgradThreads) wgrad_kernel(ALoad64 aload, const __nv_bfloat16* __restrict__ dy,
                                                              int ld_dy, int nrows, int rows_per_split,
                                                              int k_atoms_total, WgradEpi epi) {
  using Cfg = WgradCfg<MT, BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * Cfg::STAGES);
  const uint32_t tmem_ptr_addr = tmem_full_bar + 8u;
  int4* ktab = reinterpret_cast<int4*>(smem_gen + Cfg::STAGES * Cfg::STAGE_BYTES + 256);
  RowPos* rowtab = reinterpret_cast<RowPos*>(reinterpret_cast<uint8_t*>(ktab) + Cfg::KTAB_BYTES);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int group = warp >> 3;                    // producer group 0/1 (warp 16 = MMA warp)
  const int gtid = tid & 255;                     // thread index inside the producer group
  const int atom0 = blockIdx.x * MT * 2;          // first k' atom (64 wide) of this CTA
  const int n0 = blockIdx.y * BN;
  const int split = blockIdx.z;
  const int r_begin = split * rows_per_split;
  const int r_end = min(nrows, r_begin + rows_per_split);
  const int niter = (r_end > r_begin) ? (r_end - r_begin + 63) / 64 : 0;
  const int natoms = min(MT * 2, k_atoms_total - atom0);  // atoms that exist (the rest are never stored)
  const bool want_bias = (epi.bias_out != nullptr) && (blockIdx.x == 0) && (Cfg::B_ATOMS == 1);
  const bool use_rowtab = ALoad64::kNeedsTable && (rows_per_split <= Cfg::ROWTAB_ROWS);

  if (tid == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(full_bar(s), kProducerThreads);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  }
  pdl_wait();
  pdl_trigger();
  if constexpr (ALoad64::kNeedsTable) {
    // decode tables, built once per CTA: K-chunk -> tap offset (k' chunks are global: atom0*8 + ...),
    // and row -> (image, origin) for every row of this split (no integer division in the stage loop)
    aload.build_table(ktab, natoms * 8, tid, kWgradThreads, atom0 * 8);   // indexed by the CTA-local chunk
    if (use_rowtab)
      for (int lr = tid; lr < niter * 64; lr += kWgradThreads) rowtab[lr] = conv_row_pos(aload.g, r_begin + lr);
  }
  if (warp == kWgradProducerWarps) tmem_alloc(tmem_ptr_addr, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));

  float csum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (warp < kWgradProducerWarps) {
    aload.init_thread(ktab, gtid);
    for (int it = group; it < niter; it += 2) {
      const int s = it % Cfg::STAGES;
      const uint32_t ph = (it / Cfg::STAGES) & 1;
      mbar_wait(empty_bar(s), ph ^ 1, 4);
      const uint32_t a_tile = smem_base + s * Cfg::STAGE_BYTES;
      const uint32_t b_tile = a_tile + Cfg::A_BYTES;
      const int row0 = r_begin + it * 64;
      // rows beyond r_end must contribute zero: the loaders zero rows >= their nrows, and the
      // split boundary is enforced on the dY side (zero rows => zero products).
      if (use_rowtab) aload.prepare_tab(rowtab, it * 64, gtid);
      else aload.prepare(row0, gtid);
      if constexpr (Cfg::B_ATOMS == 1) {
        // dY goes through registers (its column sums are the bias gradient); its loads are issued first so
        // their latency overlaps the issue of the A-operand cp.async gathers
        constexpr int CH = Cfg::ROWB / 16;
        constexpr int NB = (64 * CH + 255) / 256;
        uint4 bv[NB];
#pragma unroll
        for (int q = 0; q < NB; ++q) {
          int i = gtid + q * 256;
          int r = i / CH, c = i % CH;
          bv[q] = make_uint4(0, 0, 0, 0);
          if (i < 64 * CH && row0 + r < r_end)
            bv[q] = __ldg(reinterpret_cast<const uint4*>(dy + (long)(row0 + r) * ld_dy + n0 + c * 8));
        }
        for (int at = 0; at < natoms; ++at)
          aload.fill_async(a_tile + at * 8192, ALoad64::kNeedsTable ? at : atom0 + at, gtid);
#pragma unroll
        for (int q = 0; q < NB; ++q) {
          int i = gtid + q * 256;
          if (i < 64 * CH) {
            int r = i / CH, c = i % CH;
            uint4 v = bv[q];
            st_shared_v4(b_tile + swz_off<Cfg::ROWB>(r, c), v);
            if (want_bias) {
              csum[0] += bf16_lo(v.x); csum[1] += bf16_hi(v.x); csum[2] += bf16_lo(v.y); csum[3] += bf16_hi(v.y);
              csum[4] += bf16_lo(v.z); csum[5] += bf16_hi(v.z); csum[6] += bf16_lo(v.w); csum[7] += bf16_hi(v.w);
            }
          }
        }
        fence_proxy_async();
      } else {
        for (int at = 0; at < natoms; ++at)
          aload.fill_async(a_tile + at * 8192, ALoad64::kNeedsTable ? at : atom0 + at, gtid);
#pragma unroll
        for (int at = 0; at < Cfg::B_ATOMS; ++at)
          fill_dense_async<64, 128>(b_tile + at * 8192, dy, ld_dy, row0, r_end, n0 + at * 64, gtid);
      }
      cp_async_mbar_arrive(full_bar(s));
      mbar_arrive(full_bar(s));
    }
  }
  if (warp < kProducerWarps) {
    // ===================== epilogue (group 0) =====================
    if (niter > 0) {
      mbar_wait(tmem_full_bar, 0, 5);
      tc_fence_after();
    }
    const int q = warp & 3;
    const int half = warp >> 2;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    constexpr int COLS_PER_WARP = BN / 2;
#pragma unroll 1
    for (int mt = 0; mt < MT; ++mt) {
      const int kprime = (atom0 + mt * 2) * 64 + q * 32 + (tid & 31);
      // K' is a multiple of 64, so this predicate is warp-uniform; the tcgen05.ld below is still
      // executed by every lane (it is .sync.aligned) and only the stores are guarded.
      const bool row_ok = kprime < epi.Kvalid;
      long orow;
      if (epi.mode == 1) {
        int hw = kprime / epi.fc_C;
        int c = kprime - hw * epi.fc_C;
        orow = (long)c * epi.fc_HW + hw;
      } else {
        orow = (long)split * epi.Kp + kprime;
      }
      constexpr int COFF = (BN >= 32) ? COLS_PER_WARP : 0;   // BN == 16: warps 0-3 own all 16 columns
      float* dst = epi.out + orow * epi.ldo + n0 + half * COFF;
      if constexpr (COLS_PER_WARP >= 32) {
#pragma unroll 1
        for (int c = 0; c < COLS_PER_WARP; c += 32) {
          uint32_t r[32];
          if (niter > 0) {
            tmem_ld32(lane_addr + mt * BN + half * COLS_PER_WARP + c, r);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = 0;
          }
          float4* d4 = reinterpret_cast<float4*>(dst + c);
          if (row_ok)
#pragma unroll
          for (int i = 0; i < 8; ++i)
            d4[i] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]),
                                __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
        }
      } else if (BN == 32 || half == 0) {
        uint32_t r16[16];
        if (niter > 0) {
          tmem_ld16(lane_addr + mt * BN + half * COFF, r16);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) r16[i] = 0;
        }
        float4* d4 = reinterpret_cast<float4*>(dst);
        if (row_ok)
#pragma unroll
        for (int i = 0; i < 4; ++i)
          d4[i] = make_float4(__uint_as_float(r16[4 * i]), __uint_as_float(r16[4 * i + 1]),
                              __uint_as_float(r16[4 * i + 2]), __uint_as_float(r16[4 * i + 3]));
      }
    }
    tc_fence_before();
  } else if (warp == kWgradProducerWarps) {
    // MMA issuer: converged warp, one elected lane issues (see common.cuh elect_one)
    constexpr uint32_t idesc = make_idesc_bf16(128, BN, 1, 1);
    constexpr uint32_t b_layout = swz_layout_type(Cfg::ROWB);
    constexpr uint32_t b_sbo = 8 * Cfg::ROWB;
    constexpr uint32_t b_kstep = 16 * Cfg::ROWB;
    const uint32_t tmem_u = make_uniform(tmem_base);
    for (int it = 0; it < niter; ++it) {
      const int s = it % Cfg::STAGES;
      const uint32_t ph = (it / Cfg::STAGES) & 1;
      mbar_wait(full_bar(s), ph, 6);
      fence_proxy_async();
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_tile = smem_base + s * Cfg::STAGE_BYTES;
        const uint32_t b_tile = a_tile + Cfg::A_BYTES;
#pragma unroll 1
        for (int mt = 0; mt < MT; ++mt) {
          // a missing second atom (odd atom count) is left unfilled: its D rows are never stored
          if (mt * 2 >= natoms) break;
          const uint32_t a0 = a_tile + (mt * 2) * 8192;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            uint64_t adesc = make_smem_desc(a0 + k * 2048, 8192, 1024, 2);
            uint64_t bdesc = make_smem_desc(b_tile + k * b_kstep, 8192, b_sbo, b_layout);
            umma_bf16(tmem_u + mt * BN, adesc, bdesc, idesc, (it > 0 || k > 0) ? 1u : 0u);
          }
        }
        umma_commit(empty_bar(s));
        if (it == niter - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
    }
  }
  __syncthreads();   // every MMA has completed (group 0 waited on tmem_full): stage memory is free
  if (warp == kWgradProducerWarps) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
  if (want_bias) {
    // bias gradient partial: column sums over this split's rows (deterministic order)
    float* scratch = reinterpret_cast<float*>(smem_gen);     // [512][8]
    if (warp < kWgradProducerWarps) {
#pragma unroll
      for (int e = 0; e < 8; ++e) scratch[tid * 8 + e] = csum[e];
    }
    __syncthreads();
    if (tid < BN) {
      constexpr int CH = Cfg::ROWB / 16;
      const int chunk = tid >> 3, e = tid & 7;
      float t = 0.f;
      for (int th = chunk; th < kWgradProducerWarps * 32; th += CH) t += scratch[th * 8 + e];
      epi.bias_out[(long)split * BN + tid] = t;
    }
  }
}

}  // namespace arl
