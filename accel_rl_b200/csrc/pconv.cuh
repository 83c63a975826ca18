// Patch-resident implicit-GEMM convolution tiles (sm_100a): the B200-native conv path.
//
// Every conv layer of the policy network is rewritten as a STRIDE-1 convolution with T x T taps over a
// "position grid" whose pixels are rows of 64 bf16 channels (128 bytes), split into P planes:
//   * a stride-s layer reads the space-to-depth(s) image of its (padded) input, so (k, s) becomes
//     T = ceil(k/s) taps over C*s*s channels (8x8/4 over 4 planes -> 2x2 over 64 ch; 4x4/2 over 32 ch ->
//     2x2 over 128 ch = 2 planes; 3x3/1 over 64 ch -> 3x3 over 64 ch);
//   * positions are numbered q = image*S + y*Wp + x (pitch Wp includes the zero padding column(s), S the
//     padding row), so tap (ty,tx) of output position q is input position q + ty*Wp + tx: a pure ROW SHIFT.
// One tile = 128 consecutive output positions.  Its input patch (128 + halo rows per plane, ~19 KB) is fetched
// ONCE by the TMA engine (cp.async.bulk, SASS UBLKCP) into a 128B-swizzled shared-memory ring; the im2col
// matrix is never built anywhere: each tap's A operand is the same patch addressed through a tcgen05 shared
// memory descriptor whose start address is advanced by `shift` rows.  L2->SM traffic drops from T*T x (im2col
// gather) to ~1.15 x the activation size, which is what lets the tensor pipe instead of LTS bound the tile.
// Global activations are stored PRE-SWIZZLED (16-byte chunk index XOR (position & 7)) so a plain 1-D bulk copy
// lands in the canonical UMMA SWIZZLE_128B layout (tiles start at multiples of 8 positions).
// Positions that are not real outputs (x >= Wo, y >= Ho: the padding columns/rows) are computed and dropped by
// the epilogue — 7-17 % extra MMA work instead of 4-9 x extra load traffic.
//
//   pconv_fwd_kernel<N>   conv forward and conv data-gradient (same kernel, different weight pack / epilogue)
//   pconv_wgrad_kernel<N> weight gradient: D[(tap,plane,ch)][co] = sum_q A[q+shift][ch] * dY[q][co], both operands
//                         MN-major straight out of the same resident patches
// Warp roles (persistent CTAs, one per SM): warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM alloc),
// warps 2..9 = epilogue (two warps per TMEM lane quadrant, each owning half of the N columns).
#pragma once
#include "common.cuh"

namespace arl {

constexpr int kPcThreads = 448;      // weight-gradient kernel: 10 warps + 4 converter warps (u8 first-layer input)
constexpr int kPcFwdThreads = 480;   // forward/dgrad kernel: + a second MMA-issuing warp (warp 10) + 4 converter warps (11..14)
constexpr int kPcMaxTaps = 16;

// First-layer input taken straight from the uint8 observations (the reference's own buffer layout, [img][C=4][H][W]):
// four converter warps load the image rows a tile touches with 16-byte global loads (two tiles ahead, in registers),
// expand them to the bf16 space-to-depth(4) patch and store it in the SWIZZLE_128B layout the MMAs read (cell (Y, X),
// channel = plane*16 + (y%4)*4 + x%4 — byte for byte what the frame kernel's bf16 mirror used to hold; u8 -> bf16 is
// exact).  No bf16 copy of the rollout exists any more: 2.18 GB of HBM and half of this layer's training reads are gone,
// and the frame kernel writes 66 KB per env-step instead of 200 KB.  (A first version staged the uint8 rows in shared
// memory with bulk copies: correct, but these tiles are bound by the shared-memory operand fetch of SS-mode MMAs, and the
// extra 23 KB of shared-memory traffic per tile doubled the kernel time — profiles/r2_u8_conv0.md.)
struct PcU8Src {
  const uint8_t* obs;     // nullptr: the patch comes from a bf16 grid (bulk copy), as for every other layer
  int H, W, Wc, Hc;       // image rows / columns, cells per row (W/4), cell rows (H/4)
  int rows_max;           // image rows per plane a tile can touch (multiple of 4)
  long img_bytes;         // C*H*W
};
constexpr int kU8Planes = 4;

// 4 packed u8 -> 4 bf16 (exact): byte b -> float(2^23 + b) - 2^23
ARL_DEVINL uint2 pc_u8x4_to_bf16x4(uint32_t w) {
  float f0 = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7650)) - 8388608.0f;
  float f1 = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7651)) - 8388608.0f;
  float f2 = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7652)) - 8388608.0f;
  float f3 = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7653)) - 8388608.0f;
  return make_uint2(pack_bf16x2(f0, f1), pack_bf16x2(f2, f3));
}

// One item = (cell row Yr, plane, pair of image rows 2j / 2j+1, 16-byte column group g): two 16-byte loads -> four
// 16-byte chunks of the patch (cells 4g .. 4g+3).  A converter thread (tid_c = 0..127) owns items tid_c + 128 k, k < 3
// (at most 9 cell rows x 4 planes x 2 x G <= 384 items); their decomposition is tile-independent and computed once.
struct PcU8Items {
  int goff[3];       // byte offset of the item's first row relative to image row 4*Y0 of plane 0: (pl*H + 4*Yr + 2*j)*W + 16*g
  int cell[3];       // Yr*Wc + 4*g: cell index relative to cell row Y0's first cell
  int chunk[3];      // 2*pl + j
  int yr[3];         // Yr (16: no such item)
};
struct PcU8Regs { uint4 r0[3], r1[3]; };

ARL_DEVINL PcU8Items pc_u8_items(const PcU8Src& u, int tid_c) {
  PcU8Items it;
  const int G = u.W >> 4;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int item = tid_c + 128 * k;
    const int g = item % G;
    int t = item / G;
    const int j = t & 1; t >>= 1;
    const int pl = t & 3;
    const int Yr = t >> 2;
    it.goff[k] = (pl * u.H + 4 * Yr + 2 * j) * u.W + 16 * g;
    it.cell[k] = Yr * u.Wc + 4 * g;
    it.chunk[k] = 2 * pl + j;
    it.yr[k] = (4 * Yr + 4 <= u.rows_max) ? Yr : 16;
  }
  return it;
}

// image rows of cells [c0, c0 + patch_rows) of image `img` -> registers (cells past the end of the image: zeros)
ARL_DEVINL void pc_u8_load(const PcU8Src& u, const PcU8Items& it, long img, int c0, int patch_rows, PcU8Regs& R) {
  const int Y0 = c0 / u.Wc;
  const int n_yr = (c0 - Y0 * u.Wc + patch_rows - 1) / u.Wc + 1;
  const uint8_t* base = u.obs + img * u.img_bytes + (long)(4 * Y0) * u.W;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    R.r0[k] = make_uint4(0u, 0u, 0u, 0u);
    R.r1[k] = R.r0[k];
    if (it.yr[k] < n_yr && Y0 + it.yr[k] < u.Hc) {
      R.r0[k] = __ldg(reinterpret_cast<const uint4*>(base + it.goff[k]));
      R.r1[k] = __ldg(reinterpret_cast<const uint4*>(base + it.goff[k] + u.W));
    }
  }
}

// registers -> bf16 patch rows [0, patch_rows) of `stage` (1024-aligned; c0 % 8 == 0, so the chunk swizzle by (row & 7)
// equals the swizzle by (cell & 7))
ARL_DEVINL void pc_u8_store(const PcU8Src& u, const PcU8Items& it, int c0, int patch_rows, const PcU8Regs& R, uint32_t stage) {
  const int Y0 = c0 / u.Wc;
  const int n_yr = (c0 - Y0 * u.Wc + patch_rows - 1) / u.Wc + 1;
  const int base = Y0 * u.Wc - c0;                         // patch row of cell row Y0's first cell (<= 0)
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (it.yr[k] >= n_yr) continue;
    const uint32_t w0[4] = {R.r0[k].x, R.r0[k].y, R.r0[k].z, R.r0[k].w}, w1[4] = {R.r1[k].x, R.r1[k].y, R.r1[k].z, R.r1[k].w};
    const int rb = base + it.cell[k];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int r = rb + b;
      if (r >= 0 && r < patch_rows) {
        const uint2 lo = pc_u8x4_to_bf16x4(w0[b]), hi = pc_u8x4_to_bf16x4(w1[b]);
        st_shared_v4(stage + swz_off<128>((uint32_t)r, (uint32_t)it.chunk[k]), make_uint4(lo.x, lo.y, hi.x, hi.y));
      }
    }
  }
}

// The converter warps' tile loop, shared by the forward and the weight-gradient kernel: tile `it` of this CTA goes to
// patch slot it % stages once `empty(s)` says the MMAs that read the slot have completed; the loads of tiles it+1 and it+2
// are already in flight (two register sets, loop unrolled by two).  Image of a tile: idx[idx_base + tile / tiles_per_img]
// (lane l looks up the image of the CTA's l-th tile, one round of latency per 32 tiles, as the TMA producer does).
template <class FullBar, class EmptyBar>
ARL_DEVINL void pc_u8_converter_loop(const PcU8Src& u, const int* idx, long idx_base, int tiles_per_img, int ntiles,
                                     int patch_rows, int stages, uint32_t stage0, uint32_t stage_bytes, FullBar full_bar,
                                     EmptyBar empty_bar, int tid_c, int lane, int code) {
  const PcU8Items items = pc_u8_items(u, tid_c);
  const int grid = (int)gridDim.x;
  int img_l = 0;
  auto image_of = [&](int it_, int tile_) -> long {
    (void)tile_;
    return __shfl_sync(0xffffffffu, img_l, it_ & 31);
  };
  auto refresh = [&](int it_, int tile_) {
    if ((it_ & 31) == 0) {
      const int tl = tile_ + lane * grid;
      const int bl = min(tl, ntiles - 1) / tiles_per_img;
      img_l = idx ? idx[idx_base + bl] : bl;
    }
  };
  auto c0_of = [&](int tile_) { return (tile_ - (tile_ / tiles_per_img) * tiles_per_img) * 128; };
  PcU8Regs A, B;
  int tile = blockIdx.x;
  // (tiles it and it+1 never straddle a refresh boundary in a way that matters: img_l holds 32 consecutive tiles)
  refresh(0, tile);
  if (tile < ntiles) pc_u8_load(u, items, image_of(0, tile), c0_of(tile), patch_rows, A);
  if (tile + grid < ntiles) pc_u8_load(u, items, image_of(1, tile + grid), c0_of(tile + grid), patch_rows, B);
  for (int it = 0; tile < ntiles; it += 2, tile += 2 * grid) {
    {
      const int s = it % stages;
      mbar_wait(empty_bar(s), ((it / stages) & 1) ^ 1, code);
      pc_u8_store(u, items, c0_of(tile), patch_rows, A, stage0 + s * stage_bytes);
      fence_proxy_async();                             // generic-proxy stores -> visible to tcgen05.mma
      __syncwarp();
      if (lane == 0) mbar_arrive(full_bar(s));
      const int t2 = tile + 2 * grid;
      refresh(it + 2, t2);                             // (it + 2) % 32 == 0 only for even it: handled here
      if (t2 < ntiles) pc_u8_load(u, items, image_of(it + 2, t2), c0_of(t2), patch_rows, A);
    }
    const int tile1 = tile + grid;
    if (tile1 >= ntiles) break;
    {
      const int s = (it + 1) % stages;
      mbar_wait(empty_bar(s), (((it + 1) / stages) & 1) ^ 1, code);
      pc_u8_store(u, items, c0_of(tile1), patch_rows, B, stage0 + s * stage_bytes);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(full_bar(s));
      const int t3 = tile1 + 2 * grid;
      if (t3 < ntiles) pc_u8_load(u, items, image_of(it + 3, t3), c0_of(t3), patch_rows, B);
    }
  }
}

struct PcOut {
  int mode;                   // 0: acc*scale + bias, ReLU -> bf16   1: acc masked by act > 0 -> bf16   2: mask + unfold
  float scale;
  const float* bias;          // [N] (mode 0)
  __nv_bfloat16* dst;
  const __nv_bfloat16* act;   // mask source (modes 1, 2)
  long dst_plane_stride;      // elements between 64-channel planes of dst
  long act_plane_stride;
  int ds_shift, us_shift;     // log2(ds), log2(us) (both are 1 or 2)
  int dS, dWp, dHc, dpad, ds; // destination grid: positions per image, pitch, rows, padding added to (y, x), space-to-depth
  int act_off;                // modes 1, 2: the mask is the forward INPUT of this layer, i.e. it lives in the source grid:
                              // position = source position + act_off, channel = output column
  int swz;                    // 1: planes of 128-byte rows, chunk-swizzled   0: dense rows of dense_ld elements
  int dense_ld;
  int fc_rows;                // > 0: destination is act_fc [pixel][fc_rows images][64] (fcgemm.cuh): position = pixel*fc_rows + image
  // mode 2 (data gradient of a stride-`us` layer -> gradient w.r.t. the PIXELS of the layer below, stored
  // position-aligned for that layer's wgrad): cell (y, x) sub-pixel (py, px) -> pixel (us*y + py - upad, ...)
  int uH, uW, uWp, uS, us, upad, uC;
};

struct PcParams {
  const __nv_bfloat16* src;   // plane 0, position 0
  long src_plane_stride;      // elements
  int planes;
  const int* idx;             // optional image gather (per-image tiling only): image b reads idx[off*nb + b]
  const int* idx_off;
  int nb;
  int S, Wp, Ho, Wo;          // source positions per image, pitch; valid output extent
  int tiles_per_img;          // > 0: tiles never cross images (layer 0, gatherable); 0: continuous over the batch
  uint32_t magic_S, magic_Wp, magic_tpi;   // floor(2^32/d)+1: q = umulhi(n, magic) replaces the integer divisions of the row decode
  int n_img, ntiles;
  int ntaps;
  int shift[kPcMaxTaps];      // row shift of each tap
  int load_rows;              // 128 + halo, multiple of 8
  const __nv_bfloat16* w;     // [ntaps*planes][N][64] bf16, rows chunk-swizzled (pack_weights_kernel PK_PCONV*)
  int stages;
  PcU8Src u8;                 // first layer: uint8 observations instead of `src` (per-image tiling only)
  PcOut out;
  // ---- wide form (pconv_fwd_kernel<N, TW>, TW = taps per filter row > 1): the taps of a row ride on the N axis ----
  //   D_tx[r][co] = sum_ty sum_k A[r + ty*Wp][k] * W[ty][tx][k][co]      one MMA per (ty, plane, k16): M = 128, N = TW*N,
  //   out[q][co]  = sum_tx D_tx[q + tx][co]                              B = the TW weight tiles of the row, contiguous
  // so one 4 KB A slice feeds TW x the math (issue cycles per the measured 25 + (4 KB + N*32 B)/128 model: 9 taps x N = 64:
  // 36 x 73 -> 12 x 105), and the epilogue adds the column groups one row apart (warp shuffles; the first rows of the next
  // lane quadrant through shared memory).  Rows 128-TW+1.. of a tile have no partner rows: tiles advance by kPcWideRows = 120
  // (a multiple of the 8-row swizzle period) and only rows < 120 are stored.
  int n_ty;                   // wide: filter rows; A row shift of row ty = shift[ty * TW]
  int src_evict_last;         // 1: the input patches are read again soon (layer 0 in training: the weight gradient re-reads
                              // the same gathered rows ~100 us later) -> ask the L2 to keep them in front of the streamed
                              // activations in between
};
constexpr int kPcWideRows = 120;
constexpr int kPcExFloats = 2 /*parity*/ * 2 /*halves*/ * 4 /*quadrants*/ * 2 /*lanes*/ * 3 /*tx*/ * 16;

__host__ __device__ inline int pc_fwd_smem(int N, int ntaps, int planes, int load_rows, int stages, int raw_stage_bytes = 0,
                                           int wide = 0) {
  return ntaps * planes * N * 128 + stages * (planes * load_rows * 128 + raw_stage_bytes) + 1024 /*align*/ + 256 /*barriers*/ + N * 4 +
         (wide ? 6144 + 16 : 0) /* cross-quadrant exchange rows: kPcExFloats floats */;
}

// n / d for d >= 1 with magic = floor(2^32/d) + 1 (exact for n < 2^32 / d; every use here is < 2^24)
ARL_DEVINL uint32_t pc_fdiv(uint32_t n, uint32_t d, uint32_t magic) { return d == 1 ? n : __umulhi(n, magic); }
__host__ inline uint32_t pc_magic(uint32_t d) { return d <= 1 ? 0u : (uint32_t)(0x100000000ull / d) + 1u; }

// dReLU mask of 32 output columns: forward activation at (source position apos, channel col0..col0+31) of the
// chunk-swizzled source grid (issued BEFORE the accumulator wait, so the loads overlap the MMAs)
ARL_DEVINL void pc_load_mask32(const PcOut& o, int col0, long apos, uint4 (&m)[4]) {
  const __nv_bfloat16* arow = o.act + (col0 >> 6) * o.act_plane_stride + apos * 64;
  const int a7 = (int)(apos & 7), ac0 = (col0 & 63) >> 3;
#pragma unroll
  for (int j = 0; j < 4; ++j) m[j] = __ldg(reinterpret_cast<const uint4*>(arow + ((ac0 + j) ^ a7) * 8));
}

// 32 (or 16) accumulator columns -> 16 (8) packed bf16 pairs: bias + ReLU (forward) or dReLU mask (gradient)
template <int NC>
ARL_DEVINL void pc_finish(const uint32_t* r, const float* bias_r, float scale, bool fwd, const uint4* m, uint32_t* packed) {
#pragma unroll
  for (int i = 0; i < NC / 2; ++i) {
    float lo = __uint_as_float(r[2 * i]), hi = __uint_as_float(r[2 * i + 1]);
    if (fwd) {
      lo = fmaxf(lo * scale + bias_r[2 * i], 0.f);
      hi = fmaxf(hi * scale + bias_r[2 * i + 1], 0.f);
    } else {
      const uint4 mm = m[i >> 2];
      const uint32_t mw = (i & 3) == 0 ? mm.x : (i & 3) == 1 ? mm.y : (i & 3) == 2 ? mm.z : mm.w;
      lo = bf16_lo(mw) > 0.f ? lo : 0.f;
      hi = bf16_hi(mw) > 0.f ? hi : 0.f;
    }
    packed[i] = pack_bf16x2(lo, hi);
  }
}

// NC/8 packed 16-byte chunks -> destination position dpos, channel ch0 of the destination cell
template <int NC>
ARL_DEVINL void pc_store(const PcOut& o, const uint32_t* packed, long dpos, int ch0) {
#ifdef ARL_DBG_NOSTORE
  if (packed[0] != 0x12345678u) return;      // timing experiment only: (almost) never store
#endif
  const int plane = ch0 >> 6, chunk0 = (ch0 & 63) >> 3;
  if (o.swz) {
    __nv_bfloat16* drow = o.dst + plane * o.dst_plane_stride + dpos * 64;
    const int x7 = (int)(dpos & 7);
#pragma unroll
    for (int j = 0; j < NC / 8; ++j)
      *reinterpret_cast<uint4*>(drow + ((chunk0 + j) ^ x7) * 8) =
          make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
  } else {
    __nv_bfloat16* drow = o.dst + dpos * o.dense_ld + ch0;
#pragma unroll
    for (int j = 0; j < NC / 8; ++j)
      *reinterpret_cast<uint4*>(drow + j * 8) = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
  }
}

// mode 2: 32 masked columns = one sub-pixel block (py, px) of a space-to-depth cell -> that pixel's row in the
// position-aligned gradient buffer of the layer below ([uS positions][uC = 32 channels], 64-byte rows chunk-swizzled
// like a SWIZZLE_64B tile)
ARL_DEVINL void pc_store_unfold(const PcOut& o, const uint32_t* packed, int b, int y, int x, int sub) {
  const int py = sub >> o.us_shift, px = sub & (o.us - 1);
  const int yy = o.us * y + py - o.upad, xx = o.us * x + px - o.upad;
  if ((unsigned)yy >= (unsigned)o.uH || (unsigned)xx >= (unsigned)o.uW) return;
  const long upos = (long)b * o.uS + yy * o.uWp + xx;
  __nv_bfloat16* drow = o.dst + upos * 32;
  const int sw = (int)((upos >> 1) & 3);
#pragma unroll
  for (int j = 0; j < 4; ++j)
    *reinterpret_cast<uint4*>(drow + (j ^ sw) * 8) = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
}

#define ARL_TP(slot) do { if (N == 32) ARL_T(slot); } while (0)
template <int N, int TW = 1>
__global__ void __launch_bounds__(kPcFwdThreads, 1) pconv_fwd_kernel(const __grid_constant__ PcParams p) {
  static_assert(N == 32 || N == 64 || N == 128, "tile width");
  static_assert(TW >= 1 && TW <= 3 && TW * N <= 256, "wide form: TW taps of a row on the N axis");
  constexpr int NW = TW * N;                                          // accumulator columns per tile
  constexpr int TCOLS = (2 * NW <= 64) ? 64 : (2 * NW <= 128) ? 128 : (2 * NW <= 256) ? 256 : 512;
  constexpr int TROWS = (TW > 1) ? kPcWideRows : 128;                  // positions a tile advances by / stores
  if ((int)blockIdx.x >= p.ntiles) return;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int nblk = p.ntaps * p.planes;
  const uint32_t w_base = smem_base;                                  // nblk tiles of [N x 128 B]
  const uint32_t stage_bytes = (uint32_t)p.planes * p.load_rows * 128;
  const uint32_t a_base = w_base + nblk * (N * 128);
  const bool u8 = p.u8.obs != nullptr;
  const uint32_t bar_base = a_base + p.stages * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (8 + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (16 + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (18 + b); };
  const uint32_t wfull_bar = bar_base + 8u * 20;
  const uint32_t tmem_ptr_addr = bar_base + 8u * 21;
  float* bias_s = reinterpret_cast<float*>(smem_raw + (bar_base + 256u - smem_u32(smem_raw)));
  float* ex_s = bias_s + ((N + 3) & ~3);                              // wide form: exchange rows (kPcExFloats floats)

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), u8 ? 4 : 1);      // u8 mode: the four converter warps fill the patch
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), 8);
    }
    mbar_init(wfull_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_addr, TCOLS);
  pdl_wait();                                  // everything below may read what the previous kernel wrote
  pdl_trigger();
  if (p.out.mode == 0)
    for (int i = tid; i < N; i += kPcFwdThreads) bias_s[i] = p.out.bias[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));

  if (warp == 0) {
    // ===================== TMA producer (converged warp, one elected lane issues) =====================
    if (elect_one()) {
      mbar_arrive_expect_tx(wfull_bar, (uint32_t)nblk * N * 128);
      // one bulk copy per (tap, plane) weight tile; wide form: the TW tiles of a (filter row, plane) sit next to each other
      // = the N = TW*N B operand of one MMA
      for (int b = 0; b < nblk; ++b) {
        int slot = b;
        if constexpr (TW > 1) {
          const int t = b / p.planes, pl = b - t * p.planes, ty = t / TW, tx = t - ty * TW;
          slot = (ty * p.planes + pl) * TW + tx;
        }
        bulk_g2s(w_base + slot * (N * 128), p.w + (long)b * N * 64, N * 128, wfull_bar);
      }
    }
    __syncwarp();
    // gathered images: the index lookups (two dependent global loads, ~1.5 us) are hoisted out of the tile loop —
    // lane l fetches the image of this CTA's l-th tile, one round of latency per 32 tiles instead of one per tile
    const long idx_base = (p.idx && p.idx_off) ? (long)p.idx_off[0] * p.nb : 0;
    int img_l = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int s = it % p.stages;
      const uint32_t ph = (it / p.stages) & 1;
      long pos0;
      if (p.tiles_per_img > 0) {
        if ((it & 31) == 0) {
          const int tl = tile + lane * (int)gridDim.x;
          const int bl = min(tl, p.ntiles - 1) / p.tiles_per_img;
          img_l = p.idx ? p.idx[idx_base + bl] : bl;
        }
        const int b = tile / p.tiles_per_img;
        const int j = tile - b * p.tiles_per_img;
        const long img = __shfl_sync(0xffffffffu, img_l, it & 31);
        pos0 = img * p.S + (long)j * TROWS;
      } else {
        pos0 = (long)tile * TROWS;
      }
      if (u8) break;                           // the converter warps build the patches: nothing to copy
      mbar_wait(empty_bar(s), ph ^ 1, 21);
#ifndef ARL_DBG_EPI
      if (lane == 0) ARL_TP(it * 6 + 0);
#endif
      if (elect_one()) {
        mbar_arrive_expect_tx(full_bar(s), stage_bytes);
        const uint32_t dst = a_base + s * stage_bytes;
        const uint64_t pol = p.src_evict_last ? l2_policy_evict_last() : 0ull;
        for (int pl = 0; pl < p.planes; ++pl)
          bulk_g2s_hint(dst + pl * p.load_rows * 128, p.src + pl * p.src_plane_stride + pos0 * 64, p.load_rows * 128, full_bar(s),
                        pol);
      }
      __syncwarp();
#ifndef ARL_DBG_EPI
      if (lane == 0) ARL_TP(it * 6 + 1);
#endif
    }
  } else if (warp == 1 || warp == 10) {
    // ===================== MMA issuers (converged warps, one elected lane issues) =====================
    // Two issuing warps alternate tiles (warp 1: even tiles / accumulator 0, warp 10: odd tiles / accumulator 1).
    // A tile's tcgen05.mma burst blocks its issuer for ~65-75 cycles per instruction (operand fetch bound, measured)
    // and the barrier waits in front of it cost ~400 cycles: with two issuers the next tile's waits are already
    // done when the tensor pipe frees up.
    const int which = (warp == 10) ? 1 : 0;
    const uint32_t tmem_u = make_uniform(tmem_base);
    constexpr uint32_t idesc = make_idesc_bf16(128, NW, 0, 0);
    // descriptor high words are constant (SBO 1024, version 1, SWIZZLE_128B); only the 14-bit address field moves
    const uint64_t desc0 = make_smem_desc(0, 16, 1024, 2);
    const uint32_t desc_hi32 = (uint32_t)(desc0 >> 32), desc_lo_flags = (uint32_t)desc0;   // LBO field lives in the low word
    auto mk = [&](uint32_t lo) { return ((uint64_t)desc_hi32 << 32) | (uint64_t)lo; };
    mbar_wait(wfull_bar, 0, 22);
    for (int it = which, tile = blockIdx.x + which * gridDim.x; tile < p.ntiles; tile += 2 * gridDim.x, it += 2) {
      const int s = it % p.stages;
      const uint32_t ph = (it / p.stages) & 1;
      const int acc = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      mbar_wait(tempty_bar(acc), aph ^ 1, 23);
      mbar_wait(full_bar(s), ph, 24);
      if (lane == 0) ARL_TP(it * 6 + 2);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_stage = a_base + s * stage_bytes;
        const uint32_t d_tmem = tmem_u + acc * NW;
        uint32_t first = 0;
        uint32_t b_lo = (w_base >> 4) | desc_lo_flags;        // shared memory addresses are < 2^18: no carry into the flags
        const int n_a = (TW > 1) ? p.n_ty : p.ntaps;          // wide: one (shifted) A view per filter row
        for (int t = 0; t < n_a; ++t) {
          uint32_t a_lo = ((a_stage + p.shift[t * TW] * 128) >> 4) | desc_lo_flags;
          for (int pl = 0; pl < p.planes; ++pl) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_bf16(d_tmem, mk(a_lo + 2 * k), mk(b_lo + 2 * k), idesc, first);
              first = 1;
            }
            a_lo += (uint32_t)p.load_rows * 8;     // next plane: load_rows * 128 bytes
            b_lo += NW * 8;                        // next weight tile (group): NW * 128 bytes
          }
        }
        umma_commit(empty_bar(s));
        umma_commit(tfull_bar(acc));
      }
      __syncwarp();
      if (lane == 0) ARL_TP(it * 6 + 3);
    }
  } else if (warp >= 11) {
    // ===================== converter warps (u8 first-layer input) =====================
    if (u8) {
      const long idx_base = (p.idx && p.idx_off) ? (long)p.idx_off[0] * p.nb : 0;
      pc_u8_converter_loop(p.u8, p.idx, idx_base, p.tiles_per_img, p.ntiles, p.load_rows, p.stages, a_base, stage_bytes,
                           full_bar, empty_bar, tid - 11 * 32, lane, 28);
    }
  } else if (warp >= 2 && warp < 10) {
    // ===================== epilogue warps =====================
    const PcOut& o = p.out;
    const int q = warp & 3;                   // TMEM lane quadrant
    const int h = (warp - 2) >> 2;            // column half
    constexpr int HC = N / 2;                 // columns per warp
    constexpr int NC = HC < 32 ? HC : 32;     // columns per register block
    constexpr int NB = HC / NC;
    const bool fwd = (o.mode == 0);
    const uint32_t row = q * 32 + lane;
    // forward: this thread's bias slice lives in registers for the whole kernel (N <= 64 on the forward path)
    float bias_r[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) bias_r[i] = (fwd && NB == 1) ? bias_s[h * HC + i] : 0.f;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t aph = (it >> 1) & 1;
#ifdef ARL_DBG_EPI
      if (warp == 2 && lane == 0) ARL_TP(it * 6 + 0);
#endif
      // decode this thread's output row (and prefetch its mask) while the MMAs run
      uint32_t b, pl_;
      if (p.tiles_per_img > 0) {
        b = pc_fdiv((uint32_t)tile, (uint32_t)p.tiles_per_img, p.magic_tpi);
        pl_ = ((uint32_t)tile - b * p.tiles_per_img) * TROWS + row;
      } else {
        const uint32_t qq = (uint32_t)tile * TROWS + row;
        b = pc_fdiv(qq, (uint32_t)p.S, p.magic_S);
        pl_ = qq - b * p.S;
      }
      const uint32_t y = pc_fdiv(pl_, (uint32_t)p.Wp, p.magic_Wp), x = pl_ - y * p.Wp;
      const bool valid = (int)b < p.n_img && (int)y < p.Ho && (int)x < p.Wo && (int)row < TROWS;
      const int Y = y + o.dpad, X = x + o.dpad;
      const int cy = Y >> o.ds_shift, cx = X >> o.ds_shift;
      const int sub = ((Y & (o.ds - 1)) << o.ds_shift) | (X & (o.ds - 1));
      const long dpos = o.fc_rows ? (long)(cy * o.dWp + cx) * o.fc_rows + b : (long)b * o.dS + cy * o.dWp + cx;
      const bool store = valid && (o.mode == 2 || (cy < o.dHc && cx < o.dWp));
      uint4 mk[NB][4];
      if (!fwd && store) {
        const long apos = (long)b * p.S + pl_ + o.act_off;
#pragma unroll
        for (int c = 0; c < NB; ++c) pc_load_mask32(o, h * HC + c * 32, apos, mk[c]);
      }
#ifdef ARL_DBG_EPI
      if (warp == 2 && lane == 0) ARL_TP(it * 6 + 1);
#endif
      mbar_wait(tfull_bar(acc), aph, 25);
      if (warp == 2 && lane == 0) ARL_TP(it * 6 + 4);
      tc_fence_after();
      if constexpr (TW > 1) {
        // ---- wide form: 16 output columns at a time; out[row] = D_0[row] + D_1[row + 1] (+ D_2[row + 2]) ----
        static_assert(HC % 16 == 0, "wide epilogue works on 16-column chunks");
        constexpr int NCH = HC / 16;
        const uint32_t tw = tmem_base + ((uint32_t)(q * 32) << 16) + acc * NW + h * HC;
#pragma unroll
        for (int cc = 0; cc < NCH; ++cc) {
          uint32_t d[TW][16];
#pragma unroll
          for (int tx = 0; tx < TW; ++tx) tmem_ld16(tw + tx * N + cc * 16, d[tx]);
          tmem_ld_wait();
          if (cc == NCH - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));  // accumulator is in registers: the next tile's MMAs may start
          }
          // Lane L needs D_tx[L + tx]: lanes L + tx < 32 get it from lane L + tx of this warp, the last tx lanes from rows
          // 0 .. tx-1 of the NEXT lane quadrant.  Lanes 0 .. TW-2 export their rows, import the next quadrant's same rows, and
          // one ROTATING shuffle serves everybody: its sources 0 .. tx-1 (never read by an in-warp partner) hand out the
          // imported rows.  One barrier per chunk among the four warps of this column half; buffers alternate by parity.
          const int par = (it * NCH + cc) & 1;
          float* ex = ex_s + ((par * 2 + h) * 4) * (2 * 3 * 16);
          if (lane < TW - 1) {
            float4* e4 = reinterpret_cast<float4*>(ex + (q * 2 + lane) * (3 * 16));
#pragma unroll
            for (int tx = 1; tx < TW; ++tx)
#pragma unroll
              for (int i = 0; i < 4; ++i)
                e4[tx * 4 + i] = make_float4(__uint_as_float(d[tx][4 * i]), __uint_as_float(d[tx][4 * i + 1]),
                                             __uint_as_float(d[tx][4 * i + 2]), __uint_as_float(d[tx][4 * i + 3]));
          }
          asm volatile("bar.sync %0, 128;" ::"r"(1 + h) : "memory");
          if (lane < TW - 1) {
            // (lane j overwrites its d[tx] for tx > j only: those were exported above and have no in-warp reader, while
            // d[tx], tx <= j, is what lane j - tx reads)
            const float4* n4 = reinterpret_cast<const float4*>(ex + (((q + 1) & 3) * 2 + lane) * (3 * 16));
#pragma unroll
            for (int tx = 1; tx < TW; ++tx)
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                if (tx <= lane) continue;
                const float4 v = (q < 3) ? n4[tx * 4 + i] : make_float4(0.f, 0.f, 0.f, 0.f);
                d[tx][4 * i] = __float_as_uint(v.x); d[tx][4 * i + 1] = __float_as_uint(v.y);
                d[tx][4 * i + 2] = __float_as_uint(v.z); d[tx][4 * i + 3] = __float_as_uint(v.w);
              }
          }
          uint32_t outw[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float acc_f = __uint_as_float(d[0][i]);
#pragma unroll
            for (int tx = 1; tx < TW; ++tx)
              acc_f += __uint_as_float(__shfl_sync(0xffffffffu, d[tx][i], (lane + tx) & 31));
            outw[i] = __float_as_uint(acc_f);
          }
          if (store) {
            uint32_t packed[8];
            pc_finish<16>(outw, bias_r + (NC > 16 ? cc * 16 : 0), o.scale, fwd, &mk[cc >> 1][2 * (cc & 1)], packed);
            pc_store<16>(o, packed, dpos, sub * N + h * HC + cc * 16);
          }
        }
        if (warp == 2 && lane == 0) ARL_TP(it * 6 + 5);
        continue;
      }
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * N + h * HC;
      uint32_t r[NB][NC];
      if constexpr (NC == 16) {
        tmem_ld16(taddr, r[0]);
      } else {
#pragma unroll
        for (int c = 0; c < NB; ++c) tmem_ld32(taddr + c * 32, r[c]);
      }
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));      // accumulator is in registers: the next tile's MMAs may start
      if (store) {
#pragma unroll
        for (int c = 0; c < NB; ++c) {
          uint32_t packed[NC / 2];
          pc_finish<NC>(r[c], bias_r, o.scale, fwd, mk[c], packed);
          if (o.mode == 2) pc_store_unfold(o, packed, (int)b, (int)y, (int)x, (h * HC + c * 32) >> 5);
          else pc_store<NC>(o, packed, dpos, sub * N + h * HC + c * NC);
        }
      }
      if (warp == 2 && lane == 0) ARL_TP(it * 6 + 5);
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TCOLS);
  }
}

// ---------------------------------------------------------------------------
// pconv_wgrad: D[(tap, plane, ch)][co] = sum over output positions q of A[q + shift(tap)][plane, ch] * dY[q][co]
// Both operands are MN-major views of resident patches: A = the layer's forward input patch (the same bulk copy
// as the forward tile), dY = 128 rows of the output-gradient grid (zero at padding positions, so the positions the
// forward pass drops contribute nothing).  One MMA covers 128 K' rows = TWO 64-channel blocks: (tap, plane 0/1) for
// two-plane layers, (tap 2m, tap 2m+1) for single-plane layers — the second block is the same patch seen through
// a different row shift, expressed as the descriptor's MN-atom stride (LBO).  Each persistent CTA accumulates all
// its tiles in TMEM (mt*N columns) and writes ONE fp32 partial [K'][N] at the end; finalize_grads_kernel sums the
// CTAs in a fixed order.  Four otherwise idle warps add up dY's columns from shared memory = the bias gradient.
// ---------------------------------------------------------------------------
constexpr int kPcMaxBlk = 2 * kPcMaxTaps;

struct PcWgradParams {
  const __nv_bfloat16* a;      // forward input grid, plane 0 position 0
  long a_plane_stride;
  int planes;
  const int* idx;              // per-image gather of A (layer 0)
  const int* idx_off;
  int nb;
  int S;                       // positions per image (per-image tiling)
  int tiles_per_img;           // > 0: per-image tiling; 0: continuous
  int ntiles;
  int nblk;                    // 64-channel K' blocks = ntaps * planes
  int blk_off[kPcMaxBlk];      // byte offset of block b's first row inside the A stage: (plane*a_rows + shift)*128
  int a_rows;                  // rows per plane in the A stage (128 + halo, multiple of 8)
  const __nv_bfloat16* dy;     // output-gradient grid [positions][N] (chunk-swizzled rows of N*2 bytes)
  int dy_off;                  // dY row of output position q is q + dy_off (padding offset of the gradient grid)
  int dy_rows;                 // rows per dY stage: 128 + 8
  float* partial;              // [gridDim.x][nblk*64][N]
  float* bias_partial;         // [gridDim.x][N]
  int stages;
  int a_evict_first;           // 1: last reader of the gathered input rows (layer 0): let the L2 drop them first
  // ---- wide form (wide = 1): the taps of one row, tx = T-1 .. 0, ride on the N axis -----------------------------------
  //   D_i[(a, ch)][(j, co)] = sum_q'' A[q'' + ty*Wp][plane, ch] * dY[q'' - tx][co]   (= dW[ty][tx][plane, ch][co], q = q'' - tx)
  // B = the SAME dY patch seen through `nb_atoms` MN-atoms one row apart (descriptor LBO = one dY row), so one MMA is
  // M = 128 x N = nb_atoms*Cout and reads its 4 KB A slice once for T taps: 105 / 89 / 73 issue cycles for 96 / 64 / 32
  // cycles of math instead of 73 / 65 for 32 / 16 (pconv.cuh header).  MMA i of a K step takes A atoms (a = 0, 1) at byte
  // offsets a_off[i], a_off[i] + a_lbo[i] of the stage (two planes, or two tap rows ty = ty0[i] + a*ty_step[i]).
  int wide, n_mma, nb_atoms, T;
  int a_off[2], a_lbo[2];
  int ty0[2], ty_step[2];      // tap row of A atom a of MMA i: ty0[i] + a*ty_step[i] (>= T: padding, dropped)
  int pl_step[2];              // plane of A atom a of MMA i: a*pl_step[i]
  int dy_bias_sub;             // stage row of dY[q0] (bias column sums); dy_sub = stage row of dY[q0 - (T-1)]
  PcU8Src u8;                  // layer 0: A patches built from the uint8 observations (see pconv_fwd_kernel)
};

__host__ __device__ inline int pc_wgrad_stage_bytes(int N, int planes, int a_rows, int dy_rows) {
  return planes * a_rows * 128 + ((dy_rows * N * 2 + 1023) / 1024) * 1024;
}
__host__ __device__ inline int pc_wgrad_smem(int N, int planes, int a_rows, int dy_rows, int stages, int raw_stage_bytes = 0) {
  return stages * (pc_wgrad_stage_bytes(N, planes, a_rows, dy_rows) + raw_stage_bytes) + 1024 + 256 + 128 * 8 * 4;
}

template <int N>
__global__ void __launch_bounds__(kPcThreads, 1) pconv_wgrad_kernel(const __grid_constant__ PcWgradParams p) {
  static_assert(N == 32 || N == 64, "gradient tile width");
  constexpr int ROWB = N * 2;                                  // bytes per dY row (64: SWIZZLE_64B, 128: SWIZZLE_128B)
  constexpr uint32_t B_LAYOUT = (ROWB == 128) ? 2u : 4u;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_bytes = (uint32_t)p.planes * p.a_rows * 128;
  const uint32_t stage_bytes = (uint32_t)pc_wgrad_stage_bytes(N, p.planes, p.a_rows, p.dy_rows);
  const bool u8 = p.u8.obs != nullptr;
  const uint32_t bar_base = smem_base + p.stages * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (8 + s); };
  const uint32_t done_bar = bar_base + 8u * 16;
  const uint32_t tmem_ptr_addr = bar_base + 8u * 17;
  float* red = reinterpret_cast<float*>(smem_raw + (bar_base + 256u - smem_u32(smem_raw)));   // [128][8] column-sum scratch

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mt = p.wide ? p.n_mma : (p.nblk + 1) >> 1;
  const int NW = p.wide ? p.nb_atoms * N : N;                  // accumulator width of one MMA
  const int acc_cols = mt * NW;
  const uint32_t tmem_cols = acc_cols <= 32 ? 32 : acc_cols <= 64 ? 64 : acc_cols <= 128 ? 128 : acc_cols <= 256 ? 256 : 512;
  const int my_tiles = ((int)blockIdx.x < p.ntiles) ? (p.ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), u8 ? 1 + 4 : 1);    // producer (dY bytes) + u8 mode: the four converter warps (A patch)
      mbar_init(empty_bar(s), 1 + 4);        // MMA commit + the four column-sum warps
    }
    mbar_init(done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_addr, tmem_cols);
  pdl_wait();
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));
  // row of the first wanted dY row inside the 8-aligned stage (wide: the row of dY[q0 - (T-1)])
  const int dy_first = p.wide ? p.dy_off - (p.T - 1) : p.dy_off;
  const int dy_sub = dy_first & 7;
  const int dy_bias_sub = p.wide ? dy_sub + (p.T - 1) : dy_sub;

  float csum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (warp == 0) {
    // ===================== TMA producer =====================
    const long idx_base = (p.idx && p.idx_off) ? (long)p.idx_off[0] * p.nb : 0;
    int img_l = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int s = it % p.stages;
      const uint32_t ph = (it / p.stages) & 1;
      long apos, dpos;
      if (p.tiles_per_img > 0) {
        if ((it & 31) == 0) {      // hoisted index lookups, see pconv_fwd_kernel
          const int tl = tile + lane * (int)gridDim.x;
          const int bl = min(tl, p.ntiles - 1) / p.tiles_per_img;
          img_l = p.idx ? p.idx[idx_base + bl] : bl;
        }
        int b = tile / p.tiles_per_img;
        int j = tile - b * p.tiles_per_img;
        long img = __shfl_sync(0xffffffffu, img_l, it & 31);
        apos = img * p.S + (long)j * 128;
        dpos = (long)b * p.S + (long)j * 128;
      } else {
        apos = (long)tile * 128;
        dpos = apos;
      }
      dpos += (dy_first & ~7);              // (dy_first may be -1 for layer 0: the grid has a zero prefix of 8 rows)
      mbar_wait(empty_bar(s), ph ^ 1, 31);
      if (u8) {
        // A patch: built by the converter warps; dY: bulk copy as before
        if (elect_one()) {
          mbar_arrive_expect_tx(full_bar(s), (uint32_t)p.dy_rows * ROWB);
          bulk_g2s(smem_base + s * stage_bytes + a_bytes, p.dy + dpos * N, p.dy_rows * ROWB, full_bar(s));
        }
        __syncwarp();
        continue;
      }
      if (elect_one()) {
        mbar_arrive_expect_tx(full_bar(s), a_bytes + (uint32_t)p.dy_rows * ROWB);
        const uint32_t dst = smem_base + s * stage_bytes;
        for (int pl = 0; pl < p.planes; ++pl)
          bulk_g2s_hint(dst + pl * p.a_rows * 128, p.a + pl * p.a_plane_stride + apos * 64, p.a_rows * 128, full_bar(s),
                        p.a_evict_first ? l2_policy_evict_first() : 0ull);
        bulk_g2s(dst + a_bytes, p.dy + dpos * N, p.dy_rows * ROWB, full_bar(s));
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t tmem_u = make_uniform(tmem_base);
    constexpr uint32_t idesc = make_idesc_bf16(128, N, 1, 1);
    const uint64_t bd0 = make_smem_desc(0, 16, 8 * ROWB, B_LAYOUT);
    const uint32_t b_hi32 = (uint32_t)(bd0 >> 32), b_flags = (uint32_t)bd0;
    const uint32_t a_hi32 = (uint32_t)(make_smem_desc(0, 16, 1024, 2) >> 32);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int s = it % p.stages;
      const uint32_t ph = (it / p.stages) & 1;
      mbar_wait(full_bar(s), ph, 32);
      tc_fence_after();
      if (elect_one()) {
       if (p.wide) {
        const uint32_t a_stage = smem_base + s * stage_bytes;
        // B: nb_atoms MN-atoms of N columns each, one dY row (ROWB bytes) apart; K groups of 8 rows 8*ROWB apart
        const uint32_t b_lo0 = ((a_stage + a_bytes + dy_sub * ROWB) >> 4) | (((uint32_t)ROWB >> 4) << 16);
        const uint32_t idesc_w = make_idesc_bf16(128, NW, 1, 1);
        for (int m = 0; m < mt; ++m) {
          const uint32_t a_lo0 = ((a_stage + p.a_off[m]) >> 4) | (((uint32_t)p.a_lbo[m] >> 4) << 16);
          const uint32_t d_tmem = tmem_u + m * NW;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint64_t ad = ((uint64_t)a_hi32 << 32) | (uint64_t)(a_lo0 + kk * 128);
            const uint64_t bd = ((uint64_t)b_hi32 << 32) | (uint64_t)(b_lo0 + kk * (ROWB));
            umma_bf16(d_tmem, ad, bd, idesc_w, (it > 0 || kk > 0) ? 1u : 0u);
          }
        }
        umma_commit(empty_bar(s));
        if (it == my_tiles - 1) umma_commit(done_bar);
       } else {
        const uint32_t a_stage = smem_base + s * stage_bytes;
        const uint32_t b_lo0 = ((a_stage + a_bytes + dy_sub * ROWB) >> 4) | b_flags;
        for (int m = 0; m < mt; ++m) {
          const int o0 = p.blk_off[2 * m];
          const int lbo = (2 * m + 1 < p.nblk) ? (p.blk_off[2 * m + 1] - o0) : 128;   // odd tail: any valid block
          const uint32_t a_lo0 = ((a_stage + o0) >> 4) | (((uint32_t)lbo >> 4) << 16);
          const uint32_t d_tmem = tmem_u + m * N;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {     // 128 positions, 16 per MMA
            const uint64_t ad = ((uint64_t)a_hi32 << 32) | (uint64_t)(a_lo0 + kk * 128);             // 16 rows * 128 B
            const uint64_t bd = ((uint64_t)b_hi32 << 32) | (uint64_t)(b_lo0 + kk * (ROWB));          // 16 rows * ROWB B >> 4
            umma_bf16(d_tmem, ad, bd, idesc, (it > 0 || kk > 0) ? 1u : 0u);
          }
        }
        umma_commit(empty_bar(s));
        if (it == my_tiles - 1) umma_commit(done_bar);
       }
      }
      __syncwarp();
    }
  } else if (warp < 6) {
    // ===================== column sums of dY (bias gradient), warps 2..5 =====================
    const int t4 = tid - 64;                       // 0..127
    constexpr int CH = ROWB / 16;                  // 16-byte chunks per row
    constexpr int RPT = 128 * CH / 128;            // rows handled per thread (= CH)
    const int c = t4 % CH, r0 = t4 / CH;           // chunk column, first row; rows r0 + i*(128/CH)
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int s = it % p.stages;
      const uint32_t ph = (it / p.stages) & 1;
      mbar_wait(full_bar(s), ph, 33);
      const uint32_t b_stage = smem_base + s * stage_bytes + a_bytes;
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const int r = dy_bias_sub + r0 + i * (128 / CH);
        const uint32_t addr = b_stage + swz_off<ROWB>(r, c);
        uint4 v;
        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
        csum[0] += bf16_lo(v.x); csum[1] += bf16_hi(v.x); csum[2] += bf16_lo(v.y); csum[3] += bf16_hi(v.y);
        csum[4] += bf16_lo(v.z); csum[5] += bf16_hi(v.z); csum[6] += bf16_lo(v.w); csum[7] += bf16_hi(v.w);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_bar(s));
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) red[t4 * 8 + e] = csum[e];
  }
  if (warp >= 10 && u8) {
    // ===================== converter warps (u8 first-layer input), warps 10..13 =====================
    const long idx_base = (p.idx && p.idx_off) ? (long)p.idx_off[0] * p.nb : 0;
    pc_u8_converter_loop(p.u8, p.idx, idx_base, p.tiles_per_img, p.ntiles, p.a_rows, p.stages, smem_base, stage_bytes,
                         full_bar, empty_bar, tid - 10 * 32, lane, 38);
  }
  __syncthreads();
  // bias partial: fixed-order sum of the 128/CH row groups per column
  if (tid < N && p.bias_partial) {
    constexpr int CH = ROWB / 16;
    const int c = tid >> 3, e = tid & 7;
    float t = 0.f;
    for (int g = 0; g < 128 / CH; ++g) t += red[(g * CH + c) * 8 + e];
    p.bias_partial[(long)blockIdx.x * N + tid] = t;
  }
  // ===================== epilogue: TMEM -> fp32 partial =====================
  if (warp >= 2 && warp < 10) {
    const int q = warp & 3, h = (warp - 2) >> 2;
    constexpr int HC = N / 2;
    if (my_tiles > 0) {
      mbar_wait(done_bar, 0, 34);
      tc_fence_after();
    }
    const int r = q * 32 + lane;
    if (p.wide) {
      // D_i row r = (atom a, channel ch), columns = (tx atom j, co): 32-column chunks never straddle a tap
      const int a = r >> 6, ch = r & 63;
      const int P = p.planes;
      for (int m = 0; m < mt; ++m) {
        const int ty = p.ty0[m] + a * p.ty_step[m], pl = a * p.pl_step[m];
        for (int c0 = h * (NW / 2); c0 < (h + 1) * (NW / 2); c0 += 32) {
          const int j = c0 / N, co0 = c0 - j * N, tx = p.T - 1 - j;
          const int blk = (ty * p.T + tx) * P + pl;
          const bool ok = ty < p.T;
          float* dst = p.partial + ((long)blockIdx.x * p.nblk * 64 + (long)(ok ? blk : 0) * 64 + ch) * N + co0;
          uint32_t v[32];
          if (my_tiles > 0) { tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + m * NW + c0, v); tmem_ld_wait(); }
          else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0;
          }
          store_rows32_coalesced(smem_base + (uint32_t)(warp - 2) * kRowStoreScratch, v, dst, ok, lane);
        }
      }
    } else
    for (int m = 0; m < mt; ++m) {
      const int blk = 2 * m + (r >> 6);
      float* dst = p.partial + ((long)blockIdx.x * p.nblk * 64 + (long)blk * 64 + (r & 63)) * N + h * HC;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + m * N + h * HC;
      if constexpr (HC == 32) {
        uint32_t v[32];
        if (my_tiles > 0) { tmem_ld32(taddr, v); tmem_ld_wait(); }
        else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0;
        }
        // (all MMAs and column sums are done: the operand stages serve as the per-warp transpose scratch)
        store_rows32_coalesced(smem_base + (uint32_t)(warp - 2) * kRowStoreScratch, v, dst, blk < p.nblk, lane);
      } else {
        uint32_t v[16];
        if (my_tiles > 0) { tmem_ld16(taddr, v); tmem_ld_wait(); }
        else {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = 0;
        }
        store_rows16_coalesced(smem_base + (uint32_t)(warp - 2) * kRowStoreScratch, v, dst, blk < p.nblk, lane);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ---------------------------------------------------------------------------
// pconv_bwd_kernel: data gradient AND weight gradient of one conv layer (l >= 1) in ONE persistent kernel.
// Both read the same tile of the output-gradient grid dY: the data gradient as its K-major A operand (rows shifted per
// tap, as in pconv_fwd_kernel), the weight gradient as its MN-major B operand (wide form: the taps of a row on the N
// axis).  Run as two kernels they share the GPU — the weight gradient holds 48 SMs while the data-gradient chain, the
// critical path, runs in two waves on the rest (profiles/r2_timeline.md).  Fused: one dY patch copy per tile serves both,
// every SM works on the critical path, and one launch / prologue / drain is paid instead of two.
//   TMEM (512 columns): [0, 2*ND) two data-gradient accumulators (epilogue overlaps the next tile),
//                       [2*ND, 2*ND + n_mma*NW) the weight-gradient accumulators, accumulated over ALL tiles of the CTA.
//   stage = dY patch (planes_out x load_rows x 128 B; wgrad B = the same bytes from row dy_first on) + forward-input patch
//           (P x a_rows x 128 B, wgrad A).
//   warps: 0 producer, 1 MMA issuer, 2..9 data-gradient epilogue per tile (+ weight-gradient epilogue at the end),
//          10..13 column sums of dY (bias gradient).
// ---------------------------------------------------------------------------
constexpr int kPcBwdThreads = 448;

struct PcBwdParams {
  PcParams d;          // data gradient: src = dY grid, w = flipped/transposed weight pack, out = mask (+ unfold)
  PcWgradParams g;     // weight gradient, wide form only (g.a = forward input grid, g.dy_off, g.a_off/a_lbo ...)
};

__host__ __device__ inline int pc_bwd_stage_bytes(int planes_out, int load_rows, int P, int a_rows) {
  return (planes_out * load_rows + P * a_rows) * 128;
}
__host__ __device__ inline int pc_bwd_smem(int ND, int ntaps, int planes_out, int load_rows, int P, int a_rows, int stages) {
  return ntaps * planes_out * ND * 128 + stages * pc_bwd_stage_bytes(planes_out, load_rows, P, a_rows) + 1024 + 256 + 128 * 8 * 4;
}

template <int ND, int NC>
__global__ void __launch_bounds__(kPcBwdThreads, 1) pconv_bwd_kernel(const __grid_constant__ PcBwdParams pp) {
  static_assert(ND == 64 || ND == 128, "data-gradient tile width");
  static_assert(NC == 64, "weight-gradient filter count");
  const PcParams& p = pp.d;
  const PcWgradParams& g = pp.g;
  if ((int)blockIdx.x >= p.ntiles) return;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int nblk = p.ntaps * p.planes;                                  // data-gradient weight tiles
  const uint32_t w_base = smem_base;
  const uint32_t dy_bytes = (uint32_t)p.planes * p.load_rows * 128;     // dY patch of a stage
  const uint32_t a_bytes = (uint32_t)g.planes * g.a_rows * 128;         // forward-input patch of a stage
  const uint32_t stage_bytes = dy_bytes + a_bytes;
  const uint32_t st_base = w_base + nblk * (ND * 128);
  const uint32_t bar_base = st_base + p.stages * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (8 + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (16 + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (18 + b); };
  const uint32_t wfull_bar = bar_base + 8u * 20;
  const uint32_t done_bar = bar_base + 8u * 21;
  const uint32_t tmem_ptr_addr = bar_base + 8u * 22;
  float* red = reinterpret_cast<float*>(smem_raw + (bar_base + 256u - smem_u32(smem_raw)));   // [128][8] column-sum scratch

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NW = g.nb_atoms * NC;                                        // width of one weight-gradient accumulator
  const int my_tiles = (p.ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1;
  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1 + 4);          // MMA commit + the four column-sum warps
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), 8);
    }
    mbar_init(wfull_bar, 1);
    mbar_init(done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_addr, 512);
  pdl_wait();
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));
  const uint32_t wg_col0 = 2 * ND;                                       // first weight-gradient column
  const int dy_first = g.dy_off - (g.T - 1);                             // patch row of dY[q0 - (T-1)] (>= 0 here)

  float csum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_arrive_expect_tx(wfull_bar, (uint32_t)nblk * ND * 128);
      for (int b = 0; b < nblk; ++b) bulk_g2s(w_base + b * (ND * 128), p.w + (long)b * ND * 64, ND * 128, wfull_bar);
    }
    __syncwarp();
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int s = it % p.stages;
      const uint32_t ph = (it / p.stages) & 1;
      const long pos0 = (long)tile * 128;
      mbar_wait(empty_bar(s), ph ^ 1, 41);
      if (elect_one()) {
        mbar_arrive_expect_tx(full_bar(s), stage_bytes);
        const uint32_t dst = st_base + s * stage_bytes;
        for (int pl = 0; pl < p.planes; ++pl)
          bulk_g2s(dst + pl * p.load_rows * 128, p.src + pl * p.src_plane_stride + pos0 * 64, p.load_rows * 128, full_bar(s));
        for (int pl = 0; pl < g.planes; ++pl)
          bulk_g2s(dst + dy_bytes + pl * g.a_rows * 128, g.a + pl * g.a_plane_stride + pos0 * 64, g.a_rows * 128, full_bar(s));
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t tmem_u = make_uniform(tmem_base);
    constexpr uint32_t idesc_d = make_idesc_bf16(128, ND, 0, 0);
    const uint32_t idesc_w = make_idesc_bf16(128, NW, 1, 1);
    const uint64_t kd0 = make_smem_desc(0, 16, 1024, 2);                 // K-major SWIZZLE_128B (data gradient A, B)
    const uint32_t k_hi32 = (uint32_t)(kd0 >> 32), k_flags = (uint32_t)kd0;
    auto mk = [&](uint32_t lo) { return ((uint64_t)k_hi32 << 32) | (uint64_t)lo; };
    const uint32_t mn_hi32 = (uint32_t)(make_smem_desc(0, 16, 1024, 2) >> 32);   // MN-major: SBO = 8 rows x 128 B
    mbar_wait(wfull_bar, 0, 42);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int s = it % p.stages;
      const uint32_t ph = (it / p.stages) & 1;
      const int acc = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      mbar_wait(tempty_bar(acc), aph ^ 1, 43);
      mbar_wait(full_bar(s), ph, 44);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t stage = st_base + s * stage_bytes;
        // ---- data gradient: every tap = the dY patch through a row-shifted descriptor ----
        {
          const uint32_t d_tmem = tmem_u + acc * ND;
          uint32_t first = 0;
          uint32_t b_lo = (w_base >> 4) | k_flags;
          for (int t = 0; t < p.ntaps; ++t) {
            uint32_t a_lo = ((stage + p.shift[t] * 128) >> 4) | k_flags;
            for (int pl = 0; pl < p.planes; ++pl) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma_bf16(d_tmem, mk(a_lo + 2 * k), mk(b_lo + 2 * k), idesc_d, first);
                first = 1;
              }
              a_lo += (uint32_t)p.load_rows * 8;
              b_lo += ND * 8;
            }
          }
          umma_commit(tfull_bar(acc));
        }
        // ---- weight gradient (wide form), accumulated over all tiles of this CTA ----
        {
          const uint32_t b_lo0 = ((stage + dy_first * 128) >> 4) | ((128u >> 4) << 16);    // atoms one dY row apart
          for (int m = 0; m < g.n_mma; ++m) {
            const uint32_t a_lo0 = ((stage + dy_bytes + g.a_off[m]) >> 4) | (((uint32_t)g.a_lbo[m] >> 4) << 16);
            const uint32_t d_tmem = tmem_u + wg_col0 + m * NW;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
              const uint64_t ad = ((uint64_t)mn_hi32 << 32) | (uint64_t)(a_lo0 + kk * 128);
              const uint64_t bd = ((uint64_t)mn_hi32 << 32) | (uint64_t)(b_lo0 + kk * 128);
              umma_bf16(d_tmem, ad, bd, idesc_w, (it > 0 || kk > 0) ? 1u : 0u);
            }
          }
        }
        umma_commit(empty_bar(s));
        if (it == my_tiles - 1) umma_commit(done_bar);
      }
      __syncwarp();
    }
  } else if (warp >= 2 && warp < 10) {
    // ===================== data-gradient epilogue (as pconv_fwd_kernel, modes 1 / 2) =====================
    const PcOut& o = p.out;
    const int q = warp & 3;
    const int h = (warp - 2) >> 2;
    constexpr int HC = ND / 2;
    constexpr int NB = HC / 32;
    const uint32_t row = q * 32 + lane;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      const uint32_t qq = (uint32_t)tile * 128 + row;
      const uint32_t b = pc_fdiv(qq, (uint32_t)p.S, p.magic_S);
      const uint32_t pl_ = qq - b * p.S;
      const uint32_t y = pc_fdiv(pl_, (uint32_t)p.Wp, p.magic_Wp), x = pl_ - y * p.Wp;
      const bool valid = (int)b < p.n_img && (int)y < p.Ho && (int)x < p.Wo;
      const int Y = y + o.dpad, X = x + o.dpad;
      const long dpos = (long)b * o.dS + Y * o.dWp + X;
      const bool store = valid && (o.mode == 2 || (Y < o.dHc && X < o.dWp));
      uint4 mk[NB][4];
      if (store) {
        const long apos = (long)b * p.S + pl_ + o.act_off;
#pragma unroll
        for (int c = 0; c < NB; ++c) pc_load_mask32(o, h * HC + c * 32, apos, mk[c]);
      }
      mbar_wait(tfull_bar(acc), aph, 45);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * ND + h * HC;
      uint32_t r[NB][32];
#pragma unroll
      for (int c = 0; c < NB; ++c) tmem_ld32(taddr + c * 32, r[c]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (store) {
#pragma unroll
        for (int c = 0; c < NB; ++c) {
          uint32_t packed[16];
          pc_finish<32>(r[c], nullptr, 1.f, false, mk[c], packed);
          if (o.mode == 2) pc_store_unfold(o, packed, (int)b, (int)y, (int)x, (h * HC + c * 32) >> 5);
          else pc_store<32>(o, packed, dpos, h * HC + c * 32);
        }
      }
    }
  } else if (warp >= 10) {
    // ===================== column sums of dY (bias gradient), warps 10..13 =====================
    const int t4 = tid - 320;                      // 0..127
    const int c = t4 & 7, r0 = t4 >> 3;            // 16-byte chunk, first row; rows r0 + 16*i
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int s = it % p.stages;
      const uint32_t ph = (it / p.stages) & 1;
      mbar_wait(full_bar(s), ph, 46);
      const uint32_t b_stage = st_base + s * stage_bytes;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = g.dy_off + r0 + i * 16;      // patch row of dY[q0 + r0 + 16 i]
        uint4 v;
        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "r"(b_stage + swz_off<128>(r, c)));
        csum[0] += bf16_lo(v.x); csum[1] += bf16_hi(v.x); csum[2] += bf16_lo(v.y); csum[3] += bf16_hi(v.y);
        csum[4] += bf16_lo(v.z); csum[5] += bf16_hi(v.z); csum[6] += bf16_lo(v.w); csum[7] += bf16_hi(v.w);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_bar(s));
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) red[t4 * 8 + e] = csum[e];
  }
  __syncthreads();
  // bias partial: fixed-order sum of the 16 row groups per column
  if (tid < NC && g.bias_partial) {
    const int c = tid >> 3, e = tid & 7;
    float t = 0.f;
    for (int gq = 0; gq < 16; ++gq) t += red[(gq * 8 + c) * 8 + e];
    g.bias_partial[(long)blockIdx.x * NC + tid] = t;
  }
  // ===================== weight-gradient epilogue: TMEM -> this CTA's fp32 partial =====================
  if (warp >= 2 && warp < 10) {
    const int q = warp & 3, h = (warp - 2) >> 2;
    mbar_wait(done_bar, 0, 47);
    tc_fence_after();
    const int r = q * 32 + lane;
    const int a = r >> 6, ch = r & 63;
    for (int m = 0; m < g.n_mma; ++m) {
      const int ty = g.ty0[m] + a * g.ty_step[m], pl = a * g.pl_step[m];
      for (int c0 = h * (NW / 2); c0 < (h + 1) * (NW / 2); c0 += 32) {
        const int j = c0 / NC, co0 = c0 - j * NC, tx = g.T - 1 - j;
        const int blk = (ty * g.T + tx) * g.planes + pl;
        const bool ok = ty < g.T;
        float* dst = g.partial + ((long)blockIdx.x * g.nblk * 64 + (long)(ok ? blk : 0) * 64 + ch) * NC + co0;
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + wg_col0 + m * NW + c0, v);
        tmem_ld_wait();
        store_rows32_coalesced(st_base + (uint32_t)(warp - 2) * kRowStoreScratch, v, dst, ok, lane);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace arl
