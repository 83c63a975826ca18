// Synchronous data-parallel gradient step fused with the optimiser update over NVLink peer memory.
//
// Reference behaviour (optimizers/sync/base.py:22-24, sync_ppo_optimizer.py:41-48, optimizers/util.py:63-76):
//   all ranks: flat grad -> NCCL all-reduce(sum) -> x 1/n_gpu -> global-norm clip -> Adam/RMSProp (identical on
//   every rank).
// Here: every rank owns 1/world of the flat vector.  ONE cooperative kernel per update
//   1. cross-GPU flag barrier (all gradients are complete),
//   2. reduce my slice by direct P2P loads from every peer's gradient buffer, average, keep the slice in a
//      local scratch, accumulate sum-of-squares,
//   3. publish my slice's sum-of-squares to every peer (P2P store) + flag barrier -> global norm (summed in
//      rank order, so bit-identical on every rank),
//   4. clip + Adam/RMSProp on my slice (m, v exist only for the slice), P2P-store the new parameters of the
//      slice into every peer's parameter vector,
//   5. flag barrier (all parameters have landed).
// Buffers are cudaMalloc'ed here and exchanged with cudaIpc handles (one process per GPU); NCCL is only the
// host-side bootstrap that ships the 64-byte handles.
#pragma once
#include <string>
#include <vector>
#include "common.cuh"

namespace arl {

constexpr int kMaxRanks = 8;
constexpr int kSyncBlocks = 148;
constexpr int kSyncThreads = 512;

struct CommDev {
  float* peer_grad[kMaxRanks];
  float* peer_param[kMaxRanks];
  unsigned int* peer_flag[kMaxRanks];     // each rank's arrival counter (written by peers)
  double* peer_norm[kMaxRanks];           // [world] slice sums of squares, slot = writer rank
  __nv_bfloat16* peer_shadow[kMaxRanks];  // bf16 FC operand copy (wfc_t / wfc_bf16) of every rank; nullptr: not shared
  int rank, world;
  long n, per;                            // vector length, slice length (multiple of 4)
  float* avg_slice;                       // local scratch [per]
  double* block_partial;                  // [kSyncBlocks]
  unsigned int* grid_counter;             // local grid barrier
  unsigned int* epoch;                    // local: number of cross-GPU barriers completed
  unsigned long long* trace;              // local [16]: device-timeline accumulators of the overlapped step (ns), see sync_trace
};

struct CommState {
  bool ready = false;
  int rank = 0, world = 1;
  long n = 0;
  void* base = nullptr;                   // local symmetric allocation
  size_t bytes = 0;
  std::vector<void*> peer_base;
  CommDev dev{};
  float* grad = nullptr;                  // local views
  float* param = nullptr;
  __nv_bfloat16* shadow = nullptr;        // local bf16 FC operand copy inside the symmetric allocation
  size_t shadow_elems = 0;
  unsigned int gen = 0;                   // grid-barrier generation carried across launches of this context
};

struct SyncUpdateArgs {
  float* param; float* grad; float* m; float* v; long n;
  long shadow_begin, shadow_end; int shadow_tiles, shadow_HW, shadow_H;   // see UpdateParams
  const float* loss_partial; int n_loss_blocks;
  const float* hyper; int* step;
  int kind; float lr, beta1, beta2, eps, rho, clip;
  float* out_norm; float* out_loss; int* log_slot; int log_cap;
};

// symmetric layout (bytes): [grad n f32][param n f32][norm kMaxRanks f64][flag u32 .. pad][shadow bf16]
inline size_t comm_layout(long n, size_t shadow_elems, size_t* off_param, size_t* off_norm, size_t* off_flag, size_t* off_shadow) {
  size_t nb = ((size_t)n * 4 + 255) / 256 * 256;
  *off_param = nb;
  *off_norm = 2 * nb;
  *off_flag = 2 * nb + 256;
  *off_shadow = 2 * nb + 512;
  return 2 * nb + 512 + (shadow_elems * 2 + 255) / 256 * 256;
}

inline int comm_local_init(CommState& s, int rank, int world, long n, size_t shadow_elems, uint8_t* handle_out, std::string& err) {
  if (world < 1 || world > kMaxRanks) { err = "world size must be in [1,8]"; return 1; }
  s.rank = rank; s.world = world; s.n = n;
  size_t op, on, of, os;
  s.shadow_elems = shadow_elems;
  s.bytes = comm_layout(n, shadow_elems, &op, &on, &of, &os);
  cudaError_t e = cudaMalloc(&s.base, s.bytes);
  if (e != cudaSuccess) { err = std::string("cudaMalloc(sym): ") + cudaGetErrorString(e); return 1; }
  cudaMemset(s.base, 0, s.bytes);
  cudaIpcMemHandle_t h;
  e = cudaIpcGetMemHandle(&h, s.base);
  if (e != cudaSuccess) { err = std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e); return 1; }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
  memcpy(handle_out, &h, 64);
  s.grad = reinterpret_cast<float*>(s.base);
  s.param = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(s.base) + op);
  s.shadow = shadow_elems ? reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(s.base) + os) : nullptr;
  return 0;
}

inline int comm_connect(CommState& s, const uint8_t* all_handles, std::string& err) {
  size_t op, on, of, os;
  comm_layout(s.n, s.shadow_elems, &op, &on, &of, &os);
  s.peer_base.assign(s.world, nullptr);
  for (int r = 0; r < s.world; ++r) {
    if (r == s.rank) { s.peer_base[r] = s.base; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, all_handles + (size_t)r * 64, 64);
    cudaError_t e = cudaIpcOpenMemHandle(&s.peer_base[r], h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { err = std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e); return 1; }
  }
  CommDev& d = s.dev;
  d.rank = s.rank; d.world = s.world; d.n = s.n;
  d.per = ((s.n + s.world - 1) / s.world + 3) / 4 * 4;
  for (int r = 0; r < s.world; ++r) {
    uint8_t* b = reinterpret_cast<uint8_t*>(s.peer_base[r]);
    d.peer_grad[r] = reinterpret_cast<float*>(b);
    d.peer_param[r] = reinterpret_cast<float*>(b + op);
    d.peer_norm[r] = reinterpret_cast<double*>(b + on);
    d.peer_flag[r] = reinterpret_cast<unsigned int*>(b + of);
    d.peer_shadow[r] = s.shadow_elems ? reinterpret_cast<__nv_bfloat16*>(b + os) : nullptr;
  }
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&d.avg_slice), (size_t)d.per * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&d.block_partial), kSyncBlocks * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&d.grid_counter), 256);
  if (e != cudaSuccess) { err = std::string("cudaMalloc(comm scratch): ") + cudaGetErrorString(e); return 1; }
  cudaMemset(d.grid_counter, 0, 256);
  d.epoch = d.grid_counter + 16;
  if (cudaMalloc(reinterpret_cast<void**>(&d.trace), 16 * sizeof(unsigned long long)) != cudaSuccess) { err = "cudaMalloc(trace)"; return 1; }
  cudaMemset(d.trace, 0, 16 * sizeof(unsigned long long));
  s.ready = true;
  return 0;
}

inline void comm_destroy(CommState& s) {
  if (!s.base) return;
  for (int r = 0; r < (int)s.peer_base.size(); ++r)
    if (r != s.rank && s.peer_base[r]) cudaIpcCloseMemHandle(s.peer_base[r]);
  cudaFree(s.dev.avg_slice); cudaFree(s.dev.block_partial); cudaFree(s.dev.grid_counter); cudaFree(s.dev.trace);
  cudaFree(s.base);
  s.base = nullptr; s.ready = false;
}

// ---- device side ---------------------------------------------------------------------------
ARL_DEVINL void st_release_sys_add(unsigned int* p) {
  asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}
// relaxed variant: the caller has already executed ONE system-scope fence; eight back-to-back release reductions
// would each wait for the previous one to be acknowledged across NVLink (measured: 19 us per barrier at 8 GPUs)
ARL_DEVINL void st_relaxed_sys_add(unsigned int* p) {
  asm volatile("red.relaxed.sys.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}
ARL_DEVINL unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// one thread: signal every peer, wait until all peers signalled me for this epoch
ARL_DEVINL void xgpu_barrier_thread(const CommDev& d) {
  unsigned int ep = *d.epoch + 1;
  __threadfence_system();
  for (int r = 0; r < d.world; ++r) st_relaxed_sys_add(d.peer_flag[r]);
  unsigned int target = ep * (unsigned int)d.world;
  long long t0 = clock64();
  while (ld_acquire_sys(d.peer_flag[d.rank]) < target) {
    if (clock64() - t0 > 20000000000LL) dev_fail(300);   // ~10 s: a peer is gone
  }
  *d.epoch = ep;
  __threadfence_system();
}

// grid-wide barrier for a co-resident (cooperative) grid; optional cross-GPU barrier by block 0
ARL_DEVINL void grid_barrier(const CommDev& d, unsigned int& gen, bool cross_gpu, bool publish_norm = false) {
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    gen += gridDim.x;
    unsigned int arrived = atomicAdd(d.grid_counter, 1u) + 1u;
    if (blockIdx.x == 0) {
      // wait for every block, then (optionally) the other GPUs, then release the grid
      long long t0 = clock64();
      while ((int)(atomicAdd(d.grid_counter, 0u) - gen) < 0) {          // wrap-safe
        if (clock64() - t0 > 20000000000LL) dev_fail(301);
      }
      if (publish_norm) {
        __threadfence();
        double t = 0.0;
        for (int b = 0; b < (int)gridDim.x; ++b) t += reinterpret_cast<volatile double*>(d.block_partial)[b];
        for (int r = 0; r < d.world; ++r) d.peer_norm[r][d.rank] = t;
      }
      if (cross_gpu) xgpu_barrier_thread(d);
      __threadfence();
      atomicExch(d.grid_counter + 1, gen);
    } else {
      (void)arrived;
      long long t0 = clock64();
      while ((int)(atomicAdd(d.grid_counter + 1, 0u) - gen) < 0) {
        if (clock64() - t0 > 20000000000LL) dev_fail(302);
      }
    }
    __threadfence();
  }
  __syncthreads();
}

// Reduce my slice over the first W ranks: all peers' loads of U float4 groups are in flight together (W*U independent
// 16-byte P2P loads per thread, ~2-3 us of NVLink latency each); the sum is taken in rank order, so every rank computes
// bit-identical averages.  Returns this thread's share of the slice's sum of squares.
template <int W, int U>
ARL_DEVINL double reduce_slice(const CommDev& d, long begin, long len4, long gtid, long gsz, float inv_world) {
  double acc = 0.0;
  for (long i0 = gtid; i0 < len4; i0 += U * gsz) {
    float4 g[W][U];
#pragma unroll
    for (int r = 0; r < W; ++r) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long i = i0 + u * gsz;
        g[r][u] = (r < d.world && i < len4) ? *reinterpret_cast<const float4*>(d.peer_grad[r] + begin + 4 * i)
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long i = i0 + u * gsz;
      if (i < len4) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < W; ++r)
          if (r < d.world) { t.x += g[r][u].x; t.y += g[r][u].y; t.z += g[r][u].z; t.w += g[r][u].w; }
        t.x *= inv_world; t.y *= inv_world; t.z *= inv_world; t.w *= inv_world;
        reinterpret_cast<float4*>(d.avg_slice)[i] = t;
        acc += (double)(t.x * t.x + t.y * t.y) + (double)(t.z * t.z + t.w * t.w);
      }
    }
  }
  return acc;
}

__global__ void __launch_bounds__(kSyncThreads) sync_allreduce_update_kernel(CommDev d, SyncUpdateArgs a) {
  // barrier generation of this launch: kept on the DEVICE (grid_counter[2]) so the launch carries no host state and
  // can be replayed from a CUDA graph.  Every block reads it before it can arrive at the first barrier; block 0
  // advances it after the last one.
  unsigned int gen = *reinterpret_cast<volatile unsigned int*>(d.grid_counter + 2);
  const unsigned int gen_next = gen + 3u * gridDim.x;
  const long begin = (long)d.rank * d.per;
  const long end = min(d.n, begin + d.per);
  const long len = end > begin ? end - begin : 0;
  const long len4 = len >> 2;
  const long gtid = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long gsz = (long)gridDim.x * blockDim.x;
  __shared__ double s_red[kSyncThreads / 32];
  __shared__ float s_scale, s_alpha;

  // 1. all gradients complete on every GPU
  grid_barrier(d, gen, true);

  // 2. reduce my slice over peers (P2P loads), average, local scratch + sum of squares.  Four independent float4
  //    groups per thread are in flight at once: a peer load is ~2-3 us of NVLink latency, the loop is latency bound.
  const float inv_world = 1.f / (float)d.world;
  double acc = (d.world <= 2) ? reduce_slice<2, 4>(d, begin, len4, gtid, gsz, inv_world)
             : (d.world <= 4) ? reduce_slice<4, 2>(d, begin, len4, gtid, gsz, inv_world)
                              : reduce_slice<kMaxRanks, 2>(d, begin, len4, gtid, gsz, inv_world);
  if (gtid == 0) {
    for (long i = len4 << 2; i < len; ++i) {
      float t = 0.f;
      for (int r = 0; r < d.world; ++r) t += d.peer_grad[r][begin + i];
      t *= inv_world;
      d.avg_slice[i] = t;
      acc += (double)t * t;
    }
  }
  acc = warp_sum_d(acc);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kSyncThreads / 32; ++w) t += s_red[w];
    d.block_partial[blockIdx.x] = t;
  }
  // 3. ONE barrier: when every block has arrived, block 0 adds up the block partials (fixed order), publishes this
  //    slice's sum of squares to every peer, runs the cross-GPU barrier and only then releases the grid
  grid_barrier(d, gen, true, true);

  // 4. global norm (rank order), clip, update my slice, P2P-store new params (+ bf16 FC operand copy) to every peer
  if (threadIdx.x == 0) {
    double t = 0.0;
    const volatile double* np = d.peer_norm[d.rank];
    for (int r = 0; r < d.world; ++r) t += np[r];
    float norm = (float)sqrt(t);
    float scale = 1.f;
    if (a.clip > 0.f) scale = fminf(norm, a.clip) / (1e-7f + norm);
    s_scale = scale;
    int tstep = a.step[0] + 1;
    float lr = a.lr * a.hyper[0];
    if (a.kind == 0) {
      double b1t = pow((double)a.beta1, (double)tstep), b2t = pow((double)a.beta2, (double)tstep);
      s_alpha = (float)((double)lr * sqrt(1.0 - b2t) / (1.0 - b1t));
    } else {
      s_alpha = lr;
    }
    if (blockIdx.x == 0) {
      int slot = a.log_slot[0];
      if (slot < a.log_cap) {
        a.out_norm[slot] = norm;
        float l = 0.f;
        for (int b = 0; b < a.n_loss_blocks; ++b) l += a.loss_partial[4 * b + 3];
        a.out_loss[slot] = l;
      }
    }
  }
  __syncthreads();
  const float scale = s_scale, alpha = s_alpha;
  const bool share_shadow = d.peer_shadow[0] != nullptr;
  for (long i = gtid; i < len4; i += gsz) {
    const long gi = begin + 4 * i;                      // begin is a multiple of 4
    const float4 g4 = reinterpret_cast<const float4*>(d.avg_slice)[i];
    const float4 p4 = *reinterpret_cast<const float4*>(a.param + gi);
    const float4 v4 = *reinterpret_cast<const float4*>(a.v + gi);
    float g[4] = {__fmul_rn(g4.x, scale), __fmul_rn(g4.y, scale), __fmul_rn(g4.z, scale), __fmul_rn(g4.w, scale)};
    float pp[4] = {p4.x, p4.y, p4.z, p4.w};
    float vv[4] = {v4.x, v4.y, v4.z, v4.w};
    if (a.kind == 0) {
      const float4 m4 = *reinterpret_cast<const float4*>(a.m + gi);
      float mm[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) opt_step_raw(0, a.beta1, a.beta2, a.eps, a.rho, pp[k], mm[k], vv[k], g[k], alpha);
      *reinterpret_cast<float4*>(a.m + gi) = make_float4(mm[0], mm[1], mm[2], mm[3]);
    } else {
      float dummy = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) opt_step_raw(1, a.beta1, a.beta2, a.eps, a.rho, pp[k], dummy, vv[k], g[k], alpha);
    }
    *reinterpret_cast<float4*>(a.v + gi) = make_float4(vv[0], vv[1], vv[2], vv[3]);
    const float4 pn = make_float4(pp[0], pp[1], pp[2], pp[3]);
    for (int r = 0; r < d.world; ++r) *reinterpret_cast<float4*>(d.peer_param[r] + gi) = pn;
    if (share_shadow && gi >= a.shadow_begin && gi + 4 <= a.shadow_end) {
      long off = gi - a.shadow_begin;
      if (a.shadow_tiles) {
        const unsigned ou = (unsigned)off, rr = ou / (unsigned)a.shadow_H;
        off = fc_tile_index(rr, (int)(ou - rr * (unsigned)a.shadow_H), a.shadow_HW, a.shadow_H);
      }
      const uint2 pk = make_uint2(pack_bf16x2(pp[0], pp[1]), pack_bf16x2(pp[2], pp[3]));
      for (int r = 0; r < d.world; ++r) *reinterpret_cast<uint2*>(d.peer_shadow[r] + off) = pk;
    }
  }
  if (gtid == 0) {
    for (long i = len4 << 2; i < len; ++i) {
      const long gi = begin + i;
      float g = __fmul_rn(d.avg_slice[i], scale);
      float p = a.param[gi], m = (a.kind == 0) ? a.m[gi] : 0.f, v = a.v[gi];
      opt_step_raw(a.kind, a.beta1, a.beta2, a.eps, a.rho, p, m, v, g, alpha);
      if (a.kind == 0) a.m[gi] = m;
      a.v[gi] = v;
      for (int r = 0; r < d.world; ++r) d.peer_param[r][gi] = p;
    }
  }
  // 5. all parameter slices have landed everywhere
  grid_barrier(d, gen, true);
  if (blockIdx.x == 0 && threadIdx.x == 0) *reinterpret_cast<volatile unsigned int*>(d.grid_counter + 2) = gen_next;
}

// ===========================================================================
// OVERLAPPED synchronous step (no global-norm clipping: PPO's default).  The monolithic kernel above starts after the
// whole backward pass and runs three cross-GPU barriers back to back; here the exchange of the FC weight gradient —
// 97.8 % of the vector, final right after the FC weight-gradient tiles, i.e. BEFORE the conv gradient chain — runs on its
// own stream beside that chain, and only the ~80 k other elements are exchanged on the critical path.  Plain launches
// only (no cooperative grid, no intra-grid barrier): every block polls a LOCAL flag word that peers advance.
//
//   sync_fc_kernel          (stream-ordered behind this rank's fc_wgrad + fc_dgrad) block 0 tells every peer "my FC gradient
//                           is final and my FC weights have been consumed" (flag A); all blocks wait for flag A from all
//                           ranks; rank r reduces slice r of the FC range by P2P loads
//                           (rank order: bit-identical everywhere), averages, Adam/RMSProp on the slice (its m, v live
//                           here only), P2P-stores the new fp32 weights + bf16 operand tiles to every rank; the LAST
//                           block publishes the slice's sum of squares to every rank, then flag B ("slice r published,
//                           my reads of your gradients are done")
//   finalize_grads_kernel   (main stream, as on one GPU) the other tensors' local gradients -> flat vector
//   sync_tail_kernel        block 0 tells every peer "my small gradients are final" (flag 2); all blocks wait for flag 2
//                           and flag B from all ranks; EVERY rank averages the small gradients of
//                           all ranks (P2P loads, rank order) and applies the update to its own replica (identical
//                           inputs, identical arithmetic -> bit-identical parameters, no broadcast), refreshes its conv
//                           operand packs; the last block logs norm / loss and advances the device counters + epochs.
// Reuse safety follows from stream order: a rank signals A(k+1) only after its tail(k), so nobody overwrites a gradient
// or a norm slot a peer may still be reading.
// ===========================================================================
// small CTAs (256 threads, <= 64 registers) so they co-reside with the persistent conv CTAs (352 threads x 118 registers)
constexpr int kSyncFcBlocks = 148;   // one per SM: 256 threads x 64 registers fit beside any conv CTA (<= 352 x 129)
constexpr int kSyncFcThreads = 256;
enum { FLAG_OLD = 0, FLAG_A = 16, FLAG_B = 32, FLAG_2 = 48 };           // u32 word offsets inside a rank's flag block
enum { EP_A = 20, EP_B = 21, EP_2 = 22, TK_FC = 24, TK_TAIL = 26 };     // words of the local counter block (grid_counter)

ARL_DEVINL unsigned long long gtimer_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
// device timeline of the overlapped step, accumulated by thread 0 of block 0 of each kernel (ns, %globaltimer):
//   [0] sync_fc: wait for flag A    [1] sync_fc: reduce + update + publish (block 0)   [2] sync_tail: wait for flags 2 / B
//   [3] sync_tail: average + update (block 0)   [4] steps   [5] slack = tail start - FC exchange end (signed; > 0: the FC
//   exchange was over before the tail began, i.e. fully hidden behind the conv gradient chain)   [8] last FC end stamp
enum { TR_FC_WAIT = 0, TR_FC_WORK = 1, TR_TAIL_WAIT = 2, TR_TAIL_WORK = 3, TR_COUNT = 4, TR_SLACK = 5, TR_FC_END = 8 };

// one thread: wait until every rank has signalled `flag_word` for the epoch after `epoch_word`
ARL_DEVINL void sync_wait_flag(const CommDev& d, int flag_word, int epoch_word, int code) {
  const unsigned int ep = reinterpret_cast<volatile unsigned int*>(d.grid_counter)[epoch_word] + 1u;
  const unsigned int target = ep * (unsigned int)d.world;
  long long t0 = clock64();
  while ((int)(ld_acquire_sys(d.peer_flag[d.rank] + flag_word) - target) < 0) {
    __nanosleep(64);
    if (clock64() - t0 > 20000000000LL) dev_fail(code);
  }
}

struct SyncFcArgs {
  float* param; float* m; float* v;
  long fc_begin, fc_len;                 // FC weight range of the flat vector (both multiples of 4)
  long per;                              // slice length per rank (multiple of 4)
  int shadow_tiles, shadow_HW, shadow_H;
  const float* hyper; const int* step;
  int kind; float lr, beta1, beta2, eps, rho;
};

template <int W>
__global__ void __launch_bounds__(kSyncFcThreads, 4) sync_fc_kernel(CommDev d, SyncFcArgs a) {
  __shared__ double s_red[kSyncFcThreads / 32];
  __shared__ float s_alpha;
  __shared__ int s_last;
  unsigned long long tr0 = 0, tr1 = 0;
  if (threadIdx.x == 0) {
    tr0 = gtimer_ns();
    if (blockIdx.x == 0) {
      // "my FC gradient is final, my FC weights have been consumed" (this kernel is stream-ordered behind both kernels)
      __threadfence_system();
      for (int r = 0; r < d.world; ++r) st_relaxed_sys_add(d.peer_flag[r] + FLAG_A);
    }
    sync_wait_flag(d, FLAG_A, EP_A, 330);
    tr1 = gtimer_ns();
    const int tstep = a.step[0] + 1;
    const float lr = a.lr * a.hyper[0];
    if (a.kind == 0) {
      const double b1t = pow((double)a.beta1, (double)tstep), b2t = pow((double)a.beta2, (double)tstep);
      s_alpha = (float)((double)lr * sqrt(1.0 - b2t) / (1.0 - b1t));
    } else {
      s_alpha = lr;
    }
  }
  __syncthreads();
  const float alpha = s_alpha;
  const float inv_world = 1.f / (float)d.world;
  const long begin = a.fc_begin + (long)d.rank * a.per;
  const long end = min(a.fc_begin + a.fc_len, begin + a.per);
  const long len4 = end > begin ? (end - begin) >> 2 : 0;
  const long gtid = (long)blockIdx.x * blockDim.x + threadIdx.x, gsz = (long)gridDim.x * blockDim.x;
  // every load of an iteration is issued before the first use: W*U gradient groups (one local, the others over NVLink,
  // ~2.5 us each) and the 3*U local state groups (p, m, v) — the loop is pure latency
  constexpr int U = (W <= 2) ? 2 : 1;
  double acc = 0.0;
  for (long i0 = gtid; i0 < len4; i0 += U * gsz) {
    float4 g[W][U], p4[U], v4[U], m4[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long i = i0 + u * gsz;
      const bool in = i < len4;
      const long gi = begin + 4 * (in ? i : 0);
#pragma unroll
      for (int r = 0; r < W; ++r)
        g[r][u] = (r < d.world && in) ? *reinterpret_cast<const float4*>(d.peer_grad[r] + gi) : make_float4(0.f, 0.f, 0.f, 0.f);
      p4[u] = *reinterpret_cast<const float4*>(a.param + gi);
      v4[u] = *reinterpret_cast<const float4*>(a.v + gi);
      m4[u] = (a.kind == 0) ? *reinterpret_cast<const float4*>(a.m + gi) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long i = i0 + u * gsz;
      if (i >= len4) continue;
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r = 0; r < W; ++r)
        if (r < d.world) { t.x += g[r][u].x; t.y += g[r][u].y; t.z += g[r][u].z; t.w += g[r][u].w; }
      t.x *= inv_world; t.y *= inv_world; t.z *= inv_world; t.w *= inv_world;
      acc += (double)(t.x * t.x + t.y * t.y) + (double)(t.z * t.z + t.w * t.w);
      const long gi = begin + 4 * i;
      float gg[4] = {t.x, t.y, t.z, t.w};
      float pp[4] = {p4[u].x, p4[u].y, p4[u].z, p4[u].w};
      float vv[4] = {v4[u].x, v4[u].y, v4[u].z, v4[u].w};
      float mm[4] = {m4[u].x, m4[u].y, m4[u].z, m4[u].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) opt_step_raw(a.kind, a.beta1, a.beta2, a.eps, a.rho, pp[k], mm[k], vv[k], gg[k], alpha);
      if (a.kind == 0) *reinterpret_cast<float4*>(a.m + gi) = make_float4(mm[0], mm[1], mm[2], mm[3]);
      *reinterpret_cast<float4*>(a.v + gi) = make_float4(vv[0], vv[1], vv[2], vv[3]);
      const float4 pn = make_float4(pp[0], pp[1], pp[2], pp[3]);
      long off = gi - a.fc_begin;
      if (a.shadow_tiles) {
        const unsigned ou = (unsigned)off, rr = ou / (unsigned)a.shadow_H;
        off = fc_tile_index(rr, (int)(ou - rr * (unsigned)a.shadow_H), a.shadow_HW, a.shadow_H);
      }
      const uint2 pk = make_uint2(pack_bf16x2(pp[0], pp[1]), pack_bf16x2(pp[2], pp[3]));
      for (int r = 0; r < d.world; ++r) {
        *reinterpret_cast<float4*>(d.peer_param[r] + gi) = pn;
        *reinterpret_cast<uint2*>(d.peer_shadow[r] + off) = pk;
      }
    }
  }
  __threadfence_system();                  // this thread's peer stores are ordered before the ticket below
  if (blockIdx.x == 0 && threadIdx.x == 0) { d.trace[TR_FC_WAIT] += tr1 - tr0; d.trace[TR_FC_WORK] += gtimer_ns() - tr1; }
  acc = warp_sum_d(acc);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kSyncFcThreads / 32; ++w) t += s_red[w];
    d.block_partial[blockIdx.x] = t;
    __threadfence();
    const unsigned int tk = atomicAdd(d.grid_counter + TK_FC, 1u);
    s_last = ((tk + 1u) % gridDim.x == 0u) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  if (threadIdx.x == 0) {
    __threadfence();
    double t = 0.0;
    for (int b = 0; b < (int)gridDim.x; ++b) t += __ldcg(d.block_partial + b);
    for (int r = 0; r < d.world; ++r) d.peer_norm[r][d.rank] = t;         // slot = writer rank
    reinterpret_cast<volatile unsigned int*>(d.grid_counter)[EP_A] += 1u;  // every block passed its flag-A wait
    __threadfence_system();
    for (int r = 0; r < d.world; ++r) st_relaxed_sys_add(d.peer_flag[r] + FLAG_B);
    d.trace[TR_FC_END] = gtimer_ns();
  }
}

struct SyncTailArgs {
  float* param; float* m; float* v; long n;
  long fc_begin, fc_len;                 // excluded range (done by sync_fc_kernel)
  const float* loss_partial; int n_loss_blocks;
  const float* hyper; int* step;
  int kind; float lr, beta1, beta2, eps, rho;
  float* out_norm; float* out_loss; int* log_slot; int log_cap; int* mb_counter;
  const unsigned long long* pk_slots; long conv_end;      // conv operand slot table (UpdateParams::pk_slots)
  double* partial;                       // [gridDim.x] block sums of squares
};

// element j of the "everything but the FC weights" index space -> index in the flat vector
ARL_DEVINL long sync_small_index(long j, long fc_begin, long fc_len) { return j < fc_begin ? j : j + fc_len; }

template <int W>
__global__ void __launch_bounds__(256) sync_tail_kernel(CommDev d, SyncTailArgs a) {
  __shared__ double s_red[8];
  __shared__ float s_alpha;
  __shared__ int s_last;
  unsigned long long tr0 = 0, tr1 = 0;
  if (threadIdx.x == 0) {
    tr0 = gtimer_ns();
    if (blockIdx.x == 0) {
      // "my small gradients are final" (this kernel is stream-ordered behind finalize_grads and the side streams)
      __threadfence_system();
      for (int r = 0; r < d.world; ++r) st_relaxed_sys_add(d.peer_flag[r] + FLAG_2);
    }
    sync_wait_flag(d, FLAG_2, EP_2, 331);
    sync_wait_flag(d, FLAG_B, EP_B, 332);
    tr1 = gtimer_ns();
    const int tstep = a.step[0] + 1;
    const float lr = a.lr * a.hyper[0];
    if (a.kind == 0) {
      const double b1t = pow((double)a.beta1, (double)tstep), b2t = pow((double)a.beta2, (double)tstep);
      s_alpha = (float)((double)lr * sqrt(1.0 - b2t) / (1.0 - b1t));
    } else {
      s_alpha = lr;
    }
  }
  __syncthreads();
  const float alpha = s_alpha;
  const float inv_world = 1.f / (float)d.world;
  const long n_small = a.n - a.fc_len;
  double acc = 0.0;
  for (long j = (long)blockIdx.x * blockDim.x + threadIdx.x; j < n_small; j += (long)gridDim.x * blockDim.x) {
    const long i = sync_small_index(j, a.fc_begin, a.fc_len);
    float gr[W];
#pragma unroll
    for (int r = 0; r < W; ++r) gr[r] = (r < d.world) ? d.peer_grad[r][i] : 0.f;
    float g = 0.f;
#pragma unroll
    for (int r = 0; r < W; ++r)
      if (r < d.world) g += gr[r];
    g *= inv_world;
    acc += (double)(g * g);
    float pv = a.param[i], mm = (a.kind == 0) ? a.m[i] : 0.f, vv = a.v[i];
    opt_step_raw(a.kind, a.beta1, a.beta2, a.eps, a.rho, pv, mm, vv, g, alpha);
    if (a.kind == 0) a.m[i] = mm;
    a.v[i] = vv;
    a.param[i] = pv;
    if (a.pk_slots && i < a.conv_end) {
      const ulonglong2 sl = __ldg(reinterpret_cast<const ulonglong2*>(a.pk_slots + 2 * i));
      const __nv_bfloat16 b = __float2bfloat16_rn(pv);
      if (sl.x) *reinterpret_cast<__nv_bfloat16*>(sl.x) = b;
      if (sl.y) *reinterpret_cast<__nv_bfloat16*>(sl.y) = b;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    d.trace[TR_TAIL_WAIT] += tr1 - tr0; d.trace[TR_TAIL_WORK] += gtimer_ns() - tr1; d.trace[TR_COUNT] += 1ULL;
    d.trace[TR_SLACK] += tr0 - d.trace[TR_FC_END];        // two's complement: read back as signed
  }
  acc = warp_sum_d(acc);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s_red[w];
    a.partial[blockIdx.x] = t;
    __threadfence();
    const unsigned int tk = atomicAdd(d.grid_counter + TK_TAIL, 1u);
    s_last = ((tk + 1u) % gridDim.x == 0u) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  // last block: fixed-assignment strided sums over all its threads, then a fixed tree (bit-reproducible)
  __threadfence();
  double a2 = 0.0;
  for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) a2 += __ldcg(a.partial + b);
  a2 = warp_sum_d(a2);
  float l = 0.f;
  for (int b = threadIdx.x; b < a.n_loss_blocks; b += blockDim.x) l += a.loss_partial[4 * b + 3];
  l = warp_sum(l);
  __shared__ float s_l[8];
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { s_red[threadIdx.x >> 5] = a2; s_l[threadIdx.x >> 5] = l; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    float tl = 0.f;
    for (int w = 0; w < 8; ++w) { t += s_red[w]; tl += s_l[w]; }
    const volatile double* np = d.peer_norm[d.rank];
    for (int r = 0; r < d.world; ++r) t += np[r];                           // FC slices, rank order
    const int slot = a.log_slot[0];
    if (slot < a.log_cap) { a.out_norm[slot] = (float)sqrt(t); a.out_loss[slot] = tl; }
    a.step[0] += 1; a.log_slot[0] += 1; a.mb_counter[0] += 1;
    volatile unsigned int* gc = reinterpret_cast<volatile unsigned int*>(d.grid_counter);
    gc[EP_2] += 1u; gc[EP_B] += 1u;
  }
}

__global__ void xgpu_barrier_kernel(CommDev d) {
  if (threadIdx.x == 0 && blockIdx.x == 0) xgpu_barrier_thread(d);
}

inline int comm_barrier(CommState& s, cudaStream_t st, std::string& err) {
  if (!s.ready) { err = "comm not connected"; return 1; }
  xgpu_barrier_kernel<<<1, 32, 0, st>>>(s.dev);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { err = cudaGetErrorString(e); return 1; }
  return 0;
}

inline int comm_sync_update(CommState& s, SyncUpdateArgs a, cudaStream_t st, std::string& err) {
  if (!s.ready) { err = "comm not connected"; return 1; }
  if (a.param != s.param || a.grad != s.grad) { err = "params/grad must be the comm's symmetric buffers"; return 1; }
  void* args[] = {(void*)&s.dev, (void*)&a};
  // Cooperative launch in both cases (attribute form inside stream capture): the driver guarantees co-residency of the
  // 148 blocks, which is what the grid barrier needs.
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(st, &cs);
  cudaError_t e;
  if (cs == cudaStreamCaptureStatusActive) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(kSyncBlocks); cfg.blockDim = dim3(kSyncThreads); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;
    at[0].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, sync_allreduce_update_kernel, s.dev, a);
  } else {
    e = cudaLaunchCooperativeKernel((const void*)sync_allreduce_update_kernel, dim3(kSyncBlocks), dim3(kSyncThreads), args, 0, st);
  }
  if (e != cudaSuccess) { err = std::string("sync kernel launch: ") + cudaGetErrorString(e); return 1; }
  return 0;
}

// ===========================================================================
// Asynchronous data parallel: central parameter / optimiser-state store with chunk-granular mutual exclusion.
//
// Reference behaviour (optimizers/async/base.py:59-104, chunked_updates.py:53-120, async_a2c_optimizer.py:96-109):
//   every learner: local gradient, clipped LOCALLY -> for each chunk: take the chunk's lock, apply RMSProp/Adam
//   with the local gradient to the CENTRAL (p, m, v)[chunk] (Adam's t is per learner), release -> copy the new
//   central p into the local parameters.  No collective; staleness between learners is allowed.
// Here the central store lives in rank 0's HBM (one cudaMalloc, shared with a cudaIpc handle); every learner runs
// ONE kernel per update: a CTA takes a region's lock with a system-scope CAS over NVLink, streams the region's
// (p, m, v) through registers with system-scope loads/stores, writes the new p both to the central store and to
// its own parameter vector (+ the bf16 FC operand copy), and releases the lock.  A CTA never holds two locks and
// never waits while holding one, so learners cannot deadlock.  The lock regions subdivide the reference's
// n_update_chunks chunks (finer regions = more CTAs in flight; the exclusion the reference relies on — an
// element's (p, m, v) triple is updated atomically — is preserved).
// ===========================================================================
struct AsyncDev {
  float* cp; float* cm; float* cv;   // central params / first moment (Adam) / second moment or RMSProp accumulator
  unsigned int* locks;               // [n_locks], 0 = free
  int n_locks;
  long per;                          // elements per lock region (multiple of 4)
};

struct AsyncState {
  bool ready = false;
  int rank = 0, world = 1;
  void* base = nullptr;              // rank 0: owning allocation; others: opened IPC mapping
  bool owner = false;
  AsyncDev dev{};
};

inline size_t async_layout(long n, int n_locks, size_t* off_m, size_t* off_v, size_t* off_locks) {
  size_t nb = ((size_t)n * 4 + 255) / 256 * 256;
  *off_m = nb; *off_v = 2 * nb; *off_locks = 3 * nb;
  return 3 * nb + ((size_t)n_locks * 4 + 255) / 256 * 256;
}

ARL_DEVINL float4 ld_sys_f4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
ARL_DEVINL void st_sys_f4(float* p, float4 v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
ARL_DEVINL float ld_sys_f(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
ARL_DEVINL void st_sys_f(float* p, float v) { asm volatile("st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }

constexpr int kAsyncThreads = 512;

__global__ void __launch_bounds__(kAsyncThreads) async_push_pull_kernel(AsyncDev d, UpdateParams p) {
  __shared__ double s_red[kAsyncThreads / 32];
  __shared__ float s_scale, s_alpha;
  // local global-norm clip (optimizers/util.py:70-76) and the per-learner step size — same arithmetic as update_kernel
  double acc = 0.0;
  for (int i = threadIdx.x; i < p.n_partial; i += blockDim.x) acc += p.sumsq_partial[i];
  acc = warp_sum_d(acc);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kAsyncThreads / 32; ++w) t += s_red[w];
    float norm = (float)sqrt(t);
    float scale = p.gscale;
    if (p.clip > 0.f) scale *= fminf(norm, p.clip) / (1e-7f + norm);
    s_scale = scale;
    int tstep = p.step[0] + 1;
    float lr = p.lr * p.hyper[0];
    if (p.kind == 0) {
      double b1t = pow((double)p.beta1, (double)tstep), b2t = pow((double)p.beta2, (double)tstep);
      s_alpha = (float)((double)lr * sqrt(1.0 - b2t) / (1.0 - b1t));
    } else {
      s_alpha = lr;
    }
    if (blockIdx.x == 0) {
      int slot = p.log_slot[0];
      if (slot < p.log_cap) {
        p.out_norm[slot] = norm;
        float l = 0.f;
        for (int b = 0; b < p.n_loss_blocks; ++b) l += p.loss_partial[4 * b + 3];
        p.out_loss[slot] = l;
      }
    }
  }
  __syncthreads();
  const float scale = s_scale, alpha = s_alpha;
  for (int c = blockIdx.x; c < d.n_locks; c += gridDim.x) {
    const long begin = (long)c * d.per;
    const long end = min(p.n, begin + d.per);
    if (threadIdx.x == 0) {
      long long t0 = clock64();
      unsigned int backoff = 32;
      while (atomicCAS_system(d.locks + c, 0u, 1u) != 0u) {
        __nanosleep(backoff);
        if (backoff < 2048) backoff <<= 1;
        if (clock64() - t0 > 40000000000LL) dev_fail(310);     // ~20 s: a lock holder died
      }
      __threadfence_system();
    }
    __syncthreads();
    const long n4 = (end - begin) >> 2;
    for (long i = threadIdx.x; i < n4; i += blockDim.x) {
      const long e0 = begin + (i << 2);
      const float4 g4 = *reinterpret_cast<const float4*>(p.grad + e0);
      float4 p4 = ld_sys_f4(d.cp + e0);
      float4 v4 = ld_sys_f4(d.cv + e0);
      float g[4] = {__fmul_rn(g4.x, scale), __fmul_rn(g4.y, scale), __fmul_rn(g4.z, scale), __fmul_rn(g4.w, scale)};
      float pp[4] = {p4.x, p4.y, p4.z, p4.w};
      float vv[4] = {v4.x, v4.y, v4.z, v4.w};
      if (p.kind == 0) {
        float4 m4 = ld_sys_f4(d.cm + e0);
        float mm[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) opt_step1(p, pp[k], mm[k], vv[k], g[k], alpha);
        st_sys_f4(d.cm + e0, make_float4(mm[0], mm[1], mm[2], mm[3]));
      } else {
        float dummy = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) opt_step1(p, pp[k], dummy, vv[k], g[k], alpha);
      }
      st_sys_f4(d.cv + e0, make_float4(vv[0], vv[1], vv[2], vv[3]));
      st_sys_f4(d.cp + e0, make_float4(pp[0], pp[1], pp[2], pp[3]));
      *reinterpret_cast<float4*>(p.param + e0) = make_float4(pp[0], pp[1], pp[2], pp[3]);   // pull
      if (p.shadow && e0 >= p.shadow_begin && e0 + 4 <= p.shadow_end) {
        long off = e0 - p.shadow_begin;
        if (p.shadow_tiles) {
          const unsigned ou = (unsigned)off, rr = ou / (unsigned)p.shadow_H;
          off = fc_tile_index(rr, (int)(ou - rr * (unsigned)p.shadow_H), p.shadow_HW, p.shadow_H);
        }
        *reinterpret_cast<uint2*>(p.shadow + off) = make_uint2(pack_bf16x2(pp[0], pp[1]), pack_bf16x2(pp[2], pp[3]));
      }
    }
    // tail of the vector (n % 4 elements) belongs to the last region
    for (long e = begin + (n4 << 2) + threadIdx.x; e < end; e += blockDim.x) {
      float g = __fmul_rn(p.grad[e], scale), pv = ld_sys_f(d.cp + e), vv = ld_sys_f(d.cv + e);
      float mm = (p.kind == 0) ? ld_sys_f(d.cm + e) : 0.f;
      opt_step1(p, pv, mm, vv, g, alpha);
      if (p.kind == 0) st_sys_f(d.cm + e, mm);
      st_sys_f(d.cv + e, vv);
      st_sys_f(d.cp + e, pv);
      p.param[e] = pv;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) atomicExch_system(d.locks + c, 0u);
  }
}

// Pull only (ActsrvAltOvrlpPollSampler, poll_sampler.py:29-39: the sampler refreshes its policy from the central
// parameters every poll_horizon rollout steps): per lock region {lock, copy central p -> local p (+ the bf16 FC operand
// copy), unlock} — the same mutual exclusion a pushing learner takes, so a region is never read half-updated.
__global__ void __launch_bounds__(kAsyncThreads) async_pull_kernel(AsyncDev d, UpdateParams p) {
  for (int c = blockIdx.x; c < d.n_locks; c += gridDim.x) {
    const long begin = (long)c * d.per;
    const long end = min(p.n, begin + d.per);
    if (threadIdx.x == 0) {
      long long t0 = clock64();
      unsigned int backoff = 32;
      while (atomicCAS_system(d.locks + c, 0u, 1u) != 0u) {
        __nanosleep(backoff);
        if (backoff < 2048) backoff <<= 1;
        if (clock64() - t0 > 40000000000LL) dev_fail(311);
      }
      __threadfence_system();
    }
    __syncthreads();
    const long n4 = (end - begin) >> 2;
    for (long i = threadIdx.x; i < n4; i += blockDim.x) {
      const long e0 = begin + (i << 2);
      const float4 p4 = ld_sys_f4(d.cp + e0);
      *reinterpret_cast<float4*>(p.param + e0) = p4;
      if (p.shadow && e0 >= p.shadow_begin && e0 + 4 <= p.shadow_end) {
        long off = e0 - p.shadow_begin;
        if (p.shadow_tiles) {
          const unsigned ou = (unsigned)off, rr = ou / (unsigned)p.shadow_H;
          off = fc_tile_index(rr, (int)(ou - rr * (unsigned)p.shadow_H), p.shadow_HW, p.shadow_H);
        }
        *reinterpret_cast<uint2*>(p.shadow + off) = make_uint2(pack_bf16x2(p4.x, p4.y), pack_bf16x2(p4.z, p4.w));
      }
    }
    for (long e = begin + (n4 << 2) + threadIdx.x; e < end; e += blockDim.x) p.param[e] = ld_sys_f(d.cp + e);
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) atomicExch_system(d.locks + c, 0u);
  }
}

inline int async_local_init(AsyncState& s, int rank, int world, long n, int n_update_chunks, const float* params,
                            uint8_t* handle_out, std::string& err) {
  if (world < 1 || world > kMaxRanks) { err = "world size must be in [1,8]"; return 1; }
  if (n_update_chunks < 1) { err = "n_update_chunks must be >= 1"; return 1; }
  s.rank = rank; s.world = world;
  // reference chunk length (chunked_updates.py:61): n // n_chunks + 1, subdivided so that ~148 regions exist
  long ppc = n / n_update_chunks + 1;
  int sub = std::max(1, 148 / n_update_chunks);
  long per = ((ppc + sub - 1) / sub + 3) / 4 * 4;
  s.dev.per = per;
  s.dev.n_locks = (int)((n + per - 1) / per);
  memset(handle_out, 0, 64);
  if (rank != 0) return 0;
  size_t om, ov, ol;
  size_t bytes = async_layout(n, s.dev.n_locks, &om, &ov, &ol);
  cudaError_t e = cudaMalloc(&s.base, bytes);
  if (e != cudaSuccess) { err = std::string("cudaMalloc(central): ") + cudaGetErrorString(e); return 1; }
  cudaMemset(s.base, 0, bytes);
  cudaMemcpy(s.base, params, (size_t)n * 4, cudaMemcpyDeviceToDevice);
  s.owner = true;
  if (world > 1) {
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, s.base);
    if (e != cudaSuccess) { err = std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e); return 1; }
    memcpy(handle_out, &h, 64);
  }
  return 0;
}

inline int async_connect(AsyncState& s, long n, const uint8_t* rank0_handle, std::string& err) {
  if (s.rank != 0) {
    cudaIpcMemHandle_t h;
    memcpy(&h, rank0_handle, 64);
    cudaError_t e = cudaIpcOpenMemHandle(&s.base, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { err = std::string("cudaIpcOpenMemHandle(central): ") + cudaGetErrorString(e); return 1; }
  }
  size_t om, ov, ol;
  async_layout(n, s.dev.n_locks, &om, &ov, &ol);
  uint8_t* b = reinterpret_cast<uint8_t*>(s.base);
  s.dev.cp = reinterpret_cast<float*>(b);
  s.dev.cm = reinterpret_cast<float*>(b + om);
  s.dev.cv = reinterpret_cast<float*>(b + ov);
  s.dev.locks = reinterpret_cast<unsigned int*>(b + ol);
  s.ready = true;
  return 0;
}

inline void async_destroy(AsyncState& s) {
  if (!s.base) return;
  if (s.owner) cudaFree(s.base); else cudaIpcCloseMemHandle(s.base);
  s.base = nullptr; s.ready = false;
}

}  // namespace arl
