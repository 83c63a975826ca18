// Shared device helpers for the accel_rl_b200 kernels (sm_100a only).
//
// PTX wrappers for mbarrier, the async-proxy fence, tcgen05 (TMEM alloc, MMA,
// commit, ld) and the UMMA shared-memory / instruction descriptors.  Every
// blocking wait carries a watchdog: a barrier that does not complete within
// ~2 s records an error code in a global flag and traps, so a protocol bug
// shows up as a CUDA error on the host instead of a hung GPU.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define ARL_DEVINL __device__ __forceinline__

namespace arl {

// ---------------------------------------------------------------------------
// error flag (device global): 0 = ok
// ---------------------------------------------------------------------------
__device__ int g_dev_error = 0;

ARL_DEVINL void dev_fail(int code) {
  atomicExch(&g_dev_error, code);
  __threadfence_system();
  __trap();
}

// ---------------------------------------------------------------------------
// shared-memory address + mbarrier
// ---------------------------------------------------------------------------
ARL_DEVINL uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

ARL_DEVINL void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

ARL_DEVINL void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

ARL_DEVINL void mbar_arrive(uint32_t bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}

ARL_DEVINL void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar),
               "r"(bytes)
               : "memory");
}

ARL_DEVINL bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}

// Wait for the phase with the given parity to complete (watchdog: ~2 s).
ARL_DEVINL void mbar_wait(uint32_t bar, uint32_t parity, int where) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3ff) == 0) {
      if (clock64() - t0 > 4000000000LL) dev_fail(100 + where);
    }
  }
}

// thread-block clusters: rank of this CTA, cluster-wide barrier (all threads of all CTAs), distributed shared memory load
ARL_DEVINL uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
ARL_DEVINL void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
ARL_DEVINL float4 ld_dsmem_f4(uint32_t local_smem_addr, uint32_t cta_rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_smem_addr), "r"(cta_rank));
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(remote) : "memory");
  return v;
}

// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma reads)
ARL_DEVINL void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 16-byte asynchronous global -> shared copy (LDGSTS); src_bytes == 0 writes 16 zero bytes (padding rows)
ARL_DEVINL void cp_async16(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
// the mbarrier phase cannot complete before this thread's prior cp.async copies have landed
// (pending count +1 now, -1 when the copies complete: pair it with a regular arrive)
ARL_DEVINL void cp_async_mbar_arrive(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

// 1-D bulk copy global -> shared (TMA engine, SASS UBLKCP), completion on mbarrier
ARL_DEVINL void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}

// the same copy with an L2 eviction-priority hint (policy from l2_policy_evict_last / _first; 0 = no hint)
ARL_DEVINL void bulk_g2s_hint(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
  if (policy == 0) { bulk_g2s(dst_smem, src, bytes, bar); return; }
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst_smem),
      "l"(src), "r"(bytes), "r"(bar), "l"(policy)
      : "memory");
}
ARL_DEVINL uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
ARL_DEVINL uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}

// ---------------------------------------------------------------------------
// tcgen05 / TMEM
// ---------------------------------------------------------------------------
ARL_DEVINL void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}

ARL_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

ARL_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
ARL_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread.
ARL_DEVINL void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// --- warp-converged single-lane issue ----------------------------------------------------------------------
// tcgen05.mma / tcgen05.commit / cp.async.bulk take their operands from UNIFORM registers.  Issued from inside a
// divergent `if (lane == 0)` region, ptxas cannot prove the operands uniform and wraps every instruction in an
// ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall (~20 SASS instructions per MMA, measured: 3 400 cycles per 16-MMA
// tile).  Keep the role warp CONVERGED, compute operands from warp-uniform values, and issue inside
// `if (elect_one()) { ... }`: ptxas recognises the elect.sync idiom and emits straight UTCHMMA / UBLKCP sequences
// fed by the uniform datapath.
ARL_DEVINL uint32_t elect_one() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(p));
  return p;
}
// a value every lane holds -> a value ptxas KNOWS is warp-uniform (REDUX writes a uniform register)
ARL_DEVINL uint32_t make_uniform(uint32_t v) { return __reduce_or_sync(0xffffffffu, v); }

// mbarrier arrives when all previously issued MMAs of this thread have completed
ARL_DEVINL void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread = lane/row)
ARL_DEVINL void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

ARL_DEVINL void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

ARL_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------
// UMMA descriptors (bit layout: cute/arch/mma_sm100_desc.hpp in the CUTLASS tree)
// ---------------------------------------------------------------------------
// layout_type: 0 none, 2 SWIZZLE_128B, 4 SWIZZLE_64B, 6 SWIZZLE_32B
ARL_DEVINL uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);            // bits [0,14)
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;   // bits [16,30)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;   // bits [32,46)
  d |= (uint64_t)1 << 46;                             // descriptor version (Blackwell)
  d |= (uint64_t)(layout_type & 7) << 61;             // bits [61,64)
  return d;
}

// kind::f16, bf16 x bf16 -> fp32.  a_mn / b_mn: 1 = MN-major operand, 0 = K-major.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn, int b_mn) {
  return (1u << 4)                      // D format F32
         | (1u << 7)                    // A format BF16
         | (1u << 10)                   // B format BF16
         | ((uint32_t)(a_mn & 1) << 15) // A major
         | ((uint32_t)(b_mn & 1) << 16) // B major
         | ((uint32_t)(N >> 3) << 17)   // N / 8
         | ((uint32_t)(M >> 4) << 24);  // M / 16
}

// swizzle mode from the byte width of one smem "row" (128 -> SW128, 64 -> SW64, 32 -> SW32)
__host__ __device__ constexpr uint32_t swz_layout_type(int row_bytes) {
  return row_bytes == 128 ? 2u : row_bytes == 64 ? 4u : row_bytes == 32 ? 6u : 0u;
}
// byte offset (within a 1024-aligned tile) of 16-byte chunk `c` of row `r`
template <int ROW_BYTES>
ARL_DEVINL uint32_t swz_off(uint32_t r, uint32_t c) {
  uint32_t off = r * ROW_BYTES + c * 16;
  constexpr uint32_t mask = ROW_BYTES == 128 ? 0x70u : ROW_BYTES == 64 ? 0x30u : ROW_BYTES == 32 ? 0x10u : 0u;
  return off ^ ((off >> 3) & mask);
}

// ---------------------------------------------------------------------------
// programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization attribute may
// start while its predecessor drains; pdl_wait() blocks until the predecessor grid has completed and its writes are
// visible.  Rule used everywhere: NOTHING that reads or writes global memory happens before pdl_wait() — only
// barrier init / TMEM allocation / shared-memory setup.  pdl_trigger() lets the successor begin its own prologue.
// Both are no-ops for a normally launched kernel.
// ---------------------------------------------------------------------------
ARL_DEVINL void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#ifdef ARL_NO_TRIGGER
ARL_DEVINL void pdl_trigger() {}
#else
ARL_DEVINL void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

// ---------------------------------------------------------------------------
// misc
// ---------------------------------------------------------------------------
ARL_DEVINL void st_shared_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

ARL_DEVINL uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

ARL_DEVINL float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
ARL_DEVINL float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

// Coalesced fp32 row stores for TMEM epilogues.  After tcgen05.ld (32x32b) every lane holds 32 consecutive columns of
// ITS OWN row; storing them directly makes each warp instruction touch 32 different 128-byte lines (32 LSU wavefronts,
// measured: the FC weight-gradient tile spent 8 200 of its 37 500 cycles in these stores).  Here the 32 x 32 block is
// transposed through a per-warp shared-memory scratch (32 rows x 36 floats: conflict-free for 16-byte accesses) so
// that one store instruction writes four rows x 128 contiguous bytes.  `dst` = this lane's row pointer (16-byte
// aligned, 32 floats stored), `valid` = whether the lane's row exists; scratch = 4608 bytes of shared memory per warp.
constexpr int kRowStoreScratch = 32 * 36 * 4;
ARL_DEVINL void store_rows32_coalesced(uint32_t scratch, const uint32_t (&v)[32], float* dst, bool valid, int lane) {
  const uint32_t my = scratch + (uint32_t)lane * 144u;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(my + 16u * j), "r"(v[4 * j]), "r"(v[4 * j + 1]),
                 "r"(v[4 * j + 2]), "r"(v[4 * j + 3])
                 : "memory");
  __syncwarp();
  const unsigned long long dp = reinterpret_cast<unsigned long long>(dst);
  const int sub = lane >> 3, ch = lane & 7;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = 4 * i + sub;
    const unsigned long long rp = __shfl_sync(0xffffffffu, dp, row);
    const int ok = __shfl_sync(0xffffffffu, valid ? 1 : 0, row);
    uint4 x;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w)
                 : "r"(scratch + (uint32_t)row * 144u + 16u * ch));
    if (ok) *reinterpret_cast<uint4*>(reinterpret_cast<float*>(rp) + 4 * ch) = x;
  }
  __syncwarp();
}

// the same for 16 words per lane (64-byte row pieces): one store instruction writes eight rows x 64 bytes.
// perm: the lane's 16-byte chunk j lands at chunk position j ^ perm of its piece (chunk-swizzled destinations).
ARL_DEVINL void store_rows16_coalesced(uint32_t scratch, const uint32_t (&v)[16], void* dst, bool valid, int lane,
                                       uint32_t perm = 0) {
  const uint32_t my = scratch + (uint32_t)lane * 80u;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(my + 16u * ((uint32_t)j ^ perm)), "r"(v[4 * j]),
                 "r"(v[4 * j + 1]), "r"(v[4 * j + 2]), "r"(v[4 * j + 3])
                 : "memory");
  __syncwarp();
  const unsigned long long dp = reinterpret_cast<unsigned long long>(dst);
  const int sub = lane >> 2, ch = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = 8 * i + sub;
    const unsigned long long rp = __shfl_sync(0xffffffffu, dp, row);
    const int ok = __shfl_sync(0xffffffffu, valid ? 1 : 0, row);
    uint4 x;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w)
                 : "r"(scratch + (uint32_t)row * 80u + 16u * ch));
    if (ok) *(reinterpret_cast<uint4*>(rp) + ch) = x;
  }
  __syncwarp();
}

// and the mirror image for loads: every lane ends up with the 64-byte piece at ITS pointer (chunk j taken from chunk
// position j ^ perm), fetched eight rows x 64 bytes per load instruction
ARL_DEVINL void load_rows16_coalesced(uint32_t scratch, uint32_t (&v)[16], const void* src, bool valid, int lane,
                                      uint32_t perm = 0) {
  const unsigned long long sp = reinterpret_cast<unsigned long long>(src);
  const int sub = lane >> 2, ch = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = 8 * i + sub;
    const unsigned long long rp = __shfl_sync(0xffffffffu, sp, row);
    const int ok = __shfl_sync(0xffffffffu, valid ? 1 : 0, row);
    uint4 x = make_uint4(0u, 0u, 0u, 0u);
    if (ok) x = __ldg(reinterpret_cast<const uint4*>(rp) + ch);
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(scratch + (uint32_t)row * 80u + 16u * ch), "r"(x.x),
                 "r"(x.y), "r"(x.z), "r"(x.w)
                 : "memory");
  }
  __syncwarp();
  const uint32_t my = scratch + (uint32_t)lane * 80u;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v[4 * j]), "=r"(v[4 * j + 1]), "=r"(v[4 * j + 2]), "=r"(v[4 * j + 3])
                 : "r"(my + 16u * ((uint32_t)j ^ perm)));
  __syncwarp();
}

ARL_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
ARL_DEVINL double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace arl
