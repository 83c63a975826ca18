// libaccelrl_b200.so — C ABI (include/accelrl_b200.h) over the sm_100a kernels.
// Single translation unit: kernels live in the .cuh files included below.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <set>
#include <vector>
#include <map>
#include <algorithm>

#include <nvtx3/nvToolsExt.h>

#include "../../include/accelrl_b200.h"
#include "common.cuh"
#include "gemm_tc.cuh"
#include "pconv.cuh"
#include "fcgemm.cuh"
#include "kernels.cuh"
#include "comm.cuh"

using namespace arl;

namespace {

std::string g_create_error;

// NVTX ranges on the phase boundaries of the path (serve / frame / fwd / sample / gae / grad /
// allreduce+adam): host-side markers around the launches (and around graph capture / replay), visible in any NVTX consumer
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

#define ARL_CHECK(ctx, call)                                                                    \
  do {                                                                                          \
    cudaError_t _e = (call);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(_e) + " @" + __FILE__ + ":" + \
                   std::to_string(__LINE__);                                                    \
      return 1;                                                                                 \
    }                                                                                           \
  } while (0)

#define ARL_FAIL(ctx, msg)  \
  do {                      \
    (ctx)->err = (msg);     \
    return 2;               \
  } while (0)

// One conv layer as the GEMM tiles see it.  Layer 0 is evaluated over the bf16 space-to-depth(stride)
// image of the observation: an 8x8/4 conv over (C,104,80) == a 2x2/1 conv over (26,20,C*16).
struct ConvLayer {
  // reference view (parameters, gradients)
  int Cin, Hin, Win, Cout, k, s, p, Ho, Wo;
  // tile view (NHWC source)
  int gC, gH, gW, gk, gs, gp;   // source channels / dims, taps per side, stride, pad
  int K;                        // gC*gk*gk == Cin*k*k
  long off_W, off_b;            // flat param offsets
  __nv_bfloat16* wpack;         // [Cout][K] forward operand
  struct DClass {               // dgrad stride-parity classes (layers >= 1)
    int ry, rx, qy0, qx0, Qh, Qw, Ty, Tx, K;
    __nv_bfloat16* wpack;       // [Cin][K]
  };
  std::vector<DClass> dclasses;
  __nv_bfloat16* act;           // [max_rows*Ho*Wo][Cout]
  __nv_bfloat16* dact;
};

// Geometry of one conv layer on the patch-resident path (pconv.cuh): the layer seen as a stride-1 T x T conv over
// a position grid of P planes x 64 channels.
struct PcLayer {
  int T = 0, P = 0, Wp = 0, Hc = 0, S = 0, N = 0;
  int ntaps = 0, shift[kPcMaxTaps] = {0};
  int load_rows = 0, tiles_per_img = 0;
  int s = 1, pad = 0;            // space-to-depth factor / padding of the input grid (how the layer below stores into it)
  int ci_major = 0;              // channel order inside a cell: 1 = (ci, py, px) [layer 0: frame kernel], 0 = (py, px, ci)
  __nv_bfloat16* in = nullptr;   // input grid [P][rows][64] (layers >= 1; layer 0 reads the obs16 buffers)
  long in_rows = 0;              // rows per plane (max_rows * S + slack)
  __nv_bfloat16* wpack = nullptr;
  // backward
  int dYpad = 0;
  __nv_bfloat16* dY = nullptr;   // gradient w.r.t. this layer's output, in this layer's grid at offset dYpad (layer 0: pixel grid)
  __nv_bfloat16* dY_base = nullptr;   // layer 0: the allocation (dY = dY_base + 64 rows)
  long dY_rows = 0;
  __nv_bfloat16* dwpack = nullptr;  // data-gradient weights [T*T*Pout][P*64][64]
  int stages_fwd = 0, stages_dgrad = 0;
};

constexpr int kSsFinCap = 4096;   // finalize blocks that own elements (preset 1: ~320)
constexpr int kSsFcCap = 4096;    // FC weight-gradient CTAs x 8 epilogue warps (preset 1: 54 x 8)
constexpr int kStreamPartials = 4096;   // block partials of update_stream_kernel (preset 1: 317 + 592 blocks)

struct TrainPlan {
  int n = 0;
  std::vector<int> conv_splits, conv_rps;
  int head_groups = 0, head_rpg = 0;
  GradJob* jobs_dev = nullptr;
  int n_jobs = 0;
  int fin_blocks = 1;
  int n_ss = 0;                 // per-block sum-of-squares slots the finalize kernel fills (GradJob::ss_off)
  long fin_total = 0;           // elements of all jobs
  std::vector<long> job_total;  // rows * cols per job
  int early_jobs = 0;           // > 0: jobs [2, n_jobs) are finalised on the side streams as soon as their partials exist;
                                // only layer 0's two jobs are left for the end of the chain
};

}  // namespace

struct arl_ctx {
  std::string err;
  arl_net_cfg cfg{};
  std::vector<ConvLayer> conv;
  std::vector<PcLayer> pc;             // patch-resident conv path (empty: geometry not supported -> gather path)
  // TMA-fed FC tiles (fcgemm.cuh), used together with the patch-resident conv path (pc_mode == 2)
  __nv_bfloat16* act_fc = nullptr;     // [HW + 2][fc_rows][64] last conv output, one plane per pixel
  __nv_bfloat16* wfc_t = nullptr;      // [HW][H/64][64][64] FC weight tiles
  __nv_bfloat16* dh_t = nullptr;       // [H/64][fc_rows][64]
  int fc_rows = 0;                     // rows per plane (max_rows rounded up to 128)
  int dh_n = 0;                        // rows of dh_t that may be non-zero
  cudaStream_t side2 = nullptr;        // second forked stream: conv weight gradients (side: head + FC weight gradients)
  cudaEvent_t ev_join2 = nullptr;
  cudaStream_t side = nullptr;         // weight-gradient kernels run here, overlapping the data-gradient chain
  cudaEvent_t ev_fork[4] = {nullptr, nullptr, nullptr, nullptr}, ev_join = nullptr;
  cudaEvent_t ev_fin[2] = {nullptr, nullptr};   // "conv weight-gradient partials of layer l are complete" (side2 -> side)
  bool no_fork = false;                // serialise everything on the caller's stream (per-kernel profiling)
  // CTA caps while a data-gradient and a weight-gradient kernel share the GPU (0 = all SMs).  Measured (B200, C2):
  // wgrad capped at 56..96 CTAs lets the concurrent dgrad chain start on the free SMs and shrinks the per-CTA partial
  // traffic: 62.6 -> 60.4 ms per iteration; capping the dgrad side as well does not help.  Re-swept after the coalesced
  // epilogues (ms per iteration at 24/32/40/48/64/80/100/148 CTAs: 56.6/56.6/55.1/54.9/55.7/55.5/56.9/59.1): 48.
  int dgrad_ctas = 0, wgrad_ctas = 48;
  int pc_dy_n = 0;                     // images whose gradient-grid rows may be non-zero
  int pc_mode = 0;                     // 0: gather path   1: pconv forward (inference)   2: pconv forward + backward
  int Kfc = 0, H = 0, A = 0, HWlast = 0, Clast = 0;
  long off_Wfc = 0, off_bfc = 0, off_Wpi = 0, off_bpi = 0, off_Wv = 0, off_bv = 0, n_params = 0;
  std::vector<long> lay_off, lay_size;
  long obs16_elems = 0;                // bf16 elements of one space-to-depth observation
  __nv_bfloat16* obs16_stage = nullptr;  // [max_rows] converted inputs (callers that hand in uint8 obs)
  __nv_bfloat16* wfc_bf16 = nullptr;   // [Kfc][H] bf16 copy of the FC weights in the reference's row order
  PackJob* pack_jobs_dev = nullptr;
  int n_pack_jobs = 0;
  long conv_pack_end = 0;              // > 0: every conv pack job is a pconv tap-tile pack; conv weights end here (flat index)
  unsigned long long* pk_slots = nullptr;   // [conv_pack_end][2] bf16 slot addresses of each conv weight (0 = none)
  // bound vectors
  float *params = nullptr, *grad = nullptr, *m = nullptr, *v = nullptr;
  // workspaces
  float* fc_partial = nullptr;
  long fc_partial_cap = 0;  // floats
  __nv_bfloat16 *h = nullptr, *dh = nullptr;
  float* dlogit = nullptr;
  std::vector<float*> wgrad_partial;  // per conv layer
  std::vector<long> wgrad_partial_cap;
  std::vector<float*> bias_partial;   // per conv layer [splits][Cout]
  float* head_partial = nullptr;   // [G][H][A+2]
  float* head_b_partial = nullptr; // [G][A+1]
  float* loss_partial = nullptr;   // [kLossBlocks][4]
  double* sumsq_partial = nullptr;
  double* sumsq_partial_fc = nullptr;  // update_range_kernel's per-block sums of squares (early FC update)
  double* ss_fin = nullptr;            // [kSsFinCap] finalize_grads_kernel's per-block sums of squares
  double* ss_fc = nullptr;             // [kSsFcCap]  FC weight-gradient tiles' per-warp sums of squares
  int pending_ss_fin = 0, pending_ss_fc = 0;   // > 0: this minibatch's global-norm partials came from the producers
  bool train_step_active = false;      // grad_minibatch is followed by the local clip_update (train_minibatches, sync == 0)
  bool in_grad_fwd = false;            // forward_trunk called by grad_minibatch: conv0's gathered rows are read again (wgrad)
  bool early_fc_done = false;          // this minibatch's FC weights were updated by update_range_kernel
  const void* pending_fin = nullptr;   // TrainPlan whose gradient finalisation clip_update must fold into its update kernel
  bool pending_stream = false;         // ... and the folding kernel is update_stream_kernel (no clipping: no barrier)
  cudaEvent_t ev_fcd = nullptr;        // "FC data gradient has read the FC weights"
  unsigned long long* ticket = nullptr; // grid-barrier ticket of update_fused_kernel
  // L2 residency of the optimiser state (ARL_L2_PERSIST): access-policy window over [params .. v] for the update kernel
  cudaAccessPolicyWindow l2win{};
  bool l2_on = false;
  size_t l2_carve = 0;                  // ARL_L2_PERSIST=3: the carve-out exists only while minibatches train
  bool l2_toggle = false, l2_carved = false, l2_grad_in_window = false;
  // split update (clip_update, ARL_SPLIT_UPDATE): the FC range's step runs on `side` beside the next minibatch's conv layers
  cudaEvent_t ev_upd_fork = nullptr, ev_updB = nullptr;
  bool split_capture = false;           // inside train_minibatches' graph capture with the local update
  bool updB_pending = false;            // the next FC forward (or the end of the capture) joins ev_updB
  float* hyper = nullptr;          // [0] lr_mult
  int* step = nullptr;             // Adam t
  int* log_slot = nullptr;
  float *log_norm = nullptr, *log_loss = nullptr;
  int log_cap = 4096;
  int* mb_counter = nullptr;       // minibatch index for graph-replayed training
  float* valid_count = nullptr;
  std::map<int, TrainPlan> plans;
  // training inputs
  const uint8_t* t_obs = nullptr; const uint8_t* t_act = nullptr; const float* t_adv = nullptr;
  const float* t_ret = nullptr; const float* t_oldv = nullptr; const float* t_oldp = nullptr;
  const int8_t* t_valids = nullptr; long t_rows = 0;
  arl_opt_cfg opt{};
  bool opt_set = false;
  // sampler
  arl_sampler_cfg sc{};
  bool sampler_set = false;
  EnvState est{};
  TrajOut tout{};
  FrameCmd* cmd = nullptr;
  int* rows_tab = nullptr;             // [T][B] row indices e*T+s
  __nv_bfloat16* step_obs16 = nullptr; // [B] bf16 space-to-depth mirror of step_obs
  __nv_bfloat16* roll_obs16 = nullptr; // [N] mirror of observations
  cudaGraphExec_t rollout_graph = nullptr;
  // the fields above are the ACTIVE sampler; arl_sampler_select parks them here and loads the other slot
  // (slot 0: training sampler, slot 1: evaluation sampler — its own envs, step buffer and trajectory records)
  struct SamplerSlot {
    arl_sampler_cfg sc{};
    bool set = false;
    EnvState est{};
    TrajOut tout{};
    FrameCmd* cmd = nullptr;
    int* rows_tab = nullptr;
    __nv_bfloat16* step_obs16 = nullptr;
    __nv_bfloat16* roll_obs16 = nullptr;
    cudaGraphExec_t rollout_graph = nullptr;
    long graph_rollout_nodes = 0;
  };
  SamplerSlot slots[2];
  int cur_slot = 0;
  // training graph cache
  cudaGraphExec_t train_graph = nullptr;
  const int* train_graph_idx = nullptr;
  int train_graph_mb = 0;
  int train_graph_sync = 0;
  int train_graph_per = 1;
  bool sync_graph_failed = false;
  long launches = 0;
  // per-kernel CUDA-event profiling (arl_profile_*): events recorded after each launch when enabled
  bool prof_on = false;
  bool prof_collect = false;             // record one label per kernel launch (arl_profile_graph)
  bool prof_timeline = false;            // stamp %globaltimer after every launch INSIDE the captured, forked minibatch graph
  unsigned long long* tl_buf = nullptr;  // [96] stamps
  std::vector<std::string> prof_labels;
  std::vector<cudaEvent_t> prof_ev;
  std::vector<std::string> prof_names;
  int prof_n = 0;
  long graph_rollout_nodes = 0, graph_train_nodes = 0;
  int n_loss_rows = 0;                 // rows of the last head_kernel<1> launch (loss partial count)
  float lr_mult_host = 1.f;
  CommState comm;
  // overlapped synchronous step (comm.cuh): the FC slice exchange runs on `cs` beside the conv gradient chain
  cudaStream_t cs = nullptr;
  cudaEvent_t ev_cs_in[2] = {nullptr, nullptr}, ev_cs_done = nullptr;
  bool sync_overlap_active = false;
  double* sync_tail_partial = nullptr;
  bool shadow_in_comm = false;         // wfc_t / wfc_bf16 lives inside comm's symmetric allocation (freed with it)
  AsyncState async_;
};

namespace {

constexpr int kMaxSplits = 160;

template <class T>
int dev_alloc(arl_ctx* c, T** p, size_t count) {
  ARL_CHECK(c, cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T)));
  ARL_CHECK(c, cudaMemset(*p, 0, count * sizeof(T)));
  return 0;
}

int roundup(int x, int m) { return (x + m - 1) / m * m; }

// Kernel launch with the programmatic-dependent-launch attribute (see common.cuh: pdl_wait / pdl_trigger): inside the
// captured graphs the next kernel's prologue may overlap this kernel's tail.  Off by default; ARL_PDL=1 enables.
bool g_pdl = false;   // measured on B200: no gain inside CUDA graphs (72.2 vs 73.1 ms/iter), early trigger slower (76.0)
// ARL_CARVEOUT=<percent>: one preferred shared-memory carve-out for every kernel launched through launch_k (an SM whose
// consecutive kernels ask for different L1 / shared splits has to drain and reconfigure between them)
int g_carveout = -1;
inline void apply_carveout(const void* kern) {
  if (g_carveout < 0) return;
  static std::set<const void*> done;
  if (done.insert(kern).second) cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, g_carveout);
}

template <class... KArgs, class... Args>
cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  apply_carveout(reinterpret_cast<const void*>(kern));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = g_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(std::forward<Args>(args))...);
}

// launch_k with an L2 access-policy window (captured into the graph's kernel node like any launch attribute)
template <class... KArgs, class... Args>
cudaError_t launch_k_win(const cudaAccessPolicyWindow* win, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                         cudaStream_t st, Args&&... args) {
  if (!win) return launch_k(kern, grid, block, smem, st, std::forward<Args>(args)...);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeAccessPolicyWindow;
  at[0].val.accessPolicyWindow = *win;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(std::forward<Args>(args))...);
}

// Cooperative launch (also inside stream capture): the driver guarantees that every block of the grid is resident at
// the same time, which is what a kernel with a grid-wide barrier needs.  Falls back to a plain launch if the driver
// refuses the attribute (the grids used here are sized to one wave anyway).
template <class... KArgs, class... Args>
cudaError_t launch_coop(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  static bool coop_ok = true;
  if (coop_ok) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;
    at[0].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
    if (e == cudaSuccess) return e;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cs);
    if (cs != cudaStreamCaptureStatusNone) return e;      // a failed call inside capture invalidates it: report
    cudaGetLastError();
    coop_ok = false;
  }
  return launch_k(kern, grid, block, smem, st, std::forward<Args>(args)...);
}

__global__ void stamp_kernel(unsigned long long* slot) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  *slot = t;
}

// The main chain of a captured training graph runs at the highest stream priority (kernel nodes inherit it), the forked
// weight-gradient / exchange streams at the default (lowest): when an SM frees up, pending CTAs of the data-gradient chain
// — the critical path — are placed before those of the side kernels, which have slack until the join.
cudaError_t create_main_stream(cudaStream_t* s) {
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  static const bool prio = !(getenv("ARL_STREAM_PRIO") && atoi(getenv("ARL_STREAM_PRIO")) == 0);
  return cudaStreamCreateWithPriority(s, cudaStreamNonBlocking, prio ? hi : lo);
}

// record an event after the launch that just happened (profiling mode only)
void prof_mark(arl_ctx* c, const char* name, cudaStream_t st) {
  if (c->prof_collect) c->prof_labels.push_back(name);
  if (c->prof_timeline) {
    // inside a captured graph events carry no timestamps: a one-thread kernel writes %globaltimer behind the launch
    if (c->prof_n < 96 && c->tl_buf) {
      if ((int)c->prof_names.size() <= c->prof_n) c->prof_names.resize(c->prof_n + 1);
      c->prof_names[c->prof_n] = name;
      stamp_kernel<<<1, 1, 0, st>>>(c->tl_buf + c->prof_n);
      c->prof_n++;
    }
    return;
  }
  if (!c->prof_on) return;
  if (c->prof_n >= (int)c->prof_ev.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    c->prof_ev.push_back(e);
    c->prof_names.push_back("");
  }
  c->prof_names[c->prof_n] = name;
  cudaEventRecord(c->prof_ev[c->prof_n], st);
  c->prof_n++;
}

// ---------------------------------------------------------------------------
// launch helpers
// ---------------------------------------------------------------------------
template <class ALoad, bool BNM, int BN>
int launch_rowgemm(arl_ctx* c, ALoad a, WeightSrc b, RowEpi e, int M, int Ntot, int num_kb, int kbps, int splits,
                   cudaStream_t st) {
  using Cfg = RowGemmCfg<BN>;
  static bool attr = false;
  if (!attr) {
    ARL_CHECK(c, cudaFuncSetAttribute(rowgemm_kernel<ALoad, BNM, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      Cfg::SMEM));
    attr = true;
  }
  dim3 grid((M + 127) / 128, Ntot / BN, splits);
  ARL_CHECK(c, launch_k(rowgemm_kernel<ALoad, BNM, BN>, dim3(grid), dim3(kGemmThreads), Cfg::SMEM, st, a, b, e, num_kb, kbps));
  c->launches++;
  ARL_CHECK(c, cudaGetLastError());
  return 0;
}

template <class ALoad, bool BNM>
int launch_rowgemm_bn(arl_ctx* c, int BN, ALoad a, WeightSrc b, RowEpi e, int M, int Ntot, int num_kb, int kbps,
                      int splits, cudaStream_t st) {
  switch (BN) {
    case 16: if constexpr (!BNM) return launch_rowgemm<ALoad, BNM, 16>(c, a, b, e, M, Ntot, num_kb, kbps, splits, st); break;
    case 32: if constexpr (!BNM) return launch_rowgemm<ALoad, BNM, 32>(c, a, b, e, M, Ntot, num_kb, kbps, splits, st); break;
    case 64: return launch_rowgemm<ALoad, BNM, 64>(c, a, b, e, M, Ntot, num_kb, kbps, splits, st);
    case 128: return launch_rowgemm<ALoad, BNM, 128>(c, a, b, e, M, Ntot, num_kb, kbps, splits, st);
  }
  ARL_FAIL(c, "unsupported tile width BN=" + std::to_string(BN));
}

template <int BN>
int launch_conv_persist(arl_ctx* c, const RowGemmMulti<ConvLoader<128>>& p, int ncls, int K, int max_tiles,
                        cudaStream_t st) {
  const int smem = conv_persist_smem<BN>(K);
  static int attr_smem = 0;
  static int occ = 1;
  if (smem > attr_smem) {
    ARL_CHECK(c, cudaFuncSetAttribute(conv_gemm_persist_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_smem = smem;
  }
  // two CTAs per SM when shared memory allows (registers are capped for 2 by __launch_bounds__); the tile loop is
  // correct for any grid size, so this is only a sizing heuristic
  occ = (2 * (smem + 1024) <= 227 * 1024) ? 2 : 1;
  int ctas = std::min(max_tiles, std::max(1, 148 * occ / ncls));
  dim3 grid(ctas, ncls, 1);
  ARL_CHECK(c, launch_k(conv_gemm_persist_kernel<BN>, dim3(grid), dim3(kPersistThreads), smem, st, p));
  c->launches++;
  ARL_CHECK(c, cudaGetLastError());
  return 0;
}

int launch_conv_persist_bn(arl_ctx* c, int BN, const RowGemmMulti<ConvLoader<128>>& p, int ncls, int K, int max_tiles,
                           cudaStream_t st) {
  switch (BN) {
    case 16: return launch_conv_persist<16>(c, p, ncls, K, max_tiles, st);
    case 32: return launch_conv_persist<32>(c, p, ncls, K, max_tiles, st);
    case 64: return launch_conv_persist<64>(c, p, ncls, K, max_tiles, st);
  }
  ARL_FAIL(c, "unsupported conv tile width BN=" + std::to_string(BN));
}

template <class ALoad, int MT, int BN>
int launch_wgrad(arl_ctx* c, ALoad a, const __nv_bfloat16* dy, int ld_dy, int nrows, int rps, int splits, int atoms,
                 int ntiles, WgradEpi e, cudaStream_t st) {
  using Cfg = WgradCfg<MT, BN>;
  static bool attr = false;
  if (!attr) {
    ARL_CHECK(c, cudaFuncSetAttribute(wgrad_kernel<ALoad, MT, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      Cfg::SMEM_TOTAL));
    attr = true;
  }
  dim3 grid((atoms + MT * 2 - 1) / (MT * 2), ntiles, splits);
  ARL_CHECK(c, launch_k(wgrad_kernel<ALoad, MT, BN>, dim3(grid), dim3(kWgradThreads), Cfg::SMEM_TOTAL, st, a, dy, ld_dy, nrows, rps, atoms, e));
  c->launches++;
  ARL_CHECK(c, cudaGetLastError());
  return 0;
}

template <class ALoad>
int launch_wgrad_conv(arl_ctx* c, ALoad a, const __nv_bfloat16* dy, int Cout, int nrows, int rps, int splits,
                      int atoms, WgradEpi e, cudaStream_t st) {
  int mt_need = (atoms + 1) / 2;
#define ARL_WG(MT, BN) return launch_wgrad<ALoad, MT, BN>(c, a, dy, Cout, nrows, rps, splits, atoms, 1, e, st)
  if (Cout == 16) {
    if (mt_need <= 1) ARL_WG(1, 16);
    if (mt_need <= 2) ARL_WG(2, 16);
    if (mt_need <= 4) ARL_WG(4, 16);
    ARL_WG(5, 16);
  } else if (Cout == 32) {
    if (mt_need <= 2) ARL_WG(2, 32);
    if (mt_need <= 4) ARL_WG(4, 32);
    ARL_WG(5, 32);
  } else if (Cout == 64) {
    if (mt_need <= 2) ARL_WG(2, 64);
    if (mt_need <= 4) ARL_WG(4, 64);
    ARL_WG(5, 64);
  }
#undef ARL_WG
  ARL_FAIL(c, "unsupported conv filter count " + std::to_string(Cout));
}

// ---------------------------------------------------------------------------
// network planning
// ---------------------------------------------------------------------------
int plan_net(arl_ctx* c) {
  const arl_net_cfg& f = c->cfg;
  if (f.n_conv < 1 || f.n_conv > ARL_MAX_CONV) ARL_FAIL(c, "n_conv out of range");
  if (f.n_actions < 1 || f.n_actions > kMaxActions) ARL_FAIL(c, "n_actions must be in [1,18]");
  int C = f.in_c, Hh = f.in_h, Ww = f.in_w;
  long off = 0;
  for (int l = 0; l < f.n_conv; ++l) {
    ConvLayer L{};
    L.Cin = C; L.Hin = Hh; L.Win = Ww;
    L.Cout = f.conv_filters[l]; L.k = f.conv_sizes[l]; L.s = f.conv_strides[l]; L.p = f.conv_pads[l];
    L.Ho = (Hh + 2 * L.p - L.k) / L.s + 1;
    L.Wo = (Ww + 2 * L.p - L.k) / L.s + 1;
    L.K = L.Cin * L.k * L.k;
    if (L.K % 64) ARL_FAIL(c, "conv layer " + std::to_string(l) + ": Cin*k*k must be a multiple of 64");
    if (L.Cout != 16 && L.Cout != 32 && L.Cout != 64)
      ARL_FAIL(c, "conv layer " + std::to_string(l) + ": filters must be 16, 32 or 64");
    if (l == 0) {
      if (L.s != 4 || L.k != 2 * L.s || L.p != 0 || (Hh % L.s) || (Ww % L.s))
        ARL_FAIL(c, "first conv layer must be 8x8, stride 4, pad 0 on dims divisible by 4 (space-to-depth path)");
      L.gC = L.Cin * L.s * L.s; L.gH = Hh / L.s; L.gW = Ww / L.s; L.gk = 2; L.gs = 1; L.gp = 0;
      c->obs16_elems = (long)L.gH * L.gW * L.gC;
    } else {
      if (L.Cin % 8) ARL_FAIL(c, "conv input channels must be a multiple of 8");
      L.gC = L.Cin; L.gH = Hh; L.gW = Ww; L.gk = L.k; L.gs = L.s; L.gp = L.p;
    }
    L.off_W = off; off += (long)L.Cout * L.Cin * L.k * L.k;
    L.off_b = off; off += L.Cout;
    c->lay_off.push_back(L.off_W); c->lay_size.push_back((long)L.Cout * L.Cin * L.k * L.k);
    c->lay_off.push_back(L.off_b); c->lay_size.push_back(L.Cout);
    if (l >= 1) {
      int Ty = (L.k + L.s - 1) / L.s;
      for (int ry = 0; ry < L.s; ++ry)
        for (int rx = 0; rx < L.s; ++rx) {
          ConvLayer::DClass d{};
          d.ry = ry; d.rx = rx; d.Ty = Ty; d.Tx = Ty;
          // u = i + p = s*q + r,  i in [0, Hin)
          auto qrange = [&](int r, int n, int& q0, int& cnt) {
            int lo = L.p - r;                       // s*q >= lo
            q0 = lo <= 0 ? 0 : (lo + L.s - 1) / L.s;
            int hi = n - 1 + L.p - r;               // s*q <= hi
            int q1 = hi < 0 ? -1 : hi / L.s;
            cnt = q1 - q0 + 1;
          };
          qrange(ry, L.Hin, d.qy0, d.Qh);
          qrange(rx, L.Win, d.qx0, d.Qw);
          d.K = d.Ty * d.Tx * L.Cout;
          if (d.K % 64) ARL_FAIL(c, "conv dgrad K not a multiple of 64");
          if (d.Qh > 0 && d.Qw > 0) L.dclasses.push_back(d);
        }
      if ((int)L.dclasses.size() > kMaxMulti) ARL_FAIL(c, "conv stride > 2 is not supported in the backward pass");
    }
    c->conv.push_back(L);
    C = L.Cout; Hh = L.Ho; Ww = L.Wo;
  }
  c->Clast = C; c->HWlast = Hh * Ww; c->Kfc = C * Hh * Ww;
  c->H = f.hidden; c->A = f.n_actions;
  if (c->Kfc % 64) ARL_FAIL(c, "flattened conv output must be a multiple of 64");
  if (c->H % 256) ARL_FAIL(c, "hidden size must be a multiple of 256");
  if (c->H > 512) ARL_FAIL(c, "hidden size must be <= 512");
  c->off_Wfc = off; off += (long)c->Kfc * c->H;
  c->off_bfc = off; off += c->H;
  c->off_Wpi = off; off += (long)c->H * c->A;
  c->off_bpi = off; off += c->A;
  c->off_Wv = off; off += c->H;
  c->off_bv = off; off += 1;
  c->n_params = off;
  long tail_off[6] = {c->off_Wfc, c->off_bfc, c->off_Wpi, c->off_bpi, c->off_Wv, c->off_bv};
  long tail_sz[6] = {(long)c->Kfc * c->H, c->H, (long)c->H * c->A, c->A, c->H, 1};
  for (int i = 0; i < 6; ++i) { c->lay_off.push_back(tail_off[i]); c->lay_size.push_back(tail_sz[i]); }
  return 0;
}


// ---------------------------------------------------------------------------
// patch-resident conv path (pconv.cuh): geometry, buffers, launches
// ---------------------------------------------------------------------------
// Fills c->pc when every conv layer fits the path's constraints (otherwise c->pc stays empty and the
// gather path of gemm_tc.cuh serves the network).
int plan_pconv(arl_ctx* c) {
  std::vector<PcLayer> pcs;
  const int nl = (int)c->conv.size();
  for (int l = 0; l < nl; ++l) {
    const ConvLayer& L = c->conv[l];
    PcLayer q{};
    q.s = L.s; q.pad = L.p; q.N = L.Cout;
    q.T = (L.k + L.s - 1) / L.s;
    int cellc = L.Cin * L.s * L.s;
    if (cellc % 64 || q.T * q.T > kPcMaxTaps) return 0;
    q.P = cellc / 64;
    if (l == 0) {
      if (L.p != 0 || (L.Hin % L.s) || (L.Win % L.s) || q.P != 1) return 0;
      if (L.Cout != 32 && L.Cout != 64) return 0;
      q.ci_major = 1;
      q.Wp = L.Win / L.s; q.Hc = L.Hin / L.s;
      q.S = roundup(q.Wp * q.Hc, 8);       // image stride in positions: tiles must start at multiples of 8 rows
                                            // (the padding positions are never written: zero)
      q.tiles_per_img = ((L.Ho - 1) * q.Wp + L.Wo + 127) / 128;
      // the weight gradient multiplies whole 128-position tiles of the gradient grid: a tile must not run into the next
      // image's rows (in the forward pass those outputs are simply dropped)
      if (q.tiles_per_img * 128 > q.S) q.S = q.tiles_per_img * 128;
    } else {
      if (L.Cout != 64 || q.P > 2) return 0;
      const bool shared = (L.s == 1 && L.p >= 1 && 2 * L.p == q.T - 1);   // 'same' conv: pad column/row shared with the neighbour
      q.Wp = shared ? L.Wo + L.p : L.Wo + q.T - 1;
      q.Hc = shared ? L.Ho + L.p : L.Ho + q.T - 1;
      q.S = q.Wp * q.Hc;
      // every input pixel the layer reads must land inside the grid
      if ((L.Hin - 1 + L.p) / L.s >= q.Hc + (shared ? 1 : 0) && !shared) return 0;
      q.dYpad = q.T - 1 - (L.s == 1 ? L.p : 0);
      if (q.dYpad < 0 || L.Ho + q.dYpad > q.Hc || L.Wo + q.dYpad > q.Wp) return 0;
      if (L.s > 1 && l != 1) return 0;                         // strided layers above layer 1: no unfold target
      if (L.s != 1 && L.s != 2 && L.s != 4) return 0;           // space-to-depth factors are decoded with shifts
      if (l == 1 && c->conv[0].Cout != 32) return 0;           // unfold epilogue: 32-channel pixels below
      if (l == 1 && L.s * L.s * c->conv[0].Cout != q.P * 64) return 0;
    }
    q.ntaps = q.T * q.T;
    for (int ty = 0; ty < q.T; ++ty)
      for (int tx = 0; tx < q.T; ++tx) q.shift[ty * q.T + tx] = ty * q.Wp + tx;
    q.load_rows = roundup(128 + (q.T - 1) * q.Wp + (q.T - 1), 8);
    pcs.push_back(q);
  }
  // shared-memory budget: weights resident + at least 2 patch stages
  for (int l = 0; l < nl; ++l) {
    PcLayer& q = pcs[l];
    for (q.stages_fwd = 4; q.stages_fwd >= 2; --q.stages_fwd)
      if (pc_fwd_smem(q.N, q.ntaps, q.P, q.load_rows, q.stages_fwd) <= 227 * 1024) break;
    if (q.stages_fwd < 2) return 0;
    if (l >= 1) {
      int Pout = q.N / 64;
      for (q.stages_dgrad = 4; q.stages_dgrad >= 2; --q.stages_dgrad)
        if (pc_fwd_smem(q.P * 64, q.ntaps, Pout, q.load_rows, q.stages_dgrad) <= 227 * 1024) break;
      if (q.stages_dgrad < 2) return 0;
    }
  }
  c->pc = pcs;
  return 0;
}

int alloc_pconv(arl_ctx* c, std::vector<PackJob>& pj) {
  const int R = c->cfg.max_rows;
  for (size_t l = 0; l < c->pc.size(); ++l) {
    PcLayer& q = c->pc[l];
    const ConvLayer& L = c->conv[l];
    if (l >= 1) {
      q.in_rows = (long)R * q.S + q.load_rows + 256;
      if (dev_alloc(c, &q.in, (size_t)q.P * q.in_rows * 64)) return 1;
      q.dY_rows = (long)R * q.S + q.load_rows + 256;
      if (dev_alloc(c, &q.dY, (size_t)(q.N / 64) * q.dY_rows * 64)) return 1;
    } else {
      // gradient w.r.t. the first layer's output PIXELS, position-aligned with the first layer's grid; 64 zero rows in
      // front: the wide weight-gradient tiles read dY[q - tx] from one row before the first position
      q.dY_rows = (long)R * q.S + q.load_rows + 256;
      if (dev_alloc(c, &q.dY_base, (size_t)(q.dY_rows + 64) * q.N)) return 1;
      q.dY = q.dY_base + (size_t)64 * q.N;
    }
    if (dev_alloc(c, &q.wpack, (size_t)q.ntaps * q.P * q.N * 64)) return 1;
    PackJob j{};
    j.dst = q.wpack; j.src_off = L.off_W; j.kind = PK_PCONV; j.rows = q.ntaps * q.P * q.N; j.cols = 64;
    j.Cout = L.Cout; j.C = L.Cin; j.kh = L.k; j.kw = L.k; j.s = L.s; j.T = q.T; j.P = q.P; j.N = q.N; j.ci_major = q.ci_major;
    pj.push_back(j);
    if (l >= 1) {
      const int Pout = q.N / 64, Nd = q.P * 64;
      if (dev_alloc(c, &q.dwpack, (size_t)q.ntaps * Pout * Nd * 64)) return 1;
      PackJob d = j;
      d.dst = q.dwpack; d.kind = PK_PCONV_DGRAD; d.rows = q.ntaps * Pout * Nd; d.P = Pout; d.N = Nd;
      pj.push_back(d);
    }
  }
  return 0;
}

template <int N, int TW = 1>
int launch_pconv(arl_ctx* c, const PcParams& p, cudaStream_t st) {
  const int smem = pc_fwd_smem(N, p.ntaps, p.planes, p.load_rows, p.stages, 0, TW > 1);
  if (smem > 227 * 1024) ARL_FAIL(c, "pconv forward: stages do not fit shared memory");
  static int attr_smem = 0;
  if (smem > attr_smem) {
    ARL_CHECK(c, (cudaFuncSetAttribute(pconv_fwd_kernel<N, TW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)));
    attr_smem = smem;
  }
  int ctas = std::min(p.ntiles, 148);
  if (p.out.mode != 0 && c->dgrad_ctas > 0) ctas = std::min(ctas, c->dgrad_ctas);   // data gradient sharing the GPU with a wgrad
  ARL_CHECK(c, (launch_k(pconv_fwd_kernel<N, TW>, dim3(ctas), dim3(kPcFwdThreads), smem, st, p)));
  c->launches++;
  ARL_CHECK(c, cudaGetLastError());
  return 0;
}

int launch_pconv_n(arl_ctx* c, int N, const PcParams& p, cudaStream_t st, int TW = 1) {
  if (TW == 2 && N == 32) return launch_pconv<32, 2>(c, p, st);
  if (TW == 2 && N == 64) return launch_pconv<64, 2>(c, p, st);
  if (TW == 3 && N == 64) return launch_pconv<64, 3>(c, p, st);
  if (TW != 1) ARL_FAIL(c, "unsupported wide pconv tile");
  switch (N) {
    case 32: return launch_pconv<32>(c, p, st);
    case 64: return launch_pconv<64>(c, p, st);
    case 128: return launch_pconv<128>(c, p, st);
  }
  ARL_FAIL(c, "unsupported pconv tile width N=" + std::to_string(N));
}

// Wide forward / data-gradient tiles (pconv.cuh, PcParams): the taps of a filter row on the N axis.  ARL_FWD_WIDE=0 keeps the
// one-tap-per-MMA tiles (A/B).  Returns TW (1: not applicable) and rewrites the tiling of p for 120-row tiles.
int pconv_make_wide(arl_ctx* c, const PcLayer& q, int N, int mode, PcParams& p, int n, int which) {
  // ARL_FWD_WIDE: bit mask of the launches that take the wide form — bits 0..2: forward of conv layer 0..2, bit 3: data
  // gradients.  OFF by default (mask 0).  Parity-green (tests/test_gpu_kernels.py runs its oracle comparisons under mask 15)
  // and the MMA burst of a tile does shrink as modelled (conv0: 1 180 -> 700 cycles, device trace), but the tile period is
  // then set by the epilogue chain (accumulator wait -> tcgen05.ld -> exchange barrier -> shuffles -> stores -> next tile:
  // ~1 650 cycles against ~1 100 for the classic tile, which is itself only just MMA-bound): conv0/1/2 forward 12.5 / 8.4 /
  // 8.7 -> 18.8 / 9.6 / 11.5 us, 51.4 -> 54.6 ms per iteration.  Needs a second epilogue warp set before it can pay.
  static const int mask = getenv("ARL_FWD_WIDE") ? atoi(getenv("ARL_FWD_WIDE")) : 0;
  const bool on = (mask >> which) & 1;
  const int TW = q.T;
  if (!on || mode == 2 || p.u8.obs || TW < 2 || TW > 3 || TW * N > 256 || (N != 32 && N != 64)) return 1;
  if (TW == 3 && N != 64) return 1;
  for (int st = p.stages; st >= 2; --st)
    if (pc_fwd_smem(N, p.ntaps, p.planes, p.load_rows, st, 0, 1) <= 227 * 1024) { p.stages = st; break; }
  if (pc_fwd_smem(N, p.ntaps, p.planes, p.load_rows, p.stages, 0, 1) > 227 * 1024) return 1;
  p.n_ty = q.T;
  if (p.tiles_per_img > 0) {
    p.tiles_per_img = ((p.Ho - 1) * p.Wp + p.Wo + kPcWideRows - 1) / kPcWideRows;
    p.ntiles = n * p.tiles_per_img;
    p.magic_tpi = pc_magic(p.tiles_per_img);
  } else {
    p.ntiles = (int)(((long)n * p.S + kPcWideRows - 1) / kPcWideRows);
  }
  return TW;
}

bool fc_tiles_ok(arl_ctx* c);

// how layer l's OUTPUT pixels are stored: into layer l+1's input grid, or dense NHWC rows for the FC layer
void pc_out_forward(arl_ctx* c, int l, PcOut& o) {
  const ConvLayer& L = c->conv[l];
  o.mode = 0; o.scale = (l == 0) ? 1.f / c->cfg.pixel_scale : 1.f;
  o.bias = c->params + L.off_b;
  if (l + 1 < (int)c->pc.size()) {
    const PcLayer& nx = c->pc[l + 1];
    o.dst = nx.in; o.dst_plane_stride = nx.in_rows * 64;
    o.dS = nx.S; o.dWp = nx.Wp; o.dHc = nx.Hc + 1; o.dpad = nx.pad; o.ds = nx.s; o.swz = 1;
  } else if (fc_tiles_ok(c)) {
    // one 64-channel plane per output pixel: the FC tiles bulk-copy [128 images x 64 ch] blocks of it
    o.dst = c->act_fc; o.swz = 1; o.fc_rows = c->fc_rows; o.dst_plane_stride = 0;
    o.dS = L.Ho * L.Wo; o.dWp = L.Wo; o.dHc = L.Ho; o.dpad = 0; o.ds = 1;
  } else {
    o.dst = L.act; o.swz = 0; o.dense_ld = L.Cout;
    o.dS = L.Ho * L.Wo; o.dWp = L.Wo; o.dHc = L.Ho; o.dpad = 0; o.ds = 1;
  }
}

// The first conv layer can take its input straight from uint8 observations (pconv.cuh PcU8Src: converter warps build the
// bf16 space-to-depth patch in shared memory): four input planes, rows of whole 16-byte groups, space-to-depth(4).
// OFF by default (ARL_U8_CONV0=1 enables; all 93 GPU tests pass with it, the bf16 mirrors of the step buffer and of the
// rollout — 2.18 GB — are then never allocated and the frame kernel writes 66 KB instead of 200 KB per env-step): measured
// 56.8 vs 51.8 ms per iteration.  The rollout gets faster (8.1 -> 7.0 ms: the frame kernel) but conv0 forward / weight
// gradient go from 12.4 / 11.9 to 24.5 / 21.6 us: these SS-mode tiles are bound by the shared-memory operand fetch of
// tcgen05.mma (65 cycles per instruction for 16 of math), and 19.5 KB of generic-proxy patch stores per tile + the
// proxy fences in front of the MMAs cost more there than the HBM bytes they save (profiles/r2_u8_conv0.md; staging the
// uint8 rows with bulk copies first was no better: 27.0 / 24.9 us).
bool u8_conv0_ok(arl_ctx* c) {
  static const bool on = getenv("ARL_U8_CONV0") && atoi(getenv("ARL_U8_CONV0")) != 0;
  return on && c->pc_mode >= 1 && !c->pc.empty() && c->pc[0].P == 1 && c->cfg.in_c == kU8Planes && c->conv[0].s == 4 &&
         (c->cfg.in_w % 16 == 0) && (c->cfg.in_h % 4 == 0);
}

PcU8Src u8_source(arl_ctx* c, const uint8_t* obs8, int patch_rows) {
  PcU8Src u{};
  u.obs = obs8; u.H = c->cfg.in_h; u.W = c->cfg.in_w; u.Wc = u.W / 4; u.Hc = u.H / 4;
  u.rows_max = 4 * ((u.Wc - 1 + patch_rows - 1) / u.Wc + 1);
  u.img_bytes = (long)c->cfg.in_c * u.H * u.W;
  return u;
}

// conv layer l forward on the patch-resident path
int pconv_forward_layer(arl_ctx* c, int l, const __nv_bfloat16* obs16, const int* idx, const int* idx_off, int n,
                        cudaStream_t st, const uint8_t* obs8 = nullptr) {
  const PcLayer& q = c->pc[l];
  const ConvLayer& L = c->conv[l];
  PcParams p{};
  if (l == 0) {
    p.src = obs16; p.src_plane_stride = 0; p.idx = idx; p.idx_off = idx_off; p.nb = n;
    p.tiles_per_img = q.tiles_per_img; p.ntiles = n * q.tiles_per_img;
  } else {
    p.src = q.in; p.src_plane_stride = q.in_rows * 64;
    p.tiles_per_img = 0; p.ntiles = (int)(((long)n * q.S + 127) / 128);
  }
  p.planes = q.P; p.S = q.S; p.Wp = q.Wp; p.Ho = L.Ho; p.Wo = L.Wo; p.n_img = n;
  p.ntaps = q.ntaps;
  for (int t = 0; t < q.ntaps; ++t) p.shift[t] = q.shift[t];
  p.load_rows = q.load_rows; p.w = q.wpack; p.stages = q.stages_fwd;
  p.magic_S = pc_magic(p.S); p.magic_Wp = pc_magic(p.Wp); p.magic_tpi = pc_magic(p.tiles_per_img);
  pc_out_forward(c, l, p.out);
  p.out.ds_shift = (p.out.ds == 2) ? 1 : (p.out.ds == 4) ? 2 : 0; p.out.us = 1;
  if (l == 0 && obs8) p.u8 = u8_source(c, obs8, p.load_rows);
  // ARL_OBS_L2_HINT=1: the gathered observation rows of a training minibatch (34 MB) are read twice, by this kernel and ~100 us
  // later by layer 0's weight gradient: first read evict_last, second read evict_first
  static const bool obs_hint = getenv("ARL_OBS_L2_HINT") && atoi(getenv("ARL_OBS_L2_HINT")) != 0;
  p.src_evict_last = (l == 0 && obs_hint && c->in_grad_fwd) ? 1 : 0;
  const int TW = pconv_make_wide(c, q, q.N, 0, p, n, std::min(l, 2));
  return launch_pconv_n(c, q.N, p, st, TW);
}


template <int N>
int launch_pconv_wgrad(arl_ctx* c, const PcWgradParams& p, int ctas, cudaStream_t st) {
  const int smem = pc_wgrad_smem(N, p.planes, p.a_rows, p.dy_rows, p.stages);
  static int attr_smem = 0;
  if (smem > attr_smem) {
    ARL_CHECK(c, cudaFuncSetAttribute(pconv_wgrad_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_smem = smem;
  }
  ARL_CHECK(c, launch_k(pconv_wgrad_kernel<N>, dim3(ctas), dim3(kPcThreads), smem, st, p));
  c->launches++;
  ARL_CHECK(c, cudaGetLastError());
  return 0;
}

bool pconv_bwd_fused_ok(arl_ctx* c, int l, int n, struct arl::PcBwdParams* out);
int pc_wgrad_ctas(arl_ctx* c, int l, int n) {
  const PcLayer& q = c->pc[l];
  int ntiles = (l == 0) ? n * q.tiles_per_img : (int)(((long)n * q.S + 127) / 128);
  if (pconv_bwd_fused_ok(c, l, n, nullptr)) return std::min(ntiles, 148);     // one partial per CTA of pconv_bwd_kernel
  // layer 0's wgrad closes the main chain; ARL_WGRAD0_CTAS leaves the SMs still held by conv1's weight gradient alone
  static const int cap0 = getenv("ARL_WGRAD0_CTAS") ? std::max(1, std::min(148, atoi(getenv("ARL_WGRAD0_CTAS")))) : 148;
  int cap = (l > 0 && c->wgrad_ctas > 0) ? c->wgrad_ctas : (l == 0 ? cap0 : 148);
  return std::min(ntiles, cap);
}

// weight (+ bias) gradient partials of conv layer l on the patch-resident path
int pconv_wgrad_params(arl_ctx* c, int l, const __nv_bfloat16* obs16, const int* idx, const int* idx_off, int n,
                       PcWgradParams& p, const uint8_t* obs8 = nullptr) {
  const PcLayer& q = c->pc[l];
  p = PcWgradParams{};
  if (l == 0) {
    p.a = obs16; p.a_plane_stride = 0; p.idx = idx; p.idx_off = idx_off; p.nb = n;
    p.tiles_per_img = q.tiles_per_img; p.ntiles = n * q.tiles_per_img; p.dy_off = 0;
    static const bool obs_hint = getenv("ARL_OBS_L2_HINT") && atoi(getenv("ARL_OBS_L2_HINT")) != 0;
    p.a_evict_first = obs_hint ? 1 : 0;
  } else {
    p.a = q.in; p.a_plane_stride = q.in_rows * 64;
    p.tiles_per_img = 0; p.ntiles = (int)(((long)n * q.S + 127) / 128);
    p.dy_off = q.dYpad * q.Wp + q.dYpad;
  }
  p.S = q.S; p.planes = q.P; p.nblk = q.ntaps * q.P;
  p.a_rows = q.load_rows + 8;
  for (int t = 0; t < q.ntaps; ++t)
    for (int pl = 0; pl < q.P; ++pl) p.blk_off[t * q.P + pl] = (pl * p.a_rows + q.shift[t]) * 128;
  p.dy = q.dY; p.dy_rows = 128 + 8;
  static const bool wide_on = !(getenv("ARL_WGRAD_WIDE") && atoi(getenv("ARL_WGRAD_WIDE")) == 0);
  // wide form: the taps of one row on the N axis (pconv.cuh).  Needs: T * Cout <= 256 columns per MMA, one or two MMAs
  // per K step ((T+1)/2 tap-row pairs for single-plane layers, T tap rows for two-plane layers), the accumulators in
  // 512 TMEM columns, and dY rows in front of the tile (dy_off >= T-1, or the zero prefix of layer 0's grid)
  const int n_mma = (q.P == 2) ? q.T : (q.T + 1) / 2;
  if (wide_on && q.T * q.N <= 256 && n_mma <= 2 && n_mma * q.T * q.N <= 512 && (q.P == 1 || q.P == 2) &&
      (l == 0 || p.dy_off >= q.T - 1)) {
    p.wide = 1; p.n_mma = n_mma; p.nb_atoms = q.T; p.T = q.T;
    p.dy_rows = 128 + 16;
    for (int i = 0; i < n_mma; ++i) {
      if (q.P == 2) {            // A atoms = the two planes, MMA i = tap row i
        p.a_off[i] = i * q.Wp * 128; p.a_lbo[i] = p.a_rows * 128;
        p.ty0[i] = i; p.ty_step[i] = 0; p.pl_step[i] = 1;
      } else {                   // A atoms = tap rows 2i, 2i+1 (Wp rows apart)
        p.a_off[i] = 2 * i * q.Wp * 128; p.a_lbo[i] = q.Wp * 128;
        p.ty0[i] = 2 * i; p.ty_step[i] = 1; p.pl_step[i] = 0;
      }
    }
    // the second A atom of the last MMA may be a padding tap row (ty = T): its rows must still lie inside the stage
    const int max_row = 127 + ((q.P == 2) ? (q.T - 1) : (2 * (n_mma - 1) + 1)) * q.Wp;
    if (max_row >= p.a_rows) p.wide = 0;
  }
  p.partial = c->wgrad_partial[l]; p.bias_partial = c->bias_partial[l];
  if (l == 0 && obs8) p.u8 = u8_source(c, obs8, p.a_rows);
  for (p.stages = 4; p.stages >= 2; --p.stages)
    if (pc_wgrad_smem(q.N, q.P, p.a_rows, p.dy_rows, p.stages) <= 227 * 1024) break;
  if (p.stages < 2) ARL_FAIL(c, "pconv wgrad: stage does not fit shared memory");
  return 0;
}

int pconv_wgrad_layer(arl_ctx* c, int l, const __nv_bfloat16* obs16, const int* idx, const int* idx_off, int n,
                      cudaStream_t st, const uint8_t* obs8 = nullptr) {
  const PcLayer& q = c->pc[l];
  PcWgradParams p;
  if (pconv_wgrad_params(c, l, obs16, idx, idx_off, n, p, obs8)) return 2;
  int ctas = pc_wgrad_ctas(c, l, n);
  if (q.N == 32) return launch_pconv_wgrad<32>(c, p, ctas, st);
  if (q.N == 64) return launch_pconv_wgrad<64>(c, p, ctas, st);
  ARL_FAIL(c, "pconv wgrad: unsupported filter count");
}

// data gradient of conv layer l (>= 1): dY_l -> gradient w.r.t. layer l-1's output, masked by that output's ReLU
void pconv_dgrad_params(arl_ctx* c, int l, int n, PcParams& p) {
  const PcLayer& q = c->pc[l];
  const PcLayer& lo = c->pc[l - 1];
  const ConvLayer& L = c->conv[l];
  p = PcParams{};
  const int Pout = q.N / 64;
  p.src = q.dY; p.src_plane_stride = q.dY_rows * 64; p.planes = Pout;
  p.S = q.S; p.Wp = q.Wp; p.n_img = n; p.tiles_per_img = 0;
  p.ntiles = (int)(((long)n * q.S + 127) / 128);
  p.ntaps = q.ntaps;
  for (int t = 0; t < q.ntaps; ++t) p.shift[t] = q.shift[t];
  p.load_rows = q.load_rows; p.w = q.dwpack; p.stages = q.stages_dgrad;
  p.magic_S = pc_magic(p.S); p.magic_Wp = pc_magic(p.Wp); p.magic_tpi = 0;
  PcOut& o = p.out;
  o.scale = 1.f; o.act = q.in; o.act_plane_stride = q.in_rows * 64;
  o.dst = lo.dY;
  if (L.s == 1) {
    // rows = input pixels (i, j) at position i*Wp + j; their forward value sits pad rows/columns further
    p.Ho = L.Hin; p.Wo = L.Win;
    o.mode = 1; o.act_off = L.p * q.Wp + L.p;
    o.dst_plane_stride = lo.dY_rows * 64;
    o.dS = lo.S; o.dWp = lo.Wp; o.dHc = lo.Hc + 1; o.dpad = lo.dYpad; o.ds = 1; o.swz = 1;
  } else {
    // rows = space-to-depth cells: every cell is an output; unfold the sub-pixels into the pixel grid below
    p.Ho = q.Hc; p.Wo = q.Wp;
    o.mode = 2; o.act_off = 0;
    o.uH = L.Hin; o.uW = L.Win; o.uWp = lo.Wp; o.uS = lo.S; o.us = L.s; o.upad = L.p; o.uC = lo.N;
    o.dS = 1; o.dWp = 1; o.dHc = 1; o.ds = 1;
    o.us_shift = (L.s == 2) ? 1 : 2;
  }
  o.ds_shift = 0; if (o.us == 0) o.us = 1;
}

int pconv_dgrad_layer(arl_ctx* c, int l, int n, cudaStream_t st) {
  PcParams p;
  pconv_dgrad_params(c, l, n, p);
  const int Nd = c->pc[l].P * 64;
  const int TW = (p.planes == 1) ? pconv_make_wide(c, c->pc[l], Nd, p.out.mode, p, n, 3) : 1;
  return launch_pconv_n(c, Nd, p, st, TW);
}

// Data AND weight gradient of layer l (>= 1) in one kernel (pconv_bwd_kernel): both read the same dY tile.  Possible when
// the weight gradient takes the wide form, the dY patch of the data gradient covers the rows the weight gradient needs,
// and weights + >= 2 stages fit shared memory.
// OFF by default (ARL_FUSED_BWD=1 enables; every parity test passes with it): measured 54.9 vs 51.0 ms per iteration.
// The fused kernels are correct and remove the SM sharing between the two gradient kernels of a layer, but each of their
// 148 CTAs handles only 3.5 tiles and pays both kernels' fixed costs (resident data-gradient weights, 512 TMEM columns,
// two epilogues): 15.5 / 16.9 us against 9.2 + 17.3 / 11.6 + 14.5 us where the weight-gradient halves ran on 48 SMs BESIDE
// the chain (10.8 tiles per CTA) — in SM-time, 148 x 15.5 = 2294 SM-us against 1362 + 830 = 2192, and the FC weight gradient
// still shares the GPU with conv2_bwd (profiles/r2_timeline.md).
bool pconv_bwd_fused_ok(arl_ctx* c, int l, int n, PcBwdParams* out) {
  static const bool on = getenv("ARL_FUSED_BWD") && atoi(getenv("ARL_FUSED_BWD")) != 0;
  if (!on || l < 1 || l >= (int)c->pc.size()) return false;
  const PcLayer& q = c->pc[l];
  if (q.N != 64) return false;
  const int ND = q.P * 64;
  if (ND != 64 && ND != 128) return false;
  PcBwdParams b{};
  pconv_dgrad_params(c, l, n, b.d);
  if (pconv_wgrad_params(c, l, nullptr, nullptr, nullptr, n, b.g)) return false;
  if (!b.g.wide || b.d.planes != 1) return false;
  if (2 * ND + b.g.n_mma * b.g.nb_atoms * 64 > 512) return false;                 // TMEM columns
  if (b.g.dy_off - (b.g.T - 1) < 0 || b.g.dy_off + 128 > b.d.load_rows) return false;   // dY rows inside the patch
  for (b.d.stages = 4; b.d.stages >= 2; --b.d.stages)
    if (pc_bwd_smem(ND, b.d.ntaps, b.d.planes, b.d.load_rows, b.g.planes, b.g.a_rows, b.d.stages) <= 227 * 1024) break;
  if (b.d.stages < 2) return false;
  if (out) *out = b;
  return true;
}

template <int ND>
int launch_pconv_bwd(arl_ctx* c, const PcBwdParams& b, cudaStream_t st) {
  const int smem = pc_bwd_smem(ND, b.d.ntaps, b.d.planes, b.d.load_rows, b.g.planes, b.g.a_rows, b.d.stages);
  static int attr_smem = 0;
  if (smem > attr_smem) {
    ARL_CHECK(c, cudaFuncSetAttribute(pconv_bwd_kernel<ND, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_smem = smem;
  }
  const int ctas = std::min(b.d.ntiles, 148);
  ARL_CHECK(c, launch_k(pconv_bwd_kernel<ND, 64>, dim3(ctas), dim3(kPcBwdThreads), smem, st, b));
  c->launches++;
  ARL_CHECK(c, cudaGetLastError());
  return 0;
}

int pconv_bwd_layer(arl_ctx* c, int l, int n, cudaStream_t st) {
  PcBwdParams b;
  if (!pconv_bwd_fused_ok(c, l, n, &b)) ARL_FAIL(c, "fused conv backward not available for this layer");
  return (c->pc[l].P == 2) ? launch_pconv_bwd<128>(c, b, st) : launch_pconv_bwd<64>(c, b, st);
}

// gradient grids must be zero outside the rows the current batch writes
int pconv_prepare_dy(arl_ctx* c, int n, cudaStream_t st) {
  if (n >= c->pc_dy_n) { c->pc_dy_n = n; return 0; }
  if (c->dh_t) ARL_CHECK(c, cudaMemsetAsync(c->dh_t, 0, (size_t)(c->H / 64) * c->fc_rows * 64 * sizeof(__nv_bfloat16), st));
  for (size_t l = 0; l < c->pc.size(); ++l) {
    PcLayer& q = c->pc[l];
    size_t elems = (l == 0) ? (size_t)q.dY_rows * q.N : (size_t)(q.N / 64) * q.dY_rows * 64;
    ARL_CHECK(c, cudaMemsetAsync(q.dY, 0, elems * sizeof(__nv_bfloat16), st));
  }
  c->pc_dy_n = n;
  return 0;
}


// ---------------------------------------------------------------------------
// TMA-fed FC tiles (fcgemm.cuh)
// ---------------------------------------------------------------------------
template <int KIND, int BN>
int launch_fc_gemm(arl_ctx* c, const FcParams& p, dim3 grid, cudaStream_t st) {
  const int smem = p.stages * p.stage_bytes + 1024 + 256;
  static int attr_smem = 0;
  if (smem > attr_smem) {
    ARL_CHECK(c, cudaFuncSetAttribute(fc_gemm_kernel<KIND, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_smem = smem;
  }
  // (the FC weight gradient's fp32 rows are inside the optimiser-state window when the caller put grad next to params:
  // the update reads them from L2, ARL_L2_PERSIST)
  const cudaAccessPolicyWindow* win = (KIND >= 2 && c->l2_on && c->l2_grad_in_window) ? &c->l2win : nullptr;
  ARL_CHECK(c, launch_k_win(win, fc_gemm_kernel<KIND, BN>, grid, dim3(kFcThreads), (size_t)smem, st, p));
  c->launches++;
  ARL_CHECK(c, cudaGetLastError());
  return 0;
}

// KIND 0 with a thread-block cluster along the split-K axis (fcgemm.cuh: FcParams.cluster)
int launch_fc_fwd_cluster(arl_ctx* c, const FcParams& p, dim3 grid, cudaStream_t st) {
  const int smem = p.stages * p.stage_bytes + 1024 + 256;
  static int attr_smem = 0;
  if (smem > attr_smem) {
    ARL_CHECK(c, (cudaFuncSetAttribute(fc_gemm_kernel<0, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)));
    attr_smem = smem;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = dim3(kFcThreads); cfg.dynamicSmemBytes = (size_t)smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = (unsigned)p.cluster;
  cfg.attrs = at; cfg.numAttrs = 1;
  ARL_CHECK(c, (cudaLaunchKernelEx(&cfg, fc_gemm_kernel<0, 128>, p)));
  c->launches++;
  ARL_CHECK(c, cudaGetLastError());
  return 0;
}

bool fc_tiles_ok(arl_ctx* c) {
  return c->pc_mode >= 2 && c->Clast == 64 && (c->HWlast % 2 == 0) && (c->H % 256 == 0) && (c->off_Wfc % 4 == 0);
}

// forward: split-K partials [S][n][H] from act_fc planes and wfc_t tiles
int fc_forward_tiles(arl_ctx* c, int n, int* fc_S, cudaStream_t st) {
  const int NT = c->H / 64, HW = c->HWlast;
  const int mt = (n + 127) / 128, nt = c->H / 128;
  int S = std::max(1, std::min(HW, (148 + mt * nt / 2) / (mt * nt)));
  int kbps = (HW + S - 1) / S;
  S = (HW + kbps - 1) / kbps;
  // ARL_FC_CLUSTER=1 (off by default): thread-block clusters of 8 splits add their partials through distributed shared
  // memory, so the head kernels read S/8 partials instead of S ("18x wasted traffic on that edge").  Parity-green,
  // measured on B200: fc_fwd 8.9 -> 21.5 us (n = 512; 196 KB-per-CTA clusters of 8 do not all become co-resident, and the
  // DSMEM pass is serial behind the last MMA) for head_loss 6.6 -> 5.8 us — the partials are L2 hits, they were never the
  // head kernel's cost.  55.4 vs 51.6 ms per iteration.
  static const bool cluster_on = getenv("ARL_FC_CLUSTER") && atoi(getenv("ARL_FC_CLUSTER")) != 0;
  int cluster = 1;
  if (cluster_on && HW >= 16) {
    int G = std::max(1, (S + 4) / 8);
    while (G > 1 && (8 * G - 1) * ((HW + 8 * G - 1) / (8 * G)) >= HW) --G;     // no empty splits
    if ((8 * G - 1) * ((HW + 8 * G - 1) / (8 * G)) < HW) {
      cluster = 8; S = 8 * G; kbps = (HW + S - 1) / S;
    }
  }
  const int S_out = S / cluster;
  if ((long)S_out * n * c->H > c->fc_partial_cap) ARL_FAIL(c, "fc partial workspace too small");
  const long plane = (long)c->fc_rows * 64;
  FcParams p{};
  p.cluster = cluster;
  p.ncopies = 2;
  p.cp[0] = FcCopy{c->act_fc, 128 * 64, 0, (long)kbps * plane, plane, 16384, 0};
  p.cp[1] = FcCopy{c->wfc_t, 0, 2 * 4096, (long)kbps * NT * 4096, (long)NT * 4096, 16384, 16384};
  p.a_bytes = 16384; p.stage_bytes = 32768; p.stages = 4;
  p.niter = kbps; p.niter_total = HW; p.M = n;
  p.out_f32 = c->fc_partial; p.ldo = c->H;
  *fc_S = S_out;
  if (cluster > 1) return launch_fc_fwd_cluster(c, p, dim3(mt, nt, S), st);
  return launch_fc_gemm<0, 128>(c, p, dim3(mt, nt, S), st);
}

// data gradient: dh_t planes x wfc_t tiles -> masked, scattered into the last conv layer's gradient grid
int fc_dgrad_tiles(arl_ctx* c, int n, cudaStream_t st) {
  const int NT = c->H / 64, HW = c->HWlast;
  const ConvLayer& LL = c->conv.back();
  const PcLayer& q = c->pc.back();
  // three pixel planes per tile (HW/3 x ceil(n/128) CTAs) when HW allows, else two
  const int PL = (HW % 3 == 0) ? 3 : 2;
  FcParams p{};
  p.ncopies = 1 + PL;
  p.cp[0] = FcCopy{c->dh_t, 128 * 64, 0, 0, (long)c->fc_rows * 64, 16384, 0};
  for (int i = 0; i < PL; ++i)
    p.cp[1 + i] = FcCopy{c->wfc_t + (long)i * NT * 4096, 0, (long)PL * NT * 4096, 0, 4096, 8192, (uint32_t)(16384 + i * 8192)};
  p.a_bytes = 16384; p.stage_bytes = 16384 + PL * 8192; p.stages = 4;
  p.niter = NT; p.niter_total = NT; p.M = n;
  p.dy = q.dY; p.act = c->act_fc; p.act_plane = (long)c->fc_rows * 64;
  p.sc_Wo = LL.Wo; p.sc_S = q.S; p.sc_Wp = q.Wp; p.sc_pad = q.dYpad;
  if (PL == 3) return launch_fc_gemm<1, 192>(c, p, dim3((n + 127) / 128, HW / 3, 1), st);
  return launch_fc_gemm<1, 128>(c, p, dim3((n + 127) / 128, HW / 2, 1), st);
}

// weight gradient: act_fc^T x dh_t -> fp32 rows of the flat gradient (reference row order)
int fc_wgrad_tiles(arl_ctx* c, int n, cudaStream_t st, bool want_ss = false) {
  const int HW = c->HWlast;
  FcParams p{};
  c->pending_ss_fc = 0;
  const long aplane = (long)c->fc_rows * 64;
  static const bool quad = !(getenv("ARL_FC_WGRAD_QUAD") && atoi(getenv("ARL_FC_WGRAD_QUAD")) == 0);
  if (quad && HW % 4 == 0) {
    // four pixel planes per CTA: 256 x 256 tiles in two accumulators, 64 KB stages (4 x 8 KB act planes + 4 x 8 KB dh planes)
    p.ncopies = 8;
    for (int i = 0; i < 4; ++i) {
      p.cp[i] = FcCopy{c->act_fc + (long)i * aplane, 4 * aplane, 0, 0, 64 * 64, 8192, (uint32_t)(i * 8192)};
      p.cp[4 + i] = FcCopy{c->dh_t + (long)i * aplane, 0, 4 * aplane, 0, 64 * 64, 8192, (uint32_t)(32768 + i * 8192)};
    }
    p.a_bytes = 32768; p.stage_bytes = 65536; p.stages = 3;
    p.niter = (n + 63) / 64; p.niter_total = p.niter; p.M = n;
    p.out_f32 = c->grad + c->off_Wfc; p.ldo = c->H; p.fc_HW = HW;
    if (want_ss && (HW / 4) * (c->H / 256) * 8 <= kSsFcCap) { p.ss_out = c->ss_fc; c->pending_ss_fc = (HW / 4) * (c->H / 256) * 8; }
    return launch_fc_gemm<3, 256>(c, p, dim3(HW / 4, c->H / 256, 1), st);
  }
  p.ncopies = 6;
  p.cp[0] = FcCopy{c->act_fc, 2 * aplane, 0, 0, 64 * 64, 8192, 0};
  p.cp[1] = FcCopy{c->act_fc + aplane, 2 * aplane, 0, 0, 64 * 64, 8192, 8192};
  for (int i = 0; i < 4; ++i)
    p.cp[2 + i] = FcCopy{c->dh_t + (long)i * aplane, 0, 4 * aplane, 0, 64 * 64, 8192, (uint32_t)(16384 + i * 8192)};
  p.a_bytes = 16384; p.stage_bytes = 16384 + 4 * 8192; p.stages = 4;
  p.niter = (n + 63) / 64; p.niter_total = p.niter; p.M = n;
  p.out_f32 = c->grad + c->off_Wfc; p.ldo = c->H; p.fc_HW = HW;
  if (want_ss && (HW / 2) * (c->H / 256) * 8 <= kSsFcCap) { p.ss_out = c->ss_fc; c->pending_ss_fc = (HW / 2) * (c->H / 256) * 8; }
  return launch_fc_gemm<2, 256>(c, p, dim3(HW / 2, c->H / 256, 1), st);
}

constexpr int kFcBN = 128;   // FC forward tile width: A is re-read H/kFcBN times, the weights ceil(n/128) times
int fc_splits(arl_ctx* c, int n, int& kbps) {
  int kb = c->Kfc / 64;
  int tiles = ((n + 127) / 128) * (c->H / kFcBN);
  int S = std::max(1, std::min(kb, (148 + tiles / 2) / tiles));
  kbps = (kb + S - 1) / S;
  S = (kb + kbps - 1) / kbps;
  return S;
}

int alloc_net(arl_ctx* c) {
  const int R = c->cfg.max_rows;
  std::vector<PackJob> pj;
  for (size_t l = 0; l < c->conv.size(); ++l) {
    ConvLayer& L = c->conv[l];
    size_t act_elems = (size_t)R * L.Ho * L.Wo * L.Cout + 8 * 1024;
    if (dev_alloc(c, &L.act, act_elems)) return 1;
    if (dev_alloc(c, &L.dact, act_elems)) return 1;
    if (dev_alloc(c, &L.wpack, (size_t)L.Cout * L.K)) return 1;
    PackJob j{};
    j.dst = L.wpack; j.src_off = L.off_W; j.kind = (l == 0) ? PK_CONV_S2D : PK_CONV_NHWC;
    j.rows = L.Cout; j.cols = L.K; j.Cout = L.Cout; j.C = L.Cin; j.kh = L.k; j.kw = L.k; j.s = L.s;
    const bool gather_packs = c->pc_mode < 2;   // the gather path's operand copies are dead weight in full pconv mode
    if (gather_packs) pj.push_back(j);
    for (auto& d : L.dclasses) {
      if (dev_alloc(c, &d.wpack, (size_t)L.Cin * d.K)) return 1;
      PackJob q{};
      q.dst = d.wpack; q.src_off = L.off_W; q.kind = PK_CONV_DGRAD;
      q.rows = L.Cin; q.cols = d.K; q.Cout = L.Cout; q.C = L.Cin; q.kh = L.k; q.kw = L.k;
      q.s = L.s; q.ry = d.ry; q.rx = d.rx; q.Tx = d.Tx;
      if (gather_packs) pj.push_back(q);
    }
    long cap = (long)kMaxSplits * std::max(L.K, c->pc.empty() ? 0 : c->pc[l].ntaps * c->pc[l].P * 64) * L.Cout;
    float* wp = nullptr;
    if (dev_alloc(c, &wp, (size_t)cap)) return 1;
    c->wgrad_partial.push_back(wp);
    c->wgrad_partial_cap.push_back(cap);
    float* bp = nullptr;
    if (dev_alloc(c, &bp, (size_t)kMaxSplits * L.Cout)) return 1;
    c->bias_partial.push_back(bp);
  }
  if (dev_alloc(c, &c->wfc_bf16, (size_t)c->H * c->Kfc)) return 1;
  if (fc_tiles_ok(c)) {
    c->fc_rows = roundup(R, 128);
    if (dev_alloc(c, &c->act_fc, (size_t)(c->HWlast + 2) * c->fc_rows * 64)) return 1;
    if (dev_alloc(c, &c->wfc_t, (size_t)c->H * c->Kfc)) return 1;
    if (dev_alloc(c, &c->dh_t, (size_t)(c->H / 64) * c->fc_rows * 64)) return 1;
  }
  {
    // LAST job = the bf16 FC operand copy (skipped by pack_weights(with_fc=false) when the update kernel refreshes it)
    PackJob j{};
    j.dst = c->wfc_bf16; j.src_off = c->off_Wfc; j.kind = PK_CAST; j.rows = c->Kfc; j.cols = c->H;
    if (fc_tiles_ok(c)) { j.dst = c->wfc_t; j.kind = PK_FC_TILES; j.HW = c->HWlast; }
    pj.push_back(j);
  }
  // the FC cast job stays LAST (pack_weights(with_fc=false) drops it); pconv packs go before it
  {
    PackJob fc = pj.back();
    pj.pop_back();
    if (alloc_pconv(c, pj)) return 1;
    pj.push_back(fc);
  }
  c->n_pack_jobs = (int)pj.size();
  {
    long end = 0;
    bool all_pconv = pj.size() >= 2;
    for (size_t i = 0; i + 1 < pj.size(); ++i) {
      all_pconv = all_pconv && (pj[i].kind == PK_PCONV || pj[i].kind == PK_PCONV_DGRAD);
      end = std::max(end, pj[i].src_off + (long)pj[i].Cout * pj[i].C * pj[i].kh * pj[i].kw);
    }
    end = (end + 3) / 4 * 4;        // whole float4 groups (the elements after the last conv weight are biases: no slot)
    c->conv_pack_end = (all_pconv && end <= c->n_params) ? end : 0;
    if (c->conv_pack_end > 0) {
      // element -> bf16 slot addresses in the forward / data-gradient packs (inverse of pack_job_body, see pack_slot_of)
      std::vector<unsigned long long> tab((size_t)c->conv_pack_end * 2, 0ULL);
      for (long e = 0; e < c->conv_pack_end; ++e) {
        int nslot = 0;
        for (size_t i = 0; i + 1 < pj.size(); ++i) {
          const long d = pack_slot_of(pj[i], e);
          if (d < 0) continue;
          if (nslot >= 2) { c->conv_pack_end = 0; break; }     // more than two packs hold this weight: keep the pack launch
          tab[(size_t)e * 2 + nslot++] = (unsigned long long)(uintptr_t)(pj[i].dst + d);
        }
        if (c->conv_pack_end == 0) break;
      }
      if (c->conv_pack_end > 0) {
        if (dev_alloc(c, &c->pk_slots, tab.size())) return 1;
        ARL_CHECK(c, cudaMemcpy(c->pk_slots, tab.data(), tab.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice));
      }
    }
  }
  if (dev_alloc(c, &c->pack_jobs_dev, pj.size())) return 1;
  ARL_CHECK(c, cudaMemcpy(c->pack_jobs_dev, pj.data(), pj.size() * sizeof(PackJob), cudaMemcpyHostToDevice));
  if (dev_alloc(c, &c->obs16_stage, (size_t)R * c->obs16_elems + 64 * 1024)) return 1;
  {
    long worst = R;
    for (int n = 1; n <= R; ++n) {
      int kbps = 0;
      long s = fc_splits(c, n, kbps);
      worst = std::max(worst, s * n);
    }
    c->fc_partial_cap = worst * c->H + 1024;
  }
  if (dev_alloc(c, &c->fc_partial, (size_t)c->fc_partial_cap)) return 1;
  if (dev_alloc(c, &c->h, (size_t)R * c->H)) return 1;
  if (dev_alloc(c, &c->dh, (size_t)R * c->H)) return 1;
  if (dev_alloc(c, &c->dlogit, (size_t)R * (c->A + 1))) return 1;
  if (dev_alloc(c, &c->head_partial, (size_t)64 * c->H * (c->A + 2))) return 1;
  if (dev_alloc(c, &c->head_b_partial, (size_t)64 * (c->A + 1))) return 1;
  if (dev_alloc(c, &c->loss_partial, (size_t)R * 4)) return 1;
  if (dev_alloc(c, &c->sumsq_partial, (size_t)kStreamPartials)) return 1;
  if (dev_alloc(c, &c->sumsq_partial_fc, (size_t)kEarlyBlocks)) return 1;
  if (dev_alloc(c, &c->ss_fin, (size_t)kSsFinCap)) return 1;
  if (dev_alloc(c, &c->ss_fc, (size_t)kSsFcCap)) return 1;
  if (dev_alloc(c, &c->ticket, 4)) return 1;
  if (dev_alloc(c, &c->hyper, 8)) return 1;
  float one = 1.f;
  ARL_CHECK(c, cudaMemcpy(c->hyper, &one, sizeof(float), cudaMemcpyHostToDevice));
  if (dev_alloc(c, &c->step, 1)) return 1;
  if (dev_alloc(c, &c->log_slot, 1)) return 1;
  if (dev_alloc(c, &c->log_norm, (size_t)c->log_cap)) return 1;
  if (dev_alloc(c, &c->log_loss, (size_t)c->log_cap)) return 1;
  if (dev_alloc(c, &c->mb_counter, 1)) return 1;
  if (dev_alloc(c, &c->valid_count, 1)) return 1;
  return 0;
}

// ---------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------
RowEpi make_epi(int mode) {
  RowEpi e{};
  e.mode = mode; e.scale = 1.f;
  return e;
}

// forward gather geometry of layer l reading `src` (NHWC bf16), n images
ConvGeom fwd_geom(const ConvLayer& L, const __nv_bfloat16* src, int n) {
  ConvGeom g{};
  g.src = src; g.Qh = L.Ho; g.Qw = L.Wo; g.Hs = L.gH; g.Ws = L.gW; g.C = L.gC;
  g.sy = L.gs; g.y0 = -L.gp; g.dty = 1; g.sx = L.gs; g.x0 = -L.gp; g.dtx = 1; g.Tx = L.gk;
  g.nrows = n * L.Ho * L.Wo;
  return g;
}

static const char* kFwdName[4] = {"conv0_fwd", "conv1_fwd", "conv2_fwd", "conv3_fwd"};
static const char* kWgradName[4] = {"conv0_wgrad", "conv1_wgrad", "conv2_wgrad", "conv3_wgrad"};
static const char* kDgradName[4] = {"conv0_dgrad", "conv1_dgrad", "conv2_dgrad", "conv3_dgrad"};

// uint8 CHW observations -> bf16 space-to-depth staging rows [0, n)
int convert_obs(arl_ctx* c, const uint8_t* obs, const int* idx, int n, bool swz, cudaStream_t st) {
  const ConvLayer& L0 = c->conv[0];
  long work = (long)n * L0.Cin * L0.Hin * (L0.Win / 4);
  int blocks = (int)std::min<long>((work + 255) / 256, 148 * 16);
  obs_to_s2d_kernel<<<blocks, 256, 0, st>>>(obs, idx, c->obs16_stage, n, L0.Cin, L0.Hin, L0.Win, swz ? 1 : 0, c->obs16_elems);
  c->launches++;
  prof_mark(c, "obs_to_s2d", st);
  ARL_CHECK(c, cudaGetLastError());
  return 0;
}

// conv stack + FC split-K partials for n observations given as bf16 space-to-depth images
// (idx/idx_off: optional image gather applied by the first layer's loader)
int forward_trunk(arl_ctx* c, const __nv_bfloat16* obs16, const int* idx, const int* idx_off, int n, int* fc_S,
                  bool pc, cudaStream_t st, const uint8_t* obs8 = nullptr) {
  NvtxRange nvtx_("fwd");
  if (n > c->cfg.max_rows) ARL_FAIL(c, "batch larger than max_rows");
  if (!c->params) ARL_FAIL(c, "parameters not bound");
  for (size_t l = 0; l < c->conv.size(); ++l) {
    if (pc) {
      // patch-resident tiles: obs16 and every intermediate activation are chunk-swizzled position grids
      if (pconv_forward_layer(c, (int)l, obs16, idx, idx_off, n, st, obs8)) return 1;
      prof_mark(c, kFwdName[l], st);
      continue;
    }
    ConvLayer& L = c->conv[l];
    int rows = n * L.Ho * L.Wo;
    RowEpi e = make_epi(EPI_BIAS_RELU_BF16);
    e.bias = c->params + L.off_b; e.out = L.act; e.ldo = L.Cout; e.M = rows;
    if (l == 0) e.scale = 1.f / c->cfg.pixel_scale;
    RowGemmMulti<ConvLoader<128>> mp{};
    mp.a[0].g = fwd_geom(L, l == 0 ? obs16 : c->conv[l - 1].act, n);
    if (l == 0) { mp.a[0].g.idx = idx; mp.a[0].g.idx_off = idx_off; }
    mp.b[0] = WeightSrc{L.wpack, (long)L.K, 0, RowPerm{0, 0}};
    mp.e[0] = e;
    mp.mtiles[0] = (rows + 127) / 128;
    mp.num_kb[0] = L.K / 64;
    if (launch_conv_persist_bn(c, L.Cout, mp, 1, L.K, mp.mtiles[0], st)) return 1;
    prof_mark(c, kFwdName[l], st);
  }
  if (c->updB_pending) {
    // the previous minibatch's FC-range update (side stream) must be complete before the FC weights are read
    ARL_CHECK(c, cudaStreamWaitEvent(st, c->ev_updB, 0));
    c->updB_pending = false;
  }
  if (pc && fc_tiles_ok(c)) {
    if (fc_forward_tiles(c, n, fc_S, st)) return 1;
    prof_mark(c, "fc_fwd", st);
    return 0;
  }
  int kbps = 0;
  int S = fc_splits(c, n, kbps);
  if ((long)S * n * c->H > c->fc_partial_cap) ARL_FAIL(c, "fc partial workspace too small");
  RowEpi e = make_epi(EPI_PARTIAL_F32);
  e.partial = c->fc_partial; e.ldo = c->H; e.M = n;
  DenseLoader<128> a{};
  a.src = c->conv.back().act; a.ld = c->Kfc; a.nrows = n;
  // B = the FC weights in the reference's (c,h,w)-row order, read N-major through the (hw,c) row permutation
  WeightSrc w{c->wfc_bf16, (long)c->H, c->Kfc, RowPerm{c->Clast, c->HWlast}};
  if (launch_rowgemm<DenseLoader<128>, true, kFcBN>(c, a, w, e, n, c->H, c->Kfc / 64, kbps, S, st)) return 1;
  prof_mark(c, "fc_fwd", st);
  *fc_S = S;
  return 0;
}


HeadParams head_base(arl_ctx* c, int n, int S) {
  HeadParams p{};
  p.partial = c->fc_partial; p.splits = S; p.M = n; p.H = c->H; p.A = c->A;
  p.fc_bias = c->params + c->off_bfc; p.w_pi = c->params + c->off_Wpi; p.b_pi = c->params + c->off_bpi;
  p.w_v = c->params + c->off_Wv; p.b_v = c->params + c->off_bv;
  p.hyper = c->hyper;
  return p;
}

// the head kernel is instantiated for action-count bounds 4 / 6 / 9 / 18 (its per-action loops are unrolled)
template <int MODE>
cudaError_t launch_head(const HeadParams& p, int n, cudaStream_t st) {
  if (p.A <= 4) return launch_k(head_kernel<MODE, 4>, dim3(n), dim3(kHeadThreads), 0, st, p);
  if (p.A <= 6) return launch_k(head_kernel<MODE, 6>, dim3(n), dim3(kHeadThreads), 0, st, p);
  if (p.A <= 9) return launch_k(head_kernel<MODE, 9>, dim3(n), dim3(kHeadThreads), 0, st, p);
  return launch_k(head_kernel<MODE, kMaxActions>, dim3(n), dim3(kHeadThreads), 0, st, p);
}

int policy_forward16(arl_ctx* c, const __nv_bfloat16* obs16, int n, const int* out_rows, float* prob, float* value,
                     const double* uniforms, uint8_t* actions, bool pc, cudaStream_t st, const EnvStepArgs* es = nullptr,
                     const uint8_t* obs8 = nullptr, const int* idx = nullptr) {
  int S = 0;
  if (forward_trunk(c, obs16, idx, nullptr, n, &S, pc, st, obs8)) return 1;
  HeadParams p = head_base(c, n, S);
  if (es) { p.es = *es; p.es_on = 1; }
  p.out_rows = out_rows; p.prob = prob; p.value = value; p.uniforms = uniforms; p.actions = uniforms ? actions : nullptr;
  ARL_CHECK(c, launch_head<0>(p, n, st));
  c->launches++;
  prof_mark(c, "head_sample", st);
  ARL_CHECK(c, cudaGetLastError());
  return 0;
}

int policy_forward(arl_ctx* c, const uint8_t* obs, const int* idx, int n, const int* out_rows, float* prob,
                   float* value, const double* uniforms, uint8_t* actions, cudaStream_t st) {
  if (n > c->cfg.max_rows) ARL_FAIL(c, "batch larger than max_rows");
  if (u8_conv0_ok(c))       // the first conv layer reads the uint8 rows itself (gathered through idx)
    return policy_forward16(c, nullptr, n, out_rows, prob, value, uniforms, actions, true, st, nullptr, obs, idx);
  if (convert_obs(c, obs, idx, n, c->pc_mode >= 1, st)) return 1;
  return policy_forward16(c, c->obs16_stage, n, out_rows, prob, value, uniforms, actions, c->pc_mode >= 1, st);
}

// ---------------------------------------------------------------------------
// training plan (depends on the minibatch size)
// ---------------------------------------------------------------------------
int get_plan(arl_ctx* c, int n, TrainPlan** out) {
  auto it = c->plans.find(n);
  if (it != c->plans.end()) { *out = &it->second; return 0; }
  TrainPlan P;
  P.n = n;
  std::vector<GradJob> jobs;
  long max_total = 1;
  for (size_t l = 0; l < c->conv.size(); ++l) {
    ConvLayer& L = c->conv[l];
    int rows = n * L.Ho * L.Wo;
    int rps = roundup((rows + 147) / 148, 64);
    int splits = (rows + rps - 1) / rps;
    if (splits > kMaxSplits) ARL_FAIL(c, "wgrad partial workspace too small");
    P.conv_splits.push_back(splits); P.conv_rps.push_back(rps);
    GradJob w{};
    w.src = c->wgrad_partial[l]; w.S = splits; w.sstride = (long)L.K * L.Cout; w.rows = L.K; w.cols = L.Cout;
    w.ld = L.Cout; w.map = (l == 0) ? GM_CONV_S2D : GM_CONV_NHWC; w.scale = (l == 0) ? 1.f / c->cfg.pixel_scale : 1.f;
    w.dst_off = L.off_W; w.C = L.Cin; w.kh = L.k; w.kw = L.k; w.s2d = L.s;
    if (c->pc_mode >= 2) {
      // patch-resident wgrad: one partial per persistent CTA, K' rows in (tap, plane, channel) order
      const PcLayer& q = c->pc[l];
      splits = pc_wgrad_ctas(c, (int)l, n);
      w.S = splits; w.rows = q.ntaps * q.P * 64; w.sstride = (long)w.rows * L.Cout; w.map = GM_PCONV;
      w.T = q.T; w.P = q.P; w.ci_major = q.ci_major;
    }
    jobs.push_back(w);
    max_total = std::max(max_total, (long)w.rows * w.cols);
    GradJob b{};
    b.src = c->bias_partial[l]; b.S = splits; b.sstride = L.Cout; b.rows = 1; b.cols = L.Cout; b.ld = L.Cout;
    b.map = GM_LINEAR; b.scale = 1.f; b.dst_off = L.off_b;
    jobs.push_back(b);
  }
  P.head_rpg = std::max(8, (n + 63) / 64);
  P.head_groups = (n + P.head_rpg - 1) / P.head_rpg;
  {
    GradJob hj{};
    hj.src = c->head_partial; hj.S = P.head_groups; hj.sstride = (long)c->H * (c->A + 2); hj.rows = c->H;
    hj.cols = c->A + 2; hj.ld = c->A + 2; hj.map = GM_HEAD; hj.scale = 1.f; hj.dst_off = c->off_Wpi;
    hj.dst_off2 = c->off_Wv; hj.dst_off3 = c->off_bfc; hj.A = c->A;
    jobs.push_back(hj);
    max_total = std::max(max_total, (long)hj.rows * hj.cols);
    GradJob bp{};
    bp.src = c->head_b_partial; bp.S = P.head_groups; bp.sstride = c->A + 1; bp.rows = 1; bp.cols = c->A;
    bp.ld = c->A + 1; bp.map = GM_LINEAR; bp.scale = 1.f; bp.dst_off = c->off_bpi;
    jobs.push_back(bp);
    GradJob bv = bp;
    bv.src = c->head_b_partial + c->A; bv.cols = 1; bv.dst_off = c->off_bv;
    jobs.push_back(bv);
  }
  P.n_jobs = (int)jobs.size();
  P.fin_blocks = (int)((max_total + kFinPerBlock - 1) / kFinPerBlock);
  for (auto& jb : jobs) { P.fin_total += (long)jb.rows * jb.cols; P.job_total.push_back((long)jb.rows * jb.cols); }
  P.early_jobs = (c->pc_mode >= 2 && fc_tiles_ok(c) && c->conv.size() >= 2) ? 1 : 0;
  ARL_CHECK(c, cudaMalloc(reinterpret_cast<void**>(&P.jobs_dev), jobs.size() * sizeof(GradJob)));
  ARL_CHECK(c, cudaMemcpy(P.jobs_dev, jobs.data(), jobs.size() * sizeof(GradJob), cudaMemcpyHostToDevice));
  auto res = c->plans.emplace(n, P);
  *out = &res.first->second;
  return 0;
}

bool early_fc_ok(arl_ctx* c);
int early_fc_update(arl_ctx* c, cudaStream_t st);

// split partials of jobs [first, first + count) -> their places in the flat gradient
int launch_finalize(arl_ctx* c, const TrainPlan* P, int first, int count, const char* label, cudaStream_t st) {
  long mx = 1;
  for (int j = first; j < first + count; ++j) mx = std::max(mx, P->job_total[j]);
  dim3 grid((unsigned)((mx + kFinPerBlock - 1) / kFinPerBlock), (unsigned)count);
  ARL_CHECK(c, launch_k(finalize_grads_kernel, dim3(grid), dim3(256), 0, st, P->jobs_dev + first, c->grad));
  c->launches++;
  prof_mark(c, label, st);
  ARL_CHECK(c, cudaGetLastError());
  return 0;
}
bool stream_update_ok(arl_ctx* c, bool fct);
int sync_fc_exchange(arl_ctx* c, cudaStream_t ws, cudaStream_t st);

// forward + loss + backward for one minibatch -> flat grad
int grad_minibatch(arl_ctx* c, const int* idx, const int* idx_off, int n, cudaStream_t st) {
  NvtxRange nvtx_("grad: fwd + loss + bwd");
  if (!c->opt_set) ARL_FAIL(c, "optimizer not configured");
  if (!c->t_obs) ARL_FAIL(c, "training inputs not bound");
  if (!c->grad) ARL_FAIL(c, "gradient vector not bound");
  TrainPlan* P = nullptr;
  if (get_plan(c, n, &P)) return 1;
  // first-layer input: the sampler's bf16 mirror of the rollout (gathered by the loader), or a conversion
  // of the caller's uint8 rows into the staging buffer
  const __nv_bfloat16* obs16;
  const uint8_t* obs8 = nullptr;
  const int *gidx, *gidx_off;
  if (c->pc_mode >= 2 && u8_conv0_ok(c)) {
    // the first conv layer (forward and weight gradient) gathers the uint8 training rows itself
    obs16 = nullptr; obs8 = c->t_obs; gidx = idx; gidx_off = idx_off;
  } else if (c->sampler_set && c->t_obs == c->sc.observations && c->roll_obs16) {
    obs16 = c->roll_obs16; gidx = idx; gidx_off = idx_off;
  } else {
    if (idx_off) ARL_FAIL(c, "graph-replayed training needs the sampler's rollout buffers as training inputs");
    if (convert_obs(c, c->t_obs, idx, n, c->pc_mode >= 2, st)) return 1;
    obs16 = c->obs16_stage; gidx = nullptr; gidx_off = nullptr;
  }
  const bool pcb = c->pc_mode >= 2;
  if (pcb && pconv_prepare_dy(c, n, st)) return 1;
  int S = 0;
  c->in_grad_fwd = true;
  const int rc_fwd = forward_trunk(c, obs16, gidx, gidx_off, n, &S, pcb, st, obs8);
  c->in_grad_fwd = false;
  if (rc_fwd) return 1;
  // ---- head: losses + dlogits + dh ----
  if (c->t_valids) {
    ARL_CHECK(c, launch_k(count_valids_idx_kernel, dim3(1), dim3(1024), 0, st, c->t_valids, idx, idx_off, n, c->valid_count));
    c->launches++;
    prof_mark(c, "count_valids", st);
  }
  HeadParams p = head_base(c, n, S);
  p.idx = idx; p.idx_off = idx_off; p.act_in = c->t_act; p.adv = c->t_adv; p.ret = c->t_ret; p.old_prob = c->t_oldp;
  p.valids = c->t_valids; p.valid_count = c->valid_count;
  p.algo = c->opt.algo; p.clip_param = c->opt.clip_param; p.v_coeff = c->opt.v_loss_coeff;
  p.tie_grad = c->opt.ppo_tie_grad == 2 ? 2.f : 1.f;
  p.ent_coeff = c->opt.ent_loss_coeff; p.inv_count = 1.f / (float)n;
  p.h_out = c->h; p.dh_out = c->dh; p.dlogit_out = c->dlogit; p.loss_partial = c->loss_partial;
  const bool fct = pcb && fc_tiles_ok(c);
  if (fct) { p.dh_t = c->dh_t; p.dh_rows = c->fc_rows; }
  c->n_loss_rows = n;
  ARL_CHECK(c, launch_head<1>(p, n, st));
  c->launches++;
  prof_mark(c, "head_loss", st);
  ARL_CHECK(c, cudaGetLastError());
  // The weight-gradient kernels only feed finalize_grads at the very end: they go to a side stream and overlap the
  // data-gradient chain (every kernel here is fixed-cost dominated at minibatch size: one kernel's ramp-up fills the
  // SMs the other's tail leaves idle).  fork k = "the gradient grid wgrad k needs is complete".
  cudaStream_t ws = st;
  if (!c->no_fork && !c->prof_on && pcb) {
    if (!c->side) {
      ARL_CHECK(c, cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking));
      for (auto& e : c->ev_fork) ARL_CHECK(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      ARL_CHECK(c, cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
      ARL_CHECK(c, cudaStreamCreateWithFlags(&c->side2, cudaStreamNonBlocking));
      ARL_CHECK(c, cudaEventCreateWithFlags(&c->ev_join2, cudaEventDisableTiming));
      ARL_CHECK(c, cudaEventCreateWithFlags(&c->ev_fcd, cudaEventDisableTiming));
      for (auto& e : c->ev_fin) ARL_CHECK(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    ws = c->side;
    ARL_CHECK(c, cudaEventRecord(c->ev_fork[0], st));
    ARL_CHECK(c, cudaStreamWaitEvent(ws, c->ev_fork[0], 0));
  }
  {
    dim3 grid((c->H + 127) / 128, P->head_groups);
    size_t sm = (size_t)P->head_rpg * (c->A + 1) * sizeof(float);
    ARL_CHECK(c, launch_k(head_wgrad_kernel, dim3(grid), dim3(128), sm, ws, c->h, c->dh, c->dlogit, n, c->H, c->A, P->head_rpg, c->head_partial,
                                              c->head_b_partial));
    c->launches++;
    prof_mark(c, "head_wgrad", ws);
    ARL_CHECK(c, cudaGetLastError());
  }
  ConvLayer& LL = c->conv.back();
  // ---- FC wgrad: dW[Kfc][H] = a_last^T dh (direct, permuted rows) ----
  if (fct) {
    if (fc_wgrad_tiles(c, n, ws)) return 1;
    prof_mark(c, "fc_wgrad", ws);
    if (c->sync_overlap_active) ARL_CHECK(c, cudaEventRecord(c->ev_cs_in[0], ws));
    // head / FC-bias gradients: their partials exist (head_wgrad, earlier on ws) -> flat vector, beside the main chain
    if (P->early_jobs && launch_finalize(c, P, 2 * (int)c->conv.size(), 3, "finalize_head", ws)) return 1;
  } else {
    DenseLoader<64> a{};
    a.src = LL.act; a.ld = c->Kfc; a.nrows = n;
    WgradEpi e{};
    e.out = c->grad + c->off_Wfc; e.mode = 1; e.Kvalid = c->Kfc; e.Kp = c->Kfc; e.ldo = c->H; e.fc_C = c->Clast;
    e.fc_HW = c->HWlast;
    if (launch_wgrad<DenseLoader<64>, 1, 256>(c, a, c->dh, c->H, n, roundup(n, 64), 1, c->Kfc / 64, c->H / 256, e, ws))
      return 1;
    prof_mark(c, "fc_wgrad", ws);
  }
  // ---- FC dgrad: da_last[n][Kfc] = dh Wfc^T, masked by a_last > 0 ----
  if (fct) {
    if (fc_dgrad_tiles(c, n, st)) return 1;
    prof_mark(c, "fc_dgrad", st);
    // synchronous learners: the FC gradient is final and the FC weights have been consumed -> its slice exchange + update
    // starts now on its own stream, beside the conv gradient chain
    if (c->sync_overlap_active && sync_fc_exchange(c, ws, st)) return 1;
    if (early_fc_ok(c)) {
      // both FC gradient kernels are issued: once the data gradient has read the weights (and the weight gradient,
      // earlier on ws, has written its rows of the flat gradient) the FC weights take their Adam/RMSProp step while
      // the conv gradient chain runs
      if (ws != st) {
        ARL_CHECK(c, cudaEventRecord(c->ev_fcd, st));
        ARL_CHECK(c, cudaStreamWaitEvent(ws, c->ev_fcd, 0));
      }
      if (early_fc_update(c, ws)) return 1;
    }
  } else {
    DenseLoader<128> a{};
    a.src = c->dh; a.ld = c->H; a.nrows = n;
    WeightSrc w{c->wfc_bf16, (long)c->H, 0, RowPerm{c->Clast, c->HWlast}};   // K-major rows n in (hw,c) order
    RowEpi e = make_epi(EPI_MASK_BF16);
    e.out = LL.dact; e.act = LL.act; e.ldo = c->Kfc; e.M = n;
    if (pcb) {
      // scatter straight into the last conv layer's gradient grid (padded, chunk-swizzled): what its dgrad / wgrad read
      const PcLayer& q = c->pc.back();
      e.out = q.dY; e.sc_on = 1; e.sc_Wo = LL.Wo; e.sc_S = q.S; e.sc_Wp = q.Wp; e.sc_pad = q.dYpad;
    }
    int BN = (c->Kfc % 128 == 0) ? 128 : 64;
    if (launch_rowgemm_bn<DenseLoader<128>, false>(c, BN, a, w, e, n, c->Kfc, c->H / 64, c->H / 64, 1, st)) return 1;
    prof_mark(c, "fc_dgrad", st);
  }
  // ---- conv layers, last to first ----
  bool used_side2 = false;
  for (int l = (int)c->conv.size() - 1; l >= 0 && pcb; --l) {
    if (pconv_bwd_fused_ok(c, l, n, nullptr)) {
      // data + weight gradient of layer l in one kernel on the main chain; its partials are finalised beside the chain
      if (pconv_bwd_layer(c, l, n, st)) return 1;
      prof_mark(c, l == 2 ? "conv2_bwd" : l == 1 ? "conv1_bwd" : "conv_bwd", st);
      if (P->early_jobs) {
        cudaStream_t fs = (ws != st) ? ws : st;
        if (fs != st) {
          ARL_CHECK(c, cudaEventRecord(c->ev_fin[l & 1], st));
          ARL_CHECK(c, cudaStreamWaitEvent(fs, c->ev_fin[l & 1], 0));
        }
        if (launch_finalize(c, P, 2 * l, 2, "finalize_conv", fs)) return 1;
      }
      continue;
    }
    // layer l's gradient grid was completed by the last kernel on `st` (FC dgrad or dgrad l+1); the first layer's
    // wgrad closes the main chain itself
    cudaStream_t wl = (l == 0 || ws == st) ? st : c->side2;
    if (wl != st) {
      used_side2 = true;
      cudaEvent_t ev = c->ev_fork[1 + (l % 3)];
      ARL_CHECK(c, cudaEventRecord(ev, st));
      ARL_CHECK(c, cudaStreamWaitEvent(wl, ev, 0));
    }
    if (pconv_wgrad_layer(c, l, obs16, gidx, gidx_off, n, wl, obs8)) return 1;
    prof_mark(c, kWgradName[l], wl);
    if (l == 0) break;
    if (P->early_jobs) {
      // its partials -> flat vector, on the head / FC side stream (idle by now), NOT inside side2's chain: the next
      // layer's weight gradient is the longest path to the join (timeline: profiles/r2_timeline.md)
      cudaStream_t fs = (wl != st && ws != st) ? ws : wl;
      if (fs != wl) {
        ARL_CHECK(c, cudaEventRecord(c->ev_fin[l & 1], wl));
        ARL_CHECK(c, cudaStreamWaitEvent(fs, c->ev_fin[l & 1], 0));
      }
      if (launch_finalize(c, P, 2 * l, 2, "finalize_conv", fs)) return 1;
    }
    if (pconv_dgrad_layer(c, l, n, st)) return 1;
    prof_mark(c, kDgradName[l], st);
  }
  if (c->sync_overlap_active) ARL_CHECK(c, cudaStreamWaitEvent(st, c->ev_cs_done, 0));
  if (ws != st) {
    ARL_CHECK(c, cudaEventRecord(c->ev_join, ws));
    ARL_CHECK(c, cudaStreamWaitEvent(st, c->ev_join, 0));
    if (used_side2) {
      ARL_CHECK(c, cudaEventRecord(c->ev_join2, c->side2));
      ARL_CHECK(c, cudaStreamWaitEvent(st, c->ev_join2, 0));
    }
  }
  for (int l = (int)c->conv.size() - 1; l >= 0 && !pcb; --l) {
    ConvLayer& L = c->conv[l];
    int rows = n * L.Ho * L.Wo;
    // weight-gradient partials (+ bias-gradient partials from the same pass over dY)
    WgradEpi e{};
    e.out = c->wgrad_partial[l]; e.mode = 0; e.Kvalid = L.K; e.Kp = L.K; e.ldo = L.Cout;
    e.bias_out = c->bias_partial[l];
    ConvLoader<64> a{};
    a.g = fwd_geom(L, l == 0 ? obs16 : c->conv[l - 1].act, n);
    if (l == 0) { a.g.idx = gidx; a.g.idx_off = gidx_off; }
    if (launch_wgrad_conv<ConvLoader<64>>(c, a, L.dact, L.Cout, rows, P->conv_rps[l], P->conv_splits[l], L.K / 64, e, st))
      return 1;
    prof_mark(c, kWgradName[l], st);
    if (l == 0) break;
    // dgrad into layer l-1's activation gradient (masked by its ReLU): all stride-parity classes in one launch
    ConvLayer& Lp = c->conv[l - 1];
    RowGemmMulti<ConvLoader<128>> mp{};
    int ncls = 0, max_tiles = 0;
    for (auto& d : L.dclasses) {
      int qrows = n * d.Qh * d.Qw;
      ConvGeom g{};
      g.src = L.dact; g.Qh = d.Qh; g.Qw = d.Qw; g.Hs = L.Ho; g.Ws = L.Wo; g.C = L.Cout;
      g.sy = 1; g.y0 = d.qy0; g.dty = -1; g.sx = 1; g.x0 = d.qx0; g.dtx = -1; g.Tx = d.Tx;
      g.nrows = qrows;
      mp.a[ncls].g = g;
      mp.b[ncls] = WeightSrc{d.wpack, (long)d.K, 0, RowPerm{0, 0}};
      RowEpi ep = make_epi(EPI_MASK_BF16);
      ep.out = Lp.dact; ep.act = Lp.act; ep.ldo = L.Cin; ep.M = qrows;
      ep.map_s = L.s; ep.map_y0 = L.s * d.qy0 + d.ry - L.p; ep.map_x0 = L.s * d.qx0 + d.rx - L.p;
      ep.map_Qh = d.Qh; ep.map_Qw = d.Qw; ep.map_H = L.Hin; ep.map_W = L.Win;
      if (L.s == 1 && ep.map_y0 == 0 && ep.map_x0 == 0 && d.Qh == L.Hin && d.Qw == L.Win) ep.map_s = 0;
      mp.e[ncls] = ep;
      mp.mtiles[ncls] = (qrows + 127) / 128;
      mp.num_kb[ncls] = d.K / 64;
      max_tiles = std::max(max_tiles, mp.mtiles[ncls]);
      ++ncls;
    }
    if (launch_conv_persist_bn(c, L.Cin, mp, ncls, L.dclasses[0].K, max_tiles, st)) return 1;
    prof_mark(c, kDgradName[l], st);
  }
  // ---- sum partials, scatter into the flat gradient ----
  if (stream_update_ok(c, fct)) {
    c->pending_fin = P;        // clip_update: update_stream_kernel sums the partials and updates in the same pass
    c->pending_stream = true;
  } else {
    if (launch_finalize(c, P, 0, (pcb && P->early_jobs) ? 2 : P->n_jobs, "finalize_grads", st)) return 1;
  }
  return 0;
}

// with_fc: also cast the FC weights (explicit set_params / sync path); the single-GPU update kernel refreshes
// that copy itself.  advance: this call closes an update -> bump the device counters.
int pack_weights(arl_ctx* c, cudaStream_t st, bool with_fc = true, bool advance = false) {
  if (!c->params) ARL_FAIL(c, "parameters not bound");
  dim3 grid(with_fc ? 148 * 4 : 32, with_fc ? c->n_pack_jobs : c->n_pack_jobs - 1);
  ARL_CHECK(c, launch_k(pack_weights_kernel, dim3(grid), dim3(256), 0, st, c->pack_jobs_dev, c->params, advance ? c->step : nullptr, c->log_slot,
                                            c->mb_counter));
  c->launches++;
  prof_mark(c, "pack_weights", st);
  ARL_CHECK(c, cudaGetLastError());
  return 0;
}

UpdateParams update_params(arl_ctx* c, float gscale, bool* fused_cast_out);

// the FC weights can be updated before the rest of the backward pass finishes when nothing couples them to the other
// gradients: no global-norm clipping, local update, operand copy refreshed by the update kernel itself.
// OFF by default (ARL_EARLY_FC=1 enables): bit-identical results (tests/test_gpu_path.py), but measured SLOWER on B200 —
// 61.8 vs 60.5 ms per PPO iteration.  The minibatch is bound by total SM time, not by its critical path: every GEMM
// kernel is a persistent 1-CTA/SM grid, so the update kernel taken off the tail only time-slices the same SMs, and
// the split costs a second launch (22.7 us range update + 8.9 us rest, against 22.4 us for the single fused update).
bool early_fc_ok(arl_ctx* c) {
  static const bool on = getenv("ARL_EARLY_FC") && atoi(getenv("ARL_EARLY_FC")) != 0;
  static const bool fused = !(getenv("ARL_FUSED_UPDATE") && atoi(getenv("ARL_FUSED_UPDATE")) == 0);
  return on && fused && c->train_step_active && c->opt_set && c->opt.grad_norm_clip <= 0.f && c->m && c->v &&
         c->pc_mode >= 2 && fc_tiles_ok(c) && (c->off_Wfc % 4 == 0) && (((long)c->Kfc * c->H) % 4 == 0);
}

// Without global-norm clipping nothing in the update depends on the norm: finalisation, update, operand refresh and
// the logs go into ONE plain launch (update_stream_kernel).  Needs the local fused step to follow (train_minibatches /
// the profiling graph), the FC weight gradient written straight into the flat vector, and the slot table that lets the
// updating thread refresh the conv operand packs.  ARL_STREAM_UPDATE=0 restores finalize_grads + update_fused.
bool stream_update_ok(arl_ctx* c, bool fct) {
  static const bool on = !(getenv("ARL_STREAM_UPDATE") && atoi(getenv("ARL_STREAM_UPDATE")) == 0);
  if (!(on && fct && c->train_step_active && c->opt_set && c->m && c->v && c->opt.grad_norm_clip <= 0.f)) return false;
  if ((c->off_Wfc % 4) || (((long)c->Kfc * c->H) % 4)) return false;
  bool fused_cast = false;
  UpdateParams u = update_params(c, 1.f, &fused_cast);
  return fused_cast && u.shadow && c->pc_mode >= 2 && c->n_pack_jobs >= 2 && c->conv_pack_end > 0 && c->pk_slots;
}

int early_fc_update(arl_ctx* c, cudaStream_t st) {
  bool fused_cast = false;
  UpdateParams u = update_params(c, 1.f, &fused_cast);
  if (!fused_cast) ARL_FAIL(c, "early FC update needs the fused operand copy");
  ARL_CHECK(c, launch_k(update_range_kernel, dim3(kEarlyBlocks), dim3(256), 0, st, u, c->off_Wfc / 4, (long)c->Kfc * c->H / 4,
                        c->sumsq_partial_fc));
  c->launches++;
  prof_mark(c, "fc_update", st);
  ARL_CHECK(c, cudaGetLastError());
  c->early_fc_done = true;
  return 0;
}

UpdateParams update_params(arl_ctx* c, float gscale, bool* fused_cast_out) {
  UpdateParams u{};
  u.param = c->params; u.grad = c->grad; u.m = c->m; u.v = c->v; u.n = c->n_params;
  u.sumsq_partial = c->sumsq_partial; u.n_partial = kSumsqBlocks;
  u.loss_partial = c->loss_partial; u.n_loss_blocks = c->n_loss_rows;
  u.hyper = c->hyper; u.step = c->step; u.kind = c->opt.update;
  u.lr = c->opt.learning_rate; u.beta1 = c->opt.beta1; u.beta2 = c->opt.beta2; u.eps = c->opt.epsilon;
  u.rho = c->opt.rho; u.clip = c->opt.grad_norm_clip; u.gscale = gscale;
  u.out_norm = c->log_norm; u.out_loss = c->log_loss; u.log_slot = c->log_slot; u.log_cap = c->log_cap;
  u.shadow = c->wfc_bf16; u.shadow_begin = c->off_Wfc; u.shadow_end = c->off_Wfc + (long)c->Kfc * c->H;
  bool fused_cast = (c->off_Wfc % 4 == 0) && (c->H % 4 == 0);
  if (fc_tiles_ok(c)) {
    u.shadow = c->wfc_t; u.shadow_tiles = 1; u.shadow_HW = c->HWlast; u.shadow_H = c->H;
    u.shadow_H_magic = pc_magic((uint32_t)c->H); u.shadow_HW_magic = pc_magic((uint32_t)c->HWlast);
    if ((long)c->Kfc * c->H >= (1L << 32) / c->H) u.shadow = nullptr;   // outside the exact range of the magic division
  }
  if (!fused_cast) u.shadow = nullptr;
  *fused_cast_out = fused_cast && (u.shadow != nullptr || !fc_tiles_ok(c));
  return u;
}

int clip_update(arl_ctx* c, float gscale, cudaStream_t st) {
  NvtxRange nvtx_("clip + update");
  if (!c->opt_set) ARL_FAIL(c, "optimizer not configured");
  if (!c->m || !c->v) ARL_FAIL(c, "optimizer state not bound");
  bool fused_cast = false, scattered = false;
  UpdateParams u = update_params(c, gscale, &fused_cast);
  if (c->early_fc_done) {
    // the FC range is done (update_range_kernel): norm over the rest + its partials, update the rest
    u.skip4_begin = c->off_Wfc / 4; u.skip4_len = (long)c->Kfc * c->H / 4;
    u.sumsq_partial2 = c->sumsq_partial_fc; u.n_partial2 = kEarlyBlocks;
    c->early_fc_done = false;
  }
  long P_total = 0;
  if (c->pending_fin) {
    const TrainPlan* P = static_cast<const TrainPlan*>(c->pending_fin);
    P_total = P->fin_total;
    u.fin_jobs = P->jobs_dev; u.n_fin_jobs = P->n_jobs;
    if (P->early_jobs) {
      // everything but layer 0 was finalised on the side streams: part A sums layer 0's partials, part A2 takes the
      // other small tensors from the flat gradient
      u.n_fin_jobs = 2; P_total = P->job_total[0] + P->job_total[1];
      u.a2_begin = c->conv[1].off_W; u.a2_mid = c->off_Wfc; u.a2_resume = c->off_Wfc + (long)c->Kfc * c->H;
    }
    u.fc4_begin = c->off_Wfc / 4; u.fc4_len = (long)c->Kfc * c->H / 4;
    c->pending_fin = nullptr;
  }
  if (c->pending_stream) {
    c->pending_stream = false;
    if (gscale != 1.f || u.n_fin_jobs == 0) ARL_FAIL(c, "update_stream: unexpected configuration");
    u.pk_slots = c->pk_slots; u.n_pk_jobs = c->n_pack_jobs - 1; u.conv_end = c->conv_pack_end;
    u.adv_done = c->ticket + 1; u.adv_log_slot = c->log_slot; u.adv_mb = c->mb_counter;
    // blocks [0, nA) sum the split partials and update everything except the FC weights, the other 4 x 148 stream the FC range
    const int nA = (int)((P_total + kFinPerBlock - 1) / kFinPerBlock);
    const long n_a2 = u.a2_resume > 0 ? (u.a2_mid - u.a2_begin) + (c->n_params - u.a2_resume) : 0;
    const int nA2 = (int)((n_a2 + 255) / 256);
    int nB = kSumsqBlocks;
    if (u.skip4_len > 0) {
      // the FC range already took its step (update_range_kernel, beside the conv gradient chain): only part A is left
      if (u.skip4_begin != u.fc4_begin || u.skip4_len != u.fc4_len) ARL_FAIL(c, "update_stream: early range is not the FC range");
      u.fc4_len = 0; nB = 0;
    }
    if (nA + nA2 + nB > kStreamPartials) ARL_FAIL(c, "update_stream: partial buffer too small");
    u.adv_done = c->ticket + 2;                 // (its own arrival counter: the grid size differs from update_fused_kernel's)
    const int G = nA + nA2 + nB;
    if (c->split_capture && c->side && c->ev_updB && nB > 0 && nA + nA2 > 0 && st != c->side) {
      // two grids over the same block index space: the FC range (98 % of the bytes) on the side stream, joined by the
      // next FC forward (forward_trunk) or the end of the capture; everything the next conv layers read on `st`
      ARL_CHECK(c, cudaEventRecord(c->ev_upd_fork, st));
      ARL_CHECK(c, cudaStreamWaitEvent(c->side, c->ev_upd_fork, 0));
      ARL_CHECK(c, launch_k(update_stream_kernel, dim3(nB), dim3(256), 0, c->side, u, c->sumsq_partial, nA, nA2, P_total, nA + nA2, G,
                            (unsigned long long*)nullptr, 0));
      ARL_CHECK(c, cudaEventRecord(c->ev_updB, c->side));
      ARL_CHECK(c, launch_k(update_stream_kernel, dim3(nA + nA2), dim3(256), 0, st, u, c->sumsq_partial, nA, nA2, P_total, 0, G,
                            c->ticket + 3, 0));
      c->updB_pending = true;
      c->launches += 2;
      ARL_CHECK(c, cudaGetLastError());
      return 0;
    }
    ARL_CHECK(c, launch_k_win(c->l2_on ? &c->l2win : nullptr, update_stream_kernel, dim3(G), dim3(256), 0, st, u, c->sumsq_partial,
                              nA, nA2, P_total, 0, G, (unsigned long long*)nullptr, 1));
    c->launches++;
    prof_mark(c, "clip_update", st);
    ARL_CHECK(c, cudaGetLastError());
    return 0;
  }
  // norm + clip + update in one launch (update_fused_kernel); ARL_FUSED_UPDATE=0 keeps the two-kernel form
  static const bool fused = !(getenv("ARL_FUSED_UPDATE") && atoi(getenv("ARL_FUSED_UPDATE")) == 0);
  if (fused) {
    // conv operand packs refreshed by the update itself + counters advanced by its last block: no pack launch at all
    static const bool scatter_on = !(getenv("ARL_SCATTER_PACK") && atoi(getenv("ARL_SCATTER_PACK")) == 0);
    if (scatter_on && fused_cast && c->pc_mode >= 2 && c->n_pack_jobs >= 2 && c->conv_pack_end > 0 && c->pk_slots) {
      u.pk_slots = c->pk_slots; u.n_pk_jobs = c->n_pack_jobs - 1;       // (the last job is the FC copy: fused_cast)
      u.conv_end = c->conv_pack_end;
      u.adv_done = c->ticket + 1; u.adv_log_slot = c->log_slot; u.adv_mb = c->mb_counter;
      scattered = true;
    }
    ARL_CHECK(c, launch_coop(update_fused_kernel, dim3(kSumsqBlocks), dim3(256), 0, st, u, c->sumsq_partial, c->ticket));
  } else {
    ARL_CHECK(c, launch_k(sumsq_kernel, dim3(kSumsqBlocks), dim3(256), 0, st, c->grad, c->n_params, gscale, c->sumsq_partial));
    c->launches++;
    prof_mark(c, "grad_sumsq", st);
    ARL_CHECK(c, launch_k(update_kernel, dim3(148 * 4), dim3(256), 0, st, u));
  }
  c->launches++;
  prof_mark(c, "clip_update", st);
  ARL_CHECK(c, cudaGetLastError());
  if (scattered) return 0;
  return pack_weights(c, st, !fused_cast, true);
}

// ---------------------------------------------------------------------------
// sampler
// ---------------------------------------------------------------------------
SynthCfg synth_cfg(const arl_sampler_cfg& s) {
  SynthCfg k{};
  k.pool_frames = s.pool_frames; k.lives0 = s.lives0; k.life_base = s.life_base; k.life_mul = s.life_mul;
  k.life_mod = s.life_mod; k.reward_mod = s.reward_mod; k.frame_stride = s.frame_stride; k.n_games = s.n_games;
  return k;
}

// frame pipeline for envs [e0, e0 + n) (n < 0: all).  The kernels index everything by env, so a sub-range is the same
// launch over base pointers advanced by e0 envs (staging is the base of the whole [B][2][frame] block).
// Lean rollout (mid_batch_reset samplers that record observations and keep the bf16 mirror): inside a batch the current
// stack of env e is rollout row e*T + s — the frame kernel reads the three kept planes from there and writes only row
// e*T + s + 1 (u8 + mirror), the policy's first conv layer gathers rows e*T + s of the mirror; the step buffer and its
// mirror are written by the batch's LAST step only (they are what extra_observations, the next batch's row 0 and the
// caller see).  Saves one u8 stack and one bf16 stack of HBM writes per env-step (291 840 -> 192 000 B).  ARL_FRAME_LEAN=0:
// every step also updates the step buffers (A/B).
bool frame_lean(arl_ctx* c) {
  static const bool on = !(getenv("ARL_FRAME_LEAN") && atoi(getenv("ARL_FRAME_LEAN")) == 0);
  const arl_sampler_cfg& s = c->sc;
  return on && s.mid_batch_reset && s.observations && c->roll_obs16 && c->step_obs16 && s.horizon > 1;
}

int launch_frame(arl_ctx* c, const uint8_t* staging, int s_next, bool to_rollout, cudaStream_t st, int e0 = 0, int n = -1) {
  NvtxRange nvtx_("frame");
  const arl_sampler_cfg& s = c->sc;
  if (n < 0) n = s.n_envs - e0;
  const long T = s.horizon;
  const long obs_bytes = (long)s.planes * c->cfg.in_h * c->cfg.in_w;
  const long fbytes = (long)kRawH * kRawW * (s.frame_mode == 1 ? 3 : 1);
  const uint8_t* stg = staging ? staging + (long)e0 * 2 * fbytes : nullptr;
  const FrameCmd* cmd = c->cmd + e0;
  uint8_t* step_obs = s.step_obs + (long)e0 * obs_bytes;
  uint8_t* roll_obs = (to_rollout && s.observations) ? s.observations + (long)e0 * T * obs_bytes : nullptr;
  __nv_bfloat16* step16 = c->step_obs16 ? c->step_obs16 + (long)e0 * c->obs16_elems : nullptr;
  __nv_bfloat16* roll16 = (to_rollout && c->roll_obs16) ? c->roll_obs16 + (long)e0 * T * c->obs16_elems : nullptr;
  const uint8_t* prev = step_obs;
  long prev_stride = obs_bytes;
  int lean = 0;
  if (frame_lean(c) && s_next >= 1) {            // s_next in [1, T]: a step of the batch (0: start / reset / warm-up)
    lean = 1;
    prev = s.observations + ((long)e0 * T + (s_next - 1)) * obs_bytes;
    prev_stride = T * obs_bytes;
    if (to_rollout) { step_obs = nullptr; step16 = nullptr; }
  }
  if (s.frame_mode == 1) {
    // north-star frames: RGB pool / staging -> gray -> 84x84 (frame_rgb_roll_kernel)
    ARL_CHECK(c, launch_k(frame_rgb_roll_kernel, dim3(n * (kNsH / kRgbRows)), dim3(kRgbThreads), 0, st, s.frame_pool, stg,
                          cmd, step_obs, roll_obs, step16, roll16, s.horizon, s_next, n, s.planes, c->pc_mode >= 2,
                          c->pc_mode >= 2, c->obs16_elems, prev, prev_stride, lean));
    c->launches++;
    prof_mark(c, "frame", st);
    ARL_CHECK(c, cudaGetLastError());
    return 0;
  }
  long items = (long)n * 520;
  int blocks = (int)((items + 255) / 256);
  ARL_CHECK(c, launch_k(frame_kernel, dim3(blocks), dim3(256), 0, st, s.frame_pool, stg, cmd, step_obs, roll_obs, step16, roll16,
                        s.horizon, s_next, n, s.planes, c->pc_mode >= 2, c->pc_mode >= 2, prev, prev_stride, lean));
  c->launches++;
  prof_mark(c, "frame", st);
  ARL_CHECK(c, cudaGetLastError());
  return 0;
}

int rollout_begin(arl_ctx* c, cudaStream_t st) {
  const arl_sampler_cfg& s = c->sc;
  int row_bytes = s.planes * c->cfg.in_h * c->cfg.in_w;
  long chunks = (long)s.n_envs * (row_bytes / 16);
  // observations[e*T + 0] = step_obs[e]   (worker.py:31-32) — and the same for the bf16 mirror
  if (s.observations) {
    copy_rows_kernel<<<(int)((chunks + 255) / 256), 256, 0, st>>>(s.step_obs, row_bytes, nullptr, s.observations, row_bytes,
                                                                 c->rows_tab, s.n_envs, row_bytes);
    c->launches += 1;
    if (c->roll_obs16) {
      int row16 = (int)(c->obs16_elems * 2);
      long chunks16 = (long)s.n_envs * (row16 / 16);
      copy_rows_kernel<<<(int)((chunks16 + 255) / 256), 256, 0, st>>>(
          reinterpret_cast<const uint8_t*>(c->step_obs16), row16, nullptr, reinterpret_cast<uint8_t*>(c->roll_obs16), row16,
          c->rows_tab, s.n_envs, row16);
      c->launches += 1;
    }
  }
  ARL_CHECK(c, cudaMemsetAsync(c->tout.count, 0, sizeof(int), st));
  ARL_CHECK(c, cudaGetLastError());
  return 0;
}

int rollout_step(arl_ctx* c, int s_idx, const uint8_t* staging, cudaStream_t st) {
  NvtxRange nvtx_("serve: fwd + sample + env step + frame");
  const arl_sampler_cfg& s = c->sc;
  const int B = s.n_envs, T = s.horizon;
  // the env step of env e runs in the head kernel's block e right after its action is sampled
  EnvStepArgs es{synth_cfg(s), c->est, c->tout, c->cmd, s.rewards, s.dones, s.raw_reward, s.need_reset, B, T, s_idx,
                 s.max_path_length, s.discount, s.mid_batch_reset, s.clip_reward, s.episodic_lives, nullptr, 0};
  const bool u8 = c->pc_mode >= 2 && u8_conv0_ok(c);
  const bool lean = frame_lean(c) && !u8;         // the current stacks are rollout rows e*T + s_idx (gathered through rows_tab)
  if (policy_forward16(c, lean ? c->roll_obs16 : c->step_obs16, B, c->rows_tab + (long)s_idx * B, s.prob, s.value,
                       s.uniforms + (long)s_idx * B, s.actions, c->pc_mode >= 2, st, &es, u8 ? s.step_obs : nullptr,
                       lean ? c->rows_tab + (long)s_idx * B : nullptr))
    return 1;
  return launch_frame(c, staging, s_idx + 1, s_idx + 1 < T, st);
}

int rollout_end(arl_ctx* c, cudaStream_t st) {
  const arl_sampler_cfg& s = c->sc;
  if (s.extra_observations) {
    size_t bytes = (size_t)s.n_envs * s.planes * c->cfg.in_h * c->cfg.in_w;
    ARL_CHECK(c, cudaMemcpyAsync(s.extra_observations, s.step_obs, bytes, cudaMemcpyDeviceToDevice, st));
  }
  if (!s.mid_batch_reset && !s.ext_emulator) {   // (external emulators: the workers reset, then arl_rollout_ingest(s = T))
    env_reset_needed_kernel<<<(s.n_envs + 127) / 128, 128, 0, st>>>(synth_cfg(s), c->est, c->cmd, s.n_envs);
    c->launches++;
    ARL_CHECK(c, cudaGetLastError());
    if (launch_frame(c, nullptr, 0, false, st)) return 1;
  }
  return 0;
}

}  // namespace

namespace {
int sync_update(arl_ctx* c, cudaStream_t st);
bool sync_overlap_ok(arl_ctx* c);
int sync_overlap_prepare(arl_ctx* c);
int sync_tail(arl_ctx* c, cudaStream_t st);
int train_minibatches(arl_ctx* c, const int* idx, int mb_size, int count, int sync, cudaStream_t st);
int async_push_pull(arl_ctx* c, cudaStream_t st);
}

// ===========================================================================
// extern "C"
// ===========================================================================
extern "C" {

int arl_create(const arl_net_cfg* cfg, arl_ctx** out) {
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e);
    return 3;
  }
  cudaDeviceProp prop{};
  int dev = 0;
  cudaGetDevice(&dev);
  cudaGetDeviceProperties(&prop, dev);
  if (prop.major != 10) {
    g_create_error = "libaccelrl_b200 requires an sm_100a device (found sm_" + std::to_string(prop.major) +
                     std::to_string(prop.minor) + ")";
    return 3;
  }
  arl_ctx* c = new arl_ctx();
  c->cfg = *cfg;
  int rc0 = plan_net(c) || plan_pconv(c);
  // conv path: the patch-resident tiles whenever the geometry allows; ARL_PCONV=0/1 forces the gather path for
  // everything / for training only (A/B measurements)
  c->pc_mode = c->pc.empty() ? 0 : 2;
  if (const char* ev = getenv("ARL_PDL")) g_pdl = atoi(ev) != 0;
  if (const char* ev = getenv("ARL_CARVEOUT")) g_carveout = atoi(ev);
  if (const char* ev = getenv("ARL_DGRAD_CTAS")) c->dgrad_ctas = atoi(ev);
  if (const char* ev = getenv("ARL_WGRAD_CTAS")) c->wgrad_ctas = atoi(ev);
  if (const char* ev = getenv("ARL_PCONV")) c->pc_mode = c->pc.empty() ? 0 : std::max(0, std::min(2, atoi(ev)));
  if (!c->pc.empty() && c->pc[0].S * 64L != c->obs16_elems) {
    // the patch-resident path pads the first layer's image stride; the gather tiles cannot read that layout
    if (c->pc_mode == 1) c->pc_mode = 0;
    if (c->pc_mode == 2) c->obs16_elems = c->pc[0].S * 64L;
  }
  if (rc0 || alloc_net(c)) {
    g_create_error = c->err;
    delete c;
    return 4;
  }
  *out = c;
  return 0;
}

void arl_destroy(arl_ctx* c) {
  if (!c) return;
  cudaDeviceSynchronize();
  if (c->l2_on && c->l2_carved) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
  comm_destroy(c->comm);
  async_destroy(c->async_);
  // device workspaces are released with the context's process lifetime; free the large ones explicitly
  for (auto& L : c->conv) {
    cudaFree(L.act); cudaFree(L.dact); cudaFree(L.wpack);
    for (auto& d : L.dclasses) cudaFree(d.wpack);
  }
  for (auto p : c->wgrad_partial) cudaFree(p);
  for (auto p : c->bias_partial) cudaFree(p);
  if (c->shadow_in_comm) { (fc_tiles_ok(c) ? c->wfc_t : c->wfc_bf16) = nullptr; }
  // patch-resident conv grids / operand packs and the FC tile planes
  for (auto& q : c->pc) { cudaFree(q.in); cudaFree(q.dY_base ? q.dY_base : q.dY); cudaFree(q.wpack); cudaFree(q.dwpack); }
  cudaFree(c->act_fc); cudaFree(c->wfc_t); cudaFree(c->dh_t); cudaFree(c->tl_buf);
  cudaFree(c->wfc_bf16); cudaFree(c->obs16_stage); cudaFree(c->step_obs16); cudaFree(c->roll_obs16); cudaFree(c->pack_jobs_dev); cudaFree(c->fc_partial); cudaFree(c->h); cudaFree(c->dh);
  cudaFree(c->dlogit); cudaFree(c->head_partial); cudaFree(c->head_b_partial); cudaFree(c->loss_partial);
  cudaFree(c->pk_slots);
  cudaFree(c->sumsq_partial); cudaFree(c->sumsq_partial_fc); cudaFree(c->ss_fin); cudaFree(c->ss_fc); cudaFree(c->hyper); cudaFree(c->step); cudaFree(c->log_slot); cudaFree(c->log_norm);
  cudaFree(c->log_loss); cudaFree(c->mb_counter); cudaFree(c->valid_count);
  for (auto& kv : c->plans) cudaFree(kv.second.jobs_dev);
  if (c->rollout_graph) cudaGraphExecDestroy(c->rollout_graph);
  if (c->train_graph) cudaGraphExecDestroy(c->train_graph);
  // forked streams / fork-join events of the training step (created lazily by grad_minibatch)
  for (auto& e : c->ev_fork) if (e) cudaEventDestroy(e);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  if (c->ev_join2) cudaEventDestroy(c->ev_join2);
  if (c->ev_fcd) cudaEventDestroy(c->ev_fcd);
  if (c->ev_upd_fork) cudaEventDestroy(c->ev_upd_fork);
  if (c->ev_updB) cudaEventDestroy(c->ev_updB);
  for (auto& e : c->ev_fin) if (e) cudaEventDestroy(e);
  for (auto& e : c->ev_cs_in) if (e) cudaEventDestroy(e);
  if (c->ev_cs_done) cudaEventDestroy(c->ev_cs_done);
  if (c->cs) cudaStreamDestroy(c->cs);
  if (c->sync_tail_partial) cudaFree(c->sync_tail_partial);
  if (c->side) cudaStreamDestroy(c->side);
  if (c->side2) cudaStreamDestroy(c->side2);
  cudaFree(c->est.f); cudaFree(c->cmd); cudaFree(c->rows_tab); cudaFree(c->tout.count);
  {
    arl_ctx::SamplerSlot& o = c->slots[1 - c->cur_slot];     // the parked sampler
    if (o.rollout_graph) cudaGraphExecDestroy(o.rollout_graph);
    cudaFree(o.est.f); cudaFree(o.cmd); cudaFree(o.rows_tab); cudaFree(o.tout.count);
    cudaFree(o.step_obs16); cudaFree(o.roll_obs16);
  }
  delete c;
}

const char* arl_last_error(arl_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int arl_device_error(arl_ctx* c) {
  int v = 0;
  cudaMemcpyFromSymbol(&v, g_dev_error, sizeof(int));
  (void)c;
  return v;
}

long arl_param_count(arl_ctx* c) { return c->n_params; }

int arl_param_layout(arl_ctx* c, long* offsets, long* sizes, int cap) {
  int n = (int)c->lay_off.size();
  for (int i = 0; i < n && i < cap; ++i) { offsets[i] = c->lay_off[i]; sizes[i] = c->lay_size[i]; }
  return n;
}

// L2 residency of the optimiser state (fp32 params, m, v: 3 x 14.5 MB, read and rewritten by the update kernel once per
// minibatch and by nothing else): a persisting carve-out of that size + an access-policy window on the update kernel's
// launches, so the three vectors stay in the 126 MB L2 between updates instead of being streamed from HBM 256 times per
// iteration.  Applies when the three vectors sit in one allocation (engine.py; not the synchronous learner, whose
// parameters live in the symmetric peer allocation).  ARL_L2_PERSIST: 0 off; 1 (2: + a line on stderr) carve-out always on (rollouts lose 43 MB of
// L2: 7.5 -> 7.9..9.2 ms); 3 (default) carve-out only while minibatches train (the runner has read the previous phase back
// before the other starts): 51.3 -> 50.3 ms per iteration; 4 also the gradient vector (no further gain).
void l2_carve(arl_ctx* c, bool on) {
  if (!c->l2_toggle || c->l2_carved == on) return;
  cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, on ? c->l2_carve : 0);
  c->l2_carved = on;
}

void l2_persist_setup(arl_ctx* c) {
  static const int mode = getenv("ARL_L2_PERSIST") ? atoi(getenv("ARL_L2_PERSIST")) : 3;
  if (c->l2_on && c->l2_carved) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);     // re-binding: start over
  c->l2_on = false; c->l2_toggle = false; c->l2_carved = false;
  if (!mode || !c->params || !c->m || !c->v) return;
  const char* lo = reinterpret_cast<const char*>(std::min(c->params, std::min(c->m, c->v)));
  const char* hi = reinterpret_cast<const char*>(std::max(c->params, std::max(c->m, c->v))) + c->n_params * sizeof(float);
  size_t bytes = (size_t)(hi - lo);
  // exactly the layout engine.py makes — [params | m | v] at one common pitch of n_params rounded up to 64 floats — and
  // nothing else: a synchronous / asynchronous learner's parameters live in the peer-shared allocation (no window there,
  // and no carve-out toggling next to NVLink traffic)
  const long pitch = (long)(c->m - c->params);
  if (pitch < c->n_params || pitch > c->n_params + 64 || (long)(c->v - c->m) != pitch) return;
  // the gradient right behind them (engine.py): the window covers it too (mode 4)
  c->l2_grad_in_window = false;
  if (mode == 4 && c->grad && reinterpret_cast<const char*>(c->grad) >= hi &&
      reinterpret_cast<const char*>(c->grad) < hi + 4096) {
    hi = reinterpret_cast<const char*>(c->grad) + c->n_params * sizeof(float);
    bytes = (size_t)(hi - lo);
    c->l2_grad_in_window = true;
  }
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceProp prop{};
  cudaGetDeviceProperties(&prop, dev);
  const size_t carve = std::min(bytes, (size_t)prop.persistingL2CacheMaxSize);
  if (carve == 0 || cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve) != cudaSuccess) { cudaGetLastError(); return; }
  c->l2win.base_ptr = const_cast<char*>(lo);
  c->l2win.num_bytes = std::min(bytes, (size_t)prop.accessPolicyMaxWindowSize);
  c->l2win.hitRatio = (float)std::min(1.0, (double)carve / (double)c->l2win.num_bytes);
  c->l2win.hitProp = cudaAccessPropertyPersisting;
  c->l2win.missProp = cudaAccessPropertyStreaming;
  c->l2_on = true;
  c->l2_carve = carve; c->l2_toggle = (mode >= 3); c->l2_carved = true;
  if (mode == 2)
    fprintf(stderr, "[accel_rl_b200] L2 persistence: window %.1f MB, carve-out %.1f MB (max %.1f MB, L2 %.1f MB), hit ratio %.2f\n",
            c->l2win.num_bytes / 1e6, carve / 1e6, prop.persistingL2CacheMaxSize / 1e6, prop.l2CacheSize / 1e6, c->l2win.hitRatio);
}

int arl_bind_params(arl_ctx* c, float* params, float* grad, float* m, float* v) {
  c->params = params; c->grad = grad; c->m = m; c->v = v;
  l2_persist_setup(c);
  if (c->train_graph) { cudaGraphExecDestroy(c->train_graph); c->train_graph = nullptr; }
  if (c->rollout_graph) { cudaGraphExecDestroy(c->rollout_graph); c->rollout_graph = nullptr; }
  for (auto& o : c->slots)
    if (o.rollout_graph) { cudaGraphExecDestroy(o.rollout_graph); o.rollout_graph = nullptr; }
  return 0;
}

int arl_pack_weights(arl_ctx* c, void* stream) { return pack_weights(c, (cudaStream_t)stream); }

int arl_policy_forward(arl_ctx* c, const uint8_t* obs, const int* idx, int n, const int* out_rows, float* prob,
                       float* value, const double* uniforms, uint8_t* actions, void* stream) {
  return policy_forward(c, obs, idx, n, out_rows, prob, value, uniforms, actions, (cudaStream_t)stream);
}

int arl_sample_actions(arl_ctx* c, const float* prob, const double* uniforms, uint8_t* actions, int n, int n_actions,
                       void* stream) {
  sample_actions_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(prob, uniforms, actions, n, n_actions);
  c->launches++;
  ARL_CHECK(c, cudaGetLastError());
  return 0;
}

int arl_frame_update(arl_ctx* c, const uint8_t* raw_a, const uint8_t* raw_b, const uint8_t* reset_mask, uint8_t* stack,
                     int n, int planes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  // frames of item i live at raw_a + i*33600 / raw_b + i*33600: reuse frame_kernel's pool addressing with
  // per-item indices, two passes of the same arithmetic are avoided by a tiny cmd table.
  FrameCmd* cmd = nullptr;
  ARL_CHECK(c, cudaMallocAsync(reinterpret_cast<void**>(&cmd), (size_t)n * sizeof(FrameCmd), st));
  make_cmd_kernel<<<(n + 255) / 256, 256, 0, st>>>(cmd, reset_mask, n, raw_a != nullptr);
  long items = (long)n * 520;
  frame_pair_kernel<<<(int)((items + 255) / 256), 256, 0, st>>>(raw_a, raw_b, cmd, stack, n, planes);
  c->launches += 2;
  ARL_CHECK(c, cudaGetLastError());
  ARL_CHECK(c, cudaFreeAsync(cmd, st));
  return 0;
}

/* north-star frame mode: RGB (210,160,3) frame pairs -> (planes,84,84) u8 stack + bf16 copy (frame_rgb_kernel) */
int arl_frame_update_rgb(arl_ctx* c, const uint8_t* raw_a, const uint8_t* raw_b, const uint8_t* reset_mask, uint8_t* stack,
                         uint16_t* stack_bf16, int n, int planes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n <= 0) return 0;
  if (planes < 1 || planes > 8) ARL_FAIL(c, "planes must be in [1,8]");
  if (!raw_b || !stack) ARL_FAIL(c, "raw_b and stack are required");
  frame_rgb_kernel<<<n * (kNsH / kRgbRows), kRgbThreads, 0, st>>>(raw_a, raw_b, reset_mask, stack, reinterpret_cast<__nv_bfloat16*>(stack_bf16),
                                                  n, planes);
  c->launches++;
  ARL_CHECK(c, cudaGetLastError());
  return 0;
}

int arl_sampler_configure(arl_ctx* c, const arl_sampler_cfg* cfg) {
  if (c->sampler_set) {
    // re-configuring the selected slot: release what the previous configuration allocated
    if (c->rollout_graph) { cudaGraphExecDestroy(c->rollout_graph); c->rollout_graph = nullptr; }
    // the training graph has the rollout mirror's address baked in (conv0 gathers from it)
    if (c->train_graph) { cudaGraphExecDestroy(c->train_graph); c->train_graph = nullptr; }
    cudaFree(c->est.f); cudaFree(c->cmd); cudaFree(c->rows_tab); cudaFree(c->tout.count);
    cudaFree(c->step_obs16); cudaFree(c->roll_obs16);
    c->est = EnvState{}; c->tout = TrajOut{}; c->cmd = nullptr; c->rows_tab = nullptr;
    c->step_obs16 = nullptr; c->roll_obs16 = nullptr;
    c->sampler_set = false;
  }
  c->sc = *cfg;
  const int B = cfg->n_envs, T = cfg->horizon;
  if (cfg->planes != c->cfg.in_c) ARL_FAIL(c, "sampler planes != network input channels");
  if (cfg->frame_mode == 0 && (c->cfg.in_h != kObsH || c->cfg.in_w != kObsW))
    ARL_FAIL(c, "frame_mode 0 (reference frames) produces 104x80 observations");
  if (cfg->frame_mode == 1 && (c->cfg.in_h != kNsH || c->cfg.in_w != kNsW))
    ARL_FAIL(c, "frame_mode 1 (RGB frames) produces 84x84 observations");
  if (cfg->frame_mode != 0 && cfg->frame_mode != 1) ARL_FAIL(c, "frame_mode must be 0 or 1");
  if (cfg->n_games > 1 && !cfg->ext_emulator && (cfg->pool_frames % cfg->n_games))
    ARL_FAIL(c, "pool_frames must be a multiple of n_games");
  if ((cfg->planes * c->cfg.in_h * c->cfg.in_w) % 16) ARL_FAIL(c, "observation bytes must be a multiple of 16");
  if (B > c->cfg.max_rows) ARL_FAIL(c, "n_envs larger than max_rows");
  // env state block: 4 int arrays + ... allocate separately for clarity
  int* iblock = nullptr;
  if (dev_alloc(c, &iblock, (size_t)B * 10)) return 1;
  c->est.f = iblock; c->est.lives_seen = iblock + B; c->est.need_reset = iblock + 2 * B; c->est.traj_len = iblock + 3 * B;
  c->est.traj_nz = iblock + 4 * B;
  float* fblock = reinterpret_cast<float*>(iblock + 5 * B);
  c->est.traj_ret = fblock; c->est.traj_raw = fblock + B; c->est.traj_disc = fblock + 2 * B;
  c->est.traj_cur = fblock + 3 * B;
  if (dev_alloc(c, &c->cmd, (size_t)B)) return 1;
  int cap = cfg->traj_cap > 0 ? cfg->traj_cap : 4 * B * 4;
  int* tblock = nullptr;
  if (dev_alloc(c, &tblock, (size_t)1 + (size_t)cap * 6)) return 1;
  c->tout.count = tblock; c->tout.cap = cap; c->tout.env = tblock + 1; c->tout.len = tblock + 1 + cap;
  c->tout.nz = tblock + 1 + 2 * cap;
  c->tout.ret = reinterpret_cast<float*>(tblock + 1 + 3 * cap);
  c->tout.raw = reinterpret_cast<float*>(tblock + 1 + 4 * cap);
  c->tout.disc = reinterpret_cast<float*>(tblock + 1 + 5 * cap);
  if (dev_alloc(c, &c->rows_tab, (size_t)B * T)) return 1;
  std::vector<int> rt((size_t)B * T);
  for (int s = 0; s < T; ++s)
    for (int e = 0; e < B; ++e) rt[(size_t)s * B + e] = e * T + s;
  ARL_CHECK(c, cudaMemcpy(c->rows_tab, rt.data(), rt.size() * sizeof(int), cudaMemcpyHostToDevice));
  // bf16 space-to-depth mirrors of the step buffer and of the rollout observations (what conv layer 0 reads)
  // (none when the first conv layer reads the uint8 stacks itself: u8_conv0_ok)
  const bool mirrors = !(c->pc_mode >= 2 && u8_conv0_ok(c) && cfg->frame_mode == 0);
  c->step_obs16 = nullptr;
  if (mirrors && dev_alloc(c, &c->step_obs16, (size_t)B * c->obs16_elems + 64 * 1024)) return 1;
  // a sampler without an observations buffer (evaluation: nothing is stored, sampler_with_eval.py:36-38) has no mirror
  c->roll_obs16 = nullptr;
  if (mirrors && cfg->observations && dev_alloc(c, &c->roll_obs16, (size_t)B * T * c->obs16_elems + 64 * 1024)) return 1;
  c->sampler_set = true;
  if (c->rollout_graph) { cudaGraphExecDestroy(c->rollout_graph); c->rollout_graph = nullptr; }
  return 0;
}

int arl_sampler_select(arl_ctx* c, int slot) {
  if (slot != 0 && slot != 1) ARL_FAIL(c, "sampler slot must be 0 (training) or 1 (evaluation)");
  if (slot == c->cur_slot) return 0;
  arl_ctx::SamplerSlot& o = c->slots[c->cur_slot];
  o.sc = c->sc; o.set = c->sampler_set; o.est = c->est; o.tout = c->tout; o.cmd = c->cmd; o.rows_tab = c->rows_tab;
  o.step_obs16 = c->step_obs16; o.roll_obs16 = c->roll_obs16; o.rollout_graph = c->rollout_graph;
  o.graph_rollout_nodes = c->graph_rollout_nodes;
  const arl_ctx::SamplerSlot& n = c->slots[slot];
  c->sc = n.sc; c->sampler_set = n.set; c->est = n.est; c->tout = n.tout; c->cmd = n.cmd; c->rows_tab = n.rows_tab;
  c->step_obs16 = n.step_obs16; c->roll_obs16 = n.roll_obs16; c->rollout_graph = n.rollout_graph;
  c->graph_rollout_nodes = n.graph_rollout_nodes;
  c->cur_slot = slot;
  return 0;
}

int arl_sampler_reset(arl_ctx* c, void* stream) {
  if (!c->sampler_set) ARL_FAIL(c, "sampler not configured");
  cudaStream_t st = (cudaStream_t)stream;
  const arl_sampler_cfg& s = c->sc;
  env_init_kernel<<<(s.n_envs + 127) / 128, 128, 0, st>>>(synth_cfg(s), c->est, c->cmd, s.n_envs);
  c->launches++;
  ARL_CHECK(c, cudaGetLastError());
  return launch_frame(c, nullptr, 0, false, st);
}

/* start_envs decorrelation (sampler/util.py:26-57): after arl_sampler_reset, env e takes n_steps[e] warm-up steps (device
   array [n_envs]; the synthetic emulator ignores actions, so the reference's random actions need no counterpart), is reset
   whenever its trajectory ends, and starts the first rollout from the observation it reached.  max_steps >= max n_steps. */
int arl_sampler_warmup(arl_ctx* c, const int* n_steps, int max_steps, void* stream) {
  if (!c->sampler_set) ARL_FAIL(c, "sampler not configured");
  if (c->sc.ext_emulator) ARL_FAIL(c, "external emulators decorrelate in their worker processes");
  cudaStream_t st = (cudaStream_t)stream;
  const arl_sampler_cfg& s = c->sc;
  for (int k = 0; k < max_steps; ++k) {
    EnvStepArgs es{synth_cfg(s), c->est, c->tout, c->cmd, s.rewards, s.dones, s.raw_reward, s.need_reset, s.n_envs, s.horizon, 0,
                   s.max_path_length, s.discount, /*mid_batch_reset=*/1, s.clip_reward, s.episodic_lives, n_steps, k};
    env_step_kernel<<<(s.n_envs + 127) / 128, 128, 0, st>>>(es);
    c->launches++;
    ARL_CHECK(c, cudaGetLastError());
    if (launch_frame(c, nullptr, 0, false, st)) return 1;
  }
  return 0;
}

int arl_rollout_begin(arl_ctx* c, void* stream) {
  if (!c->sampler_set) ARL_FAIL(c, "sampler not configured");
  l2_carve(c, false);
  return rollout_begin(c, (cudaStream_t)stream);
}
int arl_rollout_step(arl_ctx* c, int s, const uint8_t* staging, void* stream) {
  if (!c->sampler_set) ARL_FAIL(c, "sampler not configured");
  return rollout_step(c, s, staging, (cudaStream_t)stream);
}
int arl_rollout_end(arl_ctx* c, void* stream) {
  if (!c->sampler_set) ARL_FAIL(c, "sampler not configured");
  return rollout_end(c, (cudaStream_t)stream);
}

int arl_rollout_serve(arl_ctx* c, int s, int e0, int n, void* stream) {
  if (!c->sampler_set) ARL_FAIL(c, "sampler not configured");
  const arl_sampler_cfg& sc = c->sc;
  if (s < 0 || s >= sc.horizon) ARL_FAIL(c, "serve step out of range");
  if (n < 0) n = sc.n_envs - e0;
  if (e0 < 0 || n < 1 || e0 + n > sc.n_envs) ARL_FAIL(c, "serve env range out of bounds");
  const bool u8 = c->pc_mode >= 2 && u8_conv0_ok(c);
  const long obs_bytes = (long)sc.planes * c->cfg.in_h * c->cfg.in_w;
  const int* rows = c->rows_tab + (long)s * sc.n_envs + e0;
  const bool lean = frame_lean(c) && !u8;
  return policy_forward16(c, u8 ? nullptr : lean ? c->roll_obs16 : c->step_obs16 + (long)e0 * c->obs16_elems, n, rows,
                          sc.prob, sc.value, sc.uniforms + (long)s * sc.n_envs + e0, sc.actions, c->pc_mode >= 2,
                          (cudaStream_t)stream, nullptr, u8 ? sc.step_obs + (long)e0 * obs_bytes : nullptr,
                          lean ? rows : nullptr);
}

int arl_rollout_ingest(arl_ctx* c, int s, int e0, int n, const uint8_t* staging, const arl_ext_step* ext, void* stream) {
  if (!c->sampler_set) ARL_FAIL(c, "sampler not configured");
  if (!c->sc.ext_emulator) ARL_FAIL(c, "sampler was not configured for external emulators");
  if (!staging || !ext) ARL_FAIL(c, "ingest needs the staged frame pairs and the step records");
  static_assert(sizeof(ExtStep) == sizeof(arl_ext_step) && sizeof(ExtStep) == 12, "ext step record layout");
  cudaStream_t st = (cudaStream_t)stream;
  const arl_sampler_cfg& sc = c->sc;
  const long T = sc.horizon;
  if (s < -1 || s > T) ARL_FAIL(c, "ingest step out of range");
  if (n < 0) n = sc.n_envs - e0;
  if (e0 < 0 || n < 1 || e0 + n > sc.n_envs) ARL_FAIL(c, "ingest env range out of bounds");
  ext_apply_kernel<<<(n + 127) / 128, 128, 0, st>>>(reinterpret_cast<const ExtStep*>(ext) + e0, c->cmd + e0, sc.rewards + e0 * T,
                                                   sc.dones + e0 * T, sc.raw_reward + e0 * T, sc.need_reset + e0 * T, n, (int)T, s,
                                                   sc.clip_reward, sc.episodic_lives);
  c->launches++;
  ARL_CHECK(c, cudaGetLastError());
  // s in [0, T): the step's new observation goes to the step buffer and to rollout row s + 1;
  // s == -1 (start_envs) and s == T (reset_needed_envs after the batch): step buffer only
  const bool in_batch = s >= 0 && s < T;
  return launch_frame(c, staging, in_batch ? s + 1 : 0, in_batch && s + 1 < T, st, e0, n);
}

int arl_host_register(void* ptr, size_t bytes) {
  return cudaHostRegister(ptr, bytes, cudaHostRegisterPortable) == cudaSuccess ? 0 : 1;
}
int arl_host_unregister(void* ptr) { return cudaHostUnregister(ptr) == cudaSuccess ? 0 : 1; }
int arl_copy_async(arl_ctx* c, void* dst, const void* src, size_t bytes, int to_device, void* stream) {
  ARL_CHECK(c, cudaMemcpyAsync(dst, src, bytes, to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return 0;
}

int arl_rollout_run(arl_ctx* c, void* stream) {
  NvtxRange nvtx_("rollout");
  if (!c->sampler_set) ARL_FAIL(c, "sampler not configured");
  cudaStream_t st = (cudaStream_t)stream;
  l2_carve(c, false);                          // (the previous optimize() call has been read back: nothing is training)
  if (!c->rollout_graph) {
    // warm every kernel's lazy attribute setup outside capture
    cudaStream_t cap;
    ARL_CHECK(c, cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
    long l0 = c->launches;
    cudaGraph_t g = nullptr;
    ARL_CHECK(c, cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal));
    int rc = rollout_begin(c, cap);
    for (int s = 0; s < c->sc.horizon && !rc; ++s) rc = rollout_step(c, s, nullptr, cap);
    if (!rc) rc = rollout_end(c, cap);
    cudaError_t ce = cudaStreamEndCapture(cap, &g);
    c->graph_rollout_nodes = c->launches - l0;
    c->launches = l0;
    if (rc) { if (g) cudaGraphDestroy(g); cudaStreamDestroy(cap); return rc; }
    ARL_CHECK(c, ce);
    ARL_CHECK(c, cudaGraphInstantiate(&c->rollout_graph, g, 0));
    cudaGraphDestroy(g);
    cudaStreamDestroy(cap);
  }
  ARL_CHECK(c, cudaGraphLaunch(c->rollout_graph, st));
  c->launches += c->graph_rollout_nodes;
  return 0;
}

int arl_traj_read(arl_ctx* c, int* n, int* env, int* len, float* ret, float* raw, int* nz, float* disc, int cap,
                  void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  int cnt = 0;
  ARL_CHECK(c, cudaMemcpyAsync(&cnt, c->tout.count, sizeof(int), cudaMemcpyDeviceToHost, st));
  ARL_CHECK(c, cudaStreamSynchronize(st));
  cnt = std::min(cnt, std::min(cap, c->tout.cap));
  *n = cnt;
  if (cnt > 0) {
    ARL_CHECK(c, cudaMemcpyAsync(env, c->tout.env, cnt * sizeof(int), cudaMemcpyDeviceToHost, st));
    ARL_CHECK(c, cudaMemcpyAsync(len, c->tout.len, cnt * sizeof(int), cudaMemcpyDeviceToHost, st));
    ARL_CHECK(c, cudaMemcpyAsync(nz, c->tout.nz, cnt * sizeof(int), cudaMemcpyDeviceToHost, st));
    ARL_CHECK(c, cudaMemcpyAsync(ret, c->tout.ret, cnt * sizeof(float), cudaMemcpyDeviceToHost, st));
    ARL_CHECK(c, cudaMemcpyAsync(raw, c->tout.raw, cnt * sizeof(float), cudaMemcpyDeviceToHost, st));
    ARL_CHECK(c, cudaMemcpyAsync(disc, c->tout.disc, cnt * sizeof(float), cudaMemcpyDeviceToHost, st));
    ARL_CHECK(c, cudaStreamSynchronize(st));
  }
  return 0;
}

int arl_peek_frame_cmds(arl_ctx* c, int* cmd_host, int n_envs, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  ARL_CHECK(c, cudaMemcpyAsync(cmd_host, c->cmd, (size_t)n_envs * sizeof(FrameCmd), cudaMemcpyDeviceToHost, st));
  ARL_CHECK(c, cudaStreamSynchronize(st));
  return 0;
}

int arl_gae(arl_ctx* c, const float* rewards, float* values, const uint8_t* dones, const uint8_t* need_reset,
            const float* last_values, float discount, float gae_lambda, float* adv, float* ret, int8_t* valids,
            int n_envs, int horizon, int standardize, void* stream) {
  NvtxRange nvtx_("gae");
  cudaStream_t st = (cudaStream_t)stream;
  int use_gae = (gae_lambda != 1.0f) ? 1 : 0;
  gae_kernel<<<(n_envs + 3) / 4, 128, 0, st>>>(rewards, values, dones, need_reset, last_values, discount, gae_lambda,
                                                use_gae, adv, ret, valids, n_envs, horizon);
  c->launches++;
  long n = (long)n_envs * horizon;
  if (standardize) {
    standardize_adv_kernel<<<1, 1024, 0, st>>>(adv, valids, n);
    c->launches++;
  }
  ARL_CHECK(c, cudaGetLastError());
  return 0;
}

int arl_opt_configure(arl_ctx* c, const arl_opt_cfg* cfg) {
  c->opt = *cfg;
  c->opt_set = true;
  if (c->train_graph) { cudaGraphExecDestroy(c->train_graph); c->train_graph = nullptr; }
  return 0;
}

int arl_bind_train_inputs(arl_ctx* c, const uint8_t* obs, const uint8_t* actions, const float* adv, const float* ret,
                          const float* old_value, const float* old_prob, const int8_t* valids, long n_rows) {
  bool same = (c->t_obs == obs && c->t_act == actions && c->t_adv == adv && c->t_ret == ret && c->t_oldp == old_prob &&
               c->t_valids == valids);
  c->t_obs = obs; c->t_act = actions; c->t_adv = adv; c->t_ret = ret; c->t_oldv = old_value; c->t_oldp = old_prob;
  c->t_valids = valids; c->t_rows = n_rows;
  if (!same && c->train_graph) { cudaGraphExecDestroy(c->train_graph); c->train_graph = nullptr; }
  return 0;
}

int arl_set_lr_mult(arl_ctx* c, float lr_mult, void* stream) {
  c->lr_mult_host = lr_mult;
  ARL_CHECK(c, cudaMemcpyAsync(c->hyper, &c->lr_mult_host, sizeof(float), cudaMemcpyHostToDevice, (cudaStream_t)stream));
  ARL_CHECK(c, cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}

int arl_grad_minibatch(arl_ctx* c, const int* idx, int mb_size, void* stream) {
  return grad_minibatch(c, idx, nullptr, mb_size, (cudaStream_t)stream);
}

int arl_clip_update(arl_ctx* c, float gscale, void* stream) { return clip_update(c, gscale, (cudaStream_t)stream); }

int arl_train_minibatches(arl_ctx* c, const int* idx, int mb_size, int count, void* stream) {
  return train_minibatches(c, idx, mb_size, count, 0, (cudaStream_t)stream);
}
/* same loop with the synchronous data-parallel step (fused P2P all-reduce + clip + update) closing every minibatch */
int arl_train_minibatches_sync(arl_ctx* c, const int* idx, int mb_size, int count, void* stream) {
  if (!c->comm.ready) ARL_FAIL(c, "comm not connected");
  return train_minibatches(c, idx, mb_size, count, 1, (cudaStream_t)stream);
}
/* ... and for the asynchronous learner: every minibatch ends with arl_async_push_pull */
int arl_train_minibatches_async(arl_ctx* c, const int* idx, int mb_size, int count, void* stream) {
  if (!c->async_.ready) ARL_FAIL(c, "async store not connected");
  return train_minibatches(c, idx, mb_size, count, 2, (cudaStream_t)stream);
}

}  // extern "C"
namespace {
int train_minibatches(arl_ctx* c, const int* idx, int mb_size, int count, int sync, cudaStream_t st) {
  NvtxRange nvtx_("train minibatches");
  if (count > c->log_cap) ARL_FAIL(c, "more minibatches in one call than loss / grad-norm log slots (4096): read the logs in between");
  // (local update only: the cross-GPU learners' updates are other kernels; a single full-batch step per rollout, A2C, has
  // nothing to keep resident)
  if (sync == 0 && count >= 8) l2_carve(c, true);
  // sync: 0 = local clip + update, 1 = synchronous DP step, 2 = asynchronous push/pull
  const bool overlap = sync == 1 && sync_overlap_ok(c);
  if (overlap && sync_overlap_prepare(c)) return 1;
  auto step = [&](cudaStream_t s_) {
    return sync == 1 ? (overlap ? sync_tail(c, s_) : sync_update(c, s_)) : sync == 2 ? async_push_pull(c, s_) : clip_update(c, 1.f, s_);
  };
  struct Active {      // grad_minibatch may update the FC weights early only when the local clip_update follows it
    arl_ctx* c;
    Active(arl_ctx* c_, bool on, bool ov) : c(c_) { c->train_step_active = on; c->sync_overlap_active = ov; }
    ~Active() {
      c->sync_overlap_active = false;
      c->train_step_active = false; c->early_fc_done = false; c->pending_fin = nullptr; c->pending_stream = false;
      c->pending_ss_fin = c->pending_ss_fc = 0;
    }
  } active(c, sync == 0, overlap);
  const bool graphable = (c->pc_mode >= 2 && u8_conv0_ok(c) && c->t_obs) ||
                         (c->sampler_set && c->t_obs == c->sc.observations && c->roll_obs16);
  if (!graphable || (sync && c->sync_graph_failed)) {
    // training inputs that are not the sampler's rollout buffers: plain launches, one minibatch at a time
    const bool replay_idx = c->sampler_set && c->t_obs == c->sc.observations && c->roll_obs16;
    (void)replay_idx;
    for (int i = 0; i < count; ++i) {
      if (grad_minibatch(c, idx + (long)i * mb_size, nullptr, mb_size, st)) return 1;
      if (step(st)) return 1;
    }
    return 0;
  }
  // ONE graph holds `per` consecutive minibatches (all of them when count <= ARL_GRAPH_MB, default 512): the device-side
  // minibatch counter selects each one's index slice, so an iteration's 4 epochs x N/mb updates are one cudaGraphLaunch
  // instead of 256 (ARL_GRAPH_MB=1: one graph per minibatch, the round-1 behaviour)
  static const int graph_mb_cap = getenv("ARL_GRAPH_MB") ? std::max(1, atoi(getenv("ARL_GRAPH_MB"))) : 512;
  int per = std::min(count, graph_mb_cap);
  if (per < 1 || count % per != 0) per = 1;
  if (c->train_graph && (c->train_graph_idx != idx || c->train_graph_mb != mb_size || c->train_graph_sync != sync ||
                         c->train_graph_per != per)) {
    cudaGraphExecDestroy(c->train_graph);
    c->train_graph = nullptr;
  }
  if (!c->train_graph) {
    TrainPlan* P = nullptr;
    if (get_plan(c, mb_size, &P)) return 1;   // allocations happen outside capture
    cudaStream_t cap;
    ARL_CHECK(c, create_main_stream(&cap));
    long l0 = c->launches;
    cudaGraph_t g = nullptr;
    // ARL_SPLIT_UPDATE=1 (off by default): the FC range's update on the side stream, beside the next minibatch's conv
    // layers.  Bit-identical results (tests/test_gpu_path.py), measured SLOWER on B200: 43.96 vs 43.20 ms per 256
    // minibatches — the conv tiles lose more to the shared memory system than the hidden 20 us are worth
    static const bool split_on = getenv("ARL_SPLIT_UPDATE") && atoi(getenv("ARL_SPLIT_UPDATE")) != 0;
    if (split_on && sync == 0 && !c->ev_updB) {
      ARL_CHECK(c, cudaEventCreateWithFlags(&c->ev_upd_fork, cudaEventDisableTiming));
      ARL_CHECK(c, cudaEventCreateWithFlags(&c->ev_updB, cudaEventDisableTiming));
    }
    ARL_CHECK(c, cudaMemsetAsync(c->ticket + 2, 0, 2 * sizeof(unsigned long long), st));
    ARL_CHECK(c, cudaStreamSynchronize(st));
    ARL_CHECK(c, cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal));
    int rc = 0;
    c->split_capture = split_on && sync == 0;
    for (int k = 0; k < per && !rc; ++k) {
      rc = grad_minibatch(c, idx, c->mb_counter, mb_size, cap);
      if (!rc) rc = step(cap);
    }
    c->split_capture = false;
    if (c->updB_pending) {
      cudaStreamWaitEvent(cap, c->ev_updB, 0);
      c->updB_pending = false;
    }
    cudaError_t ce = cudaStreamEndCapture(cap, &g);
    c->graph_train_nodes = c->launches - l0;
    c->launches = l0;
    if (!rc && ce == cudaSuccess) ce = cudaGraphInstantiate(&c->train_graph, g, 0);
    if (g) cudaGraphDestroy(g);
    cudaStreamDestroy(cap);
    if (sync && (rc || ce != cudaSuccess)) {
      // the cooperative all-reduce kernel could not be captured on this driver: plain launches instead
      cudaGetLastError();
      fprintf(stderr, "[accel_rl_b200] synchronous minibatch could not be graph-captured (%s); launching kernels one by one\n",
              rc ? c->err.c_str() : cudaGetErrorString(ce));
      c->train_graph = nullptr;
      c->sync_graph_failed = true;
      return train_minibatches(c, idx, mb_size, count, sync, st);
    }
    if (rc) return rc;
    ARL_CHECK(c, ce);
    c->train_graph_idx = idx; c->train_graph_mb = mb_size; c->train_graph_sync = sync; c->train_graph_per = per;
  }
  ARL_CHECK(c, cudaMemsetAsync(c->mb_counter, 0, sizeof(int), st));
  for (int i = 0; i < count / per; ++i) ARL_CHECK(c, cudaGraphLaunch(c->train_graph, st));
  c->launches += (long)(count / per) * c->graph_train_nodes;
  return 0;
}
}  // namespace
extern "C" {

int arl_read_logs(arl_ctx* c, float* loss, float* grad_norm, int cap, int* n, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  int cnt = 0;
  ARL_CHECK(c, cudaMemcpyAsync(&cnt, c->log_slot, sizeof(int), cudaMemcpyDeviceToHost, st));
  ARL_CHECK(c, cudaStreamSynchronize(st));
  cnt = std::min(cnt, std::min(cap, c->log_cap));
  *n = cnt;
  if (cnt > 0) {
    ARL_CHECK(c, cudaMemcpyAsync(loss, c->log_loss, cnt * sizeof(float), cudaMemcpyDeviceToHost, st));
    ARL_CHECK(c, cudaMemcpyAsync(grad_norm, c->log_norm, cnt * sizeof(float), cudaMemcpyDeviceToHost, st));
  }
  ARL_CHECK(c, cudaMemsetAsync(c->log_slot, 0, sizeof(int), st));
  ARL_CHECK(c, cudaStreamSynchronize(st));
  return 0;
}

int arl_reset_opt_state(arl_ctx* c, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  ARL_CHECK(c, cudaMemsetAsync(c->step, 0, sizeof(int), st));
  ARL_CHECK(c, cudaMemsetAsync(c->log_slot, 0, sizeof(int), st));
  if (c->m) ARL_CHECK(c, cudaMemsetAsync(c->m, 0, c->n_params * sizeof(float), st));
  if (c->v) ARL_CHECK(c, cudaMemsetAsync(c->v, 0, c->n_params * sizeof(float), st));
  return 0;
}

int arl_opt_step_get(arl_ctx* c, int* t, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  ARL_CHECK(c, cudaMemcpyAsync(t, c->step, sizeof(int), cudaMemcpyDeviceToHost, st));
  ARL_CHECK(c, cudaStreamSynchronize(st));
  return 0;
}

int arl_opt_step_set(arl_ctx* c, int t, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (t < 0) ARL_FAIL(c, "optimizer step count must be >= 0");
  ARL_CHECK(c, cudaMemcpyAsync(c->step, &t, sizeof(int), cudaMemcpyHostToDevice, st));
  ARL_CHECK(c, cudaStreamSynchronize(st));
  return 0;
}

// ---- sync DP ------------------------------------------------------------------------------
int arl_comm_local_init(arl_ctx* c, int rank, int world, uint8_t* handle_out) {
  std::string err;
  // the bf16 FC operand copy moves into the symmetric allocation too: the owner of a parameter slice P2P-stores the
  // refreshed bf16 values next to the fp32 ones, so no rank re-packs the 3.5 M FC weights after a step
  const bool tiles = fc_tiles_ok(c);
  const size_t shadow_elems = (c->off_Wfc % 4 == 0 && c->H % 4 == 0) ? (size_t)c->Kfc * c->H : 0;
  if (comm_local_init(c->comm, rank, world, c->n_params, shadow_elems, handle_out, err)) { c->err = err; return 5; }
  if (shadow_elems) {
    __nv_bfloat16*& local = tiles ? c->wfc_t : c->wfc_bf16;
    ARL_CHECK(c, cudaMemcpy(c->comm.shadow, local, shadow_elems * sizeof(__nv_bfloat16), cudaMemcpyDeviceToDevice));
    cudaFree(local);
    local = c->comm.shadow;
    c->shadow_in_comm = true;
    // the pack job that fills this copy must follow it
    std::vector<PackJob> pj(c->n_pack_jobs);
    ARL_CHECK(c, cudaMemcpy(pj.data(), c->pack_jobs_dev, pj.size() * sizeof(PackJob), cudaMemcpyDeviceToHost));
    pj.back().dst = local;
    ARL_CHECK(c, cudaMemcpy(c->pack_jobs_dev, pj.data(), pj.size() * sizeof(PackJob), cudaMemcpyHostToDevice));
  }
  return 0;
}
int arl_comm_buffers(arl_ctx* c, float** grad_out, float** params_out) {
  if (!c->comm.base) ARL_FAIL(c, "comm not initialised");
  *grad_out = c->comm.grad; *params_out = c->comm.param;
  return 0;
}
int arl_comm_connect(arl_ctx* c, const uint8_t* all_handles) {
  std::string err;
  if (comm_connect(c->comm, all_handles, err)) { c->err = err; return 5; }
  return 0;
}
int arl_comm_barrier(arl_ctx* c, void* stream) {
  std::string err;
  if (comm_barrier(c->comm, (cudaStream_t)stream, err)) { c->err = err; return 5; }
  c->launches++;
  return 0;
}
int arl_sync_allreduce_update(arl_ctx* c, void* stream) { return sync_update(c, (cudaStream_t)stream); }

/* device timeline of the overlapped synchronous step since the last reset: out[0..5] = average microseconds per step of
   {FC exchange: wait for peers, FC exchange: reduce + update + publish, tail: wait for peers, tail: average + update,
   slack between the end of the FC exchange and the start of the tail (> 0: fully hidden)}, out[5] = steps */
int arl_comm_trace(arl_ctx* c, double* out, int reset, void* stream) {
  if (!c->comm.ready) ARL_FAIL(c, "comm not connected");
  ARL_CHECK(c, cudaStreamSynchronize((cudaStream_t)stream));
  unsigned long long h[16];
  ARL_CHECK(c, cudaMemcpy(h, c->comm.dev.trace, sizeof(h), cudaMemcpyDeviceToHost));
  const double n = h[TR_COUNT] ? (double)h[TR_COUNT] : 1.0;
  out[0] = h[TR_FC_WAIT] / n * 1e-3; out[1] = h[TR_FC_WORK] / n * 1e-3; out[2] = h[TR_TAIL_WAIT] / n * 1e-3;
  out[3] = h[TR_TAIL_WORK] / n * 1e-3; out[4] = (double)(long long)h[TR_SLACK] / n * 1e-3; out[5] = (double)h[TR_COUNT];
  if (reset) ARL_CHECK(c, cudaMemset(c->comm.dev.trace, 0, sizeof(h)));
  return 0;
}

}  // extern "C"
namespace {
int sync_update(arl_ctx* c, cudaStream_t st) {
  NvtxRange nvtx_("allreduce+adam");
  if (!c->opt_set) ARL_FAIL(c, "optimizer not configured");
  SyncUpdateArgs a{};
  a.param = c->params; a.grad = c->grad; a.m = c->m; a.v = c->v; a.n = c->n_params;
  a.loss_partial = c->loss_partial; a.n_loss_blocks = c->n_loss_rows; a.hyper = c->hyper; a.step = c->step;
  a.kind = c->opt.update; a.lr = c->opt.learning_rate; a.beta1 = c->opt.beta1; a.beta2 = c->opt.beta2;
  a.eps = c->opt.epsilon; a.rho = c->opt.rho; a.clip = c->opt.grad_norm_clip;
  a.out_norm = c->log_norm; a.out_loss = c->log_loss; a.log_slot = c->log_slot; a.log_cap = c->log_cap;
  a.shadow_begin = c->off_Wfc; a.shadow_end = c->off_Wfc + (long)c->Kfc * c->H;
  if (fc_tiles_ok(c)) { a.shadow_tiles = 1; a.shadow_HW = c->HWlast; a.shadow_H = c->H; }
  std::string err;
  if (comm_sync_update(c->comm, a, st, err)) { c->err = err; return 5; }
  c->launches += 1;
  ARL_CHECK(c, cudaGetLastError());
  return pack_weights(c, st, !c->shadow_in_comm, true);
}

// The overlapped synchronous step (comm.cuh) serves the same configuration as update_stream_kernel on one GPU: no
// global-norm clipping, FC weight gradient written straight into the flat vector, bf16 FC operand tiles inside the
// symmetric allocation, conv operand packs refreshed through the slot table.  ARL_SYNC_OVERLAP=0 keeps the monolithic kernel.
bool sync_overlap_ok(arl_ctx* c) {
  static const bool on = !(getenv("ARL_SYNC_OVERLAP") && atoi(getenv("ARL_SYNC_OVERLAP")) == 0);
  return on && c->comm.ready && c->shadow_in_comm && c->opt_set && c->opt.grad_norm_clip <= 0.f && c->m && c->v &&
         c->pc_mode >= 2 && fc_tiles_ok(c) && (c->off_Wfc % 4 == 0) && (((long)c->Kfc * c->H) % 4 == 0) &&
         c->n_pack_jobs >= 2 && c->conv_pack_end > 0 && c->pk_slots && c->params == c->comm.param && c->grad == c->comm.grad;
}

// stream, events and scratch of the overlapped step: created OUTSIDE stream capture (train_minibatches calls this first)
int sync_overlap_prepare(arl_ctx* c) {
  if (!c->cs) {
    ARL_CHECK(c, cudaStreamCreateWithFlags(&c->cs, cudaStreamNonBlocking));
    for (auto& e : c->ev_cs_in) ARL_CHECK(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ARL_CHECK(c, cudaEventCreateWithFlags(&c->ev_cs_done, cudaEventDisableTiming));
  }
  if (!c->sync_tail_partial) {
    const long n_small = c->n_params - (long)c->Kfc * c->H;
    if (dev_alloc(c, &c->sync_tail_partial, (size_t)((n_small + 255) / 256))) return 1;
  }
  return 0;
}

int sync_fc_exchange(arl_ctx* c, cudaStream_t ws, cudaStream_t st) {
  NvtxRange nvtx_("allreduce+adam: FC slice exchange");
  (void)ws;                                              // ev_cs_in[0] was recorded on it right after the FC weight gradient
  ARL_CHECK(c, cudaEventRecord(c->ev_cs_in[1], st));
  ARL_CHECK(c, cudaStreamWaitEvent(c->cs, c->ev_cs_in[0], 0));
  ARL_CHECK(c, cudaStreamWaitEvent(c->cs, c->ev_cs_in[1], 0));
  const CommDev& d = c->comm.dev;
  SyncFcArgs a{};
  a.param = c->params; a.m = c->m; a.v = c->v;
  a.fc_begin = c->off_Wfc; a.fc_len = (long)c->Kfc * c->H;
  a.per = ((a.fc_len + d.world - 1) / d.world + 3) / 4 * 4;
  a.shadow_tiles = 1; a.shadow_HW = c->HWlast; a.shadow_H = c->H;
  a.hyper = c->hyper; a.step = c->step; a.kind = c->opt.update;
  a.lr = c->opt.learning_rate; a.beta1 = c->opt.beta1; a.beta2 = c->opt.beta2; a.eps = c->opt.epsilon; a.rho = c->opt.rho;
  if (d.world <= 2) ARL_CHECK(c, launch_k(sync_fc_kernel<2>, dim3(kSyncFcBlocks), dim3(kSyncFcThreads), 0, c->cs, d, a));
  else if (d.world <= 4) ARL_CHECK(c, launch_k(sync_fc_kernel<4>, dim3(kSyncFcBlocks), dim3(kSyncFcThreads), 0, c->cs, d, a));
  else ARL_CHECK(c, launch_k(sync_fc_kernel<kMaxRanks>, dim3(kSyncFcBlocks), dim3(kSyncFcThreads), 0, c->cs, d, a));
  c->launches += 1;
  prof_mark(c, "sync_fc", c->cs);
  ARL_CHECK(c, cudaGetLastError());
  ARL_CHECK(c, cudaEventRecord(c->ev_cs_done, c->cs));
  return 0;
}

int sync_tail(arl_ctx* c, cudaStream_t st) {
  NvtxRange nvtx_("allreduce+adam: tail");
  const CommDev& d = c->comm.dev;
  const long n_small = c->n_params - (long)c->Kfc * c->H;
  const int blocks = (int)((n_small + 255) / 256);
  SyncTailArgs a{};
  a.param = c->params; a.m = c->m; a.v = c->v; a.n = c->n_params;
  a.fc_begin = c->off_Wfc; a.fc_len = (long)c->Kfc * c->H;
  a.loss_partial = c->loss_partial; a.n_loss_blocks = c->n_loss_rows; a.hyper = c->hyper; a.step = c->step;
  a.kind = c->opt.update; a.lr = c->opt.learning_rate; a.beta1 = c->opt.beta1; a.beta2 = c->opt.beta2;
  a.eps = c->opt.epsilon; a.rho = c->opt.rho;
  a.out_norm = c->log_norm; a.out_loss = c->log_loss; a.log_slot = c->log_slot; a.log_cap = c->log_cap; a.mb_counter = c->mb_counter;
  a.pk_slots = c->pk_slots; a.conv_end = c->conv_pack_end; a.partial = c->sync_tail_partial;
  if (d.world <= 2) ARL_CHECK(c, launch_k(sync_tail_kernel<2>, dim3(blocks), dim3(256), 0, st, d, a));
  else if (d.world <= 4) ARL_CHECK(c, launch_k(sync_tail_kernel<4>, dim3(blocks), dim3(256), 0, st, d, a));
  else ARL_CHECK(c, launch_k(sync_tail_kernel<kMaxRanks>, dim3(blocks), dim3(256), 0, st, d, a));
  c->launches += 1;
  prof_mark(c, "sync_tail", st);
  ARL_CHECK(c, cudaGetLastError());
  return 0;
}
}  // namespace
extern "C" {

// ---- async DP -----------------------------------------------------------------------------
int arl_async_local_init(arl_ctx* c, int rank, int world, int n_update_chunks, uint8_t* handle_out) {
  if (!c->params) ARL_FAIL(c, "parameters not bound");
  ARL_CHECK(c, cudaDeviceSynchronize());
  if (async_local_init(c->async_, rank, world, c->n_params, n_update_chunks, c->params, handle_out, c->err)) return 1;
  return 0;
}

int arl_async_connect(arl_ctx* c, const uint8_t* rank0_handle) {
  if (async_connect(c->async_, c->n_params, rank0_handle, c->err)) return 1;
  return 0;
}

int arl_async_regions(arl_ctx* c) { return c->async_.dev.n_locks; }

/* local clip -> chunk-locked update of the central (p, m, v) with the local gradient -> pull the new p */
int arl_async_push_pull(arl_ctx* c, void* stream) { return async_push_pull(c, (cudaStream_t)stream); }

}  // extern "C"
namespace {
int async_push_pull(arl_ctx* c, cudaStream_t st) {
  if (!c->async_.ready) ARL_FAIL(c, "async store not connected");
  if (!c->opt_set) ARL_FAIL(c, "optimizer not configured");
  sumsq_kernel<<<kSumsqBlocks, 256, 0, st>>>(c->grad, c->n_params, 1.f, c->sumsq_partial);
  c->launches++;
  UpdateParams u{};
  u.param = c->params; u.grad = c->grad; u.m = nullptr; u.v = nullptr; u.n = c->n_params;
  u.sumsq_partial = c->sumsq_partial; u.n_partial = kSumsqBlocks;
  u.loss_partial = c->loss_partial; u.n_loss_blocks = c->n_loss_rows;
  u.hyper = c->hyper; u.step = c->step; u.kind = c->opt.update;
  u.lr = c->opt.learning_rate; u.beta1 = c->opt.beta1; u.beta2 = c->opt.beta2; u.eps = c->opt.epsilon;
  u.rho = c->opt.rho; u.clip = c->opt.grad_norm_clip; u.gscale = 1.f;
  u.out_norm = c->log_norm; u.out_loss = c->log_loss; u.log_slot = c->log_slot; u.log_cap = c->log_cap;
  bool fused_cast = (c->off_Wfc % 4 == 0) && (c->async_.dev.per % 4 == 0) && (c->H % 4 == 0);
  u.shadow = fused_cast ? c->wfc_bf16 : nullptr;
  if (fused_cast && fc_tiles_ok(c)) { u.shadow = c->wfc_t; u.shadow_tiles = 1; u.shadow_HW = c->HWlast; u.shadow_H = c->H; }
  u.shadow_begin = c->off_Wfc; u.shadow_end = c->off_Wfc + (long)c->Kfc * c->H;
  int grid = std::min(c->async_.dev.n_locks, 148);
  async_push_pull_kernel<<<grid, kAsyncThreads, 0, st>>>(c->async_.dev, u);
  c->launches++;
  ARL_CHECK(c, cudaGetLastError());
  return pack_weights(c, st, !fused_cast, true);
}
}  // namespace
extern "C" {

/* pull only: central parameters -> local parameters + operand copies (ActsrvAltOvrlpPollSampler, poll_sampler.py:29-39) */
int arl_async_pull(arl_ctx* c, void* stream) {
  if (!c->async_.ready) ARL_FAIL(c, "async store not connected");
  cudaStream_t st = (cudaStream_t)stream;
  UpdateParams u{};
  u.param = c->params; u.n = c->n_params;
  bool fused_cast = (c->off_Wfc % 4 == 0) && (c->async_.dev.per % 4 == 0) && (c->H % 4 == 0);
  u.shadow = fused_cast ? c->wfc_bf16 : nullptr;
  if (fused_cast && fc_tiles_ok(c)) { u.shadow = c->wfc_t; u.shadow_tiles = 1; u.shadow_HW = c->HWlast; u.shadow_H = c->H; }
  u.shadow_begin = c->off_Wfc; u.shadow_end = c->off_Wfc + (long)c->Kfc * c->H;
  int grid = std::min(c->async_.dev.n_locks, 148);
  async_pull_kernel<<<grid, kAsyncThreads, 0, st>>>(c->async_.dev, u);
  c->launches++;
  ARL_CHECK(c, cudaGetLastError());
  return pack_weights(c, st, !fused_cast, false);
}

/* test hook: copy central array `which` (0 = p, 1 = m, 2 = v / accumulator) to the host */
int arl_async_read_central(arl_ctx* c, int which, float* host_out, long n, void* stream) {
  if (!c->async_.ready) ARL_FAIL(c, "async store not connected");
  ARL_CHECK(c, cudaStreamSynchronize((cudaStream_t)stream));
  const float* src = which == 0 ? c->async_.dev.cp : which == 1 ? c->async_.dev.cm : c->async_.dev.cv;
  ARL_CHECK(c, cudaMemcpy(host_out, src, (size_t)std::min(n, c->n_params) * sizeof(float), cudaMemcpyDeviceToHost));
  return 0;
}

// ---- diagnostics --------------------------------------------------------------------------
__global__ void bf16_to_f32_kernel(const __nv_bfloat16* x, float* y, long n) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = __bfloat162float(x[i]);
}

int arl_debug_activation(arl_ctx* c, int layer, float* out, long cap, long* n, void* stream) {
  // layer: 0..n_conv-1 conv activations of the last forward ([rows][Cout], NHWC), n_conv..2n_conv-1 their
  // gradients, 100 = h, 101 = dh (both need a training pass)
  cudaStream_t st = (cudaStream_t)stream;
  const __nv_bfloat16* src = nullptr;
  long cnt = 0;
  int nc = (int)c->conv.size();
  long R = c->cfg.max_rows;
  if (layer >= 0 && layer < nc) { src = c->conv[layer].act; cnt = R * c->conv[layer].Ho * c->conv[layer].Wo * c->conv[layer].Cout; }
  else if (layer >= nc && layer < 2 * nc) { auto& L = c->conv[layer - nc]; src = L.dact; cnt = R * L.Ho * L.Wo * L.Cout; }
  else if (layer == 100) { src = c->h; cnt = R * c->H; }
  else if (layer == 101) { src = c->dh; cnt = R * c->H; }
  else if (layer >= 200 && layer < 200 + (int)c->pc.size() && layer > 200) {   // raw input grid of pconv layer (plane 0)
    src = c->pc[layer - 200].in; cnt = c->pc[layer - 200].in_rows * 64;
  }
  else ARL_FAIL(c, "bad layer id");
  cnt = std::min(cnt, cap);
  *n = cnt;
  bf16_to_f32_kernel<<<(int)((cnt + 255) / 256), 256, 0, st>>>(src, out, cnt);
  ARL_CHECK(c, cudaGetLastError());
  return 0;
}

long arl_kernel_launches(arl_ctx* c) { return c->launches; }

#ifndef ARL_SRC_HASH
#define ARL_SRC_HASH "unknown"
#endif
/* sha256 of the CUDA sources this binary was built from (csrc/Makefile: api.cu + the headers, in Makefile order) */
const char* arl_source_hash(void) { return ARL_SRC_HASH; }

#ifdef ARL_TRACE
extern "C" int arl_trace_read(long long* out, int n) {
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(out, g_trace, sizeof(long long) * n);
}
#endif

// Per-kernel device times of one training minibatch (kind 0) or one rollout step (kind 1).  The sequence is
// captured into a CUDA graph exactly as the product path does; the whole graph is replayed a few times (so every
// kernel's inputs exist and are L2-warm), then every kernel node is re-launched `reps` times back to back with its
// captured arguments and timed with one event pair (GPU-bound: launch gaps are hidden behind the previous launch).
int arl_profile_graph(arl_ctx* c, int kind, const int* idx, int mb_size, int reps, char* names, int names_cap,
                      float* ms, int cap, int* n, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (kind == 0) { TrainPlan* P = nullptr; if (get_plan(c, mb_size, &P)) return 1; }
  cudaStream_t cap_s;
  ARL_CHECK(c, cudaStreamCreateWithFlags(&cap_s, cudaStreamNonBlocking));
  c->prof_labels.clear();
  c->prof_collect = true;
  long l0 = c->launches;
  cudaGraph_t g = nullptr;
  // plain (fully serialised) edges in this scratch graph: the node-introspection calls below cannot represent
  // programmatic-dependent-launch edges
  const bool pdl_saved = g_pdl;
  g_pdl = false;
  c->no_fork = true;
  ARL_CHECK(c, cudaStreamBeginCapture(cap_s, cudaStreamCaptureModeThreadLocal));
  int rc = 0;
  if (kind == 0) {
    c->train_step_active = true;
    rc = grad_minibatch(c, idx, c->mb_counter, mb_size, cap_s);
    if (!rc) rc = clip_update(c, 1.f, cap_s);
    c->train_step_active = false;
    c->early_fc_done = false;
    c->pending_fin = nullptr; c->pending_stream = false;
    c->pending_ss_fin = c->pending_ss_fc = 0;
  } else if (kind == 1) {
    rc = rollout_step(c, 0, nullptr, cap_s);
  } else {
    // kind 2: inference forward of the first mb_size rows of the staging buffer (filled by the last uint8 forward)
    rc = policy_forward16(c, c->obs16_stage, mb_size, nullptr, c->dlogit, c->dlogit + (long)mb_size * c->A, nullptr, nullptr,
                          c->pc_mode >= 1, cap_s);
  }
  cudaError_t ce = cudaStreamEndCapture(cap_s, &g);
  g_pdl = pdl_saved;
  c->no_fork = false;
  c->prof_collect = false;
  c->launches = l0;
  if (rc) { if (g) cudaGraphDestroy(g); cudaStreamDestroy(cap_s); return rc; }
  ARL_CHECK(c, ce);
  cudaGraphExec_t ge = nullptr;
  ARL_CHECK(c, cudaGraphInstantiate(&ge, g, 0));
  ARL_CHECK(c, cudaMemsetAsync(c->mb_counter, 0, sizeof(int), st));
  for (int i = 0; i < 3; ++i) ARL_CHECK(c, cudaGraphLaunch(ge, st));
  ARL_CHECK(c, cudaMemsetAsync(c->mb_counter, 0, sizeof(int), st));
  // walk the (linear) dependency chain
  size_t nroot = 1;
  cudaGraphNode_t node = nullptr;
  ARL_CHECK(c, cudaGraphGetRootNodes(g, &node, &nroot));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  std::string all;
  int cnt = 0;
  size_t li = 0;
  while (node && cnt < cap) {
    cudaGraphNodeType ty;
    ARL_CHECK(c, cudaGraphNodeGetType(node, &ty));
    if (ty == cudaGraphNodeTypeKernel) {
      cudaKernelNodeParams kp{};
      ARL_CHECK(c, cudaGraphKernelNodeGetParams(node, &kp));
      // `reps` chained copies of this kernel node in a scratch graph: device-side launch latency only (what the
      // product path pays inside its captured graphs), no host launch-rate floor
      cudaGraph_t tg = nullptr;
      ARL_CHECK(c, cudaGraphCreate(&tg, 0));
      cudaGraphNode_t prev = nullptr;
      for (int r = 0; r < reps; ++r) {
        cudaGraphNode_t nd = nullptr;
        ARL_CHECK(c, cudaGraphAddKernelNode(&nd, tg, prev ? &prev : nullptr, prev ? 1 : 0, &kp));
        ARL_CHECK(c, cudaGraphKernelNodeCopyAttributes(nd, node));   // cluster dimensions, cooperative, access window ...
        prev = nd;
      }
      cudaGraphExec_t te = nullptr;
      ARL_CHECK(c, cudaGraphInstantiate(&te, tg, 0));
      ARL_CHECK(c, cudaGraphLaunch(te, st));
      ARL_CHECK(c, cudaEventRecord(e0, st));
      ARL_CHECK(c, cudaGraphLaunch(te, st));
      ARL_CHECK(c, cudaEventRecord(e1, st));
      ARL_CHECK(c, cudaStreamSynchronize(st));
      float t = 0.f;
      ARL_CHECK(c, cudaEventElapsedTime(&t, e0, e1));
      ms[cnt++] = t / reps;
      cudaGraphExecDestroy(te);
      cudaGraphDestroy(tg);
      all += (li < c->prof_labels.size()) ? c->prof_labels[li] : std::string("kernel");
      all += ';';
      ++li;
    }
    size_t nd = 1;
    cudaGraphNode_t next = nullptr;
    cudaError_t de = cudaGraphNodeGetDependentNodes(node, &next, &nd);
    if (de != cudaSuccess || nd == 0) break;
    node = next;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  // the update kernels ran `reps` extra times: counters / optimizer state moved; restore the counters only
  ARL_CHECK(c, cudaMemsetAsync(c->mb_counter, 0, sizeof(int), st));
  ARL_CHECK(c, cudaStreamSynchronize(st));
  *n = cnt;
  snprintf(names, names_cap, "%s", all.c_str());
  cudaGraphExecDestroy(ge);
  cudaGraphDestroy(g);
  cudaStreamDestroy(cap_s);
  return 0;
}

/* Completion time of every kernel of ONE training minibatch inside the product's own forked graph (all streams), in
   microseconds after the graph's first node: events are captured behind every launch, the graph is replayed a few times and
   the events of the last replay are read.  kind 0: local update (clip_update), 1: synchronous step. */
int arl_profile_timeline(arl_ctx* c, int kind, const int* idx, int mb_size, char* names, int names_cap, float* us, int cap,
                         int* n, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  TrainPlan* P = nullptr;
  if (get_plan(c, mb_size, &P)) return 1;
  const bool overlap = kind == 1 && sync_overlap_ok(c);
  if (overlap && sync_overlap_prepare(c)) return 1;
  cudaStream_t cap_s;
  ARL_CHECK(c, create_main_stream(&cap_s));
  if (!c->tl_buf && dev_alloc(c, &c->tl_buf, 96)) return 1;
  c->prof_n = 0;
  c->prof_timeline = true;
  long l0 = c->launches;
  cudaGraph_t g = nullptr;
  ARL_CHECK(c, cudaStreamBeginCapture(cap_s, cudaStreamCaptureModeThreadLocal));
  prof_mark(c, "begin", cap_s);
  c->train_step_active = (kind == 0);
  c->sync_overlap_active = overlap;
  int rc = grad_minibatch(c, idx, c->mb_counter, mb_size, cap_s);
  if (!rc) rc = kind == 1 ? (overlap ? sync_tail(c, cap_s) : sync_update(c, cap_s)) : clip_update(c, 1.f, cap_s);
  c->train_step_active = false; c->sync_overlap_active = false; c->early_fc_done = false; c->pending_fin = nullptr;
  c->pending_stream = false;
  cudaError_t ce = cudaStreamEndCapture(cap_s, &g);
  c->prof_timeline = false;
  c->launches = l0;
  if (rc) { if (g) cudaGraphDestroy(g); cudaStreamDestroy(cap_s); return rc; }
  ARL_CHECK(c, ce);
  cudaGraphExec_t ge = nullptr;
  ARL_CHECK(c, cudaGraphInstantiate(&ge, g, 0));
  ARL_CHECK(c, cudaMemsetAsync(c->mb_counter, 0, sizeof(int), st));
  for (int i = 0; i < 4; ++i) ARL_CHECK(c, cudaGraphLaunch(ge, st));
  ARL_CHECK(c, cudaMemsetAsync(c->mb_counter, 0, sizeof(int), st));
  ARL_CHECK(c, cudaStreamSynchronize(st));
  std::string all;
  int cnt = 0;
  unsigned long long h[96];
  ARL_CHECK(c, cudaMemcpy(h, c->tl_buf, sizeof(h), cudaMemcpyDeviceToHost));
  for (int i = 1; i < c->prof_n && cnt < cap; ++i) {
    us[cnt++] = (float)((double)(long long)(h[i] - h[0]) * 1e-3);
    all += c->prof_names[i];
    all += ';';
  }
  *n = cnt;
  snprintf(names, names_cap, "%s", all.c_str());
  c->prof_n = 0;
  cudaGraphExecDestroy(ge);
  cudaGraphDestroy(g);
  cudaStreamDestroy(cap_s);
  return 0;
}

int arl_profile_begin(arl_ctx* c, void* stream) {
  c->prof_on = true;
  c->prof_n = 0;
  prof_mark(c, "begin", (cudaStream_t)stream);
  return 0;
}

int arl_profile_end(arl_ctx* c, char* names, int names_cap, float* ms, int cap, int* n, void* stream) {
  ARL_CHECK(c, cudaStreamSynchronize((cudaStream_t)stream));
  c->prof_on = false;
  int cnt = std::min(cap, c->prof_n - 1);
  std::string all;
  for (int i = 0; i < cnt; ++i) {
    float t = 0.f;
    ARL_CHECK(c, cudaEventElapsedTime(&t, c->prof_ev[i], c->prof_ev[i + 1]));
    ms[i] = t;
    all += c->prof_names[i + 1];
    all += ';';
  }
  *n = cnt < 0 ? 0 : cnt;
  snprintf(names, names_cap, "%s", all.c_str());
  return 0;
}

int arl_test_gemm(arl_ctx* c, const uint16_t* a_bf16, const uint16_t* b_bf16, float* d, int M, int N, int K,
                  int b_nmajor, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (K % 64 || N % 64) ARL_FAIL(c, "test gemm needs K%64==0 and N%64==0");
  DenseLoader<128> a{};
  a.src = reinterpret_cast<const __nv_bfloat16*>(a_bf16); a.ld = K; a.nrows = M;
  RowEpi e = make_epi(EPI_PARTIAL_F32);
  e.partial = d; e.ldo = N; e.M = M;
  if (b_nmajor) {
    WeightSrc w{reinterpret_cast<const __nv_bfloat16*>(b_bf16), (long)N, K, RowPerm{0, 0}};
    return launch_rowgemm<DenseLoader<128>, true, 64>(c, a, w, e, M, N, K / 64, K / 64, 1, st);
  }
  WeightSrc w{reinterpret_cast<const __nv_bfloat16*>(b_bf16), (long)K, 0, RowPerm{0, 0}};
  return launch_rowgemm<DenseLoader<128>, false, 64>(c, a, w, e, M, N, K / 64, K / 64, 1, st);
}

int arl_test_wgrad(arl_ctx* c, const uint16_t* a_bf16, const uint16_t* b_bf16, float* d, int rows, int Kp, int N,
                   void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (Kp % 64) ARL_FAIL(c, "test wgrad needs Kp%64==0");
  DenseLoader<64> a{};
  a.src = reinterpret_cast<const __nv_bfloat16*>(a_bf16); a.ld = Kp; a.nrows = rows;
  WgradEpi e{};
  e.out = d; e.mode = 0; e.Kvalid = Kp; e.Kp = Kp; e.ldo = N;
  const __nv_bfloat16* dy = reinterpret_cast<const __nv_bfloat16*>(b_bf16);
  int rps = roundup(rows, 64);
  int atoms = Kp / 64;
  if (N == 256) return launch_wgrad<DenseLoader<64>, 1, 256>(c, a, dy, N, rows, rps, 1, atoms, 1, e, st);
  if (N == 64) return launch_wgrad<DenseLoader<64>, 4, 64>(c, a, dy, N, rows, rps, 1, atoms, 1, e, st);
  if (N == 32) return launch_wgrad<DenseLoader<64>, 2, 32>(c, a, dy, N, rows, rps, 1, atoms, 1, e, st);
  if (N == 16) return launch_wgrad<DenseLoader<64>, 2, 16>(c, a, dy, N, rows, rps, 1, atoms, 1, e, st);
  ARL_FAIL(c, "test wgrad supports N in {16,32,64,256}");
}

}  // extern "C"

