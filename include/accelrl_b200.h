/* accelrl_b200.h — C ABI of libaccelrl_b200.so (B200 / sm_100a).
 *
 * The reference (astooke/accel_rl) has no FFI: its plugin surface is Python classes handed to a
 * Runner.  Each entry point below states the reference interface it stands in
 * for (paths relative to the reference tree).  All pointers are raw DEVICE pointers unless
 * marked host; sizes are element counts; every function returns 0 on success and a non-zero
 * code otherwise (arl_last_error() gives the message).  No function allocates device memory
 * after the corresponding *_create / *_configure call, and none takes a torch type.
 * `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 */
#ifndef ACCELRL_B200_H
#define ACCELRL_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct arl_ctx arl_ctx;

#define ARL_MAX_CONV 4

/* Network description — policies/pg/atari_cnn_policy.py:17-24 ctor args + env_spec
 * (policies/pg/networks/pg_cnn.py:17-33).  One hidden FC layer (cnn_specs 0, 1, 4 shapes). */
typedef struct {
  int n_conv;
  int conv_filters[ARL_MAX_CONV];
  int conv_sizes[ARL_MAX_CONV];
  int conv_strides[ARL_MAX_CONV];
  int conv_pads[ARL_MAX_CONV];
  int hidden;              /* hidden_sizes[0] */
  int n_actions;           /* env_spec.action_space.n (<= 18) */
  int in_c, in_h, in_w;    /* observation_space.shape = (num_img_obs, 104, 80) */
  float pixel_scale;       /* 255. */
  int max_rows;            /* largest batch any forward/backward call will use */
} arl_net_cfg;

/* Loss / optimiser description — algos/pg/{aac_base,ppo,a2c}.py ctor args and
 * optimizers/single/{ppo,a2c}_optimizer.py ctor args. */
typedef struct {
  int algo;                /* 0 = PPO (ppo.py:42-51), 1 = A2C (a2c.py:43-46) */
  float clip_param;        /* PPO ratio clip (scaled by lr_mult, ppo.py:46) */
  float v_loss_coeff, ent_loss_coeff;
  int update;              /* 0 = adam, 1 = rmsprop (optimizers/update_methods_stats.py) */
  float learning_rate, beta1, beta2, epsilon, rho;
  float grad_norm_clip;    /* <= 0: None (norm is still reported) */
  int ppo_tie_grad;        /* 0 or 1: inside the clip range, where surr_1 == surr_2 exactly, min() passes the gradient once
                              (Theano >= 0.9 `Minimum.L_op`: "gx will be gz, gy will be 0", and standard PPO);
                              2: to both branches, i.e. twice the policy-loss gradient there (older Theano's eq/eq form) */
} arl_opt_cfg;

/* Rollout buffers + synthetic-emulator description — sampler/act_server/buffers.py:7-38,
 * buffers/batch.py:36-76 (row = env*T + t) and envs/atari_env.py ctor args. */
typedef struct {
  int n_envs, horizon, planes;
  uint8_t* observations;        /* [N][planes][104][80] */
  float* rewards;               /* [N] */
  uint8_t* dones;               /* [N] bool */
  float* raw_reward;            /* [N] env_infos.raw_reward */
  uint8_t* need_reset;          /* [N] env_infos.need_reset */
  uint8_t* actions;             /* [N] */
  float* prob;                  /* [N][A] agent_infos.prob */
  float* value;                 /* [N]    agent_infos.value */
  uint8_t* extra_observations;  /* [B][planes][104][80] */
  uint8_t* step_obs;            /* [B][planes][104][80] (the sampler's step buffer) */
  double* uniforms;             /* [horizon][B] np.random.rand draws, step-major */
  const uint8_t* frame_pool;    /* [pool_frames][210][160] grayscale emulator frames */
  int pool_frames;
  int max_path_length;
  float discount;
  int mid_batch_reset, clip_reward, episodic_lives;
  /* synthetic emulator rules (oracle/synth_ale.py implements the same) */
  int lives0, life_base, life_mul, life_mod, reward_mod, frame_stride;
  int traj_cap;                 /* capacity of the completed-trajectory record buffer */
  int ext_emulator;             /* != 0: emulators run in host worker processes; steps arrive through arl_rollout_ingest */
  int n_games;                  /* > 1: game mix of the synthetic emulator (env e plays game e % n_games: own slice of the
                                   frame pool, own reward table and life clock); the policy's action count is the largest
                                   of the games' minimal action sets (BASELINE configs[2]) */
  int frame_mode;               /* 0: reference frames (210,160) gray -> (planes,104,80), atari_env.py:151-157;
                                 * 1: north-star frames (210,160,3) RGB -> gray -> (planes,84,84) (see arl_frame_update_rgb);
                                 *    frame_pool / staging then hold RGB frames */
} arl_sampler_cfg;

/* ---- lifetime ------------------------------------------------------------------------- */
int arl_create(const arl_net_cfg* cfg, arl_ctx** out);   /* AtariCnnPolicy.initialize (atari_cnn_policy.py:42-76) */
void arl_destroy(arl_ctx* ctx);
const char* arl_last_error(arl_ctx* ctx);                /* ctx may be NULL: last creation error */
int arl_device_error(arl_ctx* ctx);                      /* device-side watchdog flag (0 = ok) */

/* ---- parameters: rllab/core/parameterized.py:74-88 flat fp32 vector, Lasagne order --------- */
long arl_param_count(arl_ctx* ctx);
/* fills offsets/sizes (elements) of conv{i}.W, conv{i}.b, hidden.W, hidden.b, pi.W, pi.b, v.W, v.b */
int arl_param_layout(arl_ctx* ctx, long* offsets, long* sizes, int cap);
/* register the flat fp32 params / grad / optimiser-state vectors (owned by the caller) */
int arl_bind_params(arl_ctx* ctx, float* params, float* grad, float* m, float* v);
/* refresh the bf16 operand copies after params changed (set_param_values) */
int arl_pack_weights(arl_ctx* ctx, void* stream);

/* ---- policy: AtariCnnPolicy.get_actions / value / dist_info_value (atari_cnn_policy.py:92-111) */
/* obs [*, C, H, W] u8; idx (optional) gathers n rows; outputs written at out_rows[i] (or i).
 * prob/value/actions/uniforms may each be NULL (actions needs uniforms). */
int arl_policy_forward(arl_ctx* ctx, const uint8_t* obs, const int* idx, int n, const int* out_rows,
                       float* prob, float* value, const double* uniforms, uint8_t* actions, void* stream);
/* Discrete.weighted_sample_n (spaces/discrete.py:67-68 -> rllab/misc/special.py:22-27) */
int arl_sample_actions(arl_ctx* ctx, const float* prob, const double* uniforms, uint8_t* actions, int n, int n_actions,
                       void* stream);

/* ---- frame pipeline: AtariEnv._update_obs (envs/atari_env.py:151-157) ------------------------ */
/* raw_a/raw_b: [n][210][160] u8 grayscale screens (raw_a NULL => zeros, i.e. after _reset_obs);
 * reset_mask[n] (optional, host semantics of _reset_obs: zero the older planes);
 * stack [n][planes][104][80] updated in place. */
int arl_frame_update(arl_ctx* ctx, const uint8_t* raw_a, const uint8_t* raw_b, const uint8_t* reset_mask, uint8_t* stack,
                     int n, int planes, void* stream);
/* North-star frame mode (BASELINE.json north_star; not in the reference, whose emulator returns grayscale at
 * envs/atari_env.py:147-149): raw RGB pairs [n][210][160][3] u8 -> per-channel max -> gray (77R+150G+29B+128)>>8 ->
 * exact-area 84x84 resize -> stack [n][planes][84][84] u8 shifted oldest->newest (+ optional bf16 copy, same shape).
 * raw_a may be NULL (single frame); reset_mask[i] != 0 zeroes item i's stack and ignores raw_a (atari_env.py:159-163). */
int arl_frame_update_rgb(arl_ctx* ctx, const uint8_t* raw_a, const uint8_t* raw_b, const uint8_t* reset_mask, uint8_t* stack,
                         uint16_t* stack_bf16, int n, int planes, void* stream);

/* ---- sampler: ActsrvAltOvrlpSampler.obtain_samples (sampler/.../overlap/sampler.py:97-151) --- */
int arl_sampler_configure(arl_ctx* ctx, const arl_sampler_cfg* cfg);
/* Two sampler slots per context: 0 = training envs (default), 1 = evaluation envs (AAOEvalSampler's eval_envs /
 * eval_step_bufs, overlap/sampler_with_eval.py:6-54, worker_with_eval.py:66-99).  configure / reset / rollout_* /
 * traj_read act on the selected slot; a slot configured with observations == NULL stores no observations (evaluation
 * keeps only rewards/dones/actions/agent infos rows and the TrajInfo records). */
int arl_sampler_select(arl_ctx* ctx, int slot);
int arl_sampler_reset(arl_ctx* ctx, void* stream);                  /* start_envs (sampler/util.py:26-57) */
/* start_envs with max_decorrelation_steps > 0 (sampler/util.py:33-55): after arl_sampler_reset env e takes n_steps[e]
   (device int array [n_envs], each <= max_steps) warm-up steps, is reset whenever its trajectory ends (no TrajInfo is
   reported, nothing is recorded) and starts the first rollout from the observation it reached */
int arl_sampler_warmup(arl_ctx* ctx, const int* n_steps, int max_steps, void* stream);
int arl_rollout_begin(arl_ctx* ctx, void* stream);
/* one serve+step: forward on step_obs, sample, env step, frame update.  staging (optional):
 * host-fed raw frames [B][2][210][160] already on the device for this step */
int arl_rollout_step(arl_ctx* ctx, int s, const uint8_t* staging, void* stream);
int arl_rollout_end(arl_ctx* ctx, void* stream);
/* External-emulator feed (reference: the simulator workers of overlap/worker.py:116-153 running envs/atari_env.py:65-100
 * on host cores).  Per env-step the workers hand over one record and the raw frame pair; the device keeps the pixels
 * (frame pipeline), the policy and the rollout buffers.
 *   arl_rollout_serve(s):   policy forward on the step buffer + action sampling -> rows e*T+s of actions/prob/value
 *   arl_rollout_ingest(s, staging [B][2][210][160(x3)] u8 DEVICE, ext [B] DEVICE): files reward/done/infos into rows
 *                            e*T+s and runs the frame pipeline into the step buffer and row s+1.  s == -1: the frames of
 *                            start_envs (sampler/util.py:26-57); s == horizon: those of reset_needed_envs
 *                            (worker.py:106-113) — step buffer only. */
typedef struct {
  float reward;                 /* clipped when clip_reward */
  float raw_reward;
  uint8_t done;
  uint8_t need_reset;           /* env_info["need_reset"] */
  uint8_t flags;                /* bit0: reset / life loss: older planes zeroed, frame 1 = zeros (atari_env.py:159-163);
                                 * bit1: observation not advanced (NonResetCollector: the env finished, worker.py:84-95);
                                 * bit2: env not stepped at all, nothing is recorded (worker.py:78) */
  uint8_t pad;
} arl_ext_step;
/* both act on envs [e0, e0 + n) (n < 0: through the last env): the two alternating groups of the reference sampler are
 * served one after the other, so one group's emulators run while the other group is on the GPU.  staging / ext are the
 * bases of the whole-batch blocks. */
int arl_rollout_serve(arl_ctx* ctx, int s, int e0, int n, void* stream);
int arl_rollout_ingest(arl_ctx* ctx, int s, int e0, int n, const uint8_t* staging, const arl_ext_step* ext, void* stream);
/* page-lock / unlock host memory the workers share with the master (POSIX shared memory), and plain async copies on a
 * stream, so raw frames go pinned-host -> HBM without a bounce buffer */
int arl_host_register(void* ptr, size_t bytes);
int arl_host_unregister(void* ptr);
int arl_copy_async(arl_ctx* ctx, void* dst, const void* src, size_t bytes, int to_device, void* stream);
/* begin + horizon steps + end, replayed from a CUDA graph (resident frame pool) */
int arl_rollout_run(arl_ctx* ctx, void* stream);
/* completed TrajInfo records (sampler/util.py:75-101); host arrays of capacity cap; returns count via *n */
int arl_traj_read(arl_ctx* ctx, int* n, int* env, int* len, float* ret, float* raw, int* nz, float* disc, int cap,
                  void* stream);
/* pool indices the emulator will show env e at step s (host-fed path): fills [B][2] (-1 = zeros) + flags[B] */
int arl_peek_frame_cmds(arl_ctx* ctx, int* cmd_host, int n_envs, void* stream);

/* ---- advantages: AdvActorCriticBase.process_samples (algos/pg/aac_base.py:108-145) ---------- */
int arl_gae(arl_ctx* ctx, const float* rewards, float* values, const uint8_t* dones, const uint8_t* need_reset,
            const float* last_values, float discount, float gae_lambda, float* adv, float* ret, int8_t* valids,
            int n_envs, int horizon, int standardize, void* stream);

/* ---- learner: PpoOptimizer / A2cOptimizer (optimizers/single/*.py) --------------------------- */
int arl_opt_configure(arl_ctx* ctx, const arl_opt_cfg* cfg);
/* bind the training inputs (prep_opt_inputs, aac_base.py:147-170): rollout-length arrays */
int arl_bind_train_inputs(arl_ctx* ctx, const uint8_t* obs, const uint8_t* actions, const float* adv, const float* ret,
                          const float* old_value, const float* old_prob, const int8_t* valids, long n_rows);
int arl_set_lr_mult(arl_ctx* ctx, float lr_mult, void* stream);
/* forward + losses + backward for rows idx[mb_index*mb_size .. +mb_size) -> flat grad (no update) */
int arl_grad_minibatch(arl_ctx* ctx, const int* idx, int mb_size, void* stream);
/* total_norm_constraint + update (optimizers/util.py:70-76); gscale = gradient averaging factor */
int arl_clip_update(arl_ctx* ctx, float gscale, void* stream);
/* `count` consecutive minibatches (grad + update + repack), idx = [count*mb_size] int32; graph-replayed */
int arl_train_minibatches(arl_ctx* ctx, const int* idx, int mb_size, int count, void* stream);
/* the same loop for the synchronous multi-GPU learner: every minibatch ends with arl_sync_allreduce_update instead of
 * the local clip+update (sync_ppo_optimizer.py:56-72); one CUDA graph per minibatch incl. the cooperative all-reduce */
int arl_train_minibatches_sync(arl_ctx* ctx, const int* idx, int mb_size, int count, void* stream);
/* the asynchronous learner's PPO loop: every minibatch ends with arl_async_push_pull (one CUDA graph per minibatch) */
int arl_train_minibatches_async(arl_ctx* ctx, const int* idx, int mb_size, int count, void* stream);
/* per-update logs since the last call: losses and pre-clip grad norms (host arrays) */
int arl_read_logs(arl_ctx* ctx, float* loss, float* grad_norm, int cap, int* n, void* stream);
int arl_reset_opt_state(arl_ctx* ctx, void* stream);
/* the update count t of Adam's bias correction (update_methods_stats.py:70-73), for snapshot / resume of the optimizer
 * state next to the caller-owned m and v vectors (the reference snapshots parameters only, accel_rl_base.py:108-113) */
int arl_opt_step_get(arl_ctx* ctx, int* t, void* stream);
int arl_opt_step_set(arl_ctx* ctx, int t, void* stream);

/* ---- sync data parallel: optimizers/sync/base.py:8-24 + sync_ppo_optimizer.py:13-78 ----------- */
#define ARL_IPC_HANDLE_BYTES 64
/* allocate this rank's exchange buffers (grad mirror, flags); returns the IPC handle to publish */
int arl_comm_local_init(arl_ctx* ctx, int rank, int world, uint8_t* handle_out /*[ARL_IPC_HANDLE_BYTES] host*/);
/* the symmetric flat gradient / parameter vectors peers read and write (bind these with arl_bind_params) */
int arl_comm_buffers(arl_ctx* ctx, float** grad_out, float** params_out);
/* open every peer's handle (host array [world][ARL_IPC_HANDLE_BYTES]) */
int arl_comm_connect(arl_ctx* ctx, const uint8_t* all_handles);
/* fused: reduce my slice of the flat gradient over peers by P2P loads, average, global-norm
 * clip, Adam/RMSProp on the slice, P2P-store the new params to every peer. */
int arl_sync_allreduce_update(arl_ctx* ctx, void* stream);
/* device-side timeline (%globaltimer) of the overlapped synchronous step since the last reset, microseconds per step:
   out[0] FC exchange waiting for peers, out[1] FC exchange reduce + update + publish, out[2] tail waiting for peers,
   out[3] tail average + update, out[4] slack between FC exchange end and tail start (> 0: hidden behind the conv gradient
   chain), out[5] number of steps */
int arl_comm_trace(arl_ctx* ctx, double* out, int reset, void* stream);
int arl_comm_barrier(arl_ctx* ctx, void* stream);

/* ---- asynchronous data parallel ------------------------------------------------------------------
 * replaces BaseAsyncOptimizer / chunked_updates (accel_rl/optimizers/async/base.py:11-104,
 * chunked_updates.py:53-120): a central (params, m, v) store in rank 0's HBM, shared by a CUDA IPC handle, updated
 * under chunk-granular locks with each learner's LOCALLY clipped gradient; the learner then holds the new central
 * parameters.  local_init: rank 0 allocates the store from its bound parameters and returns the 64-byte handle
 * (zeros on the other ranks); connect: every rank passes rank 0's handle. */
int arl_async_local_init(arl_ctx* ctx, int rank, int world, int n_update_chunks, uint8_t* handle_out);
int arl_async_connect(arl_ctx* ctx, const uint8_t* rank0_handle);
int arl_async_regions(arl_ctx* ctx);   /* lock regions the chunks were subdivided into */
int arl_async_push_pull(arl_ctx* ctx, void* stream);
/* pull only: central parameters -> local parameters + operand copies, per lock region under its lock
   (ActsrvAltOvrlpPollSampler: the sampler refreshes its policy every poll_horizon rollout steps, poll_sampler.py:29-39) */
int arl_async_pull(arl_ctx* ctx, void* stream);
int arl_async_read_central(arl_ctx* ctx, int which, float* host_out, long n, void* stream);

/* ---- diagnostics / tests ---------------------------------------------------------------------- */
/* intermediate activations of the last forward (bf16 -> fp32 copies into host-visible device buffers) */
int arl_debug_activation(arl_ctx* ctx, int layer, float* out, long cap, long* n, void* stream);
long arl_kernel_launches(arl_ctx* ctx);
/* sha256 (hex) of the CUDA sources the loaded binary was compiled from; __graft_entry__.smoke() and the GPU tests compare it
 * with the sources next to the library, so a stale prebuilt .so is caught on the GPU box */
const char* arl_source_hash(void);
/* CUDA-event timing of every kernel launched (outside graphs) between begin and end, on `stream`:
 * names = ';'-separated launch labels, ms[i] = device time of launch i */
int arl_profile_begin(arl_ctx* ctx, void* stream);
/* same, but measured inside a replayed CUDA graph (event-record node after every kernel): kind 0 = one training
 * minibatch over idx[0..mb_size) (+ update), kind 1 = one rollout step; replayed `reps` times, last replay reported */
int arl_profile_graph(arl_ctx* ctx, int kind, const int* idx, int mb_size, int reps, char* names, int names_cap, float* ms,
                      int cap, int* n, void* stream);
/* completion time (microseconds after the first node) of every kernel of ONE training minibatch inside the product's own
 * forked multi-stream graph: kind 0 = local update, 1 = synchronous data-parallel step */
int arl_profile_timeline(arl_ctx* ctx, int kind, const int* idx, int mb_size, char* names, int names_cap, float* us, int cap,
                         int* n, void* stream);
int arl_profile_end(arl_ctx* ctx, char* names, int names_cap, float* ms, int cap, int* n, void* stream);   /* launches issued (graph replays count their node count) */
/* plain tcgen05 GEMM self-test: D[M][N] = A[M][K] * B (B K-major [N][K] or N-major [K][N]) */
int arl_test_gemm(arl_ctx* ctx, const uint16_t* a_bf16, const uint16_t* b_bf16, float* d, int M, int N, int K,
                  int b_nmajor, void* stream);
/* wgrad-shaped self-test: D[Kp][N] = sum_r A[r][Kp] * B[r][N] (bf16 in, fp32 out) */
int arl_test_wgrad(arl_ctx* ctx, const uint16_t* a_bf16, const uint16_t* b_bf16, float* d, int rows, int Kp, int N,
                   void* stream);

#ifdef __cplusplus
}
#endif
#endif
