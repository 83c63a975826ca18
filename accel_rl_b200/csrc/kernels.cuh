// HBM-bound and reduction kernels of the rollout + A2C/PPO path (sm_100a):
//   synthetic-emulator step + Atari frame pipeline, policy head (softmax, sampling, losses,
//   head backward), GAE scan, gradient finalisation, global-norm clip + Adam/RMSProp, weight packing.
// Reference semantics cited per kernel (paths relative to the reference tree).
#pragma once
#include "common.cuh"

namespace arl {

// ===========================================================================
// Synthetic emulator ("SynthALE") — replaces atari_py.ALEInterface for benchmarks/parity.
// Deterministic function of (env id, emulator frame counter); the oracle's fake ALE
// (oracle/synth_ale.py) implements the identical rules on the CPU.
// ===========================================================================
struct SynthCfg {
  int pool_frames;      // number of frames in the pool
  int lives0;           // lives at reset_game (5)
  int life_base;        // life period = life_base + (env * life_mul) % life_mod   [emulator frames]
  int life_mul, life_mod;
  int reward_mod;       // a frame pays when hash(env, f) % reward_mod == 0
  int frame_stride;     // pool index = (env + frame_stride * f) % pool_frames
  // game mix (BASELINE configs[2] "4-game mix"): env e plays game g = e % n_games; every game has its own slice of the
  // frame pool (pool_frames / n_games frames), its own reward table (reward_mod + 6 g) and life clock (life_base + 17 g)
  int n_games;          // <= 1: one game
};
ARL_DEVINL int synth_game(const SynthCfg& c, int e) { return c.n_games > 1 ? e % c.n_games : 0; }

ARL_DEVINL uint32_t synth_hash(uint32_t e, uint32_t f) {
  uint32_t h = e * 0x9E3779B1u + f * 0x85EBCA77u + 0x165667B1u;
  h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
  return h;
}
ARL_DEVINL float synth_reward(const SynthCfg& c, int e, int f) {
  uint32_t h = synth_hash((uint32_t)e, (uint32_t)f);
  const uint32_t rm = (uint32_t)(c.reward_mod + 6 * synth_game(c, e));
  if (h % rm != 0) return 0.f;
  uint32_t k = (h / rm) & 3u;
  return k == 2 ? 4.f : (k == 3 ? -1.f : 1.f);
}
ARL_DEVINL int synth_life_period(const SynthCfg& c, int e) {
  return c.life_base + 17 * synth_game(c, e) + (e * c.life_mul) % c.life_mod;
}
ARL_DEVINL int synth_lives(const SynthCfg& c, int e, int f) {
  int l = c.lives0 - f / synth_life_period(c, e);
  return l < 0 ? 0 : l;
}
ARL_DEVINL int synth_frame_index(const SynthCfg& c, int e, int f) {
  if (c.n_games > 1) {
    const int fpg = c.pool_frames / c.n_games;
    return synth_game(c, e) * fpg + (int)(((long)e + (long)c.frame_stride * f) % fpg);
  }
  return (int)(((long)e + (long)c.frame_stride * f) % c.pool_frames);
}

// per-env emulator + AtariEnv + collector state (struct of arrays in one int/float block)
struct EnvState {
  int* f;            // emulator frame counter since reset_game
  int* lives_seen;   // AtariEnv._lives
  int* need_reset;   // NonResetCollector: env finished earlier in this batch (not stepped again)
  int* traj_len;     // TrajInfo.Length
  float* traj_ret;   // Return (clipped)
  float* traj_raw;   // RawReturn
  int* traj_nz;      // NonzeroRewards
  float* traj_disc;  // DiscountedReturn
  float* traj_cur;   // _cur_discount
};

struct TrajOut {     // completed-episode records, appended with an atomic cursor
  int* count;        // [1]
  int cap;
  int* env;          // [cap]
  int* len;
  float* ret;
  float* raw;
  int* nz;
  float* disc;
};

// what the frame kernel must do for each env this step
struct FrameCmd {
  int src_a;   // pool index of raw frame 1 (or -1: zeros)
  int src_b;   // pool index of raw frame 2
  int flags;   // bit0: zero the older planes (reset / life loss), bit1: skip (env not stepped)
};

// One thread per env.  Mirrors AtariEnv.step (envs/atari_env.py:65-78), _done_episodic_lives
// (:185-191), reset (:93-100) and ResetCollector/NonResetCollector.collect
// (sampler/act_server/alternating/overlap/worker.py:25-113) for step s of the batch.
struct EnvStepArgs {
  SynthCfg cfg; EnvState st; TrajOut tout; FrameCmd* cmd;
  float* rewards; uint8_t* dones; float* raw_reward; uint8_t* info_need_reset;
  int n_envs, T, s, max_path_length; float discount; int mid_batch_reset, clip_reward, episodic_lives;
  // start_envs decorrelation (sampler/util.py:33-55): warm_n != nullptr -> warm-up step number warm_k; env e takes it
  // only while warm_k < warm_n[e], is reset at once when its trajectory ends, and nothing is recorded
  const int* warm_n; int warm_k;
};

ARL_DEVINL void env_step_one(const EnvStepArgs& a, int e);

__global__ void env_step_kernel(EnvStepArgs a) {
  pdl_wait();
  pdl_trigger();
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.n_envs) return;
  env_step_one(a, e);
}

// one env, one thread: also called by thread 0 of head_kernel<0> right after it sampled env e's action (the rollout
// step then needs no separate env-step launch)
ARL_DEVINL void env_step_one(const EnvStepArgs& a, int e) {
  const SynthCfg& cfg = a.cfg; const EnvState& st = a.st; const TrajOut& tout = a.tout; FrameCmd* cmd = a.cmd;
  float* rewards = a.rewards; uint8_t* dones = a.dones; float* raw_reward = a.raw_reward;
  uint8_t* info_need_reset = a.info_need_reset;
  const int T = a.T, s = a.s, max_path_length = a.max_path_length, mid_batch_reset = a.mid_batch_reset;
  const int clip_reward = a.clip_reward, episodic_lives = a.episodic_lives;
  const float discount = a.discount;
  FrameCmd c;
  c.flags = 0;
  const bool warm = a.warm_n != nullptr;
  if (warm && a.warm_k >= a.warm_n[e]) {
    c.src_a = c.src_b = -1;
    c.flags = 2;
    cmd[e] = c;
    return;
  }
  if (!mid_batch_reset && st.need_reset[e]) {
    c.src_a = c.src_b = -1;
    c.flags = 2;  // stale rows stay as they are (worker.py:78 `if not need_reset[i]`)
    cmd[e] = c;
    return;
  }
  int f = st.f[e];
  float r = 0.f;
  r += synth_reward(cfg, e, f + 1);
  r += synth_reward(cfg, e, f + 2);
  r += synth_reward(cfg, e, f + 3);
  f += 3;
  c.src_a = synth_frame_index(cfg, e, f);   // _get_screen(1) after the 3rd repeat
  f += 1;
  r += synth_reward(cfg, e, f);
  c.src_b = synth_frame_index(cfg, e, f);   // _get_screen(2) in _update_obs
  float raw = r;
  if (clip_reward) r = (r > 0.f) ? 1.f : (r < 0.f ? -1.f : 0.f);
  // _done_*: game_over first, then life check
  int lives = synth_lives(cfg, e, f);
  bool game_over = (lives == 0);
  bool lost_life = (lives < st.lives_seen[e]) && (lives > 0);
  if (lost_life) {
    f += 2;  // _life_reset: act(0), act(FIRE)
    st.lives_seen[e] = synth_lives(cfg, e, f);
    if (episodic_lives) {
      c.src_a = -1;                       // _reset_obs zeroes raw frames and the stack
      c.src_b = synth_frame_index(cfg, e, f);
      c.flags |= 1;
    }
  }
  bool done = episodic_lives ? (lost_life || game_over) : game_over;
  bool need_reset_info = game_over;       // only reported under episodic_lives (info["need_reset"])
  // TrajInfo.step (sampler/util.py:92-98)
  int len = st.traj_len[e] + 1;
  float tret = st.traj_ret[e] + r;
  float traw = st.traj_raw[e] + (clip_reward ? raw : r);
  int tnz = st.traj_nz[e] + (r != 0.f ? 1 : 0);
  float cur = st.traj_cur[e];
  float tdisc = st.traj_disc[e] + cur * r;
  cur *= discount;
  bool over_length = len > max_path_length;
  bool env_says_reset = episodic_lives ? need_reset_info : true;   // env_info.get("need_reset", True)
  if (over_length || (done && env_says_reset)) {
    done = true;
    if (over_length && episodic_lives) need_reset_info = true;
    if (!warm) {
      int slot = atomicAdd(tout.count, 1);
      if (slot < tout.cap) {
        tout.env[slot] = e; tout.len[slot] = len; tout.ret[slot] = tret; tout.raw[slot] = traw;
        tout.nz[slot] = tnz; tout.disc[slot] = tdisc;
      }
    }
    len = 0; tret = 0.f; traw = 0.f; tnz = 0; tdisc = 0.f; cur = 1.f;
    if (mid_batch_reset) {
      // env.reset(): reset_game, _reset_obs, _life_reset (2 acts), 0 start no-ops, _update_obs
      f = 2;
      st.lives_seen[e] = synth_lives(cfg, e, f);
      c.src_a = -1;
      c.src_b = synth_frame_index(cfg, e, f);
      c.flags |= 1;
    } else {
      st.need_reset[e] = 1;
      c.flags |= 2;  // observation is NOT advanced (worker.py:91-95 else-branch skipped)
    }
  }
  st.f[e] = f;
  st.traj_len[e] = len; st.traj_ret[e] = tret; st.traj_raw[e] = traw; st.traj_nz[e] = tnz;
  st.traj_disc[e] = tdisc; st.traj_cur[e] = cur;
  cmd[e] = c;
  if (warm) return;
  long row = (long)e * T + s;
  rewards[row] = r;
  dones[row] = done ? 1 : 0;
  if (clip_reward) raw_reward[row] = raw;
  if (episodic_lives) info_need_reset[row] = need_reset_info ? 1 : 0;
  cmd[e] = c;
}

// External-emulator feed: the emulators, AtariEnv's emulator control and the collector logic run in
// host worker processes (sampler/host_emulator.py, the reference's worker.py:116-153 + envs/atari_env.py:65-100); each
// step they hand over, per env, one record + the raw frame pair.  This kernel files the record into the rollout rows
// and turns its flags into the frame kernel's command (frames come from the staging pair, so src_* only say
// "present" (0) or "zeros" (-1)).
struct ExtStep {          // mirrors arl_ext_step (include/accelrl_b200.h)
  float reward;           // clipped when clip_reward
  float raw_reward;
  uint8_t done;
  uint8_t need_reset;     // env_info["need_reset"]
  uint8_t flags;          // bit0: zero the older planes and use zeros for frame 1 (reset / life loss);
                          // bit1: observation not advanced; bit2: env not stepped (nothing recorded)
  uint8_t pad;
};

__global__ void ext_apply_kernel(const ExtStep* __restrict__ ext, FrameCmd* __restrict__ cmd, float* __restrict__ rewards,
                                 uint8_t* __restrict__ dones, float* __restrict__ raw_reward,
                                 uint8_t* __restrict__ info_need_reset, int n_envs, int T, int s, int clip_reward,
                                 int episodic_lives) {
  pdl_wait();
  pdl_trigger();
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_envs) return;
  const ExtStep x = ext[e];
  FrameCmd c;
  c.flags = x.flags & 3;
  c.src_a = (x.flags & 3) ? -1 : 0;
  c.src_b = (x.flags & 2) ? -1 : 0;
  cmd[e] = c;
  if (s >= 0 && s < T && !(x.flags & 4)) {
    const long row = (long)e * T + s;
    rewards[row] = x.reward;
    dones[row] = x.done;
    if (clip_reward) raw_reward[row] = x.raw_reward;
    if (episodic_lives) info_need_reset[row] = x.need_reset;
  }
}

// After the batch when mid_batch_reset == False: reset envs flagged need_reset
// (worker.py:106-113 reset_needed_envs) — produces the first observation of the next batch.
__global__ void env_reset_needed_kernel(SynthCfg cfg, EnvState st, FrameCmd* __restrict__ cmd, int n_envs) {
  pdl_wait();
  pdl_trigger();
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_envs) return;
  FrameCmd c;
  if (st.need_reset[e]) {
    st.need_reset[e] = 0;
    st.f[e] = 2;
    st.lives_seen[e] = synth_lives(cfg, e, 2);
    c.src_a = -1; c.src_b = synth_frame_index(cfg, e, 2); c.flags = 1;
  } else {
    c.src_a = c.src_b = -1; c.flags = 2;
  }
  cmd[e] = c;
}

// initial env.reset() for every env (start_envs with max_decorrelation_steps == 0)
__global__ void env_init_kernel(SynthCfg cfg, EnvState st, FrameCmd* __restrict__ cmd, int n_envs) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_envs) return;
  st.f[e] = 2;
  st.lives_seen[e] = synth_lives(cfg, e, 2);
  st.need_reset[e] = 0;
  st.traj_len[e] = 0; st.traj_ret[e] = 0.f; st.traj_raw[e] = 0.f; st.traj_nz[e] = 0;
  st.traj_disc[e] = 0.f; st.traj_cur[e] = 1.f;
  FrameCmd c;
  c.src_a = -1; c.src_b = synth_frame_index(cfg, e, 2); c.flags = 1;
  cmd[e] = c;
}

// ===========================================================================
// Frame pipeline — AtariEnv._update_obs (envs/atari_env.py:151-157):
//   m = max(raw1, raw2); crop rows 208,209; 2x2 box mean (a+b+c+d+2)>>2 (== cv2 INTER_LINEAR at
//   an exact 2x shrink); obs = concat(obs[1:], img).  Stack planes oldest -> newest.
// One work item = 16 output pixels of one env: 8 x 16-byte loads (2 frames x 2 rows x 32 B),
// (P-1) x 16-byte plane shifts and P x 16-byte stores.  Items are flattened over all envs so
// the grid is dense.  Raw frames come either from the resident pool (src index per env) or,
// for the host-fed path, from a staging buffer [n_envs][2][210][160] (cmd.src_* >= 0 selects
// slot 0/1 of the env's staging pair).
// ===========================================================================
constexpr int kRawH = 210, kRawW = 160, kObsH = 104, kObsW = 80;

ARL_DEVINL uint32_t vmax4(uint32_t a, uint32_t b) { return __vmaxu4(a, b); }

// horizontal pair sums of 4 bytes -> two 16-bit lanes: (b0+b1) | (b2+b3)<<16
ARL_DEVINL uint32_t pair_sum(uint32_t v) { return (v & 0x00ff00ffu) + ((v >> 8) & 0x00ff00ffu); }

// 4 packed u8 -> 4 bf16 (exact): byte b -> float(2^23 + b) - 2^23, packed pairwise
ARL_DEVINL uint2 u8x4_to_bf16x4(uint32_t w) {
  float f0 = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7650)) - 8388608.0f;
  float f1 = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7651)) - 8388608.0f;
  float f2 = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7652)) - 8388608.0f;
  float f3 = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7653)) - 8388608.0f;
  return make_uint2(pack_bf16x2(f0, f1), pack_bf16x2(f2, f3));
}

// uint8 CHW observations (the reference's buffer layout) -> bf16 space-to-depth(s) NHWC:
//   dst[i][y/s][x/s][c*s*s + (y%s)*s + (x%s)] = src[idx ? idx[i] : i][c][y][x]       (s == 4)
// One thread converts 4 horizontally adjacent pixels (one aligned 32-bit load, one 8-byte store).
// swz != 0 (C*16 == 64 only): each position's eight 16-byte chunks are XOR-swizzled by (position & 7) — the layout the
// patch-resident conv tiles bulk-copy straight into SWIZZLE_128B shared memory (pconv.cuh)
__global__ void __launch_bounds__(256) obs_to_s2d_kernel(const uint8_t* __restrict__ src, const int* __restrict__ idx,
                                                         __nv_bfloat16* __restrict__ dst, int n, int C, int H, int W,
                                                         int swz, long img_elems) {
  pdl_wait();
  pdl_trigger();
  const int Wb = W >> 2, Hb = H >> 2;
  const long total = (long)n * C * H * Wb;
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    int bx = (int)(t % Wb);
    long r = t / Wb;
    int y = (int)(r % H); r /= H;
    int c = (int)(r % C);
    long i = r / C;
    long img = idx ? idx[i] : i;
    uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(src + ((img * C + c) * H + y) * (long)W) + bx);
    const int pos = (y >> 2) * Wb + bx, ch = c * 16 + (y & 3) * 4;
    long off = i * img_elems + (long)pos * (C * 16) + (swz ? ((((ch >> 3) ^ (pos & 7)) << 3) | (ch & 7)) : ch);   // img_elems >= Hb*Wb*C*16
    *reinterpret_cast<uint2*>(dst + off) = u8x4_to_bf16x4(w);
  }
}

// 16 output pixels (row oy, 16-pixel column group xc) of max(fa, fb) box-downsampled; null frame = zeros
ARL_DEVINL uint4 frame_box16(const uint8_t* __restrict__ fa, const uint8_t* __restrict__ fb, int oy, int xc) {
  const int roff = (2 * oy) * kRawW + xc * 32;
  uint4 z = make_uint4(0, 0, 0, 0);
  uint4 a00 = z, a01 = z, a10 = z, a11 = z, b00 = z, b01 = z, b10 = z, b11 = z;
  if (fa) {
    const uint4* p0 = reinterpret_cast<const uint4*>(fa + roff);
    const uint4* p1 = reinterpret_cast<const uint4*>(fa + roff + kRawW);
    a00 = __ldg(p0); a01 = __ldg(p0 + 1); a10 = __ldg(p1); a11 = __ldg(p1 + 1);
  }
  if (fb) {
    const uint4* p0 = reinterpret_cast<const uint4*>(fb + roff);
    const uint4* p1 = reinterpret_cast<const uint4*>(fb + roff + kRawW);
    b00 = __ldg(p0); b01 = __ldg(p0 + 1); b10 = __ldg(p1); b11 = __ldg(p1 + 1);
  }
  uint32_t top[8] = {vmax4(a00.x, b00.x), vmax4(a00.y, b00.y), vmax4(a00.z, b00.z), vmax4(a00.w, b00.w),
                     vmax4(a01.x, b01.x), vmax4(a01.y, b01.y), vmax4(a01.z, b01.z), vmax4(a01.w, b01.w)};
  uint32_t bot[8] = {vmax4(a10.x, b10.x), vmax4(a10.y, b10.y), vmax4(a10.z, b10.z), vmax4(a10.w, b10.w),
                     vmax4(a11.x, b11.x), vmax4(a11.y, b11.y), vmax4(a11.z, b11.z), vmax4(a11.w, b11.w)};
  uint32_t outw[4];
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    // each input word holds 4 pixels -> 2 output pixels; two words -> 4 output pixels (one out word)
    uint32_t s0 = pair_sum(top[2 * w]) + pair_sum(bot[2 * w]) + 0x00020002u;          // 2 sums (16-bit lanes)
    uint32_t s1 = pair_sum(top[2 * w + 1]) + pair_sum(bot[2 * w + 1]) + 0x00020002u;
    uint32_t o0 = (s0 >> 2) & 0x00ff00ffu;   // lanes: px0 | px1<<16
    uint32_t o1 = (s1 >> 2) & 0x00ff00ffu;
    outw[w] = (o0 & 0xffu) | ((o0 >> 16) << 8) | ((o1 & 0xffu) << 16) | ((o1 >> 16) << 24);
  }
  return make_uint4(outw[0], outw[1], outw[2], outw[3]);
}

__global__ void __launch_bounds__(256) frame_kernel(const uint8_t* __restrict__ pool, const uint8_t* __restrict__ staging,
                                                    const FrameCmd* __restrict__ cmd, uint8_t* __restrict__ step_obs,
                                                    uint8_t* __restrict__ roll_obs, __nv_bfloat16* __restrict__ step_obs16,
                                                    __nv_bfloat16* __restrict__ roll_obs16, int T, int s_next, int n_envs,
                                                    int planes, int swz_step, int swz_roll,
                                                    const uint8_t* __restrict__ prev, long prev_stride, int lean) {
  pdl_wait();
  pdl_trigger();
  // prev: where env e's CURRENT stack is read from (prev + e * prev_stride): the step buffer, or — lean rollout
  // (launch_frame) — rollout row e*T + s_next - 1, in which case the step buffer and its mirror are only written by the
  // last step of the batch (step_obs / step_obs16 == nullptr otherwise) and a skipped env's row is carried forward.
  // step_obs16 / roll_obs16: bf16 space-to-depth(4) mirrors of the same stacks, [26][20][planes*16] per
  // observation, channel = plane*16 + (y%4)*4 + (x%4) — the layout the first conv layer's tcgen05 tiles read
  // (u8 -> bf16 is exact).  They are written from the registers that already hold the u8 stack.
  // step_obs [n_envs][planes][104][80] is the sampler's step buffer (updated in place: each thread
  // shifts its own 16 pixels of every plane); roll_obs row e*T + s_next receives a copy of the new
  // stack (skipped when roll_obs == nullptr, i.e. after the last step of the batch).
  constexpr int ITEMS_PER_ROW = kObsW / 16;             // 5
  constexpr int ITEMS_PER_ENV = kObsH * ITEMS_PER_ROW;  // 520
  long item = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= (long)n_envs * ITEMS_PER_ENV) return;
  int e = (int)(item / ITEMS_PER_ENV);
  int rem = (int)(item - (long)e * ITEMS_PER_ENV);
  int oy = rem / ITEMS_PER_ROW;
  int xc = rem - oy * ITEMS_PER_ROW;
  FrameCmd c = cmd[e];
  const bool skip = (c.flags & 2) != 0;
  if (skip && !lean) return;   // env not stepped: nothing is written (rows keep their stale contents)
  const int plane_bytes = kObsH * kObsW;
  const long obs_bytes = (long)planes * plane_bytes;
  const uint8_t* prv = prev + (long)e * prev_stride;
  uint8_t* cur = step_obs ? step_obs + (long)e * obs_bytes : nullptr;
  uint8_t* dst = roll_obs ? roll_obs + ((long)e * T + s_next) * obs_bytes : nullptr;
  const int pix = oy * kObsW + xc * 16;
  // ---- new plane ----
  const long fbytes = (long)kRawH * kRawW;
  const uint8_t* fa = nullptr;
  const uint8_t* fb = nullptr;
  if (staging) {
    if (c.src_a >= 0) fa = staging + ((long)e * 2 + 0) * fbytes;
    if (c.src_b >= 0) fb = staging + ((long)e * 2 + 1) * fbytes;
  } else {
    if (c.src_a >= 0) fa = pool + (long)c.src_a * fbytes;
    if (c.src_b >= 0) fb = pool + (long)c.src_b * fbytes;
  }
  uint4 z = make_uint4(0, 0, 0, 0);
  uint4 newest;
  // ---- stack shift + store ----
  uint4 keep[3] = {z, z, z};
  if (skip) {                  // (lean) carry the unchanged stack forward
    for (int p = 0; p < planes - 1 && p < 3; ++p)
      keep[p] = *reinterpret_cast<const uint4*>(prv + p * plane_bytes + pix);
    newest = *reinterpret_cast<const uint4*>(prv + (planes - 1) * plane_bytes + pix);
  } else {
    newest = frame_box16(fa, fb, oy, xc);
    if (!(c.flags & 1)) {
      for (int p = 0; p < planes - 1 && p < 3; ++p)
        keep[p] = *reinterpret_cast<const uint4*>(prv + (p + 1) * plane_bytes + pix);
    }
  }
  for (int p = 0; p < planes - 1 && p < 3; ++p) {
    if (cur) *reinterpret_cast<uint4*>(cur + p * plane_bytes + pix) = keep[p];
    if (dst) *reinterpret_cast<uint4*>(dst + p * plane_bytes + pix) = keep[p];
  }
  if (cur) *reinterpret_cast<uint4*>(cur + (planes - 1) * plane_bytes + pix) = newest;
  if (dst) *reinterpret_cast<uint4*>(dst + (planes - 1) * plane_bytes + pix) = newest;
  if (step_obs16 || roll_obs16) {
    const int Cs = planes * 16;
    const long obs16_elems = (long)(kObsH / 4) * (kObsW / 4) * Cs;
    __nv_bfloat16* c16 = step_obs16 ? step_obs16 + (long)e * obs16_elems : nullptr;
    __nv_bfloat16* d16 = roll_obs16 ? roll_obs16 + ((long)e * T + s_next) * obs16_elems : nullptr;
    const int by = oy >> 2, dy = oy & 3;
    for (int p = 0; p < planes; ++p) {
      const uint4 v = (p == planes - 1) ? newest : keep[p < 3 ? p : 2];
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int pos = by * (kObsW / 4) + xc * 4 + b, ch = p * 16 + dy * 4;
        const long off = (long)pos * Cs + ch;
        const long offs = (long)pos * Cs + ((((ch >> 3) ^ (pos & 7)) << 3) | (ch & 7));   // chunk-swizzled (Cs == 64)
        const uint2 o = u8x4_to_bf16x4(w[b]);
        if (c16) *reinterpret_cast<uint2*>(c16 + (swz_step ? offs : off)) = o;
        if (d16) *reinterpret_cast<uint2*>(d16 + (swz_roll ? offs : off)) = o;
      }
    }
  }
}

// ===========================================================================
// North-star frame mode (BASELINE.json north_star; NOT in the reference, whose emulator hands out grayscale):
//   two raw RGB frames (210,160,3) u8 -> per-channel max -> gray -> 84x84 area resize -> 4-plane stack (u8, oldest
//   first) + the same stack as bf16 for the network, all in ONE pass.  Builder-defined arithmetic (integer, exact;
//   frozen in oracle/frame.py:rgb_*):
//     gray  Y = (77 R + 150 G + 29 B + 128) >> 8            (NTSC luma, 8-bit fixed point, weights sum to 256)
//     resize: exact area average.  Rows: 210/84 = 5/2 -> output row y covers half-rows [5y, 5y+5) (input row i covers
//             [2i, 2i+2)); columns: 160/84 = 40/21 -> output column x covers [40x, 40x+40) in units of 1/21 column (input
//             column j covers [21j, 21j+21)).  out = (sum wy*wx*Y + 100) / 200.
// One block = one env x two output rows = five input rows: the 2 x 5 x 480 raw bytes are read once with 16-byte loads
// (max of the two frames in registers), gray values go to shared memory, 168 threads produce the 2 x 84 outputs and
// shift their own pixel of the stack.  Algorithmic traffic: 2 x 100 800 B read + 7 056 B new plane.
// ===========================================================================
constexpr int kRgbH = 210, kRgbW = 160, kNsH = 84, kNsW = 84;
constexpr int kRgbRows = 4;                      // output rows per block (= 10 input rows)
constexpr int kRgbThreads = 128;
constexpr int kRgbInRows = kRgbRows * 5 / 2;     // 10
constexpr int kRgbGrayWords = kRgbInRows * kRgbW / 4;   // 400 words of gray per slab (+1 pad word: the last column's
                                                        // zero-weight third tap reads one byte past the row)

// gray value of pixel K (0..15) of a 48-byte group held in 12 words: the pixel's R,G,B bytes are moved to byte lanes
// 0..2 (byte lane 3 carries weight 0) and one dp4a forms 77 R + 150 G + 29 B + 128
template <int K>
ARL_DEVINL uint32_t rgb_gray_px(const uint32_t (&w)[12]) {
  constexpr int b = 3 * K, wi = b >> 2, sh = b & 3;
  uint32_t v;
  if constexpr (sh == 0) v = w[wi];
  else if constexpr (sh == 1) v = w[wi] >> 8;
  else if constexpr (sh == 2) v = __byte_perm(w[wi], w[(wi + 1) % 12], 0x4432);
  else v = __byte_perm(w[wi], w[(wi + 1) % 12], 0x5543);
  return __dp4a(v, 0x001D964Du, 128u) >> 8;
}

template <int I>
ARL_DEVINL uint32_t rgb_gray_word(const uint32_t (&w)[12]) {
  return rgb_gray_px<4 * I>(w) | (rgb_gray_px<4 * I + 1>(w) << 8) | (rgb_gray_px<4 * I + 2>(w) << 16) |
         (rgb_gray_px<4 * I + 3>(w) << 24);
}

// One slab = 10 input rows of an RGB frame pair (4800 bytes each, 16-byte aligned) -> 4 output rows of 84 pixels in
// s_out.  Phase A: 100 threads each own 48 bytes = 16 pixels: six 16-byte loads in flight, per-channel max and the gray
// conversion in registers, one 16-byte store of gray.  Phase B: thread x < 84 owns output column x: its three input
// columns start at j0 = 40x/21; per input row ONE dp4a over the (unaligned) byte triple with the packed column
// weights, then the four output rows are 3-row combinations (2,2,1 / 1,2,2).  Ends with a __syncthreads().
ARL_DEVINL void rgb_slab_to_rows(const uint8_t* __restrict__ fa, const uint8_t* __restrict__ fb, uint32_t* s_gray,
                                 uint8_t* s_out, int tid) {
  constexpr int GROUPS48 = kRgbInRows * kRgbW * 3 / 48;   // 100
  if (tid < GROUPS48) {
    uint32_t w[12];
    if (fb) {
      const uint4* pb = reinterpret_cast<const uint4*>(fb) + tid * 3;
      uint4 b0 = __ldg(pb), b1 = __ldg(pb + 1), b2 = __ldg(pb + 2);
      if (fa) {
        const uint4* pa = reinterpret_cast<const uint4*>(fa) + tid * 3;
        uint4 a0 = __ldg(pa), a1 = __ldg(pa + 1), a2 = __ldg(pa + 2);
        b0.x = __vmaxu4(a0.x, b0.x); b0.y = __vmaxu4(a0.y, b0.y); b0.z = __vmaxu4(a0.z, b0.z); b0.w = __vmaxu4(a0.w, b0.w);
        b1.x = __vmaxu4(a1.x, b1.x); b1.y = __vmaxu4(a1.y, b1.y); b1.z = __vmaxu4(a1.z, b1.z); b1.w = __vmaxu4(a1.w, b1.w);
        b2.x = __vmaxu4(a2.x, b2.x); b2.y = __vmaxu4(a2.y, b2.y); b2.z = __vmaxu4(a2.z, b2.z); b2.w = __vmaxu4(a2.w, b2.w);
      }
      w[0] = b0.x; w[1] = b0.y; w[2] = b0.z; w[3] = b0.w; w[4] = b1.x; w[5] = b1.y; w[6] = b1.z; w[7] = b1.w;
      w[8] = b2.x; w[9] = b2.y; w[10] = b2.z; w[11] = b2.w;
    } else if (fa) {
      const uint4* pa = reinterpret_cast<const uint4*>(fa) + tid * 3;
      uint4 a0 = __ldg(pa), a1 = __ldg(pa + 1), a2 = __ldg(pa + 2);
      w[0] = a0.x; w[1] = a0.y; w[2] = a0.z; w[3] = a0.w; w[4] = a1.x; w[5] = a1.y; w[6] = a1.z; w[7] = a1.w;
      w[8] = a2.x; w[9] = a2.y; w[10] = a2.z; w[11] = a2.w;
    } else {
#pragma unroll
      for (int i = 0; i < 12; ++i) w[i] = 0u;
    }
    uint4 g;
    g.x = rgb_gray_word<0>(w); g.y = rgb_gray_word<1>(w); g.z = rgb_gray_word<2>(w); g.w = rgb_gray_word<3>(w);
    reinterpret_cast<uint4*>(s_gray)[tid] = g;
  }
  if (tid == kRgbThreads - 1) s_gray[kRgbGrayWords] = 0u;
  __syncthreads();
  if (tid < kNsW) {
    const int c_lo = 40 * tid;
    const int j0 = c_lo / 21;
    const int w0 = 21 * (j0 + 1) - c_lo;          // 1..21
    const int w1 = min(21, 40 - w0);
    const int w2 = 40 - w0 - w1;                  // 0 for the columns whose window spans two input columns only
    const uint32_t wpack = (uint32_t)w0 | ((uint32_t)w1 << 8) | ((uint32_t)w2 << 16);
    const int wbase = j0 >> 2;
    const uint32_t sh = (uint32_t)(j0 & 3) * 8u;
    uint32_t h[kRgbInRows];
#pragma unroll
    for (int r = 0; r < kRgbInRows; ++r) {
      const uint32_t lo = s_gray[r * (kRgbW / 4) + wbase], hi = s_gray[r * (kRgbW / 4) + wbase + 1];
      h[r] = __dp4a(__funnelshift_r(lo, hi, sh), wpack, 0u);
    }
    // output rows come in pairs over 5 input rows: even row -> rows 0,1,2 (weights 2,2,1), odd row -> rows 2,3,4 (1,2,2)
#pragma unroll
    for (int q = 0; q < kRgbRows / 2; ++q) {
      const uint32_t ev = 2u * h[5 * q] + 2u * h[5 * q + 1] + h[5 * q + 2];
      const uint32_t od = h[5 * q + 2] + 2u * h[5 * q + 3] + 2u * h[5 * q + 4];
      s_out[(2 * q) * kNsW + tid] = (uint8_t)((ev + 100u) / 200u);
      s_out[(2 * q + 1) * kNsW + tid] = (uint8_t)((od + 100u) / 200u);
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kRgbThreads) frame_rgb_kernel(const uint8_t* __restrict__ raw_a,
                                                                const uint8_t* __restrict__ raw_b,
                                                                const uint8_t* __restrict__ reset_mask,
                                                                uint8_t* __restrict__ stack, __nv_bfloat16* __restrict__ stack16,
                                                                int n, int planes) {
  __shared__ __align__(16) uint32_t s_gray[kRgbGrayWords + 4];
  __shared__ __align__(4) uint8_t s_out[kRgbRows * kNsW];
  constexpr int GROUPS = kNsH / kRgbRows;        // 21 row groups per env
  constexpr int WPR = kNsW / 4;                  // 21 words per output row
  const int e = blockIdx.x / GROUPS;
  const int grp = blockIdx.x - e * GROUPS;
  const int tid = threadIdx.x;
  const bool rs = reset_mask && reset_mask[e];
  const long fbytes = (long)kRgbH * kRgbW * 3;
  const long roff = (long)grp * kRgbInRows * kRgbW * 3;
  const uint8_t* fa = (raw_a && !rs) ? raw_a + (long)e * fbytes + roff : nullptr;
  const uint8_t* fb = raw_b + (long)e * fbytes + roff;
  // stack shift + stores, 4 pixels per thread; the older planes are fetched before the frame work so their latency hides
  const bool outp = tid < kRgbRows * WPR;
  const int oy = tid / WPR, xw = tid - oy * WPR;
  const int pix = (grp * kRgbRows + oy) * kNsW + xw * 4;
  const int plane_px = kNsH * kNsW;
  uint8_t* cur = stack + (long)e * planes * plane_px + pix;
  uint32_t older[3] = {0u, 0u, 0u};
  if (outp && !rs) {
#pragma unroll
    for (int p = 1; p < 4; ++p)
      if (p < planes) older[p - 1] = *reinterpret_cast<const uint32_t*>(cur + p * plane_px);
  }
  rgb_slab_to_rows(fa, fb, s_gray, s_out, tid);
  if (!outp) return;
  __nv_bfloat16* cur16 = stack16 ? stack16 + (long)e * planes * plane_px + pix : nullptr;
  const uint32_t newest = *reinterpret_cast<const uint32_t*>(s_out + oy * kNsW + xw * 4);
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    if (p >= planes) break;
    const uint32_t v = (p < planes - 1) ? older[p < 3 ? p : 2] : newest;
    *reinterpret_cast<uint32_t*>(cur + p * plane_px) = v;
    if (cur16) *reinterpret_cast<uint2*>(cur16 + p * plane_px) = u8x4_to_bf16x4(v);
  }
}

// The same pipeline inside the rollout (arl_rollout_step, frame_mode 1): frames come from the RGB pool (or the
// host-fed staging pair) as directed by the env step's FrameCmd; writes the sampler's step buffer in place, the rollout
// row e*T + s_next, and the bf16 space-to-depth(4) mirrors the first conv layer reads ((84/4)^2 = 441 positions x
// planes*16 channels per observation, image stride img16 elements, optionally chunk-swizzled).
__global__ void __launch_bounds__(kRgbThreads) frame_rgb_roll_kernel(
    const uint8_t* __restrict__ pool, const uint8_t* __restrict__ staging, const FrameCmd* __restrict__ cmd,
    uint8_t* __restrict__ step_obs, uint8_t* __restrict__ roll_obs, __nv_bfloat16* __restrict__ step_obs16,
    __nv_bfloat16* __restrict__ roll_obs16, int T, int s_next, int n_envs, int planes, int swz_step, int swz_roll,
    long img16, const uint8_t* __restrict__ prev, long prev_stride, int lean) {
  pdl_wait();
  pdl_trigger();
  // prev / lean: as frame_kernel
  __shared__ __align__(16) uint32_t s_gray[kRgbGrayWords + 4];
  __shared__ __align__(4) uint8_t s_out[kRgbRows * kNsW];
  constexpr int GROUPS = kNsH / kRgbRows;
  constexpr int WPR = kNsW / 4;
  const int e = blockIdx.x / GROUPS;
  const int grp = blockIdx.x - e * GROUPS;
  const int tid = threadIdx.x;
  const FrameCmd c = cmd[e];
  const bool skip = (c.flags & 2) != 0;
  if (skip && !lean) return;                     // env not stepped: rows keep their stale contents
  const bool rs = (c.flags & 1) != 0;
  const long fbytes = (long)kRgbH * kRgbW * 3;
  const long roff = (long)grp * kRgbInRows * kRgbW * 3;
  const uint8_t* fa = nullptr;
  const uint8_t* fb = nullptr;
  if (staging) {
    if (c.src_a >= 0) fa = staging + ((long)e * 2 + 0) * fbytes + roff;
    if (c.src_b >= 0) fb = staging + ((long)e * 2 + 1) * fbytes + roff;
  } else {
    if (c.src_a >= 0) fa = pool + (long)c.src_a * fbytes + roff;
    if (c.src_b >= 0) fb = pool + (long)c.src_b * fbytes + roff;
  }
  const bool outp = tid < kRgbRows * WPR;
  const int oy = tid / WPR, xw = tid - oy * WPR;
  const int Y = grp * kRgbRows + oy;
  const int pix = Y * kNsW + xw * 4;
  const int plane_px = kNsH * kNsW;
  const long obs_bytes = (long)planes * plane_px;
  const uint8_t* prv = prev + (long)e * prev_stride + pix;
  uint8_t* cur = step_obs ? step_obs + (long)e * obs_bytes + pix : nullptr;
  uint32_t older[3] = {0u, 0u, 0u};
  uint32_t newest = 0u;
  if (outp && skip) {                            // (lean) carry the unchanged stack forward
#pragma unroll
    for (int p = 0; p < 3; ++p)
      if (p < planes - 1) older[p] = *reinterpret_cast<const uint32_t*>(prv + p * plane_px);
    newest = *reinterpret_cast<const uint32_t*>(prv + (planes - 1) * plane_px);
  } else if (outp && !rs) {
#pragma unroll
    for (int p = 1; p < 4; ++p)
      if (p < planes) older[p - 1] = *reinterpret_cast<const uint32_t*>(prv + p * plane_px);
  }
  if (!skip) rgb_slab_to_rows(fa, fb, s_gray, s_out, tid);     // (skip is uniform over the block)
  if (!outp) return;
  uint8_t* dst = roll_obs ? roll_obs + ((long)e * T + s_next) * obs_bytes + pix : nullptr;
  __nv_bfloat16* c16 = step_obs16 ? step_obs16 + (long)e * img16 : nullptr;
  __nv_bfloat16* d16 = roll_obs16 ? roll_obs16 + ((long)e * T + s_next) * img16 : nullptr;
  const int Cs = planes * 16;
  const int pos = (Y >> 2) * (kNsW / 4) + xw;
  if (!skip) newest = *reinterpret_cast<const uint32_t*>(s_out + oy * kNsW + xw * 4);
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    if (p >= planes) break;
    const uint32_t v = (p < planes - 1) ? older[p < 3 ? p : 2] : newest;
    if (cur) *reinterpret_cast<uint32_t*>(cur + p * plane_px) = v;
    if (dst) *reinterpret_cast<uint32_t*>(dst + p * plane_px) = v;
    if (c16 || d16) {
      const int ch = p * 16 + (Y & 3) * 4;
      const long off = (long)pos * Cs + ch;
      const long offs = (long)pos * Cs + ((((ch >> 3) ^ (pos & 7)) << 3) | (ch & 7));
      const uint2 o2 = u8x4_to_bf16x4(v);
      if (c16) *reinterpret_cast<uint2*>(c16 + (swz_step ? offs : off)) = o2;
      if (d16) *reinterpret_cast<uint2*>(d16 + (swz_roll ? offs : off)) = o2;
    }
  }
}

// standalone frame update (arl_frame_update): item i's raw frames are raw_a[i], raw_b[i]
__global__ void make_cmd_kernel(FrameCmd* cmd, const uint8_t* reset_mask, int n, int has_a) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  FrameCmd c;
  bool rs = reset_mask && reset_mask[i];
  c.src_a = (has_a && !rs) ? i : -1;
  c.src_b = i;
  c.flags = rs ? 1 : 0;
  cmd[i] = c;
}

__global__ void __launch_bounds__(256) frame_pair_kernel(const uint8_t* __restrict__ raw_a, const uint8_t* __restrict__ raw_b,
                                                         const FrameCmd* __restrict__ cmd, uint8_t* __restrict__ stack,
                                                         int n, int planes) {
  constexpr int ITEMS_PER_ROW = kObsW / 16;
  constexpr int ITEMS_PER_ENV = kObsH * ITEMS_PER_ROW;
  long item = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= (long)n * ITEMS_PER_ENV) return;
  int e = (int)(item / ITEMS_PER_ENV);
  int rem = (int)(item - (long)e * ITEMS_PER_ENV);
  int oy = rem / ITEMS_PER_ROW;
  int xc = rem - oy * ITEMS_PER_ROW;
  FrameCmd c = cmd[e];
  const long fbytes = (long)kRawH * kRawW;
  const uint8_t* fa = (c.src_a >= 0 && raw_a) ? raw_a + (long)e * fbytes : nullptr;
  const uint8_t* fb = raw_b + (long)e * fbytes;
  uint4 newest = frame_box16(fa, fb, oy, xc);
  const int plane_bytes = kObsH * kObsW;
  uint8_t* cur = stack + (long)e * planes * plane_bytes;
  const int pix = oy * kObsW + xc * 16;
  uint4 z = make_uint4(0, 0, 0, 0);
  uint4 keep[3] = {z, z, z};
  if (!(c.flags & 1))
    for (int p = 0; p < planes - 1 && p < 3; ++p) keep[p] = *reinterpret_cast<const uint4*>(cur + (p + 1) * plane_bytes + pix);
  for (int p = 0; p < planes - 1 && p < 3; ++p) *reinterpret_cast<uint4*>(cur + p * plane_bytes + pix) = keep[p];
  *reinterpret_cast<uint4*>(cur + (planes - 1) * plane_bytes + pix) = newest;
}

// rows copy: dst[drows[i]] = src[srows[i]] (row_bytes multiple of 16)
__global__ void copy_rows_kernel(const uint8_t* __restrict__ src, long sstride, const int* __restrict__ srows,
                                 uint8_t* __restrict__ dst, long dstride, const int* __restrict__ drows, int n,
                                 int row_bytes) {
  pdl_wait();
  pdl_trigger();
  int per = row_bytes / 16;
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)n * per) return;
  int r = (int)(i / per);
  int c = (int)(i - (long)r * per);
  long sr = srows ? srows[r] : r, dr = drows ? drows[r] : r;
  reinterpret_cast<uint4*>(dst + dr * dstride)[c] = __ldg(reinterpret_cast<const uint4*>(src + sr * sstride) + c);
}

// row index tables for step s: rows[e] = e*T + s
__global__ void fill_rows_kernel(int* __restrict__ rows, int n, int T, int s) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) rows[e] = e * T + s;
}

// ===========================================================================
// Policy head.  One warp per sample row.
//   h = relu(fc_bias + sum_splits partial)  (rounded to bf16: the value the backward pass sees)
//   pi = softmax(h Wpi + bpi), V = h Wv + bv        (policies/pg/networks/pg_cnn.py:70-86)
// mode 0 (rollout): write prob/value, sample the action
//     weighted_sample_n (rllab/misc/special.py:22-27): k = #{j : f64(cumsum_f32(p)_j) < u}; a = min(k, A-1)
// mode 1 (train): PPO / A2C losses and their gradient w.r.t. logits, V and h
//     (algos/pg/aac_base.py:60-70, ppo.py:42-51, a2c.py:43-46, distributions/categorical.py:35-88)
// ===========================================================================
constexpr int kMaxActions = 18;

struct HeadParams {
  const float* partial;     // [splits][M][H]
  int splits, M, H, A;
  const float* fc_bias;     // [H]
  const float* w_pi;        // [H][A]
  const float* b_pi;        // [A]
  const float* w_v;         // [H]
  const float* b_v;         // [1]
  // rollout outputs (row -> out_rows[row] or row)
  const int* out_rows;
  float* prob;              // [N][A]
  float* value;             // [N]
  uint8_t* actions;         // [N]
  const double* uniforms;   // [M]
  // train inputs (gathered through idx)
  const int* idx;           // rows of the rollout buffer: sample row reads idx[off*M + row]
  const int* idx_off;       // optional device scalar `off` (minibatch index)
  const uint8_t* act_in;    // [N]
  const float* adv;         // [N]
  const float* ret;         // [N]
  const float* old_prob;    // [N][A]
  const int8_t* valids;     // [N] or nullptr
  const float* hyper;       // device scalars: [0]=lr_mult (PPO clip scales with it)
  int algo;                 // 0 = PPO, 1 = A2C
  float clip_param, v_coeff, ent_coeff;
  float tie_grad;           // PPO: gradient multiplier inside the clip range (1, or 2 = both branches of the tied min())
  float inv_count;          // 1/M when valids == nullptr
  const float* valid_count; // device scalar sum(valids) (when valids != nullptr)
  // train outputs
  __nv_bfloat16* h_out;     // [M][H] post-ReLU hidden (bf16)
  __nv_bfloat16* dh_out;    // [M][H] gradient w.r.t. FC pre-activation (bf16)
  EnvStepArgs es; int es_on;   // mode 0: run the env step of env `row` after sampling its action
  __nv_bfloat16* dh_t;      // optional second copy as [H/64][dh_rows][64] planes, chunk-swizzled (fcgemm.cuh)
  int dh_rows;
  float* dlogit_out;        // [M][A+1]  (last column: dV)
  float* loss_partial;      // [M][4]  pi, v, ent, total per sample row
};

constexpr int kHeadThreads = 128;

// One block (4 warps) per sample row: thread t owns hidden units j = t, t+128, ... (<= 4 for H <= 512).
// The split-K partial sums are the only HBM-latency-bound part, so every thread keeps 4 x unroll
// independent loads in flight; logits/value are reduced over the block through shared memory and the
// softmax / loss arithmetic is done redundantly by every thread (a handful of flops).
// AMAX: compile-time bound on the action count (4 / 6 / 9 / 18).  Every per-action loop is fully unrolled over AMAX with
// a runtime `a < A` guard, and predicated-off instructions still issue: with one 18-wide instantiation a 4-action game
// executed 2 565 warp instructions per row (ncu), 3/4 of them dead.
template <int MODE, int AMAX>
__global__ void __launch_bounds__(kHeadThreads) head_kernel(HeadParams p) {
  pdl_wait();
  pdl_trigger();
  __shared__ float s_part[kHeadThreads / 32][AMAX + 1];
  const int row = blockIdx.x;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  constexpr int JT = 4;
  float h[JT];
#pragma unroll
  for (int i = 0; i < JT; ++i) {
    int j = t + kHeadThreads * i;
    h[i] = (j < p.H) ? __ldg(p.fc_bias + j) : 0.f;
  }
  // MODE 1: the per-sample training inputs hang off a dependent chain (index -> row -> action -> old probability);
  // start it now so its DRAM round trips overlap the split-K loads and the head arithmetic
  long src = 0;
  int act = 0;
  float adv = 0.f, ret = 0.f, w_row = 0.f;
  float old_p[AMAX];
  if (MODE == 1) {
    src = p.idx ? p.idx[(p.idx_off ? (long)p.idx_off[0] * p.M : 0) + row] : row;
    act = p.act_in[src];
    adv = p.adv[src];
    ret = p.ret[src];
    w_row = p.inv_count;
    if (p.valids) w_row = p.valids[src] ? (1.f / p.valid_count[0]) : 0.f;
#pragma unroll
    for (int a = 0; a < AMAX; ++a) old_p[a] = (p.algo == 0 && a < p.A) ? p.old_prob[src * p.A + a] : 0.f;
  }
  const float* prow = p.partial + (long)row * p.H + t;
  const long sstride = (long)p.M * p.H;
#pragma unroll 4
  for (int s = 0; s < p.splits; ++s) {
#pragma unroll
    for (int i = 0; i < JT; ++i)
      if (t + kHeadThreads * i < p.H) h[i] += prow[s * sstride + kHeadThreads * i];
  }
  float logit[AMAX];
#pragma unroll
  for (int a = 0; a < AMAX; ++a) logit[a] = 0.f;
  float v = 0.f;
#pragma unroll
  for (int i = 0; i < JT; ++i) {
    int j = t + kHeadThreads * i;
    if (j < p.H) {
      float acc = fmaxf(h[i], 0.f);
      acc = __bfloat162float(__float2bfloat16_rn(acc));     // the value the backward pass sees
      h[i] = acc;
      v += acc * __ldg(p.w_v + j);
      const float* wr = p.w_pi + (long)j * p.A;
#pragma unroll
      for (int a = 0; a < AMAX; ++a)
        if (a < p.A) logit[a] += acc * __ldg(wr + a);
    }
  }
  v = warp_sum(v);
#pragma unroll
  for (int a = 0; a < AMAX; ++a)
    if (a < p.A) logit[a] = warp_sum(logit[a]);
  if (lane == 0) {
#pragma unroll
    for (int a = 0; a < AMAX; ++a)
      if (a < p.A) s_part[warp][a] = logit[a];
    s_part[warp][AMAX] = v;
  }
  __syncthreads();
  v = __ldg(p.b_v);
#pragma unroll
  for (int w = 0; w < kHeadThreads / 32; ++w) v += s_part[w][AMAX];
  float mx = -INFINITY;
#pragma unroll
  for (int a = 0; a < AMAX; ++a)
    if (a < p.A) {
      float l = __ldg(p.b_pi + a);
#pragma unroll
      for (int w = 0; w < kHeadThreads / 32; ++w) l += s_part[w][a];
      logit[a] = l;
      mx = fmaxf(mx, l);
    }
  float prob[AMAX];
  float sum = 0.f;
#pragma unroll
  for (int a = 0; a < AMAX; ++a) {
    prob[a] = 0.f;
    if (a < p.A) { prob[a] = expf(logit[a] - mx); sum += prob[a]; }
  }
#pragma unroll
  for (int a = 0; a < AMAX; ++a) prob[a] = prob[a] / sum;

  if (MODE == 0) {
    if (t == 0) {
      long orow = p.out_rows ? p.out_rows[row] : row;
      if (p.prob) {
#pragma unroll
        for (int a = 0; a < AMAX; ++a)
          if (a < p.A) p.prob[orow * p.A + a] = prob[a];
      }
      if (p.value) p.value[orow] = v;
      if (p.actions) {
        double u = p.uniforms[row];
        float cs = 0.f;
        int k = 0;
#pragma unroll
        for (int a = 0; a < AMAX; ++a)
          if (a < p.A) { cs = __fadd_rn(cs, prob[a]); k += ((double)cs < u) ? 1 : 0; }
        p.actions[orow] = (uint8_t)min(k, p.A - 1);
      }
      if (p.es_on) env_step_one(p.es, row);
    }
  } else {
    const float w = w_row;
    const float TINY = 1e-8f;
    float g[AMAX];   // dL/dprob
    float ent = 0.f;
    float pa = 0.f;
#pragma unroll
    for (int a = 0; a < AMAX; ++a) {
      g[a] = 0.f;
      if (a < p.A) {
        float lp = logf(prob[a] + TINY);
        ent -= prob[a] * lp;
        g[a] = p.ent_coeff * w * (lp + prob[a] / (prob[a] + TINY));
        if (a == act) pa = prob[a];
      }
    }
    float l_pi;
    float gact;
    if (p.algo == 0) {
      float po = 0.f;
#pragma unroll
      for (int a = 0; a < AMAX; ++a)
        if (a == act) po = old_p[a];
      float ratio = (pa + TINY) / (po + TINY);
      float cp = p.clip_param * p.hyper[0];
      float lo = 1.f - cp, hi = 1.f + cp;
      float clipped = fminf(fmaxf(ratio, lo), hi);
      float s1 = ratio * adv, s2 = clipped * adv;
      l_pi = -fminf(s1, s2);
      float gr;
      if (ratio < lo) gr = (adv >= 0.f) ? adv : 0.f;
      else if (ratio > hi) gr = (adv <= 0.f) ? adv : 0.f;
      else gr = adv * p.tie_grad;
      gact = -w * gr / (po + TINY);
    } else {
      l_pi = -logf(pa + TINY) * adv;
      gact = -w * adv / (pa + TINY);
    }
#pragma unroll
    for (int a = 0; a < AMAX; ++a)
      if (a == act) g[a] += gact;
    float verr = v - ret;
    float dv = 2.f * p.v_coeff * w * verr;
    float dot = 0.f;
#pragma unroll
    for (int a = 0; a < AMAX; ++a) dot += prob[a] * g[a];
    float dl[AMAX];
#pragma unroll
    for (int a = 0; a < AMAX; ++a) dl[a] = prob[a] * (g[a] - dot);
    if (t == 0) {
      float lp_ = w * l_pi, lv_ = p.v_coeff * w * verr * verr, le_ = -p.ent_coeff * w * ent;
      float* o = p.loss_partial + 4 * (long)row;
      o[0] = lp_; o[1] = lv_; o[2] = le_; o[3] = lp_ + lv_ + le_;
#pragma unroll
      for (int a = 0; a < AMAX; ++a)
        if (a < p.A) p.dlogit_out[(long)row * (p.A + 1) + a] = dl[a];
      p.dlogit_out[(long)row * (p.A + 1) + p.A] = dv;
    }
#pragma unroll
    for (int i = 0; i < JT; ++i) {
      int j = t + kHeadThreads * i;
      if (j < p.H) {
        float d = dv * __ldg(p.w_v + j);
        const float* wr = p.w_pi + (long)j * p.A;
#pragma unroll
        for (int a = 0; a < AMAX; ++a)
          if (a < p.A) d += dl[a] * __ldg(wr + a);
        d = (h[i] > 0.f) ? d : 0.f;
        p.h_out[(long)row * p.H + j] = __float2bfloat16_rn(h[i]);
        p.dh_out[(long)row * p.H + j] = __float2bfloat16_rn(d);
        if (p.dh_t)
          p.dh_t[((long)(j >> 6) * p.dh_rows + row) * 64 + (((((j & 63) >> 3) ^ (row & 7)) << 3) | (j & 7))] = __float2bfloat16_rn(d);
      }
    }
  }
}

// Head weight gradients + FC bias gradient, partial over row groups (deterministic 2-stage).
//   out[g][j][0..A-1] = sum_rows h[row][j]*dlogit[row][a];  [A] = sum h*dV;  [A+1] = sum dh[row][j]
// plus (block j-group 0 only) out_b[g][0..A] = sum_rows dlogit[row][:] (bias grads of pi, v).
__global__ void __launch_bounds__(128) head_wgrad_kernel(const __nv_bfloat16* __restrict__ h,
                                                          const __nv_bfloat16* __restrict__ dh,
                                                          const float* __restrict__ dlogit, int M, int H, int A,
                                                          int rows_per_group, float* __restrict__ out,
                                                          float* __restrict__ out_b) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float s_dl[];   // [rows_per_group][A+1]
  const int g = blockIdx.y;
  const int r0 = g * rows_per_group;
  const int r1 = min(M, r0 + rows_per_group);
  const int A1 = A + 1;
  for (int i = threadIdx.x; i < (r1 - r0) * A1; i += blockDim.x) s_dl[i] = dlogit[(long)r0 * A1 + i];
  __syncthreads();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < H) {
    float acc[kMaxActions + 2];
#pragma unroll
    for (int a = 0; a < kMaxActions + 2; ++a) acc[a] = 0.f;
    for (int r = r0; r < r1; ++r) {
      float hv = __bfloat162float(h[(long)r * H + j]);
      float dv = __bfloat162float(dh[(long)r * H + j]);
      const float* dl = s_dl + (r - r0) * A1;
#pragma unroll
      for (int a = 0; a < kMaxActions + 1; ++a)
        if (a < A1) acc[a] += hv * dl[a];
      acc[kMaxActions + 1] += dv;
    }
    float* o = out + ((long)g * H + j) * (A + 2);
    for (int a = 0; a < A1; ++a) o[a] = acc[a];
    o[A1] = acc[kMaxActions + 1];
  }
  if (blockIdx.x == 0 && threadIdx.x < A1) {
    float s = 0.f;
    for (int r = r0; r < r1; ++r) s += s_dl[(r - r0) * A1 + threadIdx.x];
    out_b[g * A1 + threadIdx.x] = s;
  }
}

// ===========================================================================
// Gradient finalisation: sum split partials and scatter into the flat fp32 gradient in the
// reference's parameter order (rllab/core/parameterized.py:74-88; Lasagne W (out,in,kh,kw),
// flip_filters=True so correlation tap (ky,kx) is W[..., kh-1-ky, kw-1-kx]).
// ===========================================================================
enum GradMap { GM_LINEAR = 0, GM_CONV_NHWC = 1, GM_CONV_S2D = 2, GM_HEAD = 3, GM_PCONV = 4 };

struct GradJob {
  const float* src;   // [S][rows_pad][ld]
  int S;              // number of partials
  long sstride;       // elements between partials
  int rows, cols, ld; // logical extent of one partial (rows = k' or j, cols = cout ...)
  int map;
  float scale;
  long dst_off;       // offset into flat grad
  // conv maps: k' -> (c, ky, kx)
  int C, kh, kw;
  int s2d;            // GM_CONV_S2D: space-to-depth factor (= stride of the first conv)
  // GM_HEAD: src [S][H][A+2]; writes w_pi (H,A) at dst_off, w_v (H) at dst_off2, fc bias (H) at dst_off3
  long dst_off2, dst_off3;
  int A;
  // GM_PCONV (pconv.cuh): r = k' = ((ty*T + tx)*P + plane)*64 + ch, cell channel cc = plane*64 + ch decoded as
  // (ci, py, px) [ci_major] or (py, px, ci); ky = ty*s2d + py, kx = tx*s2d + px
  int T, P, ci_major;
  int ss_off;         // first slot of this job's per-block sums of squares (finalize_grads_kernel ss_out), see below
};

// cell channel -> (ci, py, px) of a space-to-depth(s) cell over C input channels
ARL_DEVINL void pc_decode_channel(int cc, int C, int s, int ci_major, int& ci, int& py, int& px) {
  if (ci_major) {
    ci = cc / (s * s);
    int r = cc - ci * s * s;
    py = r / s; px = r - py * s;
  } else {
    int sub = cc / C;
    ci = cc - sub * C;
    py = sub / s; px = sub - py * s;
  }
}

// Partial sums of one element, taken by a TEAM of four lanes: lanes q*8 + e (q = 0..3) of a warp hold element e of the
// warp's group of 8 consecutive elements; lane q sums partials k = q, q+4, q+8, ... with 16 independent loads in flight,
// then two butterfly shuffles combine the four sub-sums ((s_q + s_q^1) + (s_q^2 + s_q^3): the same bits in every lane).
// One thread per element with a 148-long dependent chain (ncu: profiles/r2_update_stream.md, 18 us) was pure latency;
// the team form shortens the chain 4x and keeps 32-byte-sector loads.  `valid` = the lane's element exists (all 32
// lanes must call: the shuffles are warp-wide).  The summation order is part of the parity contract between the
// training paths (single GPU / stream / synchronous), which all come through here.
ARL_DEVINL float finalize_sum4(const GradJob& jb, int r, int c, int q, bool valid) {
  float a[16];
#pragma unroll
  for (int u = 0; u < 16; ++u) a[u] = 0.f;
  if (valid) {
    const float* s = jb.src + (long)r * jb.ld + c;
    const int S = jb.S;
    const long ss = jb.sstride;
    int k = q;
    for (; k + 60 < S; k += 64) {
#pragma unroll
      for (int u = 0; u < 16; ++u) a[u] += s[(long)(k + 4 * u) * ss];
    }
#pragma unroll
    for (int u = 0; u < 16; ++u)
      if (k + 4 * u < S) a[u] += s[(long)(k + 4 * u) * ss];
  }
#pragma unroll
  for (int w = 8; w >= 1; w >>= 1)
#pragma unroll
    for (int u = 0; u < w; ++u) a[u] += a[u + w];
  float t = a[0];
  t += __shfl_xor_sync(0xffffffffu, t, 8);
  t += __shfl_xor_sync(0xffffffffu, t, 16);
  return t * jb.scale;
}

// where element (r, c) of job jb lives in the flat vector (-1: a padding tap / channel without a parameter)
ARL_DEVINL long finalize_dst(const GradJob& jb, int r, int c) {
  if (jb.map == GM_LINEAR) return jb.dst_off + (long)r * jb.cols + c;
  if (jb.map == GM_CONV_NHWC) {
    // r = k' = (ky*kw + kx)*C + ci ; c = cout
    int ci = r % jb.C;
    int t = r / jb.C;
    int kx = t % jb.kw, ky = t / jb.kw;
    return jb.dst_off + (((long)c * jb.C + ci) * jb.kh + (jb.kh - 1 - ky)) * jb.kw + (jb.kw - 1 - kx);
  }
  if (jb.map == GM_CONV_S2D) {
    // first layer over the space-to-depth input: r = k' = (ty*2 + tx)*(C*s*s) + ci*s*s + dy*s + dx
    const int s2 = jb.s2d * jb.s2d, cs = jb.C * s2;
    int ch = r % cs;
    int t = r / cs;
    int tx = t % 2, ty = t / 2;
    int ci = ch / s2, dy = (ch % s2) / jb.s2d, dx = ch % jb.s2d;
    int ky = ty * jb.s2d + dy, kx = tx * jb.s2d + dx;
    return jb.dst_off + (((long)c * jb.C + ci) * jb.kh + (jb.kh - 1 - ky)) * jb.kw + (jb.kw - 1 - kx);
  }
  if (jb.map == GM_PCONV) {
    int blk = r >> 6, ch = r & 63;
    int t = blk / jb.P, plane = blk - t * jb.P;
    int ty = t / jb.T, tx = t - ty * jb.T;
    int ci, py, px;
    pc_decode_channel(plane * 64 + ch, jb.C, jb.s2d, jb.ci_major, ci, py, px);
    int ky = ty * jb.s2d + py, kx = tx * jb.s2d + px;
    if (!(ky < jb.kh && kx < jb.kw && ci < jb.C)) return -1;
    return jb.dst_off + (((long)c * jb.C + ci) * jb.kh + (jb.kh - 1 - ky)) * jb.kw + (jb.kw - 1 - kx);
  }
  // GM_HEAD: r = j, c in [0, A+2)
  if (c < jb.A) return jb.dst_off + (long)r * jb.A + c;
  if (c == jb.A) return jb.dst_off2 + r;
  return jb.dst_off3 + r;
}

// element v of job jb for the calling lane's team -> (gradient value in every lane, flat index or -1); the team's lane
// q == 0 is the one that should consume it.  v >= rows*cols: not an element (returns -1, still joins the shuffles)
ARL_DEVINL long finalize_team(const GradJob& jb, long v, int q, float& g_out) {
  const long total = (long)jb.rows * jb.cols;
  const bool valid = v < total;
  const int r = valid ? (int)(v / jb.cols) : 0;
  const int c = valid ? (int)(v - (long)r * jb.cols) : 0;
  g_out = finalize_sum4(jb, r, c, q, valid);
  return valid ? finalize_dst(jb, r, c) : -1;
}

constexpr int kFinPerBlock = 64;     // elements per 256-thread block: 8 warps x 8 teams of 4 lanes

__global__ void __launch_bounds__(256) finalize_grads_kernel(const GradJob* __restrict__ jobs, float* __restrict__ grad) {
  pdl_wait();
  pdl_trigger();
  const GradJob& jb = jobs[blockIdx.y];
  const long total = (long)jb.rows * jb.cols;
  const int lane = threadIdx.x & 31, q = lane >> 3;
  for (long v0 = (long)blockIdx.x * kFinPerBlock; v0 < total; v0 += (long)gridDim.x * kFinPerBlock) {
    float g;
    const long dst = finalize_team(jb, v0 + (threadIdx.x >> 5) * 8 + (lane & 7), q, g);
    if (q == 0 && dst >= 0) grad[dst] = g;
  }
}

// ===========================================================================
// Global-norm clip + update.  optimizers/util.py:70-76 (Lasagne total_norm_constraint, eps 1e-7),
// update rules: optimizers/update_methods_stats.py:11-32 (rmsprop), :55-87 (adam).
//   sumsq_kernel : fixed grid, per-block partial sums of g^2 (double), no atomics
//   update_kernel: every block re-reduces the partials (identical result in every block),
//                  applies avg -> clip -> Adam/RMSProp; block 0 records grad_norm and loss.
// ===========================================================================
constexpr int kSumsqBlocks = 592;   // 4 x 148
constexpr int kEarlyBlocks = 296;   // update_range_kernel grid: 2 x 148

ARL_DEVINL void sumsq_body(const float* __restrict__ g, long n, float gscale, double* __restrict__ partial,
                           long skip4_begin = 0, long skip4_len = 0);

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long n, float gscale,
                                                     double* __restrict__ partial) {
  sumsq_body(g, n, gscale, partial);
}

// float4 groups [skip4_begin, skip4_begin + skip4_len) are left out (their sum of squares was taken by
// update_range_kernel); the remaining groups are walked in index order
ARL_DEVINL void sumsq_body(const float* __restrict__ g, long n, float gscale, double* __restrict__ partial,
                           long skip4_begin, long skip4_len) {
  pdl_wait();
  pdl_trigger();
  double acc = 0.0;
  const long n4 = (n >> 2) - skip4_len;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long j = (long)blockIdx.x * blockDim.x + threadIdx.x; j < n4; j += (long)gridDim.x * blockDim.x) {
    const long i = j < skip4_begin ? j : j + skip4_len;
    float4 v = g4[i];
    float a = v.x * gscale, b = v.y * gscale, c = v.z * gscale, d = v.w * gscale;
    acc += (double)(a * a + b * b) + (double)(c * c + d * d);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long i = (n >> 2) << 2; i < n; ++i) { float a = g[i] * gscale; acc += (double)a * a; }
  __shared__ double s[8];
  acc = warp_sum_d(acc);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s[w];
    partial[blockIdx.x] = t;
  }
}

// FC weight element (reference row r = c*HW + hw, column j) -> element offset in wfc_t [HW][H/64][64 c][64 j]
// (8 KB tiles, 16-byte chunks XOR-swizzled by (c & 7)); see fcgemm.cuh
ARL_DEVINL long fc_tile_index(long r, int j, int HW, int H) {
  const unsigned ru = (unsigned)r;                       // 32-bit division: the 64-bit one costs ~100 instructions
  const unsigned c = ru / (unsigned)HW, hw = ru - c * (unsigned)HW;
  return (((long)hw * (H >> 6) + (j >> 6)) * 64 + c) * 64 + (((((j & 63) >> 3) ^ (c & 7)) << 3) | (j & 7));
}

enum PackKind { PK_CONV_NHWC = 0, PK_CONV_S2D = 1, PK_CONV_DGRAD = 2, PK_CAST = 3, PK_PCONV = 4, PK_PCONV_DGRAD = 5, PK_FC_TILES = 6 };

struct PackJob {
  __nv_bfloat16* dst;
  long src_off;        // offset of W in the flat params
  int kind;
  int rows, cols;      // dst is [rows][cols]
  int Cout, C, kh, kw; // conv dims
  int s, ry, rx, Tx;   // dgrad class (k' = (ty*Tx+tx)*Cout + o ; ky = ry + s*ty ; kx = rx + s*tx)
  int HW;              // FC: k' = hw*C + c  <- W[(c*HW + hw)][j]
  int ldsrc;           // FC: hidden size
  // PK_PCONV / PK_PCONV_DGRAD (pconv.cuh): dst = [blocks][N][64] with 16-byte chunks XOR-swizzled by (row & 7);
  // forward: block = (ty*T + tx)*P + plane, row = cout, column ch -> cell channel plane*64 + ch;
  // dgrad  : block = (ey*T + ex)*P + plane_out, row = input cell channel, column ch -> cout = plane_out*64 + ch,
  //          tap (ty, tx) = (T-1-ey, T-1-ex)
  int T, P, N, ci_major;
};

struct UpdateParams {
  float* param; const float* grad; float* m; float* v;
  long n;
  const double* sumsq_partial; int n_partial;
  const float* loss_partial; int n_loss_blocks;   // [blocks][4]
  const float* hyper;      // [0] = lr_mult
  int* step;               // Adam t (device counter, incremented by block 0)
  int kind;                // 0 = Adam, 1 = RMSProp
  float lr, beta1, beta2, eps, rho;
  float clip;              // <= 0: no clipping (norm still reported)
  float gscale;            // gradient averaging factor (1/n_gpu for sync DP)
  float* out_norm; float* out_loss;   // [cap] logs, slot = log_slot[0]
  int* log_slot; int log_cap;
  __nv_bfloat16* shadow; long shadow_begin, shadow_end;   // bf16 copy of params[shadow_begin, shadow_end) (4-aligned)
  int shadow_tiles, shadow_HW, shadow_H;                  // != 0: the copy is the tiled wfc_t layout (H % 4 == 0)
  // float4 groups [skip4_begin, skip4_begin + skip4_len) were already updated by update_range_kernel (no clipping:
  // the update does not depend on the norm); their sum of squares arrives in sumsq_partial2
  long skip4_begin, skip4_len;
  const double* sumsq_partial2; int n_partial2;
  // update_stream_kernel: fin_jobs = the split partials of every tensor except the FC weights (summed by the kernel's
  // part-A blocks), [fc4_begin, +fc4_len) = the FC-weight range in float4 groups, which fc_gemm_kernel wrote directly
  const GradJob* fin_jobs; int n_fin_jobs;
  long fc4_begin, fc4_len;
  // a2_resume > 0: elements [a2_begin, a2_mid) and [a2_resume, n) already sit in the flat gradient (finalised on the side
  // streams while the conv gradient chain ran): update_stream_kernel's part-A2 blocks update them one element per thread
  long a2_begin, a2_mid, a2_resume;
  uint32_t shadow_H_magic, shadow_HW_magic;   // floor(2^32/d) + 1: the tile index needs two divisions per float4 group
  // n_pk_jobs > 0: the conv operand packs (pconv.cuh tap tiles, forward and data-gradient variants) are refreshed by
  // the thread that updates the weight — a scatter through the inverse of pack_weights_kernel's index map — instead of
  // a separate pack launch; elements >= conv_end are never conv weights
  const unsigned long long* pk_slots; int n_pk_jobs; long conv_end;   // pk_slots: [conv_end][2] device addresses
  // adv_done != nullptr: the last block to finish advances the device-side counters (what pack_weights_kernel did)
  unsigned long long* adv_done; int* adv_log_slot; int* adv_mb;
};

// inverse of pack_job_body's PK_PCONV / PK_PCONV_DGRAD map, evaluated ONCE on the host (pack_slot_of, api.cu) into a
// table of device addresses: element e of the flat vector -> the (at most two) bf16 slots that hold it, 0 = none.
// (Computing the inverse in the kernel was measured: the ~20 K float4 groups of conv weights sit in the first 79 blocks
// and their integer divisions made the update kernel 10 us slower than the pack launch it replaced.)
__host__ __device__ inline long pack_slot_of(const PackJob& jb, long e) {
  const long rel = e - jb.src_off;
  if (rel < 0 || rel >= (long)jb.Cout * jb.C * jb.kh * jb.kw) return -1;
  // W[co][ci][kyf][kxf], filters stored flipped: correlation tap (ky, kx) = (kh-1-kyf, kw-1-kxf)
  const int kxf = (int)(rel % jb.kw);
  long t = rel / jb.kw;
  const int kyf = (int)(t % jb.kh);
  t /= jb.kh;
  const int ci = (int)(t % jb.C), co = (int)(t / jb.C);
  const int ky = jb.kh - 1 - kyf, kx = jb.kw - 1 - kxf;
  const int ty = ky / jb.s, py = ky - ty * jb.s, tx = kx / jb.s, px = kx - tx * jb.s;
  const int cc = jb.ci_major ? (ci * jb.s * jb.s + py * jb.s + px) : ((py * jb.s + px) * jb.C + ci);   // pc_decode_channel^-1
  int pl, k, n_, blk;
  if (jb.kind == PK_PCONV) { pl = cc >> 6; k = cc & 63; n_ = co; blk = (ty * jb.T + tx) * jb.P + pl; }
  else { pl = co >> 6; k = co & 63; n_ = cc; blk = ((jb.T - 1 - ty) * jb.T + (jb.T - 1 - tx)) * jb.P + pl; }
  const long r = (long)blk * jb.N + n_;
  if (pl >= jb.P || n_ >= jb.N || r >= jb.rows) return -1;
  return r * 64 + ((((k >> 3) ^ (n_ & 7)) << 3) | (k & 7));
}

// float4 group i of the (just updated) parameters -> its bf16 slots; 8 addresses = 64 contiguous bytes of the table
ARL_DEVINL void scatter_conv_pack4(const UpdateParams& p, long i) {
  const float4 w = reinterpret_cast<const float4*>(p.param)[i];
  const float pp[4] = {w.x, w.y, w.z, w.w};
  const ulonglong2* tab = reinterpret_cast<const ulonglong2*>(p.pk_slots + 8 * i);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const ulonglong2 a = __ldg(tab + k);
    const __nv_bfloat16 v = __float2bfloat16_rn(pp[k]);
    if (a.x) *reinterpret_cast<__nv_bfloat16*>(a.x) = v;
    if (a.y) *reinterpret_cast<__nv_bfloat16*>(a.y) = v;
  }
}

ARL_DEVINL void update_body(const UpdateParams& p);

__global__ void __launch_bounds__(256) update_kernel(UpdateParams p) {
  pdl_wait();
  pdl_trigger();
  update_body(p);
}

// sum of squares + clip + update in ONE launch: phase 1 = sumsq_kernel's arithmetic (same grid, same partials, same
// order), a grid-wide ticket barrier (all kSumsqBlocks blocks are co-resident: 4 per SM), phase 2 = update_kernel's.
// The ticket counter only grows (64-bit), so the launch carries no host state and replays from a CUDA graph.
__global__ void __launch_bounds__(256, 4) update_fused_kernel(UpdateParams p, double* __restrict__ partial,
                                                              unsigned long long* __restrict__ ticket) {
  pdl_wait();
  pdl_trigger();
  sumsq_body(p.grad, p.n, p.gscale, partial, p.skip4_begin, p.skip4_len);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long t = atomicAdd(ticket, 1ULL);
    const unsigned long long target = (t / gridDim.x + 1ULL) * gridDim.x;
    long long t0 = clock64();
    while (atomicAdd(ticket, 0ULL) < target) {
      __nanosleep(40);
      if (clock64() - t0 > 20000000000LL) dev_fail(320);
    }
    __threadfence();
  }
  __syncthreads();
  update_body(p);
  if (p.adv_done) {
    // every block read the update count / log slot in update_body before it gets here: the LAST block to arrive may
    // advance them (the 64-bit arrival counter only grows, so the launch replays from a CUDA graph)
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      const unsigned long long t = atomicAdd(p.adv_done, 1ULL);
      if ((t + 1ULL) % gridDim.x == 0ULL) { p.step[0] += 1; p.adv_log_slot[0] += 1; p.adv_mb[0] += 1; }
    }
  }
}

// One Adam / RMSProp element step (optimizers/update_methods_stats.py:11-32, :55-87).  The division and the square root
// use the approximate hardware forms (MUFU.RCP / MUFU.SQRT / MUFU.RSQ, <= 2 ulp each): the IEEE sequences cost ~40
// instructions per element and made the update kernel issue-bound (ncu: ~400 warp instructions per float4 group); the
// error they add to a parameter is < 3e-7 of ONE step's change, far below the fp32 spacing of the parameter itself.
ARL_DEVINL float fast_sqrt(float x) { float r; asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
ARL_DEVINL float fast_rsqrt(float x) { float r; asm("rsqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
// Every operation is an explicit intrinsic: the compiler may not re-contract them, so every kernel that inlines this
// (single-GPU, early-FC range, stream, synchronous slice, asynchronous central update) rounds identically — the
// bit-exactness tests between those paths rely on it.
ARL_DEVINL void opt_step_raw(int kind, float beta1, float beta2, float eps, float rho, float& pp, float& mm, float& vv,
                             float g, float alpha) {
  if (kind == 0) {
    mm = __fmaf_rn(beta1, mm, __fmul_rn(1.f - beta1, g));
    vv = __fmaf_rn(beta2, vv, __fmul_rn(__fmul_rn(1.f - beta2, g), g));
    pp = __fsub_rn(pp, __fdividef(__fmul_rn(alpha, mm), __fadd_rn(fast_sqrt(vv), eps)));
  } else {
    vv = __fmaf_rn(rho, vv, __fmul_rn(__fmul_rn(1.f - rho, g), g));
    pp = __fsub_rn(pp, __fmul_rn(__fmul_rn(alpha, g), fast_rsqrt(__fadd_rn(vv, eps))));
  }
}
ARL_DEVINL void opt_step1(const UpdateParams& p, float& pp, float& mm, float& vv, float g, float alpha) {
  opt_step_raw(p.kind, p.beta1, p.beta2, p.eps, p.rho, pp, mm, vv, g, alpha);
}

// Adam / RMSProp on float4 group i (+ the bf16 operand copy of the FC weights, refreshed in the same pass)
ARL_DEVINL void update_vec4_g(const UpdateParams& p, long i, float4 g4, float scale, float alpha);
ARL_DEVINL void update_vec4(const UpdateParams& p, long i, float scale, float alpha) {
  update_vec4_g(p, i, reinterpret_cast<const float4*>(p.grad)[i], scale, alpha);
}
ARL_DEVINL void update_vec4_g(const UpdateParams& p, long i, float4 g4, float scale, float alpha) {
  float4 p4 = reinterpret_cast<float4*>(p.param)[i];
  float4 v4 = reinterpret_cast<float4*>(p.v)[i];
  float g[4] = {__fmul_rn(g4.x, scale), __fmul_rn(g4.y, scale), __fmul_rn(g4.z, scale), __fmul_rn(g4.w, scale)};
  float pp[4] = {p4.x, p4.y, p4.z, p4.w};
  float vv[4] = {v4.x, v4.y, v4.z, v4.w};
  if (p.kind == 0) {
    float4 m4 = reinterpret_cast<float4*>(p.m)[i];
    float mm[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) opt_step1(p, pp[k], mm[k], vv[k], g[k], alpha);
    reinterpret_cast<float4*>(p.m)[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
  } else {
    float dummy = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) opt_step1(p, pp[k], dummy, vv[k], g[k], alpha);
  }
  reinterpret_cast<float4*>(p.v)[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
  reinterpret_cast<float4*>(p.param)[i] = make_float4(pp[0], pp[1], pp[2], pp[3]);
  const long e0 = i << 2;
  if (p.shadow && e0 >= p.shadow_begin && e0 + 4 <= p.shadow_end) {
    long off = e0 - p.shadow_begin;
    if (p.shadow_tiles) {
      // row r = off / H, column j = off % H, r = c*HW + hw  (magic-number division: exact for off < 2^32 / H)
      const unsigned ou = (unsigned)off, rr = __umulhi(ou, p.shadow_H_magic);
      const int j = (int)(ou - rr * (unsigned)p.shadow_H);
      const unsigned c = __umulhi(rr, p.shadow_HW_magic), hw = rr - c * (unsigned)p.shadow_HW;
      off = (((long)hw * (p.shadow_H >> 6) + (j >> 6)) * 64 + c) * 64 + (((((j & 63) >> 3) ^ (c & 7)) << 3) | (j & 7));
    }
    *reinterpret_cast<uint2*>(p.shadow + off) = make_uint2(pack_bf16x2(pp[0], pp[1]), pack_bf16x2(pp[2], pp[3]));
  }
}

// step size of this update: lr * lr_mult, with Adam's bias correction for update count step + 1
ARL_DEVINL float update_alpha(const UpdateParams& p) {
  const int tstep = p.step[0] + 1;
  const float lr = p.lr * p.hyper[0];
  if (p.kind != 0) return lr;
  const double b1t = pow((double)p.beta1, (double)tstep), b2t = pow((double)p.beta2, (double)tstep);
  return (float)((double)lr * sqrt(1.0 - b2t) / (1.0 - b1t));
}

// Early update of one parameter range (the FC weights: 98 % of the vector) as soon as its gradient is final, while the
// conv data/weight-gradient chain is still running.  Only legal without global-norm clipping (PPO's default,
// algos/pg/ppo.py:29): the update then depends on nothing but its own gradient; the norm is still reported, from the
// per-block sums of squares left in `partial` for update_fused_kernel.  Does not advance the update count.
__global__ void __launch_bounds__(256) update_range_kernel(UpdateParams p, long begin4, long len4, double* __restrict__ partial) {
  pdl_wait();
  pdl_trigger();
  __shared__ float s_alpha;
  __shared__ double s[8];
  if (threadIdx.x == 0) s_alpha = update_alpha(p);
  __syncthreads();
  const float alpha = s_alpha, scale = p.gscale;
  double acc = 0.0;
  for (long j = (long)blockIdx.x * blockDim.x + threadIdx.x; j < len4; j += (long)gridDim.x * blockDim.x) {
    const long i = begin4 + j;
    const float4 v = reinterpret_cast<const float4*>(p.grad)[i];
    const float a = v.x * scale, b = v.y * scale, c = v.z * scale, d = v.w * scale;
    acc += (double)(a * a + b * b) + (double)(c * c + d * d);
    update_vec4(p, i, scale, alpha);
  }
  acc = warp_sum_d(acc);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s[w];
    partial[blockIdx.x] = t;
  }
}

ARL_DEVINL void update_body(const UpdateParams& p) {
  __shared__ double s_red[8];
  __shared__ float s_scale, s_alpha;
  // every block: reduce the partial sums in the same order -> identical norm everywhere
  double acc = 0.0;
  for (int i = threadIdx.x; i < p.n_partial; i += blockDim.x) acc += p.sumsq_partial[i];
  for (int i = threadIdx.x; i < p.n_partial2; i += blockDim.x) acc += p.sumsq_partial2[i];
  acc = warp_sum_d(acc);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s_red[w];
    float norm = (float)sqrt(t);
    float scale = p.gscale;
    if (p.clip > 0.f) scale *= fminf(norm, p.clip) / (1e-7f + norm);
    s_scale = scale;
    s_alpha = update_alpha(p);
    if (blockIdx.x == 0) {
      int slot = p.log_slot[0];
      if (slot < p.log_cap) p.out_norm[slot] = norm;
    }
  }
  __syncthreads();
  if (blockIdx.x == 0) {
    // loss of this update = sum of the per-row terms (fixed order: strided per thread, then tree)
    float l = 0.f;
    for (int b = threadIdx.x; b < p.n_loss_blocks; b += blockDim.x) l += p.loss_partial[4 * b + 3];
    l = warp_sum(l);
    __shared__ float s_l[8];
    if ((threadIdx.x & 31) == 0) s_l[threadIdx.x >> 5] = l;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tl = 0.f;
      for (int w = 0; w < 8; ++w) tl += s_l[w];
      int slot = p.log_slot[0];
      if (slot < p.log_cap) p.out_loss[slot] = tl;
    }
  }
  const float scale = s_scale, alpha = s_alpha;
  const long n4 = (p.n >> 2) - p.skip4_len;
  for (long j = (long)blockIdx.x * blockDim.x + threadIdx.x; j < n4; j += (long)gridDim.x * blockDim.x)
    update_vec4(p, j < p.skip4_begin ? j : j + p.skip4_len, scale, alpha);
  if (p.n_pk_jobs > 0) {
    // conv operand packs: every thread re-reads the conv-weight groups IT just updated (same index walk, own writes)
    // and scatters them through the slot table — a separate loop, so the main loop's registers are not squeezed
    for (long j = (long)blockIdx.x * blockDim.x + threadIdx.x; j < n4; j += (long)gridDim.x * blockDim.x) {
      const long i = j < p.skip4_begin ? j : j + p.skip4_len;
      if ((i << 2) + 4 > p.conv_end) break;          // i grows with j: nothing further is a conv weight
      scatter_conv_pack4(p, i);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (p.n & 3)) {
    long i = ((p.n >> 2) << 2) + threadIdx.x;
    float g = __fmul_rn(p.grad[i], scale);
    float pv = p.param[i], m = (p.kind == 0) ? p.m[i] : 0.f, v = p.v[i];
    opt_step1(p, pv, m, v, g, alpha);
    if (p.kind == 0) p.m[i] = m;
    p.v[i] = v;
    p.param[i] = pv;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// update_stream_kernel: gradient finalisation + Adam/RMSProp + operand refresh + logs in ONE plain launch, for updates
// WITHOUT global-norm clipping (PPO's default, algos/pg/ppo.py:29: the norm is only reported, so nothing in the update
// waits for it).  Replaces finalize_grads_kernel -> update_fused_kernel {pass over the gradient, grid barrier, update}:
//   part A  every tensor except the FC weights (~2 % of the vector): the thread that sums an element's split partials
//           (finalize_team) also takes its optimiser step and refreshes its bf16 operand slots;
//   part B  the FC weight range, float4 groups straight from the flat gradient the FC tiles wrote;
//   every thread squares what it consumed; block partials in `partial`; the LAST block to finish (ticket) adds them up
//   in index order (bit-reproducible), writes the norm / loss logs and advances the device counters.
// No grid barrier, no cooperative launch, one pass over the 101 MB of (g, p, m, v).
// ---------------------------------------------------------------------------------------------------------------
ARL_DEVINL void update_scalar(const UpdateParams& p, long i, float g, float alpha) {
  float pv = p.param[i], m = (p.kind == 0) ? p.m[i] : 0.f, v = p.v[i];
  opt_step1(p, pv, m, v, g, alpha);
  if (p.kind == 0) p.m[i] = m;
  p.v[i] = v;
  p.param[i] = pv;
  if (p.n_pk_jobs > 0 && i < p.conv_end) {
    const ulonglong2 a = __ldg(reinterpret_cast<const ulonglong2*>(p.pk_slots + 2 * i));
    const __nv_bfloat16 b = __float2bfloat16_rn(pv);
    if (a.x) *reinterpret_cast<__nv_bfloat16*>(a.x) = b;
    if (a.y) *reinterpret_cast<__nv_bfloat16*>(a.y) = b;
  }
}

// Split form (api.cu clip_update, ARL_SPLIT_UPDATE): the same block index space launched as TWO grids — parts A/A2 on the
// main stream (blk0 = 0), part B on a forked stream (blk0 = nA + nA2) so that the FC range, 98 % of the bytes, streams
// beside the NEXT minibatch's conv layers, which do not read it.  grid_total = blocks of both launches (the ticket of the
// logs); mb_ticket != nullptr: the minibatch cursor — what the next forward pass reads — is advanced by the last block of
// the launch that carries it (the main-stream one), not by the overall last block.
__global__ void __launch_bounds__(256, 4) update_stream_kernel(UpdateParams p, double* __restrict__ partial, int nA,
                                                              int nA2, long total_all, int blk0, int grid_total,
                                                              unsigned long long* mb_ticket, int advance_mb) {
  pdl_wait();
  pdl_trigger();
  __shared__ float s_alpha;
  __shared__ double s_red[8];
  __shared__ int s_last;
  if (threadIdx.x == 0) s_alpha = update_alpha(p);
  __syncthreads();
  const float alpha = s_alpha;
  double acc = 0.0;
  const int vb = (int)blockIdx.x + blk0;          // block index in the joint index space
  if (vb < nA) {
    // part A (blocks [0, nA): scheduled first): 64 elements per block, a team of four lanes per element (finalize_sum4),
    // elements numbered through the concatenated job index space.  A separate set of blocks from part B: the partial sums are
    // 48..148 dependent-latency loads per element — ncu (profiles/r2_update_stream.md) showed warps that carried both
    // parts holding their whole block at the final barrier for 40 % of the kernel
    float* grad = const_cast<float*>(p.grad);
    const int lane = threadIdx.x & 31, q = lane >> 3;
    long v = (long)vb * kFinPerBlock + (threadIdx.x >> 5) * 8 + (lane & 7);
    int jn = 0;
    if (v < total_all) {
      for (; jn < p.n_fin_jobs; ++jn) {
        const long t = (long)p.fin_jobs[jn].rows * p.fin_jobs[jn].cols;
        if (v < t) break;
        v -= t;
      }
    } else {
      v = 0x7fffffffffffL;                     // no element: the lane only joins the team shuffles
    }
    const GradJob& jb = p.fin_jobs[jn < p.n_fin_jobs ? jn : 0];      // (fields fetched where used: a register copy spills)
    float g;
    const long dst = finalize_team(jb, v, q, g);
    if (q == 0 && dst >= 0) {
      grad[dst] = g;
      acc += (double)(g * g);
      update_scalar(p, dst, g, alpha);
    }
  } else if (vb < nA + nA2) {
    // part A2: small tensors whose gradient is already in the flat vector
    const long j = (long)(vb - nA) * blockDim.x + threadIdx.x;
    const long len0 = p.a2_mid - p.a2_begin;
    const long i = j < len0 ? p.a2_begin + j : p.a2_resume + (j - len0);
    if (i < p.n) {
      const float g = p.grad[i];
      acc += (double)(g * g);
      update_scalar(p, i, g, alpha);
    }
  } else {
    // part B: the FC weights
    const int nAA = nA + nA2;
    const long gtid = (long)(vb - nAA) * blockDim.x + threadIdx.x, gsize = (long)(grid_total - nAA) * blockDim.x;
    const float4* g4p = reinterpret_cast<const float4*>(p.grad) + p.fc4_begin;
    for (long j = gtid; j < p.fc4_len; j += gsize) {
      const float4 g4 = g4p[j];
      acc += (double)(g4.x * g4.x + g4.y * g4.y) + (double)(g4.z * g4.z + g4.w * g4.w);
      update_vec4_g(p, p.fc4_begin + j, g4, 1.f, alpha);
    }
  }
  acc = warp_sum_d(acc);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s_red[w];
    partial[vb] = t;
    __threadfence();
    if (mb_ticket) {
      // (the cursor first: a block that is last in both senses must not log before it advanced it — harmless either way)
      const unsigned long long tm = atomicAdd(mb_ticket, 1ULL);
      if ((tm + 1ULL) % gridDim.x == 0ULL) p.adv_mb[0] += 1;
    }
    const unsigned long long tk = atomicAdd(p.adv_done, 1ULL);
    s_last = ((tk + 1ULL) % (unsigned long long)grid_total == 0ULL) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  // last block: every other block has published its partial and read the update count / log slot
  __threadfence();
  double a2 = 0.0;
  for (int i = threadIdx.x; i < grid_total; i += blockDim.x) a2 += __ldcg(partial + i);
  for (int i = threadIdx.x; i < p.n_partial2; i += blockDim.x) a2 += __ldcg(p.sumsq_partial2 + i);   // early FC update's share
  a2 = warp_sum_d(a2);
  float l = 0.f;
  for (int b = threadIdx.x; b < p.n_loss_blocks; b += blockDim.x) l += p.loss_partial[4 * b + 3];
  l = warp_sum(l);
  __shared__ float s_l[8];
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { s_red[threadIdx.x >> 5] = a2; s_l[threadIdx.x >> 5] = l; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    float tl = 0.f;
    for (int w = 0; w < 8; ++w) { t += s_red[w]; tl += s_l[w]; }
    const int slot = p.log_slot[0];
    if (slot < p.log_cap) { p.out_norm[slot] = (float)sqrt(t); p.out_loss[slot] = tl; }
    p.step[0] += 1; p.adv_log_slot[0] += 1;
    if (advance_mb) p.adv_mb[0] += 1;
  }
}

// runs after update_kernel (stream order): advance the device-side counters
__global__ void advance_counters_kernel(int* step, int* log_slot, int* mb_counter) {
  step[0] += 1;
  log_slot[0] += 1;
  mb_counter[0] += 1;
}

// ===========================================================================
// Weight packing: fp32 master (reference layout) -> bf16 operand matrices for the GEMM tiles
// ===========================================================================
ARL_DEVINL void pack_job_body(const PackJob& jb, const float* __restrict__ params, long first, long stride);

__global__ void __launch_bounds__(256) pack_weights_kernel(const PackJob* __restrict__ jobs,
                                                            const float* __restrict__ params, int* step,
                                                            int* log_slot, int* mb_counter) {
  pdl_wait();
  pdl_trigger();
  // last kernel of an update: also advances the device-side counters (Adam t, log slot, minibatch index)
  if (step && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
    step[0] += 1; log_slot[0] += 1; mb_counter[0] += 1;
  }
  const PackJob jb = jobs[blockIdx.y];
  pack_job_body(jb, params, (long)blockIdx.x * blockDim.x + threadIdx.x, (long)gridDim.x * blockDim.x);
}

// parameter reads bypass L1 (__ldcg): inside step_fused_kernel the values were written by other SMs earlier in the
// same launch
ARL_DEVINL void pack_job_body(const PackJob& jb, const float* __restrict__ params, long first, long stride) {
  const long total = (long)jb.rows * jb.cols;
  const float* W = params + jb.src_off;
  for (long i = first; i < total; i += stride) {
    int r = (int)(i / jb.cols);
    int k = (int)(i - (long)r * jb.cols);
    float v = 0.f;
    if (jb.kind == PK_CONV_NHWC) {          // r = cout, k = (ky*kw+kx)*C + c
      int c = k % jb.C; int t = k / jb.C; int kx = t % jb.kw, ky = t / jb.kw;
      v = __ldcg(W + ((((long)r * jb.C + c) * jb.kh + (jb.kh - 1 - ky)) * jb.kw + (jb.kw - 1 - kx)));
    } else if (jb.kind == PK_CONV_S2D) {    // r = cout, k = (ty*2+tx)*(C*s*s) + c*s*s + dy*s + dx
      const int s2 = jb.s * jb.s, cs = jb.C * s2;
      int ch = k % cs; int t = k / cs; int tx = t % 2, ty = t / 2;
      int c = ch / s2, dy = (ch % s2) / jb.s, dx = ch % jb.s;
      int ky = ty * jb.s + dy, kx = tx * jb.s + dx;
      v = __ldcg(W + ((((long)r * jb.C + c) * jb.kh + (jb.kh - 1 - ky)) * jb.kw + (jb.kw - 1 - kx)));
    } else if (jb.kind == PK_CONV_DGRAD) {  // r = cin, k = (ty*Tx+tx)*Cout + o
      int o = k % jb.Cout; int t = k / jb.Cout; int tx = t % jb.Tx, ty = t / jb.Tx;
      int ky = jb.ry + jb.s * ty, kx = jb.rx + jb.s * tx;
      if (ky < jb.kh && kx < jb.kw)
        v = __ldcg(W + ((((long)o * jb.C + r) * jb.kh + (jb.kh - 1 - ky)) * jb.kw + (jb.kw - 1 - kx)));
    } else if (jb.kind == PK_PCONV || jb.kind == PK_PCONV_DGRAD) {
      // rows = blocks*N, cols = 64
      int blk = r / jb.N, n_ = r - blk * jb.N;
      int t = blk / jb.P, pl = blk - t * jb.P;
      int ty = t / jb.T, tx = t - ty * jb.T;
      int co, cc;
      if (jb.kind == PK_PCONV) { co = n_; cc = pl * 64 + k; }
      else { co = pl * 64 + k; cc = n_; ty = jb.T - 1 - ty; tx = jb.T - 1 - tx; }
      int ci, py, px;
      pc_decode_channel(cc, jb.C, jb.s, jb.ci_major, ci, py, px);
      int ky = ty * jb.s + py, kx = tx * jb.s + px;
      if (ky < jb.kh && kx < jb.kw && ci < jb.C && co < jb.Cout)
        v = __ldcg(W + ((((long)co * jb.C + ci) * jb.kh + (jb.kh - 1 - ky)) * jb.kw + (jb.kw - 1 - kx)));
      jb.dst[(long)r * 64 + ((((k >> 3) ^ (n_ & 7)) << 3) | (k & 7))] = __float2bfloat16_rn(v);
      continue;
    } else if (jb.kind == PK_FC_TILES) {     // FC weights -> wfc_t tiles (rows = Kfc in reference order, cols = H)
      jb.dst[fc_tile_index(r, k, jb.HW, jb.cols)] = __float2bfloat16_rn(W[i]);
      continue;
    } else {                                 // PK_CAST: same layout, fp32 -> bf16 (FC weights, reference order)
      v = __ldcg(W + (i));
    }
    jb.dst[i] = __float2bfloat16_rn(v);
  }
}

// Grid-wide ticket barrier for a co-resident (cooperatively launched) grid.  The 64-bit ticket only grows, so a launch
// carries no host state and replays from a CUDA graph.
ARL_DEVINL void ticket_barrier(unsigned long long* ticket, int code) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long t = atomicAdd(ticket, 1ULL);
    const unsigned long long target = (t / gridDim.x + 1ULL) * gridDim.x;
    long long t0 = clock64();
    while (atomicAdd(ticket, 0ULL) < target) {
      if (clock64() - t0 > 20000000000LL) dev_fail(code);
    }
    __threadfence();
  }
  __syncthreads();
}

// (a variant that also folded the conv operand re-pack behind a second barrier was measured: 30.0 us against
// 22.4 + 5 us for update_fused_kernel + pack_weights_kernel — no gain, dropped)

// ===========================================================================
// GAE / discounted returns — algos/pg/util.py:6-37, aac_base.py:108-145.
// Both recurrences are affine: x_t = a_t + g_t * x_{t+1}, g_t = coef * (1 - done_t):
//   GAE      : a_t = r_t + gamma*V_{t+1}*(1-d_t) - V_t, coef = gamma*lambda, x = advantage
//   lambda=1 : a_t = r_t,                                coef = gamma,        x = return
// One warp per env: each lane composes its contiguous chunk of steps, a reverse warp-shuffle
// scan combines the chunks, then each lane replays its chunk from the incoming value.
// With valids (mid_batch_reset == False): valids[t] = t < t_inv, adv/ret/value zeroed after
// (update_valids / zero_after_reset, util.py:40-63).
// ===========================================================================
__global__ void __launch_bounds__(128) gae_kernel(const float* __restrict__ rewards, float* __restrict__ values,
                                                   const uint8_t* __restrict__ dones,
                                                   const uint8_t* __restrict__ need_reset,
                                                   const float* __restrict__ last_values, float gamma, float lambda,
                                                   int use_gae, float* __restrict__ adv, float* __restrict__ ret,
                                                   int8_t* __restrict__ valids, int n_envs, int T) {
  const int lane = threadIdx.x & 31;
  const int env = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (env >= n_envs) return;
  const int chunk = (T + 31) / 32;
  const int t0 = lane * chunk;
  const int t1 = min(T, t0 + chunk);
  const long base = (long)env * T;
  const float coef = use_gae ? gamma * lambda : gamma;
  // compose chunk map x_{t0} = A + G * x_{t1}
  float A = 0.f, G = 1.f;
  for (int t = t1 - 1; t >= t0; --t) {
    float nd = 1.f - (float)dones[base + t];
    float a;
    if (use_gae) {
      float vnext = (t + 1 < T) ? values[base + t + 1] : last_values[env];
      a = rewards[base + t] + gamma * vnext * nd - values[base + t];
    } else {
      a = rewards[base + t];
    }
    float g = coef * nd;
    A = a + g * A;
    G = g * G;
  }
  // reverse inclusive scan over lanes: after it, (A, G) maps x at the END of the last lane to x_{t0}
  for (int o = 1; o < 32; o <<= 1) {
    float A2 = __shfl_down_sync(0xffffffffu, A, o);
    float G2 = __shfl_down_sync(0xffffffffu, G, o);
    if (lane + o < 32) { A = A + G * A2; G = G * G2; }
  }
  // value entering this lane's chunk from the right = x_{t1} = result of lane+1's composed map
  const float x_end = use_gae ? 0.f : last_values[env];
  float xin = __shfl_down_sync(0xffffffffu, A + G * x_end, 1);
  if (lane == 31) xin = x_end;
  // first need_reset index (for valids)
  int t_inv = T;
  if (valids) {
    int first = T;
    for (int t = t0; t < t1; ++t)
      if (need_reset[base + t]) { first = t; break; }
    for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
    t_inv = (first < T) ? first + 1 : T;
  }
  float x = xin;
  for (int t = t1 - 1; t >= t0; --t) {
    float nd = 1.f - (float)dones[base + t];
    float v = values[base + t];
    float a;
    if (use_gae) {
      float vnext = (t + 1 < T) ? values[base + t + 1] : last_values[env];
      a = rewards[base + t] + gamma * vnext * nd - v;
    } else {
      a = rewards[base + t];
    }
    x = a + coef * nd * x;
    float ad = use_gae ? x : x - v;
    float rt = use_gae ? x + v : x;
    if (valids) {
      bool ok = t < t_inv;
      valids[base + t] = ok ? 1 : 0;
      if (!ok) { ad = 0.f; rt = 0.f; }
    }
    adv[base + t] = ad;
    ret[base + t] = rt;
  }
  if (valids) {
    __syncwarp();
    for (int t = max(t0, t_inv); t < t1; ++t) values[base + t] = 0.f;   // zero_after_reset on values
  }
}

// standardize_adv (aac_base.py:136-143): (adv - mean) / (std + 1e-6), population std, optional valids.
// Single block (N is at most a few hundred thousand): two passes in double.
__global__ void __launch_bounds__(1024) standardize_adv_kernel(float* __restrict__ adv, const int8_t* __restrict__ valids,
                                                                long n) {
  __shared__ double s_a[32], s_b[32];
  __shared__ double s_mean, s_cnt, s_std;
  double sum = 0.0, cnt = 0.0;
  for (long i = threadIdx.x; i < n; i += blockDim.x)
    if (!valids || valids[i]) { sum += adv[i]; cnt += 1.0; }
  sum = warp_sum_d(sum); cnt = warp_sum_d(cnt);
  if ((threadIdx.x & 31) == 0) { s_a[threadIdx.x >> 5] = sum; s_b[threadIdx.x >> 5] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int w = 0; w < 32; ++w) { a += s_a[w]; b += s_b[w]; }
    s_cnt = b; s_mean = a / b;
  }
  __syncthreads();
  const double mean = s_mean;
  double var = 0.0;
  for (long i = threadIdx.x; i < n; i += blockDim.x)
    if (!valids || valids[i]) { double d = adv[i] - mean; var += d * d; }
  var = warp_sum_d(var);
  if ((threadIdx.x & 31) == 0) s_a[threadIdx.x >> 5] = var;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0;
    for (int w = 0; w < 32; ++w) a += s_a[w];
    s_std = sqrt(a / s_cnt);
  }
  __syncthreads();
  const float fm = (float)mean, fs = (float)s_std + 1e-6f;
  for (long i = threadIdx.x; i < n; i += blockDim.x)
    if (!valids || valids[i]) adv[i] = (adv[i] - fm) / fs;
}

// sum of valids -> device scalar (float)
__global__ void __launch_bounds__(1024) count_valids_kernel(const int8_t* __restrict__ valids, long n, float* out) {
  __shared__ int s[32];
  int c = 0;
  for (long i = threadIdx.x; i < n; i += blockDim.x) c += valids[i] ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 32; ++w) t += s[w];
    out[0] = (float)t;
  }
}

// sum of valids over the rows of one minibatch (valids_mean denominators are per minibatch, util.py:49-53)
__global__ void __launch_bounds__(1024) count_valids_idx_kernel(const int8_t* __restrict__ valids,
                                                                const int* __restrict__ idx,
                                                                const int* __restrict__ idx_off, int n, float* out) {
  pdl_wait();
  pdl_trigger();
  __shared__ int s[32];
  const int* ip = idx;
  if (ip && idx_off) ip += (long)idx_off[0] * n;
  int c = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) c += valids[ip ? ip[i] : i] ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 32; ++w) t += s[w];
    out[0] = (float)t;
  }
}

// standalone action sampling (policy.get_actions on host-provided probabilities)
__global__ void sample_actions_kernel(const float* __restrict__ prob, const double* __restrict__ u,
                                      uint8_t* __restrict__ act, int n, int A) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float cs = 0.f;
  int k = 0;
  for (int a = 0; a < A; ++a) { cs = __fadd_rn(cs, prob[(long)i * A + a]); k += ((double)cs < u[i]) ? 1 : 0; }
  act[i] = (uint8_t)min(k, A - 1);
}

}  // namespace arl
