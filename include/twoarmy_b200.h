/*
 * twoarmy_b200.h -- C ABI of the B200 (sm_100a) batched Twoarmy hot path.
 *
 * This is the drop-in boundary: every entry point names the reference interface it
 * replaces (paths relative to the reference repository root).  The reference is pure
 * Python with no FFI of its own; the binding a maintainer would add is the ctypes stub in
 * INTEGRATION.md (and the one this repo ships in
 * goal-conditioned-reinforcement-learning-with-environmental-and-policy-priors_b200/_capi.py).
 *
 * Conventions
 *   - plain C: pointers and sizes only, no torch / C++ types.
 *   - unless a function name ends in _host, every array argument is a DEVICE pointer owned
 *     by the caller; the library never frees it.  obs pointers must be 16-byte aligned.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Kernels
 *     are enqueued on it and the call returns without synchronising.
 *   - a handle owns the per-env state (HBM, struct-of-arrays), is bound to one device and
 *     is not thread-safe: one handle per GPU / process.
 *   - return value: 0 on success, a negative TA_E_* code otherwise; nothing throws.
 *   - there is no CPU fallback: without a CUDA device ta_create fails with TA_E_CUDA.
 */
#ifndef TWOARMY_B200_H
#define TWOARMY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TA_ABI_VERSION 1

/* error codes */
#define TA_OK 0
#define TA_E_INVALID (-1)   /* bad argument (version, view, n_envs, null pointer, alignment) */
#define TA_E_CUDA (-2)      /* a CUDA runtime call failed; ta_last_cuda_error() has the text */
#define TA_E_NOMEM (-3)
#define TA_E_UNSUPPORTED (-4)

/* per-env sticky error bits (reference raises a Python exception in these situations) */
#define TA_ENV_ERR_BAD_ACTION 1 /* action 4,5 or <0: AttributeError at gym_minigrid/minigrid.py:1397 */
#define TA_ENV_ERR_NONE_POS 2   /* patrol set but patrol balls unplaced (reset() mid-episode in v4):
                                   TypeError at gym_minigrid/envs/twoarmy_v4.py:122/:243 */
#define TA_ENV_ERR_OOB_MOVE 4   /* Grid.get assertion, gym_minigrid/minigrid.py:604-606 */

/* ta_step flags */
#define TA_STEP_AUTORESET 1 /* after a done step run reset() on that env (what the reference's
                               caller does next, soa/train_ppo.py:104-105).  The obs returned for
                               that step is still the terminal one, as env.step returns it. */

/* action dtypes */
#define TA_ACT_I32 0
#define TA_ACT_U8 1
#define TA_ACT_I64 2

/* flag bits inside ta_env_state.flags (gym_minigrid/envs/twoarmy_v4.py:14-24) */
#define TA_F_PONE 1
#define TA_F_PATROL 2
#define TA_F_UP1 4
#define TA_F_RIGHT2 8
#define TA_F_UPD_H 16
#define TA_F_UPD_L 32
#define TA_F_FIRST_ROOM2 64

/* cell codes: 0 empty (None), 1 wall, 2 ball (yellow), 3 goal */

/* One env as exported / imported for differential testing against the reference
 * (everything Twoarmy.step reads or writes). 328 bytes. */
typedef struct {
    uint8_t grid[289];     /* reference order: index y*17+x (minigrid.py:599-607) */
    uint8_t agent_x, agent_y;
    uint8_t flags;         /* TA_F_* */
    uint8_t risk_count;
    uint8_t error;         /* TA_ENV_ERR_* */
    uint8_t balls[10][2];  /* obstacles[0..2], obstacles1[0..2], obstacles2[0..3]: (x,y); 0xFF = cur_pos None */
    uint8_t pad_[2];
    int32_t step_count;    /* minigrid.py:972,1334 */
    int32_t step_move;     /* twoarmy_v4.py:14,87,296 */
    uint32_t t;            /* env steps since creation (Philox counter word) */
} ta_env_state;

typedef struct ta_batch *ta_handle;

/* Replaces gym.make(id, agent_view_size=view) x n_envs
 * (gym_minigrid/__init__.py:10-20 -> Twoarmy_v4.__init__ twoarmy_v4.py:9-36 ->
 * MiniGridEnv.__init__ minigrid.py:866-945, which ends in reset()).
 * version 4|6; view odd in [3,17]; env_id0 = global index of env 0 (sharding across
 * GPUs keeps draws independent of the shard layout); seed keys the Philox stream. */
int ta_create(ta_handle *out, int version, int64_t n_envs, int view, int device, uint64_t seed,
              uint64_t env_id0);
int ta_destroy(ta_handle h);

int64_t ta_num_envs(ta_handle h);
int ta_view(ta_handle h);
int ta_version(ta_handle h);

/* Replaces MiniGridEnv.reset (minigrid.py:947-980 -> _gen_grid twoarmy_v4.py:38-80) on the
 * envs whose mask byte is non-zero (mask NULL = all).  Like the reference it rebuilds grid,
 * balls, agent and step_count only; the Twoarmy flags carry over.  hard != 0 additionally
 * re-runs __init__'s flag initialisation (a freshly constructed env) and clears the error
 * bits.  obs_out (nullable): uint8 [n][V][V][3], the gen_obs() of EVERY env after the call. */
int ta_reset(ta_handle h, const uint8_t *mask, int hard, uint8_t *obs_out, void *stream);

/* Replaces MiniGridEnv.gen_obs in its GENERAL form (minigrid.py:1443-1496): any agent_dir, and
 * see_through_walls=False -> Grid.process_vis (minigrid.py:795-832).  The registered Twoarmy envs hard-code
 * agent_dir 3 / see_through_walls=True (twoarmy_v4.py:34,68), for which ta_reset / ta_step produce the
 * observation; this entry point exists so that the whole of gen_obs has a device form.
 *   agent_dirs   nullable uint8 [n] (0 right, 1 down, 2 left, 3 up); NULL: agent_dir for every env
 *   see_through  nullable uint8 [n]; NULL: see_through_all for every env
 *   obs_out      uint8 [n][V][V][3]; unseen cells are (0,0,0) as Grid.encode leaves them */
int ta_observe_general(ta_handle h, const uint8_t *agent_dirs, int agent_dir, const uint8_t *see_through, int see_through_all,
                       uint8_t *obs_out, void *stream);

/* Replaces Twoarmy_v4.step / Twoarmy_v6.step (twoarmy_v4.py:82-322, twoarmy_v6.py:83-325)
 * including MiniGridEnv.step (minigrid.py:1333-1441) and gen_obs (minigrid.py:1443-1496),
 * for all n envs in one fused kernel.
 *   actions      [n] of action_dtype; >=7 is clamped to 0 (twoarmy_v4.py:84-85); 4, 5 and
 *                negatives set TA_ENV_ERR_BAD_ACTION and leave that env untouched
 *   draws        nullable.  NULL: production mode, draws come from Philox4x32-10 keyed by
 *                (seed, global env id, env step index).  Non-NULL: verification mode, uint8
 *                [n][8], the value np.random.choice returned at call-site slot s (SURVEY.md
 *                section 3.5) is read from draws[i][s]
 *   obs_out      uint8 [n][V][V][3], image[x][y][c] as Grid.encode lays it out
 *   reward_out   float32 [n]  (the reference's Python float cast to float32)
 *   term_out, trunc_out  uint8 [n]
 *   consumed_out nullable uint8 [n]: bit s set iff call-site slot s executed this step */
int ta_step(ta_handle h, const void *actions, int action_dtype, const uint8_t *draws, int flags,
            uint8_t *obs_out, float *reward_out, uint8_t *term_out, uint8_t *trunc_out,
            uint8_t *consumed_out, void *stream);

/* ta_step with the results in TRANSFER form (what ta_step_host ships over PCIe):
 *   codes_out   uint32 [npad/32][2*V*V]: the observations of a 32-env tile as one cell stream (env-major, then
 *               image[x][y]) of 2-bit cell codes, 16 cells per word (cell q of a word = bits 2q, 2q+1); expanding
 *               code c to (type, color, 0) = ((1,0),(2,5),(6,4),(8,1))[c] yields exactly obs_out of ta_step
 *   status_out  uint8 [npad]: reward index (0..4 = -0.01, -0.1, -0.9, 0.2, 0.9) | terminated << 3 | truncated << 4
 * npad = n rounded up to 32. */
int ta_step_packed(ta_handle h, const void *actions, int action_dtype, const uint8_t *draws, int flags,
                   uint32_t *codes_out, uint8_t *status_out, uint8_t *consumed_out, void *stream);

/* The reference-facing call: env.step for every env with HOST arrays in and out (what a gym caller holds,
 * gym_minigrid/minigrid.py:1439-1441: obs["image"] is a numpy array).  H2D actions, the fused kernel, D2H results,
 * synchronised before returning; this is the end-to-end path bench.py times.
 *   default            the observations cross PCIe in the packed transfer form above (72.25 B instead of 867 B per
 *                      env at V = 17), in pieces, and a pool of host threads (TA_HOST_THREADS, default: the cores of
 *                      this process / LOCAL_WORLD_SIZE) expands each piece into obs_out while the next one is on the
 *                      wire.  obs_out may be pageable; 16-byte alignment enables streaming stores.
 *   TA_STEP_HOST_DMA   the expanded observations are copied straight into obs_out by the DMA engine (fastest when
 *                      obs_out is pinned and host cores are scarce); reward / flags still travel as one status byte.
 * When the process is the only rank on the node (LOCAL_WORLD_SIZE unset or 1), has at most 4 host threads and obs_out
 * is pinned, the default switches itself to the DMA form (decided on the handle's first call; TA_STEP_HOST_AUTO=0
 * keeps the threads).  With several ranks per node the packed form is always used: the DMA form saturates the host.
 * Same bytes either way (tests/test_gpu_parity.py::test_step_host_equals_device_step). */
#define TA_STEP_HOST_DMA 2
int ta_step_host(ta_handle h, const void *actions, int action_dtype, int flags, uint8_t *obs_out,
                 float *reward_out, uint8_t *term_out, uint8_t *trunc_out);
/* The decode stage of ta_step_host on its own (pure host code, calling thread only): HOST copies of
 * ta_step_packed's codes / status -> the arrays ta_step would have produced.  For consumers that keep the transfer
 * form (replay storage, sockets) and expand it where the observations are needed. */
int ta_decode_packed_host(const uint32_t *codes, const uint8_t *status, int64_t n, int view, uint8_t *obs_out,
                          float *reward_out, uint8_t *term_out, uint8_t *trunc_out);
/* device->host bytes the last ta_step_host call moved */
int64_t ta_step_host_d2h_bytes(ta_handle h);

/* T consecutive steps with a pre-sampled action tensor [T][n] (random-policy rollouts,
 * soa/datacol_predictor.py:106 style).  Outputs are [T][n]... ; always autoreset, Philox draws. */
int ta_rollout(ta_handle h, const void *actions, int action_dtype, int T, uint8_t *obs_out,
               float *reward_out, uint8_t *term_out, uint8_t *trunc_out, void *stream);

/* Replaces Env_transact.matrix_env + data_env + the frame-stack roll
 * (soa/env_buffer.py:300-334, soa/train_ppo.py:116-121).
 *   codes_out  nullable uint8 [n][289], index y*17+x: 0 empty/goal(0.9) 1 wall(-0.9)
 *              2 ball(-0.5) 4 agent(0.3)   (compact form of the float LUT)
 *   matrix_out nullable float32 [n][289], the LUT applied
 *   place_out  nullable float32 [n][2] = (agent y, agent x);  goal is the constant (2,14) */
int ta_state_matrix(ta_handle h, uint8_t *codes_out, float *matrix_out, float *place_out, void *stream);

/* Frame-stack roll fused with matrix_env: s_stack float32 [n][5][289], p_stack [n][5][2] are
 * shifted by one frame (oldest dropped) and the current frame appended, exactly
 * np.delete(x,0,0); np.append(x,[new],0) of train_ppo.py:116-121.  With init != 0 all five
 * frames are set to the current one (np.tile in Env_transact.reset, env_buffer.py:420-423)
 * for envs whose mask byte is non-zero (mask NULL = all). */
int ta_stack_roll(ta_handle h, float *s_stack, float *p_stack, const uint8_t *init_mask, int init,
                  void *stream);
/* Same roll on the compact codes (uint8 [n][5][289], the codes of ta_state_matrix): the form
 * the device rollout buffer stores (1 byte per cell instead of Buffer_gridworld's float32,
 * soa/train_ppo.py:93-97); the float LUT of matrix_env is applied when a minibatch is loaded. */
int ta_stack_roll_codes(ta_handle h, uint8_t *s_codes, float *p_stack, const uint8_t *init_mask, int init,
                        void *stream);

/* The same frame-stack update OUT OF PLACE, fused with the episode-start tiling -- what the
 * vectorised rollout uses (one pass: read 4 frames, write 5):
 *   base  = tile(reset frame) x5 / (15,3) x5   for envs with init_all != 0 or prev_done[e] != 0
 *           (Env_transact.reset after MiniGridEnv.reset, soa/env_buffer.py:413-428)
 *         = s_prev[e] / p_prev[e]               otherwise
 *   s_out[e] = [base[1], base[2], base[3], base[4], matrix_env(current state)]   (train_ppo.py:116-121)
 * dtype: TA_STACK_F32 (float32 LUT values, the reference record) or TA_STACK_U8 (compact codes).
 * s_prev/s_out [n][5][289] must be 16-byte aligned and distinct; p_prev/p_out float32 [n][5][2]
 * (p_out nullable); prev_done uint8 [n] nullable. */
#define TA_STACK_F32 0
#define TA_STACK_U8 1
int ta_stack_push(ta_handle h, const void *s_prev, void *s_out, const float *p_prev, float *p_out,
                  const uint8_t *prev_done, int init_all, int dtype, void *stream);

/* Replaces MiniGridEnv.get_full_render (gym_minigrid/minigrid.py:1514-1563, Grid.render :712-747) for the
 * envs env_ids[0..m) (NULL = envs 0..m-1): rgb_out uint8 [m][17*ts][17*ts][3].
 *   atlas  uint8 [16][ts][ts][3] on the device: the tile images render_tile (:662-710) caches, index =
 *          cell code | agent here << 2 | highlighted << 3 (built on the host, render.tile_atlas)
 *   highlight != 0: cells inside the agent's view window use the highlighted tiles (env.highlight) */
int ta_render(ta_handle h, const uint8_t *atlas, int tile_size, int highlight, const int64_t *env_ids, int64_t m,
              uint8_t *rgb_out, void *stream);

/* Export / import the full env state (device buffers of n ta_env_state records). */
int ta_export_state(ta_handle h, ta_env_state *out, void *stream);
int ta_import_state(ta_handle h, const ta_env_state *in, void *stream);

/* Advantage / returns over a rollout laid out [T][n] (time-major), fp32.
 * Replaces soa/agent/PPO.py:112-115 (the lambda = 0, use_mask = 0, normalize = 0 mode:
 * target_v = r + gamma*V(s'), adv = target_v - V(s)) and generalises it to GAE(gamma,lambda).
 *   v_next  nullable [T][n]: V(s') per sample as the reference computes it; when NULL,
 *           V(s_{t+1}) = v[t+1] and last_v [n] bootstraps the final step
 *   done    nullable uint8 [T][n], used only when use_mask != 0
 *   normalize != 0: adv <- (adv - mean) / (std + 1e-8) over all T*n (unbiased std, the
 *           commented-out line PPO.py:115); stats [3] float64 device scratch receives
 *           (sum, sum of squares, count) so several ranks can combine them first
 *   adv_out, ret_out  float32 [T][n]  (ret = target_v in the reference's naming) */
int ta_gae(const float *reward, const float *v, const float *v_next, const float *last_v,
           const uint8_t *done, float gamma, float lam, int use_mask, int T, int64_t n,
           float *adv_out, float *ret_out, void *stream);
/* ta_gae that also leaves (sum, sum of squares, count) of adv_out in stats3 (overwritten) from the same launch:
 * normalising then costs ta_adv_normalize only. */
int ta_gae_stats(const float *reward, const float *v, const float *v_next, const float *last_v,
                 const uint8_t *done, float gamma, float lam, int use_mask, int T, int64_t n,
                 float *adv_out, float *ret_out, double *stats3, void *stream);
/* ta_gae followed by adv <- (adv - mean) / (std + 1e-8) (PPO.py:115) for ONE rank.  work: TA_GAE_WORK_DOUBLES float64
 * of device scratch (overwritten; work[0..2] end as sum, sum of squares, count of the un-normalised adv).  A rollout whose
 * whole grid is resident at once (T <= 128 and about n <= 18944 on a B200: BASELINE configs[3]'s 128 x 16384 is) is
 * normalised inside the GAE launch -- adv is written once, and the moments are added in a fixed order (bitwise
 * reproducible); larger ones take the fused moments plus one ta_adv_normalize pass. */
#define TA_GAE_WORK_CTAS 1024
#define TA_GAE_WORK_DOUBLES (8 + 2 * TA_GAE_WORK_CTAS)
int ta_gae_normalized(const float *reward, const float *v, const float *v_next, const float *last_v,
                      const uint8_t *done, float gamma, float lam, int use_mask, int T, int64_t n,
                      float *adv_out, float *ret_out, double *work, void *stream);
int ta_adv_stats(const float *adv, int64_t count, double *stats3, void *stream);
int ta_adv_normalize(float *adv, int64_t count, const double *stats3, void *stream);

/* Hindsight relabelling of a [T][n] rollout: Buffer_gridworld.her_func (soa/env_buffer.py:101-143)
 * for every episode that ends inside the window, without materialising the copies.
 *   first_record  0 for her_func.  4 = Buffer_gridworld.pre_her_func (soa/env_buffer.py:145-210): its 9-frame
 *             records exist from an episode's 5th step on (train_ppo_predictor.py:140-142), so only records
 *             >= first_record of an episode are candidates, indices are still counted from the episode start,
 *             and a prefix is kept if index > first_record (its `index > 0`)
 *   p         float32 [T][n][5][2], the record's p field; [4] = (y, x) reached by the step
 *   done      uint8 [T][n], the episode ended with this record (terminated | truncated)
 *   chosen_in nullable uint8 [T][n][4].  NULL: the up-to-4 relabel indices are drawn with Philox
 *             keyed by (seed, global env id, t).  Non-NULL (verification): read at each
 *             episode-end record, the record indices np.random.choice(indices, k, replace=False)
 *             returned (env_buffer.py:115), 0xFF = none
 *   uniq_out  nullable uint8 [T][n][64], m_out nullable uint8 [T][n]: at each episode-end record the
 *             `indices` np.unique returns (env_buffer.py:108; first occurrence of every distinct
 *             position in sorted (y,x) order) and their count -- what the host feeds np.random.choice
 *   plan_out  uint16 [T][n][4]: 0xFFFF = the record is not part of relabel slot c; otherwise the
 *             new goal y<<5|x, with bit 15 set on the prefix's last record (r = 0.9, d = 1 there,
 *             env_buffer.py:126-127).  Prefixes with index 0 are dropped like line 121. */
int ta_her_plan(const float *p, const uint8_t *done, int T, int64_t n, int first_record, uint64_t seed, uint64_t env_id0,
                const uint8_t *chosen_in, uint8_t *uniq_out, uint8_t *m_out, uint16_t *plan_out, void *stream);

/* First layer of TINet fused for the device rollout buffer: matrix_env LUT decode + UpsamplingNearest2d(4)
 * + Conv2d(4, 64, kernel 4, stride 2) + bias + ReLU (soa/agent/net/all_net.py:142-143,157,180-181),
 * folded exactly into a K = 16 product over 2x2 patches of the 17x17 frames (see csrc/ta_conv1.cuh).
 *   x        4 frames x 289 per sample, samples x_stride ELEMENTS apart: TA_X_F32 float32 matrix_env
 *            values, or TA_X_U8 the compact codes (decoded on the fly)
 *   w4, b4   float32 [256][16] / [256]: the conv weight / bias folded per output phase by the host
 *            (row (py*2+px)*64 + o, column (dy*2+dx)*4 + c)
 *   y_bf16   bfloat16 [batch][33][33][64] (channels-last), = relu(conv + bias)
 * Three forward kernels (TA_CONV1_TC): 2 (default) warp-specialised tcgen05 with tensor-map stores (csrc/ta_conv1_fwd_ws.cuh;
 * needs y_bf16 16-byte aligned, else the next one runs), 1 single-role tcgen05 (csrc/ta_conv1_tc.cuh) -- the two are
 * bit-identical --, 0 FP32 FMA (csrc/ta_conv1.cuh).
 * ta_conv1_bwd: gradients of w4 / b4 (overwritten) given y and dL/dy; x needs no gradient. */
#define TA_X_F32 0
#define TA_X_U8 1
int ta_conv1_fwd(const void *x, int x_dtype, int64_t x_stride, const float *w4, const float *b4, int64_t batch,
                 void *y_bf16, void *stream);
int ta_conv1_bwd(const void *x, int x_dtype, int64_t x_stride, const void *y_bf16, const void *dy_bf16,
                 int64_t batch, float *dw4, float *db4, void *stream);
/* ta_conv1_fwd with an addend and an optional ReLU: y = act(b4 + w4 . patch(x) + addend), act = ReLU (relu != 0) or the
 * identity.  The folded first layer is linear in its input channels, so a first convolution with 8 input channels
 * (Net_PPO_Predictor_actor / _critic, soa/agent/net/all_net.py:249-305: 4 current + 4 predicted frames) is two passes
 * over 4 channels each: the second half without bias / ReLU, then the first half with the result as addend. */
int ta_conv1_fwd_add(const void *x, int x_dtype, int64_t x_stride, const float *w4, const float *b4, int64_t batch,
                     const void *addend_bf16, int relu, void *y_bf16, void *stream);
/* Parity planes.  The data gradient of a k x k (k = 3, 4) stride-2 unpadded convolution splits by the parity (pa, pb)
 * of the input pixel: dx[2i+pa][2j+pb] only receives the taps ky = pa, kx = pb (mod 2), from dz[i - ky/2][j - kx/2] --
 * a stride-1 convolution of dz with a <= 2x2 sub-kernel.  All four classes come out of ONE stride-1 convolution
 * (2x2 kernel, padding 1, 4*cin output channels): "merged planes" bfloat16 [batch][OH+1][OW+1][4][cin], class
 * pa*2+pb in channel block c.  (TINet's second and third convolutions, all_net.py:144-147.)
 *
 * ta_parity_class_weights: w bfloat16 [cout][cin][k][k] with the given element strides -> out = that convolution's
 *   weight [4*cin][cout][2][2] in channels-last memory (16*cin*cout elements).
 * ta_planes_to_dense_relu: merged planes -> the dense channels-last gradient [batch][H][W][C], masked with the ReLU
 *   of the layer below: dz_out = y > 0 ? planes[b][h>>1][w>>1][(h&1)*2+(w&1)] : 0.
 * ta_conv1_bwd_planes: ta_conv1_bwd with dL/dy given as merged planes [batch][17][17][4][64] (block py*2+px of
 *   position (m, n) = gradient of output pixel (2m+py, 2n+px)): the planes are the phase structure of the folded
 *   first layer, so the 33x33x64 gradient tensor is never materialised. */
int ta_parity_class_weights(const void *w_bf16, int64_t stride_o, int64_t stride_i, int64_t stride_y, int64_t stride_x,
                            int cout, int cin, int ksize, void *out_bf16, void *stream);
int ta_planes_to_dense_relu(const void *planes_bf16, const void *y_bf16, void *dz_bf16, int64_t batch, int H, int W,
                            int C, int ksize, void *stream);
int ta_conv1_bwd_planes(const void *x, int x_dtype, int64_t x_stride, const void *y_bf16, const uint32_t *relu_mask,
                        const void *planes_bf16, int class_major, int64_t batch, float *dw4, float *db4, void *stream);
/* ta_conv1_fwd that also writes the layer's ReLU mask as bits: relu_mask uint32 [batch*289 positions][4 phases][2
 * halves of 32 channels]; in a word, bit q / bit 16+q = channel 2q / 2q+1 of the half of output pixel (2m+py, 2n+px)
 * is non-zero.  Given to ta_conv1_bwd_planes (y_bf16 may then be NULL) the weight gradient reads 8 bytes per pixel
 * instead of the 128 bytes of y. */
int ta_conv1_fwd_mask(const void *x, int x_dtype, int64_t x_stride, const float *w4, const float *b4, int64_t batch,
                      void *y_bf16, uint32_t *relu_mask, void *stream);

/* Data gradient of TINet's second convolution (Conv2d(64, 64, 3, stride 2), all_net.py:144-145) as merged parity planes
 * on the tensor cores (tcgen05), every parity class with its own tap list (9 tap products per position instead of the 16
 * of the merged stride-1 convolution):
 *   ta_conv2_dgrad_prep    w bf16 [64][64][3][3] with the given element strides -> wimg, the 73728-byte operand image
 *                          (once per optimiser step)
 *   ta_conv2_dgrad_planes  dz bf16 [batch][16][16][64] (channels-last) -> planes bf16; class_major = 1: [4][batch*289][64]
 *                          (the warp-specialised kernel: cp.async producer warp, MMA warp, four epilogue warps writing 4 KB
 *                          TMA bulk stores), class_major = 0: [batch][17][17][4][64] like the cuDNN merged-plane convolution
 *                          (single-role kernel).  relu_mask (nullable): the first layer's ReLU bit mask (ta_conv1_fwd_mask)
 *                          applied in the epilogue, so the planes are already d loss / d (conv1 pre-activation).
 * ta_conv1_bwd_planes takes either layout (its class_major argument). */
int ta_conv2_dgrad_prep(const void *w_bf16, int64_t stride_o, int64_t stride_i, int64_t stride_y, int64_t stride_x,
                        void *wimg_bf16, void *stream);
int ta_conv2_dgrad_planes(const void *dz_bf16, const void *wimg_bf16, const uint32_t *relu_mask, int64_t batch,
                          int class_major, void *planes_bf16, void *stream);

/* ta_conv2_dgrad_planes followed by ta_conv1_bwd_planes in ONE kernel (csrc/ta_stem_bwd_tc.cuh): the gradient of the first
 * layer's output goes from the tensor core's accumulators through shared memory into the weight-gradient GEMM and is never
 * written to memory.  dz_bf16 [batch][16][16][64] = the (ReLU-masked) gradient of the second convolution's output, wimg_bf16
 * from ta_conv2_dgrad_prep, relu_mask from ta_conv1_fwd_mask, x / x_dtype / x_stride as in ta_conv1_bwd;
 * dw4 float [256][16], db4 float [256] are overwritten (all_net.py:142-145 backward). */
int ta_conv2_dgrad_conv1_bwd(const void *dz_bf16, const void *wimg_bf16, const uint32_t *relu_mask, const void *x, int x_dtype,
                             int64_t x_stride, int64_t batch, float *dw4, float *db4, void *stream);

/* Data gradient helper for TINet's stride-2 unpadded convolutions (all_net.py:144-149) in channels-last
 * bf16: dcols [batch*OH*OW][k*k*C] (= dY x W from a plain GEMM, columns (ky,kx,c)) -> dx [batch][H][W][C].
 * k in {3, 4}, C a multiple of 8, OH = (H-k)/2+1. */
int ta_col2im_s2(const void *dcols_bf16, void *dx_bf16, int64_t batch, int H, int W, int C, int ksize, void *stream);
/* Its forward: x [batch][H][W][C] -> cols [batch*OH*OW][k*k*C] (columns (ky,kx,c)), the operand of the plain GEMM
 * TINet's last convolution (all_net.py:150, 128 -> 256, 3x3 stride 2 on 7x7) runs as. */
int ta_im2col_s2(const void *x_bf16, void *cols_bf16, int64_t batch, int H, int W, int C, int ksize, void *stream);
/* Bias gradient of those convolutions: out[c] = sum over rows of x [rows][C] (channels-last bf16, C in
 * {64, 128, 256}); out float32 [C] is overwritten. */
/* The frozen frame predictor's convolution stacks in eval mode (BASELINE configs[4]; PPO_Predictor.pred_states,
 * soa/agent/PPO_Predictor.py:70-83), one fused kernel each, every activation in shared memory (csrc/ta_pred.cuh).
 * ta_pred_encoder = Net_Encoder (soa/agent/net/all_net.py:7-51): x [M][289] uint8 matrix codes (TA_STACK_U8) or float32
 *   LUT values (TA_STACK_F32) -> z bf16 [M][1024] (64 x 4 x 4).  Per layer: convolution weights (w1 [16][4][4];
 *   w2 [5][5][16 in][16 out]; w3 [2][2][16 in][64 out]) and the folded eval-mode BatchNorm + bias as scale s / shift t
 *   per output channel (y = relu(conv * s + t)).
 * ta_pred_decoder = Net_Decoder (all_net.py:100-137) up to and including the average pool: z bf16 [M][1024] -> out
 *   float32 [M][289].  w1 [2][2][64 in][16 out], w2 [5][5][16 in][16 out] (transposed-convolution taps), w3 [3][3][16] =
 *   ConvTranspose2d(16,1,4,2) + AvgPool2d(4) folded into a 3x3 stride-2 padding-1 convolution, b3 its bias.
 * twoarmy_b200.predictor builds these arrays from the modules' parameters. */
int ta_pred_encoder(const void *x, int x_dtype, int64_t M, const float *w1, const float *s1, const float *t1,
                    const float *w2, const float *s2, const float *t2, const float *w3, const float *s3,
                    const float *t3, void *z_bf16, void *stream);
int ta_pred_decoder(const void *z_bf16, int64_t M, const float *w1, const float *b1, const float *w2,
                    const float *b2, const float *w3, float b3, float *out, void *stream);
/* The cell update of torch.nn.LSTM (gate order i, f, g, o) for the frozen frame predictor of BASELINE configs[4]
 * (soa/agent/net/all_net.py:53-98; twoarmy_b200.predictor.LSTM runs the gate GEMMs and calls this per layer and step):
 *   pre = gx + gh + bias;  c <- sigmoid(f) c + sigmoid(i) tanh(g);  h <- sigmoid(o) tanh(c)
 * gx, gh float32 [B][4H] (gh nullable), bias float32 [4H] (b_ih + b_hh), c float32 [B][H] in place, h_out bf16 [B][ld_h].
 * H % 4 == 0; gx / gh / bias / c 16-byte aligned, h_out 8-byte aligned with ld_h % 4 == 0. */
int ta_lstm_gates(const float *gx, const float *gh, const float *bias, float *c, void *h_out_bf16, int64_t ld_h,
                  int64_t B, int H, void *stream);
int ta_channel_sum_bf16(const void *x_bf16, int64_t rows, int C, float *out, void *stream);

/* ---- the hand-scheduled PPO optimiser step (soa/agent/PPO.py:124-144 without autograd; fused_step.py) ----------------
 * Everything around the convolutions / GEMMs of one optimiser step, as single kernels with deterministic reductions.
 *
 * ta_relu_bwd_bias: dz = dy * [y > 0] (bf16 [rows][C]; dy rows ld_dy elements apart) and db_out[c] = sum over rows of
 *   dz[.][c] (float32) in one pass: a Linear / conv layer's ReLU backward and bias gradient.  y_bf16 == dz_bf16 == NULL:
 *   plain column sum of dy.  C a multiple of 8 with 256 % (C/8) == 0.  scratch: ta_relu_bwd_bias_scratch_floats(rows, C)
 *   floats, zero before the first use (the kernel leaves its counter zero); one scratch per concurrent stream.
 * ta_ppo_actor_loss: logits bf16 [B][8] (5 used) -> loss_out[0] = mean_i( -min(r_i A_i, clamp(r_i, 1-clip, 1+clip) A_i)
 *   - ent_coef H_i ), r_i = exp(log softmax(logits_i)[a_i] - old_logp_i) (PPO.py:124-132); dlogits bf16 [B][8] its
 *   gradient, db_head[5] the column sums of dlogits; step_counter[0] += 1 (the optimiser step count ta_adam_shadow reads).
 * ta_ppo_critic_loss: v bf16 [B][8] (column 0) -> mean smooth_l1(v, target) (PPO.py:133), dv, db_head[1].
 * ta_adam_shadow: torch.optim.Adam's update (PPO.py:57-58: lr, eps = 1e-5, betas (0.9, 0.999), no weight decay) on flat
 *   float32 p / m / v of n elements from gradient g * grad_scale, at step t = step_counter[0]; also writes the bf16
 *   copy p_bf16 the forward / backward kernels read.
 * ta_planes_relu_bwd_bias: ta_relu_bwd_bias with dy given as merged parity planes bf16 [batch][(H+1)/2][(W+1)/2][4][C]
 *   of the map [batch][H][W][C]: ta_planes_to_dense_relu and the bias gradient in one pass (rows = batch*H*W). */
int64_t ta_relu_bwd_bias_scratch_floats(int64_t rows, int C);
int ta_planes_relu_bwd_bias(const void *planes_bf16, const void *y_bf16, void *dz_bf16, int64_t batch, int H, int W, int C,
                            float *db_out, float *scratch, void *stream);
int ta_relu_bwd_bias(const void *dy_bf16, int64_t ld_dy, const void *y_bf16, void *dz_bf16, int64_t rows, int C,
                     float *db_out, float *scratch, void *stream);
int ta_ppo_actor_loss(const void *logits_bf16, const int32_t *act, const float *old_logp, const float *adv, int B,
                      float clip, float ent_coef, void *dlogits_bf16, float *loss_out, float *db_head,
                      float *step_counter, void *stream);
int ta_ppo_critic_loss(const void *v_bf16, const float *target, int B, void *dv_bf16, float *loss_out, float *db_head,
                       float *step_counter, void *stream);
int ta_adam_shadow(float *p, const float *g, float *m, float *v, void *p_bf16, int64_t n, const float *step_counter,
                   float lr, double beta1, double beta2, float eps, float grad_scale, void *stream);

/* Per-step weight forms of one network derived from its master copy (device pointers; w1 strides in elements):
 * the folded first layer (all_net.py:142-143,157 -> [256][16] / [256], see ta_conv1_fwd), fc0 with its input features
 * permuted to the (pixel, channel) order conv4's GEMM produces, positionnet / head padded to 16 columns / 8 rows. */
typedef struct {
    const float *w1;
    int64_t s_o, s_c, s_y, s_x;
    const float *b1;
    float *w4, *b4;
    const void *fc0;      /* bf16 [256][2304] */
    void *fc0p;           /* bf16 [256][9][256] */
    const void *pos;      /* bf16 [128][10] */
    void *pos16;          /* bf16 [128][16] */
    const void *head, *head_b;   /* bf16 [nh][512], [nh] */
    void *head8, *head_b8;       /* bf16 [8][512], [8] */
    int nh;
} ta_tinet_prep_args;
int ta_tinet_prep(const ta_tinet_prep_args *args, void *stream);

/* The inverse for the gradients: folded conv1 gradients, the bf16 weight gradients the GEMMs / cuDNN leave (4 dense
 * segments), the permuted fc0 and padded positionnet / head gradients -> the flat float32 gradient buffer.  Every group is
 * optional (dw4 / fc0p / pos16 / head8 NULL, n[k] 0): the late layers are finalised first so that their all-reduce
 * overlaps the convolution stem's backward. */
typedef struct {
    const float *dw4, *db4;
    float *g_w1;
    int64_t s_o, s_c, s_y, s_x;
    float *g_b1;
    const void *src[4];   /* bf16 */
    float *dst[4];
    int64_t n[4];
    const void *fc0p;
    float *g_fc0;
    const void *pos16;
    float *g_pos;
    const void *head8;
    float *g_head;
    int nh;
} ta_tinet_grad_args;
int ta_tinet_grad(const ta_tinet_grad_args *args, void *stream);

/* Minibatch gather for the update (PPO.py:122-127): sample idx[i] (through src for hindsight-relabelled samples) ->
 * sb uint8 [bs][4][289] (frames 0..3), pg16 bf16 [bs][16] (4 positions + goal, zero padded), a_mb int32, old_logp /
 * adv / target_v float32 [bs]. */
int ta_gather_minibatch(const uint8_t *s, const float *p, const float *g, const int64_t *a, const float *old_logp,
                        const float *adv, const float *target_v, const int64_t *idx, const int64_t *src, int bs,
                        uint8_t *sb, void *pg16_bf16, int32_t *a_mb, float *old_mb, float *adv_mb, float *tv_mb,
                        void *stream);

/* introspection */
int ta_abi_version(void);
const char *ta_strerror(int code);
const char *ta_last_cuda_error(void);
/* kernels launched by this library since load (bench.py's gpu_launches claim) */
int64_t ta_launch_count(void);
/* duration in ms of the last `ta_step` kernel when timing is enabled (CUDA events on the
 * launching stream); enable with ta_set_timing(h, 1). Synchronises. */
int ta_set_timing(ta_handle h, int on);
int ta_last_step_ms(ta_handle h, float *ms);

#ifdef __cplusplus
}
#endif
#endif /* TWOARMY_B200_H */
