/* njode_b200 -- C ABI of the B200-native NJ-ODE hot path (libnjode_b200.so).
 *
 * Plain C, raw device pointers and sizes, no torch types.  Every entry point returns 0 on success
 * and a negative code on failure; njode_last_error() returns a human-readable message for the last
 * failure of the calling thread.  Kernels never allocate: the caller owns every buffer (sizes come
 * from njode_plan()).  All work is enqueued on the given CUDA stream (cudaStream_t passed as void*).
 *
 * The reference (HerreraKrachTeichmann/NJODE) is pure Python and has no FFI; these entry points
 * replace, as a unit, the Python functions named next to each of them (paths relative to the
 * reference tree).  The reference-side binding is the ctypes stub in INTEGRATION.md /
 * njode_b200/_ext.py.
 */
#ifndef NJODE_B200_H
#define NJODE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NJODE_ABI_VERSION 6
#define NJODE_MAX_LINEAR 8          /* max number of Linear layers per network */

enum { NJODE_ACT_NONE = 0, NJODE_ACT_TANH = 1, NJODE_ACT_RELU = 2 };
enum { NJODE_LOSS_STANDARD = 0, NJODE_LOSS_EASY = 1 };     /* NJODE/models.py:129-132 LOSS_FUN_DICT */
enum { NJODE_NET_ODE = 0, NJODE_NET_ENC = 1, NJODE_NET_RO = 2,
       /* use_rnn=True only: the two affine maps of torch.nn.GRUCell (NJODE/models.py:202-217), one Linear each:
        * weight_ih [3H, input] / bias_ih and weight_hh [3H, H] / bias_hh, gate order (r, z, n) */
       NJODE_NET_GRU_IH = 3, NJODE_NET_GRU_HH = 4, NJODE_NUM_NETS = 5 };

/* one feed-forward network as built by get_ffnn (NJODE/models.py:140-166): n_linear Linear layers,
 * an activation (+dropout) after every layer but the last.  Weights are nn.Linear layout
 * [out, in] row-major fp32 at float offset w_off[l] of the flat parameter buffer; b_off[l] = -1
 * when bias=False. */
typedef struct njode_mlp {
    int32_t n_linear;
    int32_t dims[NJODE_MAX_LINEAR + 1];
    int32_t act[NJODE_MAX_LINEAR];
    int64_t w_off[NJODE_MAX_LINEAR];
    int64_t b_off[NJODE_MAX_LINEAR];
} njode_mlp_t;

/* constructor state of NJODE (NJODE/models.py:284-362) that the kernels need */
typedef struct njode_model {
    int32_t input_size, hidden_size, output_size;
    int32_t masked;            /* options['masked']            NJODE/models.py:339-341 */
    int32_t input_current_t;   /* options['input_current_t']   NJODE/models.py:334-337 */
    int32_t loss_kind;         /* options['which_loss']        NJODE/models.py:322-326 */
    int32_t residual;          /* options['residual_enc_dec']  NJODE/models.py:329-332 */
    int32_t training;          /* module.training (dropout on) */
    float   weight;            /* model.weight, NJODE/models.py:316,364-367 */
    float   dropout_p;
    uint64_t dropout_seed;     /* counter-based masks keyed (seed, path, event, net, layer, neuron) */
    int64_t n_params;          /* floats in the flat parameter / gradient buffer */
    int32_t use_rnn;           /* use_rnn=True: the jump is h[i_obs] = GRUCell(tanh(X_obs), tanh(h[i_obs]))
                                  (NJODE/models.py:202-217,460-461) instead of the encoder */
    int32_t reserved1;
    njode_mlp_t net[NJODE_NUM_NETS];   /* NJODE_NET_ODE / ENC / RO (/ GRU_IH / GRU_HH when use_rnn) */
} njode_model_t;

/* one batch in the reference's collate contract (NJODE/data_utils.py:311-315) plus the host-built
 * event schedule (restating the float64 loop conditions of NJODE/models.py:430-439,497-505) and the
 * per-path CSR of observation rows.  All pointers are DEVICE pointers. */
typedef struct njode_batch {
    int32_t B;                 /* paths on this rank */
    int32_t N;                 /* observation rows  = time_ptr[-1] */
    int32_t K;                 /* observation times = len(times) */
    int32_t S;                 /* Euler steps the reference executes for this batch */
    int32_t E;                 /* recorded events (len(path_t)) when return_path, else 0 */
    int32_t batch_size_norm;   /* batch size used in the loss normalisation (global B under DP) */
    int32_t path_id_offset;    /* global id of local path 0 (dropout keys are rank-invariant) */
    int32_t n_units;           /* work units, see unit_desc */
    int32_t unit_kind;         /* 0: whole paths; 1: (path, inter-observation segment) units -- each loss
                                  unit ends with exactly one jump (c1 = c0 + 1, s1 = jump_step of its row),
                                  tail units have none; enables the segment fast path (non-masked model) */
    int32_t n_loss_units;      /* unit_kind 1: units [0, n_loss_units) end with a jump (sorted longest first),
                                  units [n_loss_units, n_units) are the tails (sorted longest first) */
    int32_t seg_n1[2];         /* per run (loss, tail): number of units with length >= T1 ... */
    int32_t seg_n2[2];         /* ... and >= T2 (T1 >= T2, chosen by the host from the batch's total work):
                                  long units are marched in lower tiles so that no tile's sequential chain
                                  of Euler steps dominates the makespan */
    int32_t reserved0;
    const float*   X;          /* [N, input_size] */
    const float*   M;          /* [N, input_size] or NULL */
    const float*   start_X;    /* [B, input_size] */
    const float*   n_obs_ot;   /* [B] as float, or NULL when no loss is requested */
    const int32_t* path_ptr;   /* [B+1]  CSR over paths ... */
    const int32_t* path_rows;  /* [N]    ... row ids of each path in time order */
    const int32_t* row_jump;   /* [N]    observation-time index i of each row */
    const float*   step_dt;    /* [S]    fl32(delta_t_) of Euler step k */
    const float*   step_t;     /* [S]    fl32(current_time) at the start of step k */
    const int32_t* jump_step;  /* [K]    number of Euler steps executed before jump i */
    const float*   jump_tau;   /* [K]    fl32(times[i]) */
    const int32_t* step_event; /* [S]    index into path_t of the record after step k (return_path) */
    const int32_t* jump_event; /* [K]    index into path_t of the record after jump i  (return_path) */
    /* work units: 6 int32 each = {path, s0, s1, c0, c1, start_code}; a unit integrates path `path`
     * over Euler steps [s0, s1), applying the path's jumps path_rows[c0..c1) where they fall,
     * starting from enc(X[start_row]) or, when start_row = -1, enc(start_X[path]).
     * start_code = (start_row + 1) | NJODE_UNIT_WRITES_HT if the unit ends its path (writes hT). */
    const int32_t* unit_desc;  /* [n_units, 6], sorted by length (longest first) */
} njode_batch_t;

#define NJODE_UNIT_WRITES_HT (1 << 30)

/* launch plan / buffer sizes for one (model, batch-shape) pair */
typedef struct njode_plan {
    int32_t tile_paths;        /* units marched in lockstep by one CTA */
    int32_t threads;
    int32_t grid_fwd, grid_bwd;
    int32_t weights_in_smem, grads_in_smem;
    int64_t smem_fwd_bytes, smem_bwd_bytes;
    int64_t image_floats;      /* padded parameter image (also the per-CTA gradient partial) */
    int64_t workspace_bytes;   /* scratch the caller must provide to forward/backward */
    int64_t recompute_bytes;   /* > 0: njode_backward can run WITHOUT anything saved by the forward pass (saved members NULL):
                                  it recomputes the forward of every segment from its checkpoint at the observation time; this
                                  many bytes of the workspace hold the h chains of the tiles in flight.  0: the backward of this
                                  (model, batch) needs the history written by njode_forward */
    int64_t act_bytes;         /* > 0: the kernels of this (model, batch) can keep the ODE network's hidden activations of every
                                  Euler step (njode_saved_t.act_hist, this many bytes): the backward then reads them instead
                                  of recomputing the hidden layers (optional: trade HBM for backward time) */
} njode_plan_t;

/* buffers produced by the forward pass that the backward pass re-reads; all NULL when the backward recomputes
 * (njode_plan_t.recompute_bytes > 0): nothing of size O(S * B) exists then */
typedef struct njode_saved {
    float* h_hist;             /* [S, B, hidden]  h at the start of every Euler step, or NULL (no grad / recompute) */
    float* h_before;           /* [N, hidden]     h just before the jump of row r, or NULL */
    float* y_after;            /* [N, output]     Y = readout(h after jump) of row r, or NULL */
    float* act_hist;           /* njode_plan_t.act_bytes bytes or NULL: hidden activations of the ODE network per Euler step
                                  (the same pointer, or NULL, in njode_forward and njode_backward) */
} njode_saved_t;

const char* njode_last_error(void);
int njode_abi_version(void);

/* sizes and launch geometry.  `device` = CUDA ordinal. */
int njode_plan(const njode_model_t* model, const njode_batch_t* batch_shape, int device,
               njode_plan_t* plan_out);

/* NJODE.forward (NJODE/models.py:379-518) incl. ode_step (369-377), ODEFunc.forward (188-199),
 * FFNN.forward (261-276), the jump scatter (449-489), compute_loss / compute_loss_2 (71-126).
 *   params    [n_params] flat fp32 parameters (layout given by model->net[*].w_off/b_off)
 *   hT        [B, hidden]             out
 *   loss      [1]                     out (sum over rows / batch_size_norm), or NULL (get_loss=False)
 *   path_h    [E, B, hidden] / path_y [E, B, output]   out when batch->E > 0 (return_path=True)
 *   saved     history for njode_backward, members may be NULL when no gradient is needed
 *   workspace [plan.workspace_bytes] */
int njode_forward(const njode_model_t* model, const njode_batch_t* batch, const float* params,
                  float* hT, float* loss, float* path_h, float* path_y,
                  const njode_saved_t* saved, void* workspace, void* stream);

/* reverse pass of the same computation = the autograd tape replay of loss.backward()
 * (NJODE/train.py:522) for all parameter tensors.
 *   grad_loss [1] device scalar dL/dloss ; grad_hT [B, hidden] or NULL
 *   grads     [n_params] out (overwritten), same layout as params */
int njode_backward(const njode_model_t* model, const njode_batch_t* batch, const float* params,
                   const njode_saved_t* saved, const float* grad_loss, const float* grad_hT,
                   float* grads, void* workspace, void* stream);

/* ---- tensor-core path (tcgen05 / TMEM, bf16 operands, fp32 accumulation and fp32 hidden state) ----
 * Serves the non-masked training / loss call of models whose MLPs are real dense contractions
 * (BASELINE config 5: d=16, H=256, 4x256 nets): input_size = output_size in {1,2,4,8,16}, hidden_size a
 * multiple of 16 up to 256, layer widths up to 256, segment units (batch->unit_kind == 1), no path
 * recording.  Same arguments and outputs as njode_forward / njode_backward; results agree with the fp32
 * path within the bf16 tolerance stated in tests/test_gpu_wide.py.
 *   njode_wide_supported        1 when the model qualifies (else 0, reason in njode_last_error())
 *   njode_wide_workspace_bytes  scratch for either call (not preserved between calls)
 *   njode_wide_saved_bytes      size of the caller-owned blob that carries the forward's operand tiles
 *                               (bf16 activation images per tile and chain step), h_before, Y, Y_bj and the
 *                               encoder outputs to njode_wide_backward
 * njode_wide_forward: `wide_saved` NULL = no gradient bookkeeping.  `saved` (may be NULL) optionally receives
 * h_hist / h_before / y_after in the layout njode_backward reads, so the fp32 backward can follow a
 * tensor-core forward. */
int njode_wide_supported(const njode_model_t* model);
int64_t njode_wide_workspace_bytes(const njode_model_t* model, const njode_batch_t* batch);
int64_t njode_wide_saved_bytes(const njode_model_t* model, const njode_batch_t* batch);
int njode_wide_forward(const njode_model_t* model, const njode_batch_t* batch, const float* params,
                       float* hT, float* loss, const njode_saved_t* saved, void* wide_saved,
                       void* workspace, void* stream);
int njode_wide_backward(const njode_model_t* model, const njode_batch_t* batch, const float* params,
                        void* wide_saved, const float* grad_loss, const float* grad_hT, float* grads,
                        void* workspace, void* stream);
/* elapsed ms of the encoder / Euler-chain / readout passes of the last njode_wide_forward, and of the
 * chain passes / dW pass of the last njode_wide_backward (timing on) */
int njode_wide_get_timing(float* enc_ms, float* ode_ms, float* ro_ms);
int njode_wide_get_timing_bwd(float* chain_ms, float* dw_ms);
/* debugging aid: clock64 stamps of CTA 0 of the forward Euler-chain kernel (buf: 4 x 256 uint64 device words, or NULL) */
void njode_wide_set_profile(void* buf);

/* Euler-Maruyama generators + Bernoulli observation mask (NJODE/stock_model.py:181-221, 288-335,
 * 356-375, 397-418 and NJODE/data_utils.py:73-81), one Philox-4x32-10 subsequence per global path id. */
enum { NJODE_SDE_BLACK_SCHOLES = 0, NJODE_SDE_ORNSTEIN_UHLENBECK = 1, NJODE_SDE_HESTON = 2,
       NJODE_SDE_HESTON_WO_FELLER = 3 };

typedef struct njode_sde {
    int32_t model;             /* NJODE_SDE_* */
    int32_t dimension;         /* independent copies per path (np.size(S0), stock_model.py:28) */
    int32_t nb_steps;
    int32_t return_vol;        /* HestonWOFeller: append variance coordinates (stock_model.py:329-330) */
    double  drift, volatility, mean, speed, correlation, v0, maturity;
    double  sine_coeff;        /* periodic_coeff(t) = 1 + sin(sine_coeff t); NaN = constant 1 (stock_model.py:29-32) */
    double  obs_perc;
    double  t0;                /* time of column 0 (combined datasets chain models, data_utils.py:141-157) */
    uint64_t seed;
} njode_sde_t;

/*   first_path  global id of path 0 of this call (data are identical for any sharding)
 *   S0          [dimension] start values, or start_X [n_paths, dimension] when per_path_start != 0
 *   paths       [n_paths, out_dim, nb_steps+1] float64 out (out_dim = dimension * (1 + return_vol))
 *   observed    [n_paths, nb_steps+1] int32 out or NULL; nb_obs [n_paths] int32 out or NULL */
int njode_sde_generate(const njode_sde_t* sde, int64_t first_path, int64_t n_paths,
                       const double* S0, int per_path_start, double* paths, int32_t* observed,
                       int32_t* nb_obs, void* stream);

/* on-device restatement of custom_collate_fn (NJODE/data_utils.py:278-316) for a set of paths of a
 * device-resident dataset: rows ordered (time ascending, batch position ascending).
 *   sel [B] dataset indices of the batch; outputs sized for the worst case N <= B * nb_steps.
 *   counts_out [2] = {K, N} (device) */
int njode_collate(const double* paths, const int32_t* observed, int64_t n_paths_total, int32_t dim,
                  int32_t nb_steps, const int64_t* sel, int32_t B, float* X, int32_t* obs_idx,
                  int32_t* time_ptr, int32_t* time_idx, float* start_X, int32_t* n_obs_ot,
                  int32_t* counts_out, void* workspace, int64_t workspace_bytes, void* stream);

/* analytic conditional-expectation path of the data-generating model on the event schedule of a return_path batch
 * (StockModel.compute_cond_exp, NJODE/stock_model.py:50-151 with next_cond_exp 178-179, 277-286, 353-354, 393-395):
 * the reference target that NJODE.evaluate (NJODE/models.py:551-558) compares the model's path_y with.
 *   batch   the prepared batch of the forward(return_path=True) call (E > 0, whole-path units)
 *   path_y  [E, B, d] out, d = dimension * (1 + return_vol), same record order as NJODE.forward's path_t */
int njode_cond_exp(const njode_sde_t* sde, const njode_batch_t* batch, float* path_y, void* stream);

/* per-path CSR of observation rows + sorted work units of one batch (the njode_batch_t index arrays), built on the device
 * from the raw collate arrays: replaces the per-observation-time slicing of X / obs_idx by time_ptr that NJODE.forward does
 * in Python (NJODE/models.py:449-456) and the n_obs_ot bookkeeping of NJODE/train.py:501-507.
 *   obs [N] path index of every row (rows time-major, path ascending inside a time: NJODE/data_utils.py:298-307),
 *   time_ptr [K+1], jump_step [K] (host schedule); segments != 0: (path, inter-observation segment) units, else whole paths
 *   path_ptr [B+1], path_rows [N], row_jump [N], unit_desc [(segments ? N + B : B) * 6] out
 *   stats [6] out: units of length >= T1 / >= T2 among the loss units, among the tail units (njode_batch_t.seg_n1/seg_n2),
 *                  duplicate (time, path) flag, index-out-of-range flag */
int64_t njode_index_workspace_bytes(int32_t N, int32_t B);
int njode_build_index(const int32_t* obs, int32_t N, const int32_t* time_ptr, int32_t K, const int32_t* jump_step,
                      int32_t B, int32_t S, int32_t segments, int32_t T1, int32_t T2,
                      int32_t* path_ptr, int32_t* path_rows, int32_t* row_jump, int32_t* unit_desc, int32_t* stats,
                      void* workspace, int64_t workspace_bytes, void* stream);

/* ---- measurement helpers (bench.py) -------------------------------------------------------- */
/* device-side timing of the main forward / backward kernel of the most recent call (cudaEvents on
 * the launching stream); enable with njode_set_timing(1) or NJODE_TIMING=1. */
void njode_set_timing(int on);
int njode_get_timing(float* fwd_ms, float* bwd_ms);
/* kernels launched by this library since it was loaded (bench.py: gpu_launches) */
long long njode_launch_count(void);
/* name of the main kernel launched by the most recent forward (which = 0) / backward (which = 1) call of either path */
const char* njode_last_kernel(int which);
/* fp32 FMA-pipe microbenchmark: dependent chains of FFMA on every SM; *fmas = lane-FMAs issued */
int njode_fma_peak_launch(float* scratch, int iters, double* fmas, void* stream);
/* legacy tensor path (mma.sync m16n8k8 tf32) microbenchmark, 8 independent accumulator tiles per warp on every SM;
 * *macs = multiply-accumulates issued.  Measures whether a 3xTF32 mma.sync dW phase could beat the FFMA pipe. */
int njode_mma_tf32_peak_launch(float* scratch, int iters, double* macs, void* stream);
/* writes `bytes` of buf (size it > L2) so the next kernel starts with a cold L2 */
int njode_l2_flush(void* buf, int64_t bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NJODE_B200_H */
