// njode_sde.cu -- on-device dataset generation for the NJ-ODE hot path:
//   * Euler-Maruyama generators of NJODE/stock_model.py (BlackScholes 356-375, OrnsteinUhlenbeck
//     397-418, Heston 181-221, HestonWOFeller 288-335) + the Bernoulli observation mask of
//     NJODE/data_utils.py:73-81 -- njode_sde_generate
//   * the collate of NJODE/data_utils.py:278-316 for a batch drawn from a device-resident dataset
//     -- njode_collate
// Randomness: Philox-4x32-10 (Salmon et al. 2011), key = seed, counter = (path id lo, hi, step,
// stream), so every path depends only on (seed, global path id): data are identical for any
// sharding over GPUs.  Normals by Box-Muller in fp64 from 32-bit uniforms.  The reference draws from
// numpy's Mersenne Twister in (path, step) order, which cannot be reproduced on a GPU; parity is
// pinned (a) bit-for-bit against oracle/sde_oracle.py, which restates the update rules on the same
// Philox stream, and (b) in distribution against the reference generator and the closed-form moments
// of the Euler scheme (tests/test_sde_*.py).
//
// Both kernels are HBM-bound byte movers: the generator writes paths*dim*(steps+1)*8 B (+ the mask);
// path rows are staged through shared memory so that global stores are 128 B coalesced.
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <string>
#include "../../include/njode_b200.h"

extern "C" const char* njode_last_error(void);
int nj_set_error(int code, const char* msg);       // njode_api.cu
void nj_count_launches(int n);                     // njode_api.cu

// ------------------------------------------------------------------------------------------------
// Philox-4x32-10
// ------------------------------------------------------------------------------------------------
struct nj_u4 { unsigned x, y, z, w; };

__host__ __device__ inline nj_u4 nj_philox(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1) {
    for (int r = 0; r < 10; ++r) {
        const unsigned long long p0 = 0xD2511F53ull * c0, p1 = 0xCD9E8D57ull * c2;
        const unsigned hi0 = (unsigned)(p0 >> 32), lo0 = (unsigned)p0, hi1 = (unsigned)(p1 >> 32), lo1 = (unsigned)p1;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    nj_u4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

#define NJ_STREAM_MASK 0xFFFFFFFFu

__device__ inline double nj_box_muller(unsigned a, unsigned b) {
    const double u1 = ((double)a + 1.0) * (1.0 / 4294967296.0);      // (0, 1]
    const double u2 = (double)b * (1.0 / 4294967296.0);              // [0, 1)
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

// ------------------------------------------------------------------------------------------------
// generator: one lane per (path, coordinate) row; a warp stages 32 rows x 32 steps in shared memory
// ------------------------------------------------------------------------------------------------
struct NjSdeArgs {
    njode_sde_t p;
    long long first_path, n_paths;
    const double* S0; int per_path_start;
    double* paths; int32_t* observed; int32_t* nb_obs;
};

__device__ inline double nj_coeff(const njode_sde_t& p, double t) {
    return isnan(p.sine_coeff) ? 1.0 : 1.0 + sin(p.sine_coeff * t);       // NJODE/stock_model.py:29-32
}

__global__ void __launch_bounds__(128) nj_sde_kernel(const __grid_constant__ NjSdeArgs a) {
    __shared__ double tile[4][32][33];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const njode_sde_t& p = a.p;
    const int dim = p.dimension, n1 = p.nb_steps + 1;
    const int out_dim = dim * (p.return_vol ? 2 : 1);
    const long long n_rows = a.n_paths * dim;
    const long long row0 = ((long long)blockIdx.x * 4 + warp) * 32;
    const long long row = row0 + lane;
    const bool valid = row < n_rows;
    const long long pl = valid ? row / dim : 0;       // local path
    const int j = valid ? (int)(row % dim) : 0;
    const unsigned long long gid = (unsigned long long)(a.first_path + pl);
    const unsigned k0 = (unsigned)(p.seed & 0xFFFFFFFFull), k1 = (unsigned)(p.seed >> 32);
    const double dt = p.maturity / (double)p.nb_steps, sq = sqrt(dt);
    const bool heston = p.model == NJODE_SDE_HESTON, hwf = p.model == NJODE_SDE_HESTON_WO_FELLER;
    double S = 0.0, v = 0.0;
    if (valid) {
        S = a.per_path_start ? a.S0[pl * dim + j] : a.S0[j];
        v = heston ? p.mean : p.v0;
    }
    const double rho2 = sqrt(1.0 - p.correlation * p.correlation);
    for (int kc = 0; kc < n1; kc += 32) {
        // each lane advances its row through up to 32 columns
        for (int kk = 0; kk < 32 && kc + kk < n1; ++kk) {
            const int k = kc + kk;
            if (k > 0 && valid) {
                const nj_u4 r = nj_philox((unsigned)gid, (unsigned)(gid >> 32), (unsigned)k, (unsigned)j, k0, k1);
                const double nrm1 = nj_box_muller(r.x, r.y);
                const double dW = nrm1 * sq;
                const double tprev = p.t0 + (double)(k - 1) * dt;
                if (p.model == NJODE_SDE_BLACK_SCHOLES) {                    // stock_model.py:371-374
                    S = S + p.drift * nj_coeff(p, tprev) * S * dt + p.volatility * S * dW;
                } else if (p.model == NJODE_SDE_ORNSTEIN_UHLENBECK) {        // stock_model.py:414-417
                    S = S + (-p.speed * nj_coeff(p, tprev) * (S - p.mean)) * dt + p.volatility * dW;
                } else {
                    const double nrm2 = nj_box_muller(r.z, r.w);
                    const double dZ = (p.correlation * nrm1 + rho2 * nrm2) * sq;          // stock_model.py:206-207
                    if (heston) {                                            // stock_model.py:209-219: spot uses the NEW variance
                        const double vn = v + (-p.speed * (v - p.mean)) * dt + p.volatility * sqrt(v) * dZ;
                        S = S + p.drift * nj_coeff(p, tprev) * S * dt + sqrt(vn) * S * dW;
                        v = vn;
                    } else {                                                 // stock_model.py:317-328: log-Euler, v+ = max(v, 0)
                        const double vp = fmax(v, 0.0);
                        S = exp(log(S) + (p.drift * nj_coeff(p, tprev) - 0.5 * vp) * dt + sqrt(vp) * dW);
                        v = v + (-p.speed * (vp - p.mean)) * dt + p.volatility * sqrt(vp) * dZ;
                    }
                }
            }
            tile[warp][lane][kk] = S;
            // variance coordinates (HestonWOFeller with return_vol, stock_model.py:329-330): rare, stored directly
            if (valid && p.return_vol && hwf) a.paths[(pl * out_dim + dim + j) * n1 + k] = v;
        }
        __syncwarp();
        // coalesced stores: lane = column
        const int k = kc + lane;
        for (int rr = 0; rr < 32; ++rr) {
            const long long r2 = row0 + rr;
            if (r2 >= n_rows || k >= n1) continue;
            const long long pl2 = r2 / dim; const int j2 = (int)(r2 % dim);
            a.paths[(pl2 * out_dim + j2) * n1 + k] = tile[warp][rr][lane];
        }
        __syncwarp();
    }
}

// observation mask: observed[p][k] = (u < obs_perc) for every column incl. column 0, which the reference also draws at
// random and never uses as an observation (NJODE/data_utils.py:79-81; the collate starts at column 1, 292-307)
__global__ void __launch_bounds__(256) nj_mask_kernel(const __grid_constant__ NjSdeArgs a) {
    const int n1 = a.p.nb_steps + 1;
    const unsigned k0 = (unsigned)(a.p.seed & 0xFFFFFFFFull), k1 = (unsigned)(a.p.seed >> 32);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long pl = (long long)blockIdx.x * 8 + warp;
    if (pl >= a.n_paths) return;
    const unsigned long long gid = (unsigned long long)(a.first_path + pl);
    int cnt = 0;
    for (int kb = 0; kb < n1; kb += 128) {
        const int k4 = (kb >> 2) + lane;               // one Philox call serves 4 columns
        const nj_u4 r = nj_philox((unsigned)gid, (unsigned)(gid >> 32), (unsigned)k4, NJ_STREAM_MASK, k0, k1);
        const unsigned w[4] = {r.x, r.y, r.z, r.w};
        int4 o;
        int* op = &o.x;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = 4 * k4 + q;
            int ob = ((double)w[q] * (1.0 / 4294967296.0)) < a.p.obs_perc ? 1 : 0;
            op[q] = ob;
            if (k >= 1 && k < n1) cnt += ob;
        }
        if (a.observed) {
            for (int q = 0; q < 4; ++q) { const int k = 4 * k4 + q; if (k < n1) a.observed[pl * n1 + k] = op[q]; }
        }
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
    if (lane == 0 && a.nb_obs) a.nb_obs[pl] = cnt;      // nb_obs = observed[:, 1:].sum(1), data_utils.py:81
}

extern "C" int njode_sde_generate(const njode_sde_t* sde, int64_t first_path, int64_t n_paths, const double* S0,
                                  int per_path_start, double* paths, int32_t* observed, int32_t* nb_obs, void* stream) {
    if (!sde || !S0 || !paths) return nj_set_error(-1, "njode_sde_generate: null argument");
    if (sde->model < 0 || sde->model > NJODE_SDE_HESTON_WO_FELLER) return nj_set_error(-1, "njode_sde_generate: unknown model");
    if (sde->dimension < 1 || sde->nb_steps < 1 || n_paths < 0) return nj_set_error(-1, "njode_sde_generate: bad sizes");
    if (n_paths == 0) return 0;
    NjSdeArgs a;
    a.p = *sde; a.first_path = first_path; a.n_paths = n_paths; a.S0 = S0; a.per_path_start = per_path_start;
    a.paths = paths; a.observed = observed; a.nb_obs = nb_obs;
    cudaStream_t st = (cudaStream_t)stream;
    const long long rows = n_paths * sde->dimension;
    nj_sde_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, st>>>(a);
    if (observed || nb_obs) nj_mask_kernel<<<(unsigned)((n_paths + 7) / 8), 256, 0, st>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return nj_set_error(-2, cudaGetErrorString(e));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// collate: rows ordered (grid time ascending, batch position ascending), NJODE/data_utils.py:292-315
// ------------------------------------------------------------------------------------------------
// pass 1: per warp (32 consecutive batch positions) and grid column t: number of observations
__global__ void __launch_bounds__(256) nj_collate_count(const int32_t* __restrict__ observed, const int64_t* __restrict__ sel,
                                                        int B, int n1, int32_t* __restrict__ warp_cnt, int32_t* __restrict__ n_obs_ot) {
    const int w = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int nwarps = (B + 31) / 32;
    if (w >= nwarps) return;
    const int b = w * 32 + lane;
    const int32_t* row = b < B ? observed + sel[b] * (long long)n1 : nullptr;
    int total = 0;
    for (int t = 1; t < n1; ++t) {
        const int ob = row ? (row[t] == 1) : 0;
        const unsigned m = __ballot_sync(0xFFFFFFFFu, ob);
        if (lane == 0) warp_cnt[(size_t)(t - 1) * nwarps + w] = __popc(m);
        total += ob;
    }
    if (b < B && n_obs_ot) n_obs_ot[b] = total;
}

// pass 2: one block per grid column: exclusive scan of the warp counts (in place), column total
__global__ void __launch_bounds__(256) nj_collate_scan_cols(int32_t* __restrict__ warp_cnt, int nwarps, int32_t* __restrict__ col_cnt) {
    __shared__ int sh[256];
    __shared__ int carry;
    int32_t* col = warp_cnt + (size_t)blockIdx.x * nwarps;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nwarps; base += 256) {
        const int i = base + threadIdx.x;
        const int v = i < nwarps ? col[i] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 256; o <<= 1) {
            const int x = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
            __syncthreads();
            sh[threadIdx.x] += x;
            __syncthreads();
        }
        const int incl = sh[threadIdx.x];
        if (i < nwarps) col[i] = carry + incl - v;
        __syncthreads();
        if (threadIdx.x == 255) carry += incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) col_cnt[blockIdx.x] = carry;
}

// pass 3: single block: compact the observed grid columns -> time_idx, time_ptr, {K, N}
__global__ void __launch_bounds__(256) nj_collate_scan_times(const int32_t* __restrict__ col_cnt, int nsteps, int32_t* __restrict__ time_ptr,
                                                             int32_t* __restrict__ time_idx, int32_t* __restrict__ col_slot,
                                                             int32_t* __restrict__ col_off, int32_t* __restrict__ counts) {
    // nsteps is small (<= a few thousand): serial scan by one thread keeps it simple and exact
    if (threadIdx.x != 0) return;
    int K = 0, N = 0;
    time_ptr[0] = 0;
    for (int t = 0; t < nsteps; ++t) {
        const int c = col_cnt[t];
        col_off[t] = N;
        if (c > 0) { time_idx[K] = t + 1; col_slot[t] = K; N += c; ++K; time_ptr[K] = N; }
        else col_slot[t] = -1;
    }
    counts[0] = K; counts[1] = N;
}

// pass 4: scatter rows
__global__ void __launch_bounds__(256) nj_collate_scatter(const double* __restrict__ paths, const int32_t* __restrict__ observed,
                                                          const int64_t* __restrict__ sel, int B, int dim, int n1,
                                                          const int32_t* __restrict__ warp_off, const int32_t* __restrict__ col_off,
                                                          float* __restrict__ X, int32_t* __restrict__ obs_idx, float* __restrict__ start_X) {
    const int w = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int nwarps = (B + 31) / 32;
    if (w >= nwarps) return;
    const int b = w * 32 + lane;
    const long long src = b < B ? sel[b] : 0;
    const int32_t* row = b < B ? observed + src * n1 : nullptr;
    const double* pp = paths + src * (long long)dim * n1;
    if (b < B) for (int j = 0; j < dim; ++j) start_X[(size_t)b * dim + j] = (float)pp[(size_t)j * n1];
    for (int t = 1; t < n1; ++t) {
        const int ob = row ? (row[t] == 1) : 0;
        const unsigned m = __ballot_sync(0xFFFFFFFFu, ob);
        if (ob) {
            const int r = col_off[t - 1] + warp_off[(size_t)(t - 1) * nwarps + w] + __popc(m & ((1u << lane) - 1u));
            obs_idx[r] = b;
            for (int j = 0; j < dim; ++j) X[(size_t)r * dim + j] = (float)pp[(size_t)j * n1 + t];
        }
    }
}

extern "C" int njode_collate(const double* paths, const int32_t* observed, int64_t n_paths_total, int32_t dim,
                             int32_t nb_steps, const int64_t* sel, int32_t B, float* X, int32_t* obs_idx,
                             int32_t* time_ptr, int32_t* time_idx, float* start_X, int32_t* n_obs_ot,
                             int32_t* counts_out, void* workspace, int64_t workspace_bytes, void* stream) {
    (void)n_paths_total;
    if (!paths || !observed || !sel || !X || !obs_idx || !time_ptr || !time_idx || !start_X || !counts_out || !workspace)
        return nj_set_error(-1, "njode_collate: null argument");
    if (B < 1 || dim < 1 || nb_steps < 1) return nj_set_error(-1, "njode_collate: bad sizes");
    const int nwarps = (B + 31) / 32, n1 = nb_steps + 1;
    const size_t need = ((size_t)nb_steps * nwarps + 3 * (size_t)nb_steps + 16) * 4;
    if ((size_t)workspace_bytes < need) return nj_set_error(-1, "njode_collate: workspace too small (need (nb_steps*ceil(B/32) + 3*nb_steps + 16) * 4 bytes)");
    int32_t* warp_cnt = (int32_t*)workspace;
    int32_t* col_cnt = warp_cnt + (size_t)nb_steps * nwarps;
    int32_t* col_slot = col_cnt + nb_steps;
    int32_t* col_off = col_slot + nb_steps;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned gw = (unsigned)((nwarps + 7) / 8);
    nj_collate_count<<<gw, 256, 0, st>>>(observed, sel, B, n1, warp_cnt, n_obs_ot);
    nj_collate_scan_cols<<<nb_steps, 256, 0, st>>>(warp_cnt, nwarps, col_cnt);
    nj_collate_scan_times<<<1, 32, 0, st>>>(col_cnt, nb_steps, time_ptr, time_idx, col_slot, col_off, counts_out);
    nj_collate_scatter<<<gw, 256, 0, st>>>(paths, observed, sel, B, dim, n1, warp_cnt, col_off, X, obs_idx, start_X);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return nj_set_error(-2, cudaGetErrorString(e));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// analytic conditional expectation path (StockModel.compute_cond_exp, NJODE/stock_model.py:50-151, with the models'
// next_cond_exp: 178-179, 277-286, 353-354, 393-395) on the event schedule of a return_path batch:
//   y <- start_X;  every Euler step: y <- E[X_{t+dt} | X_t = y];  every observation: y[i_obs] <- X_obs;
// one record of y for the whole batch after every step and every observation time, in the order of NJODE.forward's
// path_t.  One thread per (path, coordinate): each coordinate evolves on its own (growth y e^{mu c(t) dt}, or mean
// reversion y e + m (1 - e), e = e^{-kappa c(t) dt}; HestonWOFeller's variance coordinates revert without c(t)).
// fp64 arithmetic, fp32 records.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) nj_cond_exp_kernel(const __grid_constant__ njode_sde_t sde, const __grid_constant__ njode_batch_t b,
                                                          int d, float* __restrict__ path_y) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= b.B * d) return;
    const int p = idx / d, c = idx - p * d;
    const bool periodic = !isnan(sde.sine_coeff);
    // coordinate kind: 0 growth, 1 mean reversion with the periodic coefficient (OU), 2 plain mean reversion (variance)
    int kind = sde.model == NJODE_SDE_ORNSTEIN_UHLENBECK ? 1 : 0;
    if (sde.model == NJODE_SDE_HESTON_WO_FELLER && sde.return_vol && c >= sde.dimension) kind = 2;
    double y = (double)b.start_X[idx];
    const size_t rec = (size_t)b.B * d;
    path_y[idx] = (float)y;
    int cur = b.path_ptr[p];
    const int cend = b.path_ptr[p + 1];
    int gi = 0;
    for (int k = 0; k <= b.S; ++k) {
        while (gi < b.K && b.jump_step[gi] == k) {
            if (cur < cend) {
                const int r = b.path_rows[cur];
                if (b.row_jump[r] == gi) { y = (double)b.X[(size_t)r * d + c]; ++cur; }
            }
            path_y[(size_t)b.jump_event[gi] * rec + idx] = (float)y;
            ++gi;
        }
        if (k == b.S) break;
        const double dt = (double)b.step_dt[k], t = (double)b.step_t[k];
        const double coef = periodic ? 1.0 + sin(sde.sine_coeff * t) : 1.0;
        if (kind == 0) y = y * exp(sde.drift * coef * dt);
        else {
            const double e = exp(-sde.speed * (kind == 1 ? coef : 1.0) * dt);
            y = y * e + sde.mean * (1.0 - e);
        }
        path_y[(size_t)b.step_event[k] * rec + idx] = (float)y;
    }
}

extern "C" int njode_cond_exp(const njode_sde_t* sde, const njode_batch_t* batch, float* path_y, void* stream) {
    if (!sde || !batch || !path_y) return nj_set_error(-1, "njode_cond_exp: null argument");
    if (batch->E <= 0 || batch->unit_kind != 0) return nj_set_error(-1, "njode_cond_exp: needs the event schedule of a return_path batch");
    if (sde->model < 0 || sde->model > NJODE_SDE_HESTON_WO_FELLER) return nj_set_error(-1, "njode_cond_exp: unknown model");
    const int d = sde->dimension * (sde->return_vol ? 2 : 1);
    const long long n = (long long)batch->B * d;
    if (n <= 0) return 0;
    nj_cond_exp_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*sde, *batch, d, path_y);
    nj_count_launches(1);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return nj_set_error(-2, cudaGetErrorString(e));
    return 0;
}
