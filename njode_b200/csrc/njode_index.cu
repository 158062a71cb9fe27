// njode_index.cu -- the per-path CSR of observation rows and the sorted work units of one batch, built on the
// device from the raw collate arrays (obs_idx, time_ptr) and the host schedule (jump_step): njode_build_index.
//
// Replaces, as one call, the per-batch bookkeeping the reference does in Python inside NJODE.forward
// (NJODE/models.py:449-456: slicing X / obs_idx by time_ptr at every observation time) and train.py:501-507, restated
// for a kernel that marches (path, segment) units instead of observation times: see njode_batch_t in
// include/njode_b200.h for the layout.  Results are bit-identical to the NumPy builder
// njode_b200/schedule.py::build_csr + build_units (tests/test_gpu_index.py).
//
// All of it is HBM-bound integer work on N ~ 10 rows per path: two histogram/scan passes and three stable LSD radix
// sorts (cub::DeviceRadixSort on the minimal number of key bits), ~10 launches instead of ~65 tensor-op launches.
#include <cuda_runtime.h>
#include <stdint.h>
#include <cub/cub.cuh>
#include "../../include/njode_b200.h"

int nj_set_error(int code, const char* msg);       // njode_api.cu
void nj_count_launches(int n);                     // njode_api.cu

namespace {

struct IdxWs {          // workspace carve-up (int32 words unless noted)
    int32_t *counts, *sorted_path, *iota, *key_l, *key_l_s, *o1, *key_t, *key_t_s, *o2, *loss_tmp, *tail_tmp;
    void* cub_tmp; size_t cub_bytes;
    size_t total;
};

inline size_t al(size_t x) { return (x + 255) & ~(size_t)255; }

IdxWs idx_layout(char* base, int N, int B) {
    IdxWs w;
    size_t o = 0;
    auto take = [&](size_t words) { int32_t* p = reinterpret_cast<int32_t*>(base + o); o += al(words * 4); return p; };
    const int Nn = N > 0 ? N : 1, Bn = B > 0 ? B : 1;
    w.counts = take(Bn + 1);
    w.sorted_path = take(Nn); w.iota = take(Nn > Bn ? Nn : Bn);
    w.key_l = take(Nn); w.key_l_s = take(Nn); w.o1 = take(Nn);
    w.key_t = take(Bn); w.key_t_s = take(Bn); w.o2 = take(Bn);
    w.loss_tmp = take((size_t)Nn * 6); w.tail_tmp = take((size_t)Bn * 6);
    size_t b1 = 0, b2 = 0, b3 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b1, (const int32_t*)nullptr, (int32_t*)nullptr, (const int32_t*)nullptr, (int32_t*)nullptr, Nn, 0, 32);
    cub::DeviceRadixSort::SortPairs(nullptr, b2, (const int32_t*)nullptr, (int32_t*)nullptr, (const int32_t*)nullptr, (int32_t*)nullptr, Bn, 0, 32);
    cub::DeviceScan::InclusiveSum(nullptr, b3, (const int32_t*)nullptr, (int32_t*)nullptr, Bn + 1);
    w.cub_bytes = al(b1 > b2 ? (b1 > b3 ? b1 : b3) : (b2 > b3 ? b2 : b3));
    w.cub_tmp = base + o; o += w.cub_bytes;
    w.total = o;
    return w;
}

inline int bits_for(int n) { int b = 1; while (b < 31 && (1 << b) < n) ++b; return b; }

// stats: [0..3] units of length >= T1 / >= T2 in the loss run and the tail run, [4] duplicate (time, path), [5] bad index
__global__ void idx_rows_kernel(const int32_t* __restrict__ obs, int N, const int32_t* __restrict__ time_ptr, int K, int B,
                                int32_t* __restrict__ counts, int32_t* __restrict__ obs_c, int32_t* __restrict__ iota,
                                int32_t* __restrict__ row_jump, int32_t* __restrict__ stats) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= N) return;
    int p = obs[r];
    if (p < 0 || p >= B) { atomicOr(stats + 5, 1); p = p < 0 ? 0 : B - 1; }
    obs_c[r] = p;
    iota[r] = r;
    atomicAdd(counts + p + 1, 1);
    // row_jump[r] = (number of time_ptr entries <= r) - 1 : index of the observation time the row belongs to
    int lo = 0, hi = K + 1;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (time_ptr[mid] <= r) lo = mid + 1; else hi = mid; }
    row_jump[r] = lo - 1;
}

__global__ void idx_iota_kernel(int32_t* __restrict__ iota, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) iota[i] = i;
}

__device__ __forceinline__ void warp_count(bool pred, int32_t* dst) {
    const unsigned m = __ballot_sync(0xFFFFFFFFu, pred);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(dst, __popc(m));
}

// one thread per sorted position q (rows of a path are contiguous and in time order): the loss unit that ends with row q
__global__ void idx_loss_units_kernel(const int32_t* __restrict__ sorted_path, const int32_t* __restrict__ path_rows,
                                      const int32_t* __restrict__ row_jump, const int32_t* __restrict__ jump_step, int N, int S,
                                      int T1, int T2, int32_t* __restrict__ loss_tmp, int32_t* __restrict__ key_l,
                                      int32_t* __restrict__ stats) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    bool ge1 = false, ge2 = false;
    if (q < N) {
        const int p = sorted_path[q], r = path_rows[q], i = row_jump[r], js = jump_step[i];
        const bool first = q == 0 || sorted_path[q - 1] != p;
        int prev_js = 0, prev_row = -1;
        if (!first) {
            prev_row = path_rows[q - 1];
            const int pi = row_jump[prev_row];
            prev_js = jump_step[pi];
            if (pi == i) atomicOr(stats + 4, 1);          // at most one row per (time, path): NJODE/data_utils.py:302-306
        }
        int32_t* u = loss_tmp + (size_t)q * 6;
        u[0] = p; u[1] = prev_js; u[2] = js; u[3] = q; u[4] = q + 1; u[5] = prev_row + 1;
        const int len = js - prev_js;
        int key = S - len; key = key < 0 ? 0 : (key > S ? S : key);
        key_l[q] = key;
        ge1 = len >= T1; ge2 = len >= T2;
    }
    warp_count(ge1, stats + 0);
    warp_count(ge2, stats + 1);
}

// one thread per path: the tail unit after the path's last observation (or the whole path when it has none)
__global__ void idx_tail_units_kernel(const int32_t* __restrict__ path_ptr, const int32_t* __restrict__ path_rows,
                                      const int32_t* __restrict__ row_jump, const int32_t* __restrict__ jump_step, int B, int S,
                                      int T1, int T2, int32_t* __restrict__ tail_tmp, int32_t* __restrict__ key_t,
                                      int32_t* __restrict__ stats) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    bool ge1 = false, ge2 = false;
    if (p < B) {
        const int a = path_ptr[p], b = path_ptr[p + 1];
        int s0 = 0, start = 0;
        if (b > a) { const int r = path_rows[b - 1]; s0 = jump_step[row_jump[r]]; start = r + 1; }
        int32_t* u = tail_tmp + (size_t)p * 6;
        u[0] = p; u[1] = s0; u[2] = S; u[3] = b; u[4] = b; u[5] = start | NJODE_UNIT_WRITES_HT;
        const int len = S - s0;
        int key = S - len; key = key < 0 ? 0 : (key > S ? S : key);
        key_t[p] = key;
        ge1 = len >= T1; ge2 = len >= T2;
    }
    warp_count(ge1, stats + 2);
    warp_count(ge2, stats + 3);
}

// whole-path units (masked model, return_path): unit p = path p over all steps with all its rows
__global__ void idx_path_units_kernel(const int32_t* __restrict__ path_ptr, int B, int S, int32_t* __restrict__ units) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B) return;
    int32_t* u = units + (size_t)p * 6;
    u[0] = p; u[1] = 0; u[2] = S; u[3] = path_ptr[p]; u[4] = path_ptr[p + 1]; u[5] = NJODE_UNIT_WRITES_HT;
}

__global__ void idx_dup_kernel(const int32_t* __restrict__ sorted_path, const int32_t* __restrict__ path_rows,
                               const int32_t* __restrict__ row_jump, int N, int32_t* __restrict__ stats) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < 1 || q >= N) return;
    if (sorted_path[q - 1] == sorted_path[q] && row_jump[path_rows[q - 1]] == row_jump[path_rows[q]]) atomicOr(stats + 4, 1);
}

// units[j] = src[order[j]] (6 words each); one thread per word
__global__ void idx_gather_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ order, int n,
                                  int32_t* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * 6) return;
    const int j = i / 6, f = i - j * 6;
    dst[i] = src[(size_t)order[j] * 6 + f];
}

}  // namespace

extern "C" int64_t njode_index_workspace_bytes(int32_t N, int32_t B) {
    if (N < 0 || B < 0) return -1;
    return (int64_t)idx_layout(nullptr, N, B).total;
}

extern "C" int njode_build_index(const int32_t* obs, int32_t N, const int32_t* time_ptr, int32_t K,
                                 const int32_t* jump_step, int32_t B, int32_t S, int32_t segments, int32_t T1, int32_t T2,
                                 int32_t* path_ptr, int32_t* path_rows, int32_t* row_jump, int32_t* unit_desc,
                                 int32_t* stats, void* workspace, int64_t workspace_bytes, void* stream) {
    if (N < 0 || B < 1 || K < 0 || S < 0) return nj_set_error(-1, "njode_build_index: bad sizes");
    if (!path_ptr || !path_rows || !row_jump || !unit_desc || !stats || !workspace || (N > 0 && (!obs || !time_ptr)))
        return nj_set_error(-1, "njode_build_index: null argument");
    IdxWs w = idx_layout(reinterpret_cast<char*>(workspace), N, B);
    if ((int64_t)w.total > workspace_bytes) return nj_set_error(-1, "njode_build_index: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int T = 256;
    int launches = 0;
    cudaMemsetAsync(w.counts, 0, (size_t)(B + 1) * 4, st);
    cudaMemsetAsync(stats, 0, 6 * 4, st);
    if (N > 0) { idx_rows_kernel<<<(N + T - 1) / T, T, 0, st>>>(obs, N, time_ptr, K, B, w.counts, w.key_l, w.iota, row_jump, stats); ++launches; }
    size_t cb = w.cub_bytes;
    cub::DeviceScan::InclusiveSum(w.cub_tmp, cb, w.counts, path_ptr, B + 1, st); ++launches;
    if (N > 0) {
        // stable sort of the rows by path: rows of a path come out contiguous and in time order (rows are time-major)
        cb = w.cub_bytes;
        cub::DeviceRadixSort::SortPairs(w.cub_tmp, cb, w.key_l, w.sorted_path, w.iota, path_rows, N, 0, bits_for(B), st); launches += 2;
    }
    if (!segments) {
        idx_path_units_kernel<<<(B + T - 1) / T, T, 0, st>>>(path_ptr, B, S, unit_desc); ++launches;
        if (N > 1) { idx_dup_kernel<<<(N + T - 1) / T, T, 0, st>>>(w.sorted_path, path_rows, row_jump, N, stats); ++launches; }
    } else {
        if (N > 0) {
            idx_loss_units_kernel<<<(N + T - 1) / T, T, 0, st>>>(w.sorted_path, path_rows, row_jump, jump_step, N, S, T1, T2,
                                                                  w.loss_tmp, w.key_l, stats); ++launches;
        }
        idx_tail_units_kernel<<<(B + T - 1) / T, T, 0, st>>>(path_ptr, path_rows, row_jump, jump_step, B, S, T1, T2,
                                                              w.tail_tmp, w.key_t, stats); ++launches;
        // longest first (descending length, stable): ascending stable sort of S - length
        const int kb = bits_for(S + 1);
        if (N > 0) {
            idx_iota_kernel<<<(N + T - 1) / T, T, 0, st>>>(w.iota, N);
            cb = w.cub_bytes;
            cub::DeviceRadixSort::SortPairs(w.cub_tmp, cb, w.key_l, w.key_l_s, w.iota, w.o1, N, 0, kb, st);
            idx_gather_kernel<<<(N * 6 + T - 1) / T, T, 0, st>>>(w.loss_tmp, w.o1, N, unit_desc);
            launches += 4;
        }
        idx_iota_kernel<<<(B + T - 1) / T, T, 0, st>>>(w.iota, B);
        cb = w.cub_bytes;
        cub::DeviceRadixSort::SortPairs(w.cub_tmp, cb, w.key_t, w.key_t_s, w.iota, w.o2, B, 0, kb, st);
        idx_gather_kernel<<<(B * 6 + T - 1) / T, T, 0, st>>>(w.tail_tmp, w.o2, B, unit_desc + (size_t)N * 6);
        launches += 4;
    }
    nj_count_launches(launches);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return nj_set_error(-2, cudaGetErrorString(e));
    return 0;
}
