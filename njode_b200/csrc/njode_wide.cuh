// njode_wide.cuh -- tensor-core ("wide") path of the NJ-ODE hot path for sm_100a: tcgen05.mma with
// TMEM accumulators, weights streamed through shared memory by TMA bulk copies, bf16 operands and
// fp32 accumulation.  Used when the MLPs are real dense contractions (BASELINE config 5:
// d=16, H=256, 4x256 tanh nets); the narrow demo nets stay on the fp32 FMA kernels (njode_seg.cuh).
//
// Same segment decomposition as njode_seg.cuh (NJODE/models.py:463-470: h after a jump = enc(X_obs)):
// every (path, inter-observation segment) is an independent unit.  One CTA marches a tile of 128
// units; unit r of the tile is TMEM lane r / row r of every operand tile.
//
//   pass ENC : h_start[u] = encoder(start value of unit u)                  (FFNN.forward, models.py:261-276)
//   pass ODE : Euler steps of the unit, h kept in TMEM columns [256,512) in fp32; h_hist / h_before / hT
//              written on the fly                                           (ode_step, models.py:369-377)
//   pass RO  : Y_bj = readout(h_before[row]), Y = readout(enc(X_obs)), loss row (compute_loss, models.py:71-126)
//
// Every Linear layer over the tile is D[128 x N] = A[128 x K] . W[N x K]^T:
//   * A (activations, bf16) lives in shared memory as K-blocks of 64 columns, each block [128 rows x 128 B]
//     in the canonical K-major SWIZZLE_128B layout (16-byte chunk c of row r stored at chunk c ^ (r & 7));
//     it is written by the epilogue threads of the previous layer (generic proxy -> fence.proxy.async).
//   * W comes from a bf16 image in global memory that nj_wide_pack_kernel has already laid out in exactly
//     that shared-memory image, one [n16 rows x 128 B] block per (layer, K-block), so a stage is filled by ONE
//     cp.async.bulk (TMA, no tensor map) completing on an mbarrier.
//   * D accumulates in TMEM columns [0,256) (tcgen05.mma.cta_group::1.kind::f16, M=128, N=n16, K=16 per
//     instruction, issued by one thread); the epilogue warps read it with tcgen05.ld.32x32b, add the bias,
//     apply tanh (MUFU.TANH) / relu and dropout, and write the next layer's A operand.
// Warp roles (384 threads): warp 0 = weight producer, warp 1 = MMA issuer + TMEM allocator, warps 4..11 =
// epilogue (warp w reads TMEM lanes 32*(w%4).., column half (w-4)/4).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/njode_b200.h"
#include "njode_hash.cuh"

namespace njw {

constexpr int TILE_M = 128;
constexpr int NSTAGE = 5;
constexpr int STAGE_BYTES = 128 * 128;          // one K-block of one 128-row half of a weight matrix
constexpr int A_BLOCK_BYTES = TILE_M * 128;     // one K-block of the activation tile
constexpr int A_BLOCKS = 7;                     // 3 rotating pairs of main K-blocks + the auxiliary block
constexpr int AUX_BLOCK = 6;
#ifndef NJW_EPI_GROUPS
#define NJW_EPI_GROUPS 2
#endif
constexpr int EPI_GROUPS = NJW_EPI_GROUPS;                   // column groups of the epilogue: warp w = lane quadrant w % 4, group (w - 4) / 4
constexpr int EPI_WARPS = 4 * EPI_GROUPS;
constexpr int EPI_THREADS = 32 * EPI_WARPS;
constexpr int EPI_WARP0 = 4;
constexpr int NUM_THREADS = 32 * EPI_WARP0 + EPI_THREADS;
constexpr int MAX_W = 256;                      // widest layer / hidden state
constexpr int MAX_D = 16;                       // widest input / output
constexpr int AUX_X0 = 8;                       // auxiliary block: columns 0..5 = tau, tdiff, tau+tdiff as bf16 hi/lo
                                                // pairs, columns 8..8+d = tanh(x)
constexpr int TMEM_COLS = 512;
constexpr int TMEM_H = 256;                     // first TMEM column of the fp32 hidden state

enum { MODE_ENC = 0, MODE_ODE = 1, MODE_RO = 2 };
enum { KIND_PLAIN = 0, KIND_ODE0 = 1, KIND_ENC0 = 2 };

struct WLayer {
    int kb_main;        // K-blocks read from the main activation blocks
    int has_aux;        // + the auxiliary block (layer 0 of the ODE / encoder nets)
    int aux_ksteps;     // UMMA K-steps (of 16 columns) issued on the auxiliary block
    int n, n16;         // outputs, padded to the UMMA N granularity
    int in_dim;         // inputs of the nn.Linear
    int act, drop, kind;
    unsigned img_off;   // byte offset of the layer's bf16 image: for each 128-row half of the outputs, its K-blocks
                        // ([rows of the half x 128 B] each, main blocks then the auxiliary block)
    int bias_off;       // float offset inside the bias image
    long long w_src, b_src;
    // backward: acc[128 x k] = G_l[128 x n] . W_l[n x k]  (B operand = W_l^T, K-major over the outputs)
    unsigned wt_off;    // byte offset of the transposed image: for each 128-row half of nt16, kt_blocks K-blocks
    int kt_blocks;      // K-blocks of the transposed GEMM = ceil(n16 / 64)
    int kt_last_ksteps; // K-steps issued on the last of them
    int nt, nt16;       // N of the transposed GEMM: main inputs of the layer (ODE layer 0: H), padded to 16
    int act_off;        // byte offset of this layer's input image inside an activation record (forward spill)
    int g_off;          // byte offset of this layer's output-gradient image inside a gradient record
    int dw_slot0[2];    // first partial-accumulator slot of the (main, aux) dW items, -1 = no such item
    int dw_J[2];        // CTAs sharing the records of the item
};
struct WNet { int n; WLayer l[NJODE_MAX_LINEAR]; };
struct WCfg {
    WNet net[3];
    int d, H, dout, curt, loss_kind, residual, training, has_drop;
    float w, keep_scale;
    unsigned thr, seed_lo, seed_hi;
    unsigned img_bytes; int bias_floats;
    unsigned wt_bytes;
    int act_rec[3], g_rec[3];      // bytes per activation / gradient record of each net
    int dw_items, dw_slots;
};

struct WArgs {
    njode_batch_t b;
    const unsigned char* wimg;     // packed bf16 weight image
    const float* bias;             // packed fp32 biases
    float* h_start;                // [n_units, H]  encoder output per unit
    int* row_unit;                 // [N] unit that starts at observation row r
    float *hT, *row_loss, *h_hist, *h_before, *y_after;
    int mode, n_tiles_loss, n_tiles, get_loss;
    // training: operand tiles spilled for the backward pass (bf16 images, one record per tile and chain step)
    float* y_before;               // [N, d] readout before the jump
    unsigned char* act;            // activation records of the pass's net (forward: written, backward: read)
    unsigned char* gsp;            // gradient records of the pass's net (backward: written; dW: read)
    int* tile_base;                // [n_tiles + 1] first ODE record of every tile
    int spill;
    // backward
    const unsigned char* wt;       // transposed bf16 weight image
    float* g_before;               // [N, H]       dL/dh_before[row]
    float* g_start;                // [n_units, H] dL/dh_start[u]
    const float* grad_loss; const float* grad_hT;
    float* dw_part;                // [dw_slots][256 x 256] partial dW accumulators, then [dw_slots][256] bias partials
    int n_tiles_all;
    unsigned long long* prof;      // debugging: clock64 stamps of CTA 0 (4 per layer GEMM: a_ready seen, MMAs issued, accumulator seen, epilogue done)
};

// dynamic shared memory map (bytes from the 1024-aligned base)
constexpr int SM_A = 0;
constexpr int SM_W = SM_A + A_BLOCKS * A_BLOCK_BYTES;
constexpr int SM_BIAS = SM_W + NSTAGE * STAGE_BYTES;                 // [8][256] fp32
constexpr int SM_RES = SM_BIAS + NJODE_MAX_LINEAR * MAX_W * 4;       // [128][16] fp32 readout residual (summed over the column groups)
constexpr int SM_YBJ = SM_RES + TILE_M * MAX_D * 4;                  // [128][16] fp32
constexpr int SM_BAR = SM_YBJ + TILE_M * MAX_D * 4;                  // mbarriers (SM_BAR is a multiple of 1024)
constexpr int SM_TOTAL = SM_BAR + 256;
constexpr int SMEM_BYTES = SM_TOTAL + 1024;                          // + alignment slack

#if !defined(NJODE_HOST_SIM)
// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
// bounded spin: a protocol error traps (with the barrier's offset) instead of hanging the device.  `sleep_ns` > 0
// backs the polling off: the single-thread roles (producer, MMA issuer, spill) share their scheduler with an
// epilogue warp and must not take its issue slots while they wait.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, unsigned sleep_ns = 0) {
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; ++spins) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (!done) {
            if (sleep_ns) __nanosleep(sleep_ns);
            if (spins > (1u << 22)) {
                printf("njode_wide: mbarrier wait timed out (cta %d thread %d barrier@%u parity %u)\n", (int)blockIdx.x, (int)threadIdx.x, bar & 255u, parity);
                __trap();
            }
        }
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                 "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                 "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                 :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                    "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
                    "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
                    "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" :: "n"(EPI_THREADS) : "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&p);
}
// bf16 hi/lo split of an fp32 scalar: hi + lo reproduces v to ~16 mantissa bits
__device__ __forceinline__ uint32_t split_bf16(float v) {
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    return (uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(lo) << 16);
}

// shared-memory matrix descriptor of a K-major SWIZZLE_128B operand tile (rows of 128 B, 8-row groups 1024 B
// apart); advancing along K inside the 64-column block = adding the byte offset to the start address.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (1u << 16);               // start address, LBO (unused) = 1
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);              // SBO = 1024 B, version 1, SWIZZLE_128B
    return (uint64_t)lo | ((uint64_t)hi << 32);
}
// instruction descriptor: D fp32, A/B bf16, both K-major, M = 128, N = n16
__device__ __forceinline__ uint32_t make_idesc(int n16) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n16 >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------
// tile bookkeeping shared by the three roles
// ------------------------------------------------------------------------------------------------
struct Tile { int u0, u1, reps; };
__device__ __forceinline__ Tile tile_of(const WCfg& c, const WArgs& a, int t) {
    Tile T;
    const int nl = a.b.n_loss_units;
    if (t < a.n_tiles_loss) { T.u0 = t * TILE_M; T.u1 = min(T.u0 + TILE_M, nl); }
    else { T.u0 = nl + (t - a.n_tiles_loss) * TILE_M; T.u1 = min(T.u0 + TILE_M, a.b.n_units); }
    if (a.mode == MODE_ODE) {
        const int32_t* dsc = a.b.unit_desc + (size_t)T.u0 * 6;       // units are sorted longest first within a run
        T.reps = __ldg(dsc + 2) - __ldg(dsc + 1);
    } else T.reps = (a.mode == MODE_RO) ? 2 : 1;
    return T;
}

// record (spilled operand images) of chain step `rep` of tile t
__device__ __forceinline__ size_t record_of(const WArgs& a, int t, int rep) {
    if (a.mode == MODE_ODE) return (size_t)(__ldg(a.tile_base + t) + rep);
    return a.mode == MODE_RO ? (size_t)(2 * t + rep) : (size_t)t;
}

// ------------------------------------------------------------------------------------------------
// epilogue pieces (thread = TMEM lane r, column range of its half)
// ------------------------------------------------------------------------------------------------
// The A operand of consecutive layer GEMMs rotates through three pairs of K-blocks: GEMM g reads its K-blocks
// 0-1 from pair X_g and 2-3 from pair Y_g while the epilogue of its first output half already writes the next
// GEMM's K-blocks 0-1 into the third pair Z_g; the second half's epilogue (all MMAs of g complete) writes K-blocks
// 2-3 into X_g.  Hence (X, Y, Z)_{g+1} = (Z, X, Y)_g.
__device__ __forceinline__ int pair_x(int g) { return (3 - g % 3) % 3; }
__device__ __forceinline__ int pair_y(int g) { return (pair_x(g) + 1) % 3; }
__device__ __forceinline__ int pair_z(int g) { return (pair_x(g) + 2) % 3; }
// physical block of logical K-block kb of GEMM g's A operand
__device__ __forceinline__ uint32_t ablock_of(uint32_t a_base, int g, int kb) {
    return a_base + (uint32_t)((kb < 2 ? pair_x(g) : pair_y(g)) * 2 + (kb & 1)) * A_BLOCK_BYTES;
}
// physical block written by the epilogue of output half p of GEMM g for columns [128 p + 64 j, +64) (= K-block 2p + j
// of GEMM g + 1)
__device__ __forceinline__ uint32_t wblock_of(uint32_t a_base, int g, int p, int j) {
    return a_base + (uint32_t)((p == 0 ? pair_z(g) : pair_x(g)) * 2 + j) * A_BLOCK_BYTES;
}

// stores 32 consecutive bf16 values (packed in p[16]) of row r at columns [cin, cin+32) (cin = 0 or 32) of a 64-column block
__device__ __forceinline__ void store_a_chunk(uint32_t block, int r, int cin, const uint32_t* p) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t addr = block + (uint32_t)r * 128u + ((uint32_t)(((cin >> 3) + q) ^ (r & 7)) << 4);
        st_shared_v4(addr, p[4 * q], p[4 * q + 1], p[4 * q + 2], p[4 * q + 3]);
    }
}

// hidden layer: act(acc + bias) (+ dropout) of columns [c0, c0+32) -> bf16 -> A.  ACT / DROP are compile-time so
// that the 32-element loops are branch-free; the caller dispatches on the (warp-uniform) layer description.
template <int ACT, bool DROP>
__device__ __forceinline__ void epi_hidden_chunk_t(uint32_t taddr, const float* bias_s, int c0, int n,
                                                   const WCfg& c, unsigned lk, uint32_t block, int r) {
    uint32_t v[32];
    tmem_ld32(taddr + (uint32_t)c0, v);
    float b[32];
#pragma unroll
    for (int q = 0; q < 8; ++q) {                       // broadcast LDS.128 of the bias (overlaps the TMEM load)
        const float4 t = *reinterpret_cast<const float4*>(bias_s + c0 + 4 * q);
        b[4 * q] = t.x; b[4 * q + 1] = t.y; b[4 * q + 2] = t.z; b[4 * q + 3] = t.w;
    }
    tmem_wait_ld();
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const float x = __uint_as_float(v[j]) + b[j];
        f[j] = ACT == NJODE_ACT_TANH ? tanh_fast(x) : (ACT == NJODE_ACT_RELU ? fmaxf(x, 0.f) : x);
    }
    if (c0 + 32 > n) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = (c0 + j < n) ? f[j] : 0.f;
    }
    if (DROP) {
        // neurons o and o ^ 8 share one hash word (nj_keep, njode_hash.cuh): 16 words serve the 32 columns
        const float ks = c.keep_scale;
        const unsigned thr = c.thr;
#pragma unroll
        for (int g = 0; g < 2; ++g)
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const unsigned word = nj_keep_word(lk, (unsigned)jj | ((unsigned)((c0 >> 4) + g) << 3));
                const int j0 = g * 16 + jj, j1 = j0 + 8;
                f[j0] = (word & 0xFFFFu) >= thr ? f[j0] * ks : -0.f;
                f[j1] = (word >> 16) >= thr ? f[j1] * ks : -0.f;
            }
    }
    uint32_t p[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) p[j] = pack_bf16(f[2 * j], f[2 * j + 1]);
    store_a_chunk(block, r, c0 & 32, p);
}
__device__ __forceinline__ void epi_hidden_chunk(uint32_t taddr, const float* bias_s, int c0, const WLayer& L,
                                                 const WCfg& c, unsigned lk, uint32_t block, int r) {
    if (L.act == NJODE_ACT_TANH) {
        if (L.drop) epi_hidden_chunk_t<NJODE_ACT_TANH, true>(taddr, bias_s, c0, L.n, c, lk, block, r);
        else epi_hidden_chunk_t<NJODE_ACT_TANH, false>(taddr, bias_s, c0, L.n, c, lk, block, r);
    } else if (L.act == NJODE_ACT_RELU) {
        if (L.drop) epi_hidden_chunk_t<NJODE_ACT_RELU, true>(taddr, bias_s, c0, L.n, c, lk, block, r);
        else epi_hidden_chunk_t<NJODE_ACT_RELU, false>(taddr, bias_s, c0, L.n, c, lk, block, r);
    } else epi_hidden_chunk_t<NJODE_ACT_NONE, false>(taddr, bias_s, c0, L.n, c, lk, block, r);
}

// tanh of 32 fp32 values -> bf16 -> A main blocks at columns [c0, c0+32); columns >= n are written as 0
__device__ __forceinline__ void store_tanh_chunk(uint32_t block, int r, int c0, int n, const float* h) {
    float t[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) t[j] = tanh_fast(h[j]);
    if (c0 + 32 > n) {
#pragma unroll
        for (int j = 0; j < 32; ++j) t[j] = (c0 + j < n) ? t[j] : 0.f;
    }
    uint32_t p[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) p[j] = pack_bf16(t[2 * j], t[2 * j + 1]);
    store_a_chunk(block, r, c0 & 32, p);
}

// 32 floats of a global row (columns [c0, c0+32) of a width-n row; n % 4 == 0), zero beyond n or when src == null
__device__ __forceinline__ void load_row_chunk(const float* src, int c0, int n, float* h) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (src && c0 + 4 * q < n) t = *reinterpret_cast<const float4*>(src + c0 + 4 * q);
        h[4 * q] = t.x; h[4 * q + 1] = t.y; h[4 * q + 2] = t.z; h[4 * q + 3] = t.w;
    }
}
__device__ __forceinline__ void store_row_chunk(float* dst, int c0, int n, const float* h) {
#pragma unroll
    for (int q = 0; q < 8; ++q)
        if (c0 + 4 * q < n) *reinterpret_cast<float4*>(dst + c0 + 4 * q) = make_float4(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]);
}

// the auxiliary block row of unit r: time columns (chunk 0) and tanh(x) (columns 8..8+d)
__device__ __forceinline__ void store_aux_time(uint32_t a_base, int r, float tau, float tdiff, int curt) {
    const uint32_t addr = a_base + AUX_BLOCK * A_BLOCK_BYTES + (uint32_t)r * 128u + ((uint32_t)(0 ^ (r & 7)) << 4);
    st_shared_v4(addr, split_bf16(tau), split_bf16(tdiff), curt ? split_bf16(tau + tdiff) : 0u, 0u);
}
__device__ __forceinline__ void store_aux_x(uint32_t a_base, int r, const float* x, int d) {
    // columns 8..23 (chunks 1, 2) hold tanh(x[0..16)); chunks 3..7 are zero
    uint32_t p[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float t0 = (2 * j < d) ? nj_tanh(x[2 * j]) : 0.f, t1 = (2 * j + 1 < d) ? nj_tanh(x[2 * j + 1]) : 0.f;
        p[j] = pack_bf16(t0, t1);
    }
    const uint32_t row = a_base + AUX_BLOCK * A_BLOCK_BYTES + (uint32_t)r * 128u;
    st_shared_v4(row + ((uint32_t)(1 ^ (r & 7)) << 4), p[0], p[1], p[2], p[3]);
    st_shared_v4(row + ((uint32_t)(2 ^ (r & 7)) << 4), p[4], p[5], p[6], p[7]);
#pragma unroll
    for (int q = 3; q < 8; ++q) st_shared_v4(row + ((uint32_t)(q ^ (r & 7)) << 4), 0u, 0u, 0u, 0u);
}

// d is a power of two <= 16.  res[m] holds the sum over the columns = m (mod 16); fold it to column mod d.
__device__ __forceinline__ void fold_res(float* res, int d) {
#pragma unroll
    for (int w = 8; w >= 1; w >>= 1)
        if (d <= w) {
#pragma unroll
            for (int j = 0; j < w; ++j) res[j] += res[j + w];
        }
}
// xe[m] = x[m mod d]
__device__ __forceinline__ void expand_x(const float* x, float* xe, int d) {
#pragma unroll
    for (int m = 0; m < MAX_D; ++m) xe[m] = x[m];
#pragma unroll
    for (int w = 1; w < MAX_D; w <<= 1)
        if (d <= w) {
#pragma unroll
            for (int j = 0; j < w; ++j) xe[j + w] = xe[j];
        }
}

__device__ __forceinline__ unsigned ev_start(const WArgs& a, int sr) {
    return sr < 0 ? NJ_EVENT_INIT : NJ_EVENT_JUMP_BASE + 3u * (unsigned)__ldg(a.b.row_jump + sr) + 1u;
}
__device__ __forceinline__ unsigned ev_jump(const WArgs& a, int row, unsigned which) {
    return NJ_EVENT_JUMP_BASE + 3u * (unsigned)__ldg(a.b.row_jump + row) + which;
}

// ------------------------------------------------------------------------------------------------
// the forward kernel
// ------------------------------------------------------------------------------------------------
// Every layer GEMM is issued as two output halves of <= 128 columns (TMEM columns [0,128) and [128,256)), each
// committed to its own mbarrier, so that the epilogue of half 0 overlaps the MMAs of half 1 and the MMAs of the
// next layer's half 0 (which only need K-blocks 0-1 = the columns half 0's epilogue has just written) overlap the
// epilogue of half 1.  Barriers: full/empty per weight stage, acc[2] (MMA -> epilogue), a_ready[2] (epilogue ->
// MMA: "columns of half p written and TMEM half p drained"), spill (spill warp -> epilogue).
struct Bars { uint32_t full, empty, acc, a, spill; };
__device__ __forceinline__ Bars bars_of(uint32_t bar0) {
    Bars b;
    b.full = bar0; b.empty = bar0 + 8 * NSTAGE; b.acc = bar0 + 16 * NSTAGE; b.a = b.acc + 16; b.spill = b.a + 16;
    return b;
}
constexpr int SM_TMEM_SLOT = SM_BAR + 16 * NSTAGE + 48;

__device__ __forceinline__ void cta_setup(const Bars& B, uint32_t* tmem_slot, uint32_t zero_base, int zero_bytes) {
    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(B.full + 8 * s, 1); mbar_init(B.empty + 8 * s, 1); }
        mbar_init(B.acc, 1); mbar_init(B.acc + 8, 1);
        mbar_init(B.a, EPI_WARPS); mbar_init(B.a + 8, EPI_WARPS);
        mbar_init(B.spill, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if ((threadIdx.x >> 5) == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // zero the operand blocks once: padding columns only ever meet zero weights but must stay finite
    for (int i = threadIdx.x; i < zero_bytes / 16; i += blockDim.x)
        st_shared_v4(zero_base + 16u * i, 0u, 0u, 0u, 0u);
}

__device__ __forceinline__ void wide_cta(const WCfg& c, const WArgs& a, unsigned char* smem_raw) {
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    const uint32_t a_base = sbase + SM_A, w_base = sbase + SM_W;
    float* bias_s = reinterpret_cast<float*>(smem + SM_BIAS);
    float* res_s = reinterpret_cast<float*>(smem + SM_RES);
    float* ybj_s = reinterpret_cast<float*>(smem + SM_YBJ);
    const Bars B = bars_of(sbase + SM_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_TMEM_SLOT);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int netid = a.mode == MODE_ENC ? NJODE_NET_ENC : (a.mode == MODE_ODE ? NJODE_NET_ODE : NJODE_NET_RO);
    const WNet& net = c.net[netid];

    cta_setup(B, tmem_slot, a_base, A_BLOCKS * A_BLOCK_BYTES);
    for (int i = threadIdx.x; i < net.n * MAX_W; i += NUM_THREADS) {
        const int l = i / MAX_W, o = i % MAX_W;
        bias_s[i] = o < net.l[l].n16 ? __ldg(a.bias + net.l[l].bias_off + o) : 0.f;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ===================== weight producer: (half, K-block) tiles in the order the MMAs consume them =====================
        if (lane == 0) {
            int stage = 0; uint32_t ph = 0;
            for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
                const Tile T = tile_of(c, a, t);
                for (int rep = 0; rep < T.reps; ++rep)
                    for (int l = 0; l < net.n; ++l) {
                        const WLayer& L = net.l[l];
                        const int nkb = L.kb_main + L.has_aux;
                        const unsigned char* src = a.wimg + L.img_off;
                        for (int p = 0; p * 128 < L.n16; ++p) {
                            const uint32_t bytes = (uint32_t)min(128, L.n16 - 128 * p) * 128u;
                            for (int kb = 0; kb < nkb; ++kb) {
                                mbar_wait(B.empty + 8 * stage, ph ^ 1, 128);
                                mbar_expect_tx(B.full + 8 * stage, bytes);
                                bulk_g2s(w_base + stage * STAGE_BYTES, src, bytes, B.full + 8 * stage);
                                src += bytes;
                                if (++stage == NSTAGE) { stage = 0; ph ^= 1; }
                            }
                        }
                    }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            int stage = 0, g = 0; uint32_t ph = 0, pa0 = 0, pa1 = 0;
            for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
                const Tile T = tile_of(c, a, t);
                for (int rep = 0; rep < T.reps; ++rep)
                    for (int l = 0; l < net.n; ++l, ++g) {
                        const WLayer& L = net.l[l];
                        const int nkb = L.kb_main + L.has_aux;
                        bool got1 = false;
                        mbar_wait(B.a, pa0, 32); pa0 ^= 1;               // K-blocks 0-1 written, TMEM half 0 drained
                        tc_fence_after();
                        if (a.prof && blockIdx.x == 0 && g < 256) a.prof[4 * g] = clock64();
                        for (int p = 0; p * 128 < L.n16; ++p) {
                            const uint32_t idesc = make_idesc(min(128, L.n16 - 128 * p));
                            for (int kb = 0; kb < nkb; ++kb) {
                                const bool aux = kb >= L.kb_main;
                                if ((aux || kb >= 2) && !got1) {     // K-blocks 2-3 / the auxiliary block, TMEM half 1 drained
                                    mbar_wait(B.a + 8, pa1, 32); pa1 ^= 1; got1 = true;
                                    tc_fence_after();
                                }
                                const int ksteps = aux ? L.aux_ksteps : 4;
                                mbar_wait(B.full + 8 * stage, ph, 32);
                                tc_fence_after();
                                const uint32_t ab = aux ? a_base + AUX_BLOCK * A_BLOCK_BYTES : ablock_of(a_base, g, kb);
                                const uint32_t wb = w_base + (uint32_t)stage * STAGE_BYTES;
                                for (int k = 0; k < ksteps; ++k)
                                    tc_mma(tmem + 128u * p, make_desc(ab + 32u * k), make_desc(wb + 32u * k), idesc, (kb | k) ? 1u : 0u);
                                tc_commit(B.empty + 8 * stage);      // stage free once these MMAs have read it
                                if (++stage == NSTAGE) { stage = 0; ph ^= 1; }
                            }
                            if (p == 0 && !got1) { mbar_wait(B.a + 8, pa1, 32); pa1 ^= 1; got1 = true; tc_fence_after(); }
                            tc_commit(B.acc + 8 * p);                // output half p complete
                        }
                        if (L.n16 <= 128) tc_commit(B.acc + 8);      // no second half: keep the barrier phases in step
                        if (a.prof && blockIdx.x == 0 && g < 256) a.prof[4 * g + 1] = clock64();
                    }
            }
        }
    } else if (warp == 2) {
        // ===================== operand spill (training): every layer's A image -> its activation record =====================
        if (lane == 0 && a.spill) {
            uint32_t pa1 = 0; int g = 0;
            for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
                const Tile T = tile_of(c, a, t);
                for (int rep = 0; rep < T.reps; ++rep) {
                    unsigned char* rec = a.act + record_of(a, t, rep) * (size_t)c.act_rec[netid];
                    for (int l = 0; l < net.n; ++l, ++g) {
                        const WLayer& L = net.l[l];
                        mbar_wait(B.a + 8, pa1, 128); pa1 ^= 1;      // the whole image is written (half 1 arrives last)
                        for (int kb = 0; kb < L.kb_main; ++kb)
                            bulk_s2g(rec + L.act_off + (size_t)kb * A_BLOCK_BYTES, ablock_of(a_base, g, kb), A_BLOCK_BYTES);
                        if (L.has_aux)
                            bulk_s2g(rec + L.act_off + (size_t)L.kb_main * A_BLOCK_BYTES, a_base + AUX_BLOCK * A_BLOCK_BYTES, A_BLOCK_BYTES);
                        bulk_commit();
                        bulk_wait_read();                 // the image has been read: the epilogue may overwrite it
                        mbar_arrive(B.spill);
                    }
                }
            }
            bulk_wait_all();
        }
    } else if (warp >= EPI_WARP0) {
        // ===================== epilogue warps =====================
        const int q = warp & 3, cg = (warp - EPI_WARP0) >> 2;
        const int r = q * 32 + lane;
        const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
        uint32_t pf0 = 0, pf1 = 0, ps = 0;
        int g = 0;                         // layer GEMMs issued so far (all roles count alike)
        int n_arr = 0, n_spw = 0;          // complete a_ready arrivals made / spill completions consumed
        // the A image of an arrival may be overwritten only after the spill warp has read it
        auto spill_sync = [&]() {
            if (a.spill)
                while (n_spw < n_arr) { mbar_wait(B.spill, ps); ps ^= 1; ++n_spw; }
        };
        auto arrive = [&](int p) {
            tc_fence_before(); fence_proxy_async(); __syncwarp();
            if (lane == 0) mbar_arrive(B.a + 8 * p);
        };
        const int H = c.H, d = c.d;
        const int hch = (H + 31) >> 5;
        // chunk c (32 columns) of an H-wide row belongs to the column group (c % 4) % EPI_GROUPS
        auto mine = [&](int ch) { return ((ch & 3) % EPI_GROUPS) == cg; };
        for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
            const Tile T = tile_of(c, a, t);
            const int u = T.u0 + r;
            const bool valid = u < T.u1;
            int path = 0, s0 = 0, len = 0, row = -1, flag = 0, sr = -1;
            if (valid) {
                const int32_t* dsc = a.b.unit_desc + (size_t)u * 6;
                path = __ldg(dsc); s0 = __ldg(dsc + 1); len = __ldg(dsc + 2) - s0;
                const int c0_ = __ldg(dsc + 3), c1_ = __ldg(dsc + 4), sc = __ldg(dsc + 5);
                row = c1_ > c0_ ? __ldg(a.b.path_rows + c0_) : -1;
                flag = (sc & NJODE_UNIT_WRITES_HT) ? 1 : 0;
                sr = (sc & ~NJODE_UNIT_WRITES_HT) - 1;
            }
            const unsigned gpath = (unsigned)(path + a.b.path_id_offset);
            float xr[MAX_D], xe[MAX_D];
            float tau = 0.f;
            unsigned rk = 0;
            // tanh of my chunks of an H-wide fp32 row -> A image of GEMM `gg`; optionally the readout residual sums
            auto stage_h_row = [&](const float* src, int gg, bool to_tmem, bool with_res) {
                float res[MAX_D];
#pragma unroll
                for (int j = 0; j < MAX_D; ++j) res[j] = 0.f;
                for (int ch = 0; ch < hch; ++ch) {
                    if (!mine(ch)) continue;
                    float h[32];
                    load_row_chunk(src, ch * 32, H, h);
                    if (to_tmem) tmem_st32(tlane + TMEM_H + ch * 32, reinterpret_cast<const uint32_t*>(h));
                    store_tanh_chunk(ablock_of(a_base, gg, ch >> 1), r, ch * 32, H, h);
                    if (with_res) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) if (ch * 32 + j < H) res[j & 15] += h[j];     // column mod 16
                    }
                }
                if (with_res) {
                    fold_res(res, d);
#pragma unroll
                    for (int j = 0; j < MAX_D; ++j) if (j < d) atomicAdd(res_s + r * MAX_D + j, res[j]);
                }
            };
            auto zero_res = [&]() {
                if (cg == 0) {
#pragma unroll
                    for (int j = 0; j < MAX_D; ++j) res_s[r * MAX_D + j] = 0.f;
                }
                epi_bar_sync();
            };
            // ---------------- prologue: the A image of the tile's first GEMM ----------------
            spill_sync();
            if (a.mode == MODE_ENC) {
#pragma unroll
                for (int j = 0; j < MAX_D; ++j)
                    xr[j] = (valid && j < d) ? (sr < 0 ? __ldg(a.b.start_X + (size_t)path * d + j) : __ldg(a.b.X + (size_t)sr * d + j)) : 0.f;
                if (cg == 0) {
                    store_aux_x(a_base, r, xr, d);
                    store_aux_time(a_base, r, 0.f, 0.f, 0);
                    if (valid && sr >= 0) a.row_unit[sr] = u;
                }
                expand_x(xr, xe, d);
                rk = nj_row_key(c.seed_lo, c.seed_hi, gpath, ev_start(a, sr));
            } else if (a.mode == MODE_ODE) {
#pragma unroll
                for (int j = 0; j < MAX_D; ++j)
                    xr[j] = (valid && j < d) ? (sr < 0 ? __ldg(a.b.start_X + (size_t)path * d + j) : __ldg(a.b.X + (size_t)sr * d + j)) : 0.f;
                tau = (valid && sr >= 0) ? __ldg(a.b.jump_tau + __ldg(a.b.row_jump + sr)) : 0.f;
                const float* hs = valid ? a.h_start + (size_t)u * H : nullptr;
                stage_h_row(hs, g, true, false);
                tmem_wait_st();
                if (valid && len == 0) {               // a unit without Euler steps (observation at t = 0)
                    for (int ch = 0; ch < hch; ++ch) {
                        if (!mine(ch)) continue;
                        float h[32];
                        load_row_chunk(hs, ch * 32, H, h);
                        if (row >= 0 && a.h_before) store_row_chunk(a.h_before + (size_t)row * H, ch * 32, H, h);
                        if (flag) store_row_chunk(a.hT + (size_t)path * H, ch * 32, H, h);
                    }
                }
                if (cg == 0) {
                    store_aux_x(a_base, r, xr, d);
                    const float t0 = (valid && len > 0) ? __ldg(a.b.step_t + s0) : 0.f;
                    store_aux_time(a_base, r, tau, t0 - tau, c.curt);
                }
                rk = nj_row_key(c.seed_lo, c.seed_hi, gpath, (unsigned)s0);
            } else {
                zero_res();
                stage_h_row((valid && row >= 0) ? a.h_before + (size_t)row * H : nullptr, g, false, c.residual != 0);
                epi_bar_sync();                        // residual sums complete before a group-0 thread reads them
                rk = (valid && row >= 0) ? nj_row_key(c.seed_lo, c.seed_hi, gpath, ev_jump(a, row, 0u)) : 0u;
            }
            if (T.reps == 0) continue;
            arrive(0); arrive(1); ++n_arr;
            // ---------------- layers ----------------
            for (int rep = 0; rep < T.reps; ++rep) {
                for (int l = 0; l < net.n; ++l, ++g) {
                    const WLayer& L = net.l[l];
                    const bool last = (l == net.n - 1);
                    const bool tile_done = last && rep + 1 == T.reps;
                    const bool ro_chain1_end = a.mode == MODE_RO && last && rep == 0;
                    const unsigned lk = (!last && L.drop) ? nj_layer_key(rk, (unsigned)(netid * 16 + l + 1)) : 0u;
                    // ODE: per-step scalars of the last layer
                    const int i = rep;
                    const bool active = valid && i < len;
                    const float dt = (a.mode == MODE_ODE && last && active) ? __ldg(a.b.step_dt + s0 + i) : 0.f;
                    const bool more = i + 1 < T.reps;
                    for (int p = 0; p < 2; ++p) {
                        if (p == 0) { mbar_wait(B.acc, pf0, 32); pf0 ^= 1; }
                        else { mbar_wait(B.acc + 8, pf1, 32); pf1 ^= 1; }
                        tc_fence_after();
                        if (p == 0 && a.prof && blockIdx.x == 0 && threadIdx.x == EPI_WARP0 * 32 && g < 256) a.prof[4 * g + 2] = clock64();
                        if (p == 1) spill_sync();                 // half 1 overwrites K-blocks 0-1 of this GEMM's own A image
                        const int ncols = min(128, L.n16 - 128 * p);              // <= 0: no such half
                        const int nchp = ncols > 0 ? (ncols + 31) >> 5 : 0;
                        for (int ch = cg; ch < nchp; ch += EPI_GROUPS) {
                            const int c0 = 128 * p + 32 * ch;                     // first output column of the chunk
                            const uint32_t tcol = tlane + (uint32_t)c0;
                            if (!last) {
                                epi_hidden_chunk(tcol - c0, bias_s + l * MAX_W, c0, L, c, lk, wblock_of(a_base, g, p, ch >> 1), r);
                            } else if (a.mode == MODE_ENC) {
                                uint32_t v[32];
                                tmem_ld32(tcol, v);
                                tmem_wait_ld();
                                float e[32];
#pragma unroll
                                for (int j = 0; j < 32; ++j) {
                                    e[j] = __uint_as_float(v[j]) + bias_s[l * MAX_W + c0 + j];
                                    if (c.residual) e[j] += xe[j & 15];                       // case 1: x.repeat(1, H / d)
                                }
                                if (valid) store_row_chunk(a.h_start + (size_t)u * H, c0, H, e);
                            } else if (a.mode == MODE_ODE) {
                                uint32_t v[32];
                                float h[32];
                                tmem_ld32(tcol, v);
                                tmem_ld32(tlane + TMEM_H + c0, reinterpret_cast<uint32_t*>(h));
                                tmem_wait_ld();
                                if (active) {
                                    if (a.h_hist) store_row_chunk(a.h_hist + ((size_t)(s0 + i) * a.b.B + path) * H, c0, H, h);
#pragma unroll
                                    for (int j = 0; j < 32; ++j)
                                        h[j] = fmaf(dt, __uint_as_float(v[j]) + bias_s[l * MAX_W + c0 + j], h[j]);
                                    if (i == len - 1) {
                                        if (row >= 0 && a.h_before) store_row_chunk(a.h_before + (size_t)row * H, c0, H, h);
                                        if (flag) store_row_chunk(a.hT + (size_t)path * H, c0, H, h);
                                    }
                                }
                                // .sync.aligned: every lane of the warp stores (finished / padding rows write h back unchanged)
                                tmem_st32(tlane + TMEM_H + c0, reinterpret_cast<const uint32_t*>(h));
                                if (more) store_tanh_chunk(wblock_of(a_base, g, p, ch >> 1), r, c0, H, h);
                            } else if (p == 0 && ch == 0) {
                                // readout: y = acc + bias + residual (mean over the H / d chunks of h, or h itself when H == d)
                                uint32_t v[32];
                                tmem_ld32(tcol, v);
                                tmem_wait_ld();
                                const float rmul = c.residual ? 1.f / (float)(H / d) : 0.f;
                                float y[MAX_D];
#pragma unroll
                                for (int j = 0; j < MAX_D; ++j)
                                    y[j] = __uint_as_float(v[j]) + bias_s[l * MAX_W + j] + rmul * res_s[r * MAX_D + j];
                                if (rep == 0) {
#pragma unroll
                                    for (int j = 0; j < MAX_D; ++j) {
                                        ybj_s[r * MAX_D + j] = y[j];
                                        if (a.y_before && valid && row >= 0 && j < d) a.y_before[(size_t)row * d + j] = y[j];
                                    }
                                } else if (valid && row >= 0) {
                                    float sa = 0.f, sb = 0.f;
#pragma unroll
                                    for (int j = 0; j < MAX_D; ++j) {
                                        if (j < d) {
                                            if (a.y_after) a.y_after[(size_t)row * d + j] = y[j];
                                            const float x = __ldg(a.b.X + (size_t)row * d + j), yb = ybj_s[r * MAX_D + j];
                                            const float da = x - y[j], db = (c.loss_kind == NJODE_LOSS_STANDARD) ? (yb - y[j]) : (yb - x);
                                            sa = fmaf(da, da, sa); sb = fmaf(db, db, sb);
                                        }
                                    }
                                    if (a.get_loss) {
                                        const float ra = sqrtf(sa + 1e-10f), rb = sqrtf(sb + 1e-10f);
                                        const float sm = (c.loss_kind == NJODE_LOSS_STANDARD) ? (2.f * c.w * ra + 2.f * (1.f - c.w) * rb)
                                                                                              : (c.w * ra + (1.f - c.w) * rb);
                                        a.row_loss[row] = sm * sm / __ldg(a.b.n_obs_ot + path);
                                    }
                                }
                            }
                        }
                        if (p == 1) {
                            if (a.mode == MODE_ODE && last) {
                                tmem_wait_st();
                                if (more) {
                                    if (cg == 0) {
                                        const float tn = (valid && i + 1 < len) ? __ldg(a.b.step_t + s0 + i + 1) : 0.f;
                                        store_aux_time(a_base, r, tau, tn - tau, c.curt);
                                    }
                                    rk = nj_row_key(c.seed_lo, c.seed_hi, gpath, (unsigned)(s0 + i + 1));
                                }
                            }
                            if (ro_chain1_end) {
                                // second chain of the tile: Y = readout(h after the jump), h = encoder output of the unit that
                                // starts at this row.  All MMAs of this GEMM are complete: both block pairs may be written.
                                epi_bar_sync();                       // chain 1's residual sums are consumed
                                zero_res();
                                stage_h_row((valid && row >= 0) ? a.h_start + (size_t)a.row_unit[row] * H : nullptr, g + 1, false, c.residual != 0);
                                epi_bar_sync();
                                rk = (valid && row >= 0) ? nj_row_key(c.seed_lo, c.seed_hi, gpath, ev_jump(a, row, 2u)) : 0u;
                            }
                            if (a.prof && blockIdx.x == 0 && threadIdx.x == EPI_WARP0 * 32 && g < 256) a.prof[4 * g + 3] = clock64();
                        }
                        if (!tile_done) {
                            // chain 1 of the readout writes the next image only in its second phase: both arrivals then
                            if (p == 0 && !ro_chain1_end) arrive(0);
                            if (p == 1) { if (ro_chain1_end) arrive(0); arrive(1); ++n_arr; }
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(TMEM_COLS) : "memory");
    }
}

// ================================================================================================
// backward chain: per tile and chain step, G_{l-1} = (G_l . W_l) * act'(A_l) for l = L-1 .. 0, with the
// forward's spilled activation images A_l supplying act' and every G_l image spilled for the dW pass.
//   BWD_RO  : two chains per loss tile (Y_bj, Y): G_{L-1} = dLoss/dY..., ends in g_before[row] / g_start[unit]
//   BWD_ODE : reverse Euler steps, adjoint dL/dh in TMEM (fp32): G_{L-1} = dt * adjoint,
//             adjoint += (G_0 . W_0)[tanh(h) columns] * (1 - tanh(h)^2); ends in g_start[unit]
//   BWD_ENC : G_{L-1} = g_start[unit]; no input gradient (layer 0 only spills G_0)
// Same warp roles, barriers and operand layouts as the forward kernel; the B operand is the transposed weight
// image (W_l^T, K-major over the layer's outputs).
// ================================================================================================
// my row's 32 columns [c0, c0+32) of a spilled image (global memory), as 16 packed bf16 pairs
__device__ __forceinline__ void load_img_chunk(const unsigned char* img, int r, int c0, uint32_t* p) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int col = c0 + 8 * q;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(img + (size_t)(col >> 6) * A_BLOCK_BYTES + (size_t)r * 128
                                                              + ((((col & 63) >> 3) ^ (r & 7)) << 4)));
        p[4 * q] = v.x; p[4 * q + 1] = v.y; p[4 * q + 2] = v.z; p[4 * q + 3] = v.w;
    }
}
// 32 fp32 gradients -> bf16 -> G image columns [c0, c0+32); columns >= n are written as 0
__device__ __forceinline__ void store_g_chunk(uint32_t block, int r, int c0, int n, float* g) {
    if (c0 + 32 > n) {
#pragma unroll
        for (int j = 0; j < 32; ++j) g[j] = (c0 + j < n) ? g[j] : 0.f;
    }
    uint32_t p[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) p[j] = pack_bf16(g[2 * j], g[2 * j + 1]);
    store_a_chunk(block, r, c0 & 32, p);
}
// g[j] = acc[j] * act'(a_j) with the (warp-uniform) activation kind hoisted out of the element loop
template <int ACT, bool DROP>
__device__ __forceinline__ void apply_act_grad_t(const uint32_t* v, const uint32_t* av, float ks, float inv_ks, float* g) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const uint32_t b16 = (j & 1) ? (av[j >> 1] >> 16) : (av[j >> 1] & 0xFFFFu);
        const float a = __uint_as_float(b16 << 16);
        float f;
        if (ACT == NJODE_ACT_TANH) {
            if (DROP) { const float t = a * inv_ks; f = b16 == 0x8000u ? 0.f : ks * (1.f - t * t); }   // dropped: stored as -0
            else f = 1.f - a * a;
        } else if (ACT == NJODE_ACT_RELU) f = a > 0.f ? (DROP ? ks : 1.f) : 0.f;
        else f = 1.f;
        g[j] = __uint_as_float(v[j]) * f;
    }
}
__device__ __forceinline__ void apply_act_grad(const uint32_t* v, const uint32_t* av, int act, int drop, float ks, float inv_ks, float* g) {
    if (act == NJODE_ACT_TANH) { if (drop) apply_act_grad_t<NJODE_ACT_TANH, true>(v, av, ks, inv_ks, g); else apply_act_grad_t<NJODE_ACT_TANH, false>(v, av, ks, inv_ks, g); }
    else if (act == NJODE_ACT_RELU) { if (drop) apply_act_grad_t<NJODE_ACT_RELU, true>(v, av, ks, inv_ks, g); else apply_act_grad_t<NJODE_ACT_RELU, false>(v, av, ks, inv_ks, g); }
    else apply_act_grad_t<NJODE_ACT_NONE, false>(v, av, ks, inv_ks, g);
}

__device__ __forceinline__ void wide_bwd_cta(const WCfg& c, const WArgs& a, unsigned char* smem_raw) {
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    const uint32_t a_base = sbase + SM_A, w_base = sbase + SM_W;
    const Bars B = bars_of(sbase + SM_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_TMEM_SLOT);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int netid = a.mode == MODE_ENC ? NJODE_NET_ENC : (a.mode == MODE_ODE ? NJODE_NET_ODE : NJODE_NET_RO);
    const WNet& net = c.net[netid];
    const int l_min = a.mode == MODE_ENC ? 1 : 0;          // the encoder input needs no gradient

    cta_setup(B, tmem_slot, a_base, A_BLOCKS * A_BLOCK_BYTES);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t ph = 0;
            for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
                const Tile T = tile_of(c, a, t);
                for (int rep = 0; rep < T.reps; ++rep)
                    for (int l = net.n - 1; l >= l_min; --l) {
                        const WLayer& L = net.l[l];
                        const unsigned char* src = a.wt + L.wt_off;
                        for (int p = 0; p * 128 < L.nt16; ++p) {
                            const uint32_t bytes = (uint32_t)min(128, L.nt16 - 128 * p) * 128u;
                            for (int kb = 0; kb < L.kt_blocks; ++kb) {
                                mbar_wait(B.empty + 8 * stage, ph ^ 1, 128);
                                mbar_expect_tx(B.full + 8 * stage, bytes);
                                bulk_g2s(w_base + stage * STAGE_BYTES, src, bytes, B.full + 8 * stage);
                                src += bytes;
                                if (++stage == NSTAGE) { stage = 0; ph ^= 1; }
                            }
                        }
                    }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int stage = 0, g = 0; uint32_t ph = 0, pa0 = 0, pa1 = 0;
            for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
                const Tile T = tile_of(c, a, t);
                for (int rep = 0; rep < T.reps; ++rep)
                    for (int l = net.n - 1; l >= 0; --l, ++g) {
                        const WLayer& L = net.l[l];
                        mbar_wait(B.a, pa0, 32); pa0 ^= 1;               // G_l K-blocks 0-1 written, TMEM half 0 drained
                        if (l < l_min) { mbar_wait(B.a + 8, pa1, 32); pa1 ^= 1; continue; }     // spilled only, no GEMM
                        tc_fence_after();
                        bool got1 = false;
                        for (int p = 0; p * 128 < L.nt16; ++p) {
                            const uint32_t idesc = make_idesc(min(128, L.nt16 - 128 * p));
                            for (int kb = 0; kb < L.kt_blocks; ++kb) {
                                if (kb >= 2 && !got1) { mbar_wait(B.a + 8, pa1, 32); pa1 ^= 1; got1 = true; tc_fence_after(); }
                                const int ksteps = (kb == L.kt_blocks - 1) ? L.kt_last_ksteps : 4;
                                mbar_wait(B.full + 8 * stage, ph, 32);
                                tc_fence_after();
                                const uint32_t ab = ablock_of(a_base, g, kb);
                                const uint32_t wb = w_base + (uint32_t)stage * STAGE_BYTES;
                                for (int k = 0; k < ksteps; ++k)
                                    tc_mma(tmem + 128u * p, make_desc(ab + 32u * k), make_desc(wb + 32u * k), idesc, (kb | k) ? 1u : 0u);
                                tc_commit(B.empty + 8 * stage);
                                if (++stage == NSTAGE) { stage = 0; ph ^= 1; }
                            }
                            if (p == 0 && !got1) { mbar_wait(B.a + 8, pa1, 32); pa1 ^= 1; got1 = true; tc_fence_after(); }
                            tc_commit(B.acc + 8 * p);
                        }
                        if (L.nt16 <= 128) tc_commit(B.acc + 8);
                    }
            }
        }
    } else if (warp == 2) {
        // every G_l image -> its gradient record (read again by the dW pass)
        if (lane == 0) {
            uint32_t pa1 = 0; int g = 0;
            for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
                const Tile T = tile_of(c, a, t);
                for (int rep = 0; rep < T.reps; ++rep) {
                    const int step = a.mode == MODE_ODE ? T.reps - 1 - rep : rep;
                    unsigned char* rec = a.gsp + record_of(a, t, step) * (size_t)c.g_rec[netid];
                    for (int l = net.n - 1; l >= 0; --l, ++g) {
                        const WLayer& L = net.l[l];
                        mbar_wait(B.a + 8, pa1, 128); pa1 ^= 1;
                        for (int kb = 0; kb < L.kt_blocks; ++kb)
                            bulk_s2g(rec + L.g_off + (size_t)kb * A_BLOCK_BYTES, ablock_of(a_base, g, kb), A_BLOCK_BYTES);
                        bulk_commit();
                        bulk_wait_read();
                        mbar_arrive(B.spill);
                    }
                }
            }
            bulk_wait_all();
        }
    } else if (warp >= EPI_WARP0) {
        const int q = warp & 3, cg = (warp - EPI_WARP0) >> 2;
        const int r = q * 32 + lane;
        const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
        uint32_t pf0 = 0, pf1 = 0, ps = 0;
        int g = 0, n_arr = 0, n_spw = 0;
        auto spill_sync = [&]() { while (n_spw < n_arr) { mbar_wait(B.spill, ps); ps ^= 1; ++n_spw; } };
        auto arrive = [&](int p) {
            tc_fence_before(); fence_proxy_async(); __syncwarp();
            if (lane == 0) mbar_arrive(B.a + 8 * p);
        };
        const int H = c.H, d = c.d;
        const int hch = (H + 31) >> 5;
        auto mine = [&](int ch) { return ((ch & 3) % EPI_GROUPS) == cg; };
        const float ks = c.keep_scale, inv_ks = c.keep_scale > 0.f ? 1.f / c.keep_scale : 0.f;
        const float gl = a.grad_loss ? __ldg(a.grad_loss) / (float)a.b.batch_size_norm : 0.f;
        const int Ln = net.n;
        for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
            const Tile T = tile_of(c, a, t);
            const int u = T.u0 + r;
            const bool valid = u < T.u1;
            int path = 0, s0 = 0, len = 0, row = -1, flag = 0, sr = -1;
            if (valid) {
                const int32_t* dsc = a.b.unit_desc + (size_t)u * 6;
                path = __ldg(dsc); s0 = __ldg(dsc + 1); len = __ldg(dsc + 2) - s0;
                const int c0_ = __ldg(dsc + 3), c1_ = __ldg(dsc + 4), sc = __ldg(dsc + 5);
                row = c1_ > c0_ ? __ldg(a.b.path_rows + c0_) : -1;
                flag = (sc & NJODE_UNIT_WRITES_HT) ? 1 : 0;
                sr = (sc & ~NJODE_UNIT_WRITES_HT) - 1;
            }
            float gy[MAX_D];                     // BWD_RO: dL/dy of the current chain (every thread of the row holds it)
#pragma unroll
            for (int j = 0; j < MAX_D; ++j) gy[j] = 0.f;
            if (a.mode == MODE_ODE) {
                // adjoint at the end of the unit: dL/dh_before[row] (loss units) or the caller's gradient into hT
                const float* src = nullptr;
                if (valid && row >= 0) src = a.g_before + (size_t)row * H;
                else if (valid && flag && a.grad_hT) src = a.grad_hT + (size_t)path * H;
                for (int ch = 0; ch < hch; ++ch) {
                    if (!mine(ch)) continue;
                    float gv[32];
                    load_row_chunk(src, ch * 32, H, gv);
                    tmem_st32(tlane + TMEM_H + ch * 32, reinterpret_cast<const uint32_t*>(gv));
                }
                tmem_wait_st();
            }
            for (int rep = 0; rep < T.reps; ++rep) {
                const int step = a.mode == MODE_ODE ? T.reps - 1 - rep : rep;
                const unsigned char* arec = a.act + record_of(a, t, step) * (size_t)c.act_rec[netid];
                // ---------------- G_{L-1}: gradient of the chain's output = the A image of this step's first GEMM ----------------
                spill_sync();
                if (a.mode == MODE_ODE) {
                    const bool active = valid && step < len;
                    const float dt = active ? __ldg(a.b.step_dt + s0 + step) : 0.f;
                    for (int ch = 0; ch < hch; ++ch) {
                        if (!mine(ch)) continue;
                        float gv[32];
                        tmem_ld32(tlane + TMEM_H + ch * 32, reinterpret_cast<uint32_t*>(gv));
                        tmem_wait_ld();
#pragma unroll
                        for (int j = 0; j < 32; ++j) gv[j] *= dt;                     // h += dt * f  (0 for finished / padding rows)
                        store_g_chunk(ablock_of(a_base, g, ch >> 1), r, ch * 32, H, gv);
                    }
                } else if (a.mode == MODE_ENC) {
                    const float* src = valid ? a.g_start + (size_t)u * H : nullptr;
                    for (int ch = 0; ch < hch; ++ch) {
                        if (!mine(ch)) continue;
                        float gv[32];
                        load_row_chunk(src, ch * 32, H, gv);
                        store_g_chunk(ablock_of(a_base, g, ch >> 1), r, ch * 32, H, gv);
                    }
                } else {
                    // loss row (compute_loss / compute_loss_2, NJODE/models.py:71-126): chain 0 = Y_bj, chain 1 = Y
                    if (valid && row >= 0) {
                        float sa = 0.f, sb = 0.f, x[MAX_D], y[MAX_D], yb[MAX_D];
#pragma unroll
                        for (int j = 0; j < MAX_D; ++j) {
                            x[j] = y[j] = yb[j] = 0.f;
                            if (j < d) {
                                x[j] = __ldg(a.b.X + (size_t)row * d + j); y[j] = a.y_after[(size_t)row * d + j]; yb[j] = a.y_before[(size_t)row * d + j];
                                const float da = x[j] - y[j], db = (c.loss_kind == NJODE_LOSS_STANDARD) ? (yb[j] - y[j]) : (yb[j] - x[j]);
                                sa = fmaf(da, da, sa); sb = fmaf(db, db, sb);
                            }
                        }
                        const float ra = sqrtf(sa + 1e-10f), rb = sqrtf(sb + 1e-10f);
                        const bool stdl = c.loss_kind == NJODE_LOSS_STANDARD;
                        const float wa = stdl ? 2.f * c.w : c.w, wb = stdl ? 2.f * (1.f - c.w) : (1.f - c.w);
                        const float sm = wa * ra + wb * rb;
                        const float k0 = gl * 2.f * sm / __ldg(a.b.n_obs_ot + path);
#pragma unroll
                        for (int j = 0; j < MAX_D; ++j) {
                            if (j < d) {
                                if (rep == 0) gy[j] = k0 * wb * (stdl ? (yb[j] - y[j]) : (yb[j] - x[j])) / rb;
                                else gy[j] = k0 * (wa * (y[j] - x[j]) / ra + (stdl ? wb * (y[j] - yb[j]) / rb : 0.f));
                            } else gy[j] = 0.f;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < MAX_D; ++j) gy[j] = 0.f;
                    }
                    if (cg == 0) {
                        uint32_t p[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) p[j] = pack_bf16(gy[2 * j], gy[2 * j + 1]);
                        const uint32_t rowa = ablock_of(a_base, g, 0) + (uint32_t)r * 128u;
                        st_shared_v4(rowa + ((uint32_t)(0 ^ (r & 7)) << 4), p[0], p[1], p[2], p[3]);
                        st_shared_v4(rowa + ((uint32_t)(1 ^ (r & 7)) << 4), p[4], p[5], p[6], p[7]);
                    }
                }
                arrive(0); arrive(1); ++n_arr;
                // ---------------- layers, last to first ----------------
                for (int l = Ln - 1; l >= 0; --l, ++g) {
                    const WLayer& L = net.l[l];
                    if (l < l_min) continue;                                 // G_0 of the encoder chain is only spilled (counts as a GEMM slot)
                    const int ncol = L.nt;                                    // gradient width = main inputs of layer l
                    const WLayer& Lp = net.l[l > 0 ? l - 1 : 0];             // layer whose activation A_l is
                    // my chunks of this layer's input image A_l (forward spill), one per output half: prefetched before the
                    // accumulator waits
                    uint32_t av[2][16];
#pragma unroll
                    for (int p = 0; p < 2; ++p) {
                        const int c0 = 128 * p + 32 * cg;
                        if (EPI_GROUPS == 4 && c0 < L.nt16) load_img_chunk(arec + L.act_off, r, c0, av[p]);
                    }
#pragma unroll
                    for (int p = 0; p < 2; ++p) {
                        if (p == 0) { mbar_wait(B.acc, pf0, 32); pf0 ^= 1; }
                        else { mbar_wait(B.acc + 8, pf1, 32); pf1 ^= 1; }
                        tc_fence_after();
                        if (p == 1) spill_sync();
                        const int ncols = min(128, L.nt16 - 128 * p);
                        const int nchp = ncols > 0 ? (ncols + 31) >> 5 : 0;
                        for (int ch = cg; ch < nchp; ch += EPI_GROUPS) {
                            const int c0 = 128 * p + 32 * ch;
                            uint32_t v[32];
                            float gv[32];
                            if (EPI_GROUPS != 4) load_img_chunk(arec + L.act_off, r, c0, av[p]);
                            tmem_ld32(tlane + (uint32_t)c0, v);
                            if (l == 0 && a.mode == MODE_ODE) tmem_ld32(tlane + TMEM_H + c0, reinterpret_cast<uint32_t*>(gv));
                            tmem_wait_ld();
                            if (l > 0) {
                                apply_act_grad(v, av[p], Lp.act, Lp.drop, ks, inv_ks, gv);
                                store_g_chunk(wblock_of(a_base, g, p, ch >> 1), r, c0, ncol, gv);
                            } else if (a.mode == MODE_ODE) {
                                // input of the net = tanh(h): d/dh = 1 - tanh(h)^2, tanh(h) from the spilled A_0
#pragma unroll
                                for (int j = 0; j < 32; ++j) {
                                    const uint32_t b16 = (j & 1) ? (av[p][j >> 1] >> 16) : (av[p][j >> 1] & 0xFFFFu);
                                    const float th = __uint_as_float(b16 << 16);
                                    gv[j] = fmaf(__uint_as_float(v[j]), 1.f - th * th, gv[j]);
                                }
                                tmem_st32(tlane + TMEM_H + c0, reinterpret_cast<const uint32_t*>(gv));
                            } else {
                                // readout residual (FFNN.forward, models.py:268-276): y[j] += mean_k h[k d + j]
                                float ge[MAX_D];
                                expand_x(gy, ge, d);
                                const float rmul = c.residual ? 1.f / (float)(H / d) : 0.f;
#pragma unroll
                                for (int j = 0; j < 32; ++j) {
                                    const uint32_t b16 = (j & 1) ? (av[p][j >> 1] >> 16) : (av[p][j >> 1] & 0xFFFFu);
                                    const float th = __uint_as_float(b16 << 16);
                                    gv[j] = __uint_as_float(v[j]) * (1.f - th * th) + rmul * ge[j & 15];
                                }
                                if (valid && row >= 0) {
                                    float* dst = rep == 0 ? a.g_before + (size_t)row * H : a.g_start + (size_t)a.row_unit[row] * H;
                                    store_row_chunk(dst, c0, H, gv);
                                }
                            }
                        }
                        if (l == 0 && a.mode == MODE_ODE && p == 1) tmem_wait_st();
                        if (l > 0) { arrive(p); if (p == 1) ++n_arr; }        // G_{l-1} is the next GEMM's operand (or only spilled)
                    }
                }
            }
            if (a.mode == MODE_ODE) {
                // dL/dh_start[u] = adjoint at the start of the unit (+ the readout-after-jump part written by BWD_RO);
                // every lane executes the (warp-collective) TMEM load
                for (int ch = 0; ch < hch; ++ch) {
                    if (!mine(ch)) continue;
                    float gv[32], o[32];
                    tmem_ld32(tlane + TMEM_H + ch * 32, reinterpret_cast<uint32_t*>(gv));
                    tmem_wait_ld();
                    if (valid) {
                        if (sr >= 0) {
                            load_row_chunk(a.g_start + (size_t)u * H, ch * 32, H, o);
#pragma unroll
                            for (int j = 0; j < 32; ++j) gv[j] += o[j];
                        }
                        store_row_chunk(a.g_start + (size_t)u * H, ch * 32, H, gv);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(TMEM_COLS) : "memory");
    }
}

// ================================================================================================
// dW pass: dW_l[n x k] = sum over records of G_l^T[n x 128] . A_l[128 x k], db_l = column sums of G_l.
// The spilled images are [64-column block][128 rows x 128 B] SWIZZLE_128B tiles; read as MN-major operands
// (64 contiguous elements of the M / N dimension per 128-byte row, K = the 128 units of the tile) they feed
// tcgen05.mma directly: A operand = G_l (M = layer outputs, two M = 128 halves in TMEM columns [0,256) and
// [256,512)), B operand = A_l (N = layer inputs).  One CTA owns one (net, layer, main|aux part) item and a
// contiguous share of the net's records; half records (64 units = 4 K-steps) stream through a 3-stage ring.
// ================================================================================================
struct DwItem { int net, layer, part, j, J, slot; };
constexpr int DW_NSTAGE = 3;
constexpr int DW_THREADS = 384;                   // producer, MMA, 2 idle, 8 bias / write-out warps
constexpr int DW_EPI_WARPS = 8;
constexpr int DW_STAGE_BYTES = 65536;             // [4 G half-blocks][4 A half-blocks] of 8 KB
constexpr int DW_HALF_BLOCK = 8192;

__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr) {
    const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | ((uint32_t)(DW_HALF_BLOCK >> 4) << 16);   // LBO: next 64 elements of M / N
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);                                // SBO: next 8 units of K
    return (uint64_t)lo | ((uint64_t)hi << 32);
}

__device__ __forceinline__ void wide_dw_cta(const WCfg& c, const WArgs& a, const DwItem& item, unsigned char* smem_raw) {
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar0 = sbase + SM_BAR;
    const uint32_t bar_full = bar0, bar_empty = bar0 + 8 * DW_NSTAGE, bar_acc = bar0 + 16 * DW_NSTAGE;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_TMEM_SLOT);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const DwItem it = item;
    const WLayer& L = c.net[it.net].l[it.layer];
    const int R = it.net == NJODE_NET_ENC ? a.n_tiles_all : (it.net == NJODE_NET_RO ? 2 * a.n_tiles_loss : __ldg(a.tile_base + a.n_tiles_all));
    const int r0 = (int)((long long)R * it.j / it.J), r1 = (int)((long long)R * (it.j + 1) / it.J);
    const int nhalf = 2 * (r1 - r0);
    const bool aux = it.part == 1;
    const int nb = aux ? 1 : L.kb_main;                       // B blocks
    const int b0 = aux ? L.kb_main : 0;                       // first B block inside the activation image
    const int Nn = aux ? 16 * L.aux_ksteps : L.nt16;
    const int Mh = L.n16 > 128 ? 2 : 1;
    const bool do_bias = (aux ? L.kb_main == 0 : true) && L.b_src >= 0;
    const unsigned char* act = a.act;                         // records of it.net (set by the host per launch)
    const unsigned char* gsp = a.gsp;
    const size_t arec = (size_t)c.act_rec[it.net], grec = (size_t)c.g_rec[it.net];

    if (threadIdx.x == 0) {
        for (int s = 0; s < DW_NSTAGE; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1 + DW_EPI_WARPS); }
        mbar_init(bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < DW_NSTAGE * DW_STAGE_BYTES / 16; i += DW_THREADS)
        st_shared_v4(sbase + 16u * i, 0u, 0u, 0u, 0u);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t ph = 0;
            for (int hidx = 0; hidx < nhalf; ++hidx) {
                const int rec = r0 + (hidx >> 1), h = hidx & 1;
                const unsigned char* gi = gsp + (size_t)rec * grec + L.g_off + (size_t)h * DW_HALF_BLOCK;
                const unsigned char* ai = act + (size_t)rec * arec + L.act_off + (size_t)b0 * A_BLOCK_BYTES + (size_t)h * DW_HALF_BLOCK;
                const uint32_t sb = sbase + (uint32_t)stage * DW_STAGE_BYTES;
                mbar_wait(bar_empty + 8 * stage, ph ^ 1);
                mbar_expect_tx(bar_full + 8 * stage, (uint32_t)(L.kt_blocks + nb) * DW_HALF_BLOCK);
                for (int kb = 0; kb < L.kt_blocks; ++kb)
                    bulk_g2s(sb + kb * DW_HALF_BLOCK, gi + (size_t)kb * A_BLOCK_BYTES, DW_HALF_BLOCK, bar_full + 8 * stage);
                for (int b = 0; b < nb; ++b)
                    bulk_g2s(sb + 4 * DW_HALF_BLOCK + b * DW_HALF_BLOCK, ai + (size_t)b * A_BLOCK_BYTES, DW_HALF_BLOCK, bar_full + 8 * stage);
                if (++stage == DW_NSTAGE) { stage = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int stage = 0; uint32_t ph = 0;
            const uint32_t idesc = make_idesc(Nn) | (1u << 15) | (1u << 16);        // both operands MN-major
            for (int hidx = 0; hidx < nhalf; ++hidx) {
                const uint32_t sb = sbase + (uint32_t)stage * DW_STAGE_BYTES;
                mbar_wait(bar_full + 8 * stage, ph);
                tc_fence_after();
                for (int mh = 0; mh < Mh; ++mh)
                    for (int k = 0; k < 4; ++k)
                        tc_mma(tmem + (uint32_t)mh * 256u, make_desc_mn(sb + (uint32_t)mh * 2 * DW_HALF_BLOCK + 2048u * k),
                               make_desc_mn(sb + 4 * DW_HALF_BLOCK + 2048u * k), idesc, (hidx | k) ? 1u : 0u);
                tc_commit(bar_empty + 8 * stage);
                if (++stage == DW_NSTAGE) { stage = 0; ph ^= 1; }
            }
            tc_commit(bar_acc);
        }
    } else if (warp >= EPI_WARP0) {
        const int q = warp & 3, hf = (warp - EPI_WARP0) >> 2;
        const int o = threadIdx.x - EPI_WARP0 * 32;               // bias: this thread's output column
        float bsum = 0.f;
        {
            int stage = 0; uint32_t ph = 0;
            for (int hidx = 0; hidx < nhalf; ++hidx) {
                mbar_wait(bar_full + 8 * stage, ph);
                if (do_bias && o < L.n16) {
                    const unsigned char* g = smem + (size_t)stage * DW_STAGE_BYTES + (size_t)(o >> 6) * DW_HALF_BLOCK + (o & 7) * 2;
                    const int cch = (o & 63) >> 3;
#pragma unroll 8
                    for (int row = 0; row < 64; ++row) {
                        const unsigned short b = *reinterpret_cast<const unsigned short*>(g + row * 128 + ((cch ^ (row & 7)) << 4));
                        bsum += __uint_as_float((uint32_t)b << 16);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_empty + 8 * stage);
                if (++stage == DW_NSTAGE) { stage = 0; ph ^= 1; }
            }
        }
        float* part = a.dw_part + (size_t)it.slot * (256 * 256);
        float* bpart = a.dw_part + (size_t)c.dw_slots * (256 * 256) + (size_t)it.slot * 256;
        bpart[o] = bsum;
        if (nhalf > 0) { mbar_wait(bar_acc, 0); tc_fence_after(); }
        const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
        const int nch = (Nn + 31) >> 5, per = (nch + 1) >> 1;
        const int c_lo = hf * per, c_hi = min(nch, c_lo + per);
        for (int mh = 0; mh < Mh; ++mh)
            for (int ch = c_lo; ch < c_hi; ++ch) {
                float v[32];
                if (nhalf > 0) { tmem_ld32(tlane + (uint32_t)mh * 256u + ch * 32, reinterpret_cast<uint32_t*>(v)); tmem_wait_ld(); }
                else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = 0.f;
                }
                store_row_chunk(part + (size_t)(mh * 128 + q * 32 + lane) * 256, ch * 32, 256, v);
            }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(TMEM_COLS) : "memory");
    }
}

#endif  // !NJODE_HOST_SIM

}  // namespace njw
