// njode_hash.cuh -- the counter-based dropout keys and the fast tanh shared by the fp32 kernels
// (njode_core.cuh / njode_seg.cuh) and the tensor-core kernels (njode_wide.cuh).
#pragma once
#include <stdint.h>
#include <math.h>
#ifndef NJ_HD
#if defined(NJODE_HOST_SIM)
#define NJ_HD inline
#else
#define NJ_HD __device__ __forceinline__
#endif
#endif

// ------------------------------------------------------------------------------------------------
// dropout: counter-based keep-mask (murmur3 finaliser chain); restated in oracle/njode_oracle.py
// ------------------------------------------------------------------------------------------------
NJ_HD unsigned nj_fmix32(unsigned h) {
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}
NJ_HD unsigned nj_row_key(unsigned seed_lo, unsigned seed_hi, unsigned path, unsigned event) {
    return nj_fmix32(nj_fmix32(path ^ seed_lo) + event * 0x9E3779B9u) ^ seed_hi;
}
NJ_HD unsigned nj_layer_key(unsigned row_key, unsigned tag) { return nj_fmix32(row_key + tag * 0x85EBCA77u); }
// one 32-bit hash word serves the two neurons o and o ^ 8 (16-bit fields): word index
// (o & 7) | ((o >> 4) << 3), field (o >> 3) & 1; keep iff field >= thr16 = floor(p * 65536).
NJ_HD unsigned nj_keep_word(unsigned layer_key, unsigned widx) { return nj_fmix32(layer_key + widx * 0xC2B2AE3Du); }
NJ_HD bool nj_keep(unsigned layer_key, unsigned neuron, unsigned thr16) {
    const unsigned w = nj_keep_word(layer_key, (neuron & 7u) | ((neuron >> 4) << 3));
    return (((neuron >> 3) & 1u) ? (w >> 16) : (w & 0xFFFFu)) >= thr16;
}
#define NJ_EVENT_JUMP_BASE 0x40000000u
#define NJ_EVENT_PATH_RO_BASE 0x20000000u
#define NJ_EVENT_INIT 0x7FFFFFFFu

// tanh(x) = 1 - 2 / (exp(2x) + 1): two MUFU ops (ex2, rcp) + three FMA-pipe ops; absolute error
// <= ~2e-7 over the whole range (saturates correctly at +-1), versus ~20 instructions for tanhf.
NJ_HD float nj_tanh(float x) {
#if defined(NJODE_HOST_SIM)
    return 1.f - 2.f / (exp2f(x * 2.8853900817779268f) + 1.f);
#else
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.8853900817779268f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.f));
    return fmaf(-2.f, r, 1.f);
#endif
}
