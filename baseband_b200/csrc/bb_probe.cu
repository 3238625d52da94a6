// Bandwidth probes (sm_100a): the ceilings bench.py reports next to the
// decode kernels, measured in the same run with the same launch shape as the
// kernels themselves (one-shot grid, 256 threads, 8 float4 per thread, every
// warp store 512 contiguous bytes).  A write-dominated decoder is bounded by
// the pure-write rate, not by the 50/50 copy rate.
#include "bb_runtime.cuh"

namespace bb {

constexpr int kProbeBlock = 256, kProbeF4 = 8;

// pattern 0: contiguous (warp store = 512 B); pattern 1: the ROWGROUP shape
// (a warp store is 8 pieces of 64 B, 512 B apart -- what the 16-thread VDIF
// decoder did before the TILE mode).
template <int PATTERN>
__global__ void __launch_bounds__(kProbeBlock)
k_probe_fill(float4 *dst, unsigned long long n4, float value) {
    const unsigned long long base =
        (unsigned long long)blockIdx.x * (kProbeBlock * kProbeF4);
    const float4 v = make_float4(value, value, value, value);
#pragma unroll
    for (int j = 0; j < kProbeF4; ++j) {
        unsigned long long i;
        if (PATTERN == 0) {
            i = base + threadIdx.x + (unsigned long long)j * kProbeBlock;
        } else {
            // lane (piece p = lane / 4, part = lane % 4) -> float4
            // (p * 8 + j) * 4 + part of the warp's 256-float4 run
            const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
            i = base + warp * (32 * kProbeF4)
                + ((lane >> 2) * kProbeF4 + j) * 4 + (lane & 3u);
        }
        if (i < n4) dst[i] = v;
    }
}

__global__ void __launch_bounds__(kProbeBlock)
k_probe_copy(float4 *dst, const float4 *src, unsigned long long n4) {
    const unsigned long long base =
        (unsigned long long)blockIdx.x * (kProbeBlock * kProbeF4);
    float4 v[kProbeF4];
#pragma unroll
    for (int j = 0; j < kProbeF4; ++j) {
        const unsigned long long i = base + threadIdx.x
            + (unsigned long long)j * kProbeBlock;
        v[j] = i < n4 ? src[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < kProbeF4; ++j) {
        const unsigned long long i = base + threadIdx.x
            + (unsigned long long)j * kProbeBlock;
        if (i < n4) dst[i] = v[j];
    }
}

}  // namespace bb

using namespace bb;

extern "C" int bb_probe_fill(void *dst, int64_t nbytes, int32_t pattern,
                             void *stream) {
    if (!dst || nbytes < 0 || (nbytes & 15) || !aligned(dst, 16))
        return set_error(BB_ERR_ARGUMENT,
                         "dst must be 16-byte aligned, nbytes a multiple of 16");
    if (nbytes == 0) return BB_OK;
    const unsigned long long n4 = (unsigned long long)nbytes / 16;
    const unsigned long long per = kProbeBlock * kProbeF4;
    const unsigned long long grid = (n4 + per - 1) / per;
    if (grid > 0x7fffffffull)
        return set_error(BB_ERR_ARGUMENT, "buffer too large for one launch");
    if (pattern == 1)
        k_probe_fill<1><<<(unsigned)grid, kProbeBlock, 0, as_stream(stream)>>>(
            (float4 *)dst, n4, 1.0f);
    else
        k_probe_fill<0><<<(unsigned)grid, kProbeBlock, 0, as_stream(stream)>>>(
            (float4 *)dst, n4, 1.0f);
    BB_CHECK_LAUNCH("bb_probe_fill");
    return BB_OK;
}

extern "C" int bb_probe_copy(void *dst, const void *src, int64_t nbytes,
                             void *stream) {
    if (!dst || !src || nbytes < 0 || (nbytes & 15) || !aligned(dst, 16)
        || !aligned(src, 16))
        return set_error(BB_ERR_ARGUMENT,
                         "buffers must be 16-byte aligned, nbytes a multiple "
                         "of 16");
    if (nbytes == 0) return BB_OK;
    const unsigned long long n4 = (unsigned long long)nbytes / 16;
    const unsigned long long per = kProbeBlock * kProbeF4;
    const unsigned long long grid = (n4 + per - 1) / per;
    if (grid > 0x7fffffffull)
        return set_error(BB_ERR_ARGUMENT, "buffer too large for one launch");
    k_probe_copy<<<(unsigned)grid, kProbeBlock, 0, as_stream(stream)>>>(
        (float4 *)dst, (const float4 *)src, n4);
    BB_CHECK_LAUNCH("bb_probe_copy");
    return BB_OK;
}
