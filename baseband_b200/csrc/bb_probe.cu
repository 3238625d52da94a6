// Bandwidth probes (sm_100a): the ceilings bench.py reports next to the
// decode kernels, measured in the same run with the same launch shape as the
// kernels themselves (one-shot grid, 256 threads, 8 float4 per thread, every
// warp store 512 contiguous bytes).  A write-dominated decoder is bounded by
// the pure-write rate, not by the 50/50 copy rate.
#include <stdlib.h>
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

// 1:16 expansion (2 bit -> float32) with ideal access patterns and no
// arithmetic to speak of: the ceiling for a stream that reads one byte for
// every sixteen it writes.  A warp reads 256 contiguous bytes (pattern 0) or
// -- pattern 1, the shape of 16 interleaved VDIF threads -- 16 pieces of 16
// bytes from 16 streams `stream_stride` bytes apart, and writes 4 KiB
// contiguous, 512 bytes per store instruction.  Pattern 2 = pattern 0 with
// the input wrapped to one MiB (always an L2 hit: no DRAM reads at all).
template <int PATTERN>
__global__ void __launch_bounds__(kProbeBlock)
k_probe_expand(float4 *dst, const uint32_t *src, unsigned long long n4,
               unsigned long long stream_stride_words, unsigned lead,
               unsigned ahead) {
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    if (PATTERN == 4 && threadIdx.x == 0 && blockIdx.x % lead == 0) {
        // pattern 4 = pattern 0 with the reads clustered in time: one thread
        // in `lead` CTAs asks L2 for the input of `lead` CTAs, `ahead` CTAs
        // before they run (a CTA reads 2 KiB), so that DRAM sees read bursts
        // of lead * 2 KiB instead of a trickle mixed into the writes
        const unsigned long long first =
            ((unsigned long long)blockIdx.x + ahead) * (kProbeBlock / 32) * 256;
        const unsigned long long total = n4 * 16 / 16;      // input bytes
        if (first < total) {
            unsigned long long nb = (unsigned long long)lead
                * (kProbeBlock / 32) * 256;
            if (first + nb > total) nb = (total - first) & ~15ull;
            if (nb)
                asm volatile(
                    "cp.async.bulk.prefetch.L2.global [%0], %1;"
                    :: "l"(reinterpret_cast<const char *>(src) + first),
                       "r"((unsigned)nb) : "memory");
        }
    }
    const unsigned long long chunk =
        (unsigned long long)blockIdx.x * (kProbeBlock / 32) + warp;
    const unsigned long long base = chunk * (32 * kProbeF4);
    if (base >= n4) return;
    uint32_t w0, w1;
    if (PATTERN == 0 || PATTERN == 2 || PATTERN == 4) {
        // pattern 2: the input wraps every MiB, i.e. it stays in L2
        const unsigned long long at = chunk * 32 + lane;
        const uint2 v = reinterpret_cast<const uint2 *>(src)[
            PATTERN == 2 ? (at & 0x1ffffull) : at];
        w0 = v.x;
        w1 = v.y;
    } else {
        // chunk = 4 word positions x 16 streams; lane -> (stream pair, word)
        const unsigned long long k = chunk * 4 + (lane & 3u);
        const unsigned st = (lane >> 2) * 2;
        w0 = src[st * stream_stride_words + k];
        w1 = src[(st + 1) * stream_stride_words + k];
    }
#pragma unroll
    for (int j = 0; j < kProbeF4; ++j) {
        const unsigned long long i = base + j * 32 + lane;
        const uint32_t a = (j & 1 ? w1 : w0) >> (8 * (j >> 1));
        const float4 v = make_float4(__uint_as_float(0x3f800000u | (a & 3u)),
                                     __uint_as_float(0x3f800000u | (a & 12u)),
                                     __uint_as_float(0x3f800000u | (a & 48u)),
                                     __uint_as_float(0x3f800000u | (a & 192u)));
        if (i < n4) dst[i] = v;
    }
}

// 1:4 expansion (8 bit -> float32), ideal patterns: a warp reads 1 KiB
// contiguous (two 16-byte loads per lane) and writes 4 KiB contiguous.
__global__ void __launch_bounds__(kProbeBlock)
k_probe_expand4(float4 *dst, const uint4 *src, unsigned long long n4) {
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const unsigned long long chunk =
        (unsigned long long)blockIdx.x * (kProbeBlock / 32) + warp;
    const unsigned long long base = chunk * (32 * kProbeF4);
    if (base >= n4) return;
    const uint4 a = src[chunk * 64 + lane], b = src[chunk * 64 + 32 + lane];
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < kProbeF4; ++j) {
        const unsigned long long i = base + j * 32 + lane;
        const float4 v = make_float4(
            __uint_as_float(0x3f800000u | (w[j] & 0xffu)),
            __uint_as_float(0x3f800000u | ((w[j] >> 8) & 0xffu)),
            __uint_as_float(0x3f800000u | ((w[j] >> 16) & 0xffu)),
            __uint_as_float(0x3f800000u | (w[j] >> 24)));
        if (i < n4) dst[i] = v;
    }
}

// Pull `nbytes` of src into L2 with evict_last priority (a pure-read phase
// ahead of a kernel that then finds its input in L2).
__global__ void __launch_bounds__(256)
k_probe_prefetch(const uint8_t *src, unsigned long long nbytes) {
    const unsigned long long stride = (unsigned long long)gridDim.x * 256 * 128;
    for (unsigned long long at = ((unsigned long long)blockIdx.x * 256
                                  + threadIdx.x) * 128;
         at < nbytes; at += stride)
        asm volatile("prefetch.global.L2::evict_last [%0];" :: "l"(src + at));
}

// Pure read: every thread loads kProbeF4 16-byte vectors (a warp load covers
// 512 contiguous bytes, all loads in flight before they are combined) and
// stores nothing -- the ceiling of the consumers that only read packed bytes.
__device__ unsigned int g_probe_sink;

__global__ void __launch_bounds__(kProbeBlock)
k_probe_read(const uint4 *src, unsigned long long n4) {
    const unsigned long long base =
        (unsigned long long)blockIdx.x * (kProbeBlock * kProbeF4);
    uint4 v[kProbeF4];
#pragma unroll
    for (int j = 0; j < kProbeF4; ++j) {
        const unsigned long long i = base + j * kProbeBlock + threadIdx.x;
        v[j] = i < n4 ? src[i] : make_uint4(0u, 0u, 0u, 0u);
    }
    unsigned int x = 0u;
#pragma unroll
    for (int j = 0; j < kProbeF4; ++j) x ^= v[j].x ^ v[j].y ^ v[j].z ^ v[j].w;
    if (x == 0x9e3779b9u && n4 == 1ull) g_probe_sink = x;   // never in practice
}

}  // namespace bb

using namespace bb;

extern "C" int bb_probe_read(const void *src, int64_t nbytes, void *stream) {
    if (!src || nbytes < 0 || (nbytes & 15) || !aligned(src, 16))
        return set_error(BB_ERR_ARGUMENT,
                         "src must be 16-byte aligned, nbytes a multiple of 16");
    if (nbytes == 0) return BB_OK;
    const unsigned long long n4 = (unsigned long long)nbytes / 16;
    const unsigned long long per = (unsigned long long)kProbeBlock * kProbeF4;
    const unsigned long long grid = (n4 + per - 1) / per;
    if (grid > 0x7fffffffull)
        return set_error(BB_ERR_ARGUMENT, "buffer too large for one launch");
    k_probe_read<<<(unsigned)grid, kProbeBlock, 0, as_stream(stream)>>>(
        (const uint4 *)src, n4);
    BB_CHECK_LAUNCH("bb_probe_read");
    return BB_OK;
}

extern "C" int bb_probe_prefetch(const void *src, int64_t nbytes,
                                 void *stream) {
    if (!src || nbytes < 0) return set_error(BB_ERR_ARGUMENT, "bad arguments");
    if (nbytes == 0) return BB_OK;
    const unsigned long long lines = ((unsigned long long)nbytes + 127) / 128;
    unsigned long long grid = (lines + 255) / 256;
    const unsigned long long cap = (unsigned long long)sm_count() * 8;
    if (grid > cap) grid = cap;
    k_probe_prefetch<<<(unsigned)grid, 256, 0, as_stream(stream)>>>(
        (const uint8_t *)src, (unsigned long long)nbytes);
    BB_CHECK_LAUNCH("bb_probe_prefetch");
    return BB_OK;
}


extern "C" int bb_probe_expand(void *dst, int64_t nbytes, const void *src,
                               int32_t pattern, void *stream) {
    if (!dst || !src || nbytes < 0 || (nbytes & 4095) || !aligned(dst, 16)
        || !aligned(src, 8))
        return set_error(BB_ERR_ARGUMENT,
                         "dst 16-byte, src 8-byte aligned, nbytes a multiple "
                         "of 4096");
    if (nbytes == 0) return BB_OK;
    const unsigned long long n4 = (unsigned long long)nbytes / 16;
    const unsigned long long per = kProbeBlock * kProbeF4;
    const unsigned long long grid = (n4 + per - 1) / per;
    if (grid > 0x7fffffffull)
        return set_error(BB_ERR_ARGUMENT, "buffer too large for one launch");
    // pattern 1: 16 streams of nbytes / 16 / 16 bytes each
    const unsigned long long stride_words = (unsigned long long)nbytes / 1024;
    unsigned lead = 512, ahead = 2048;
    if (const char *e = getenv("BB_PROBE_LEAD")) lead = (unsigned)atoi(e);
    if (const char *e = getenv("BB_PROBE_AHEAD")) ahead = (unsigned)atoi(e);
    if (lead < 1) lead = 1;
    if (pattern == 4) {
        k_probe_expand<4><<<(unsigned)grid, kProbeBlock, 0,
                            as_stream(stream)>>>(
            (float4 *)dst, (const uint32_t *)src, n4, stride_words, lead,
            ahead);
    } else if (pattern == 3) {
        if (!aligned(src, 16))
            return set_error(BB_ERR_ALIGNMENT, "src must be 16-byte aligned");
        k_probe_expand4<<<(unsigned)grid, kProbeBlock, 0, as_stream(stream)>>>(
            (float4 *)dst, (const uint4 *)src, n4);
    } else if (pattern == 1)
        k_probe_expand<1><<<(unsigned)grid, kProbeBlock, 0,
                            as_stream(stream)>>>(
            (float4 *)dst, (const uint32_t *)src, n4, stride_words, lead, ahead);
    else if (pattern == 2)
        k_probe_expand<2><<<(unsigned)grid, kProbeBlock, 0,
                            as_stream(stream)>>>(
            (float4 *)dst, (const uint32_t *)src, n4, stride_words, lead, ahead);
    else
        k_probe_expand<0><<<(unsigned)grid, kProbeBlock, 0,
                            as_stream(stream)>>>(
            (float4 *)dst, (const uint32_t *)src, n4, stride_words, lead, ahead);
    BB_CHECK_LAUNCH("bb_probe_expand");
    return BB_OK;
}


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
