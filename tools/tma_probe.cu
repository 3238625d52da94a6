// Evidence microbenchmark (not product code): does staging the packed VDIF
// payloads in shared memory with TMA bulk copies (cp.async.bulk + mbarrier)
// beat plain coalesced global loads for the 2-bit, 16-thread decode (C2)?
// Both kernels produce the same (nsample, 16) float32 output.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_probe tma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int NTHREAD = 16, PAYLOAD = 8000, FRAME = 8032, NWORD = 2000;
constexpr int SPF = 32000;
struct alignas(16) F4 { float x, y, z, w; };
struct alignas(8) F2 { float x, y; };

__device__ __forceinline__ F2 pair(const float *lut, uint32_t w, int m) {
    return reinterpret_cast<const F2 *>(lut)[(w >> (4 * m)) & 15u];
}

__device__ __forceinline__ void emit16(const float *lut, const uint32_t w[4],
                                       float *dst) {
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        F2 a = pair(lut, w[0], m), b = pair(lut, w[1], m);
        F2 c = pair(lut, w[2], m), d = pair(lut, w[3], m);
        *reinterpret_cast<F4 *>(dst) = F4{a.x, b.x, c.x, d.x};
        dst += NTHREAD;
        *reinterpret_cast<F4 *>(dst) = F4{a.y, b.y, c.y, d.y};
        dst += NTHREAD;
    }
}

__device__ void build_lut(float *lut) {
    const float lv[4] = {-3.316505f, -1.f, 1.f, 3.316505f};
    if (threadIdx.x < 32) lut[threadIdx.x] =
        lv[(threadIdx.x & 1) ? (threadIdx.x >> 3) : ((threadIdx.x >> 1) & 3)];
    __syncthreads();
}

// A: direct loads (production decomposition: item = (set, word, group of 4)).
__global__ void __launch_bounds__(256) k_direct(const uint8_t *src, float *out,
                                                uint32_t nitems) {
    __shared__ __align__(128) float lut[32];
    build_lut(lut);
    uint32_t item = blockIdx.x * 256 + threadIdx.x;
    if (item >= nitems) return;
    uint32_t g = item & 3u, lw = item >> 2;
    uint32_t set = lw / NWORD, k = lw - set * NWORD;
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
        w[j] = *reinterpret_cast<const uint32_t *>(
            src + ((size_t)set * NTHREAD + g * 4 + j) * FRAME + 32 + 4 * k);
    emit16(lut, w, out + ((size_t)lw * 16) * NTHREAD + g * 4);
}

// B: one CTA stages KW words of all 16 thread payloads of one set with 16
// TMA bulk copies, then decodes from shared memory.
constexpr int KW = 200;                       // words per slot per CTA (800 B)
constexpr int PITCH = KW + 4;                 // words; keeps 16-byte alignment
__global__ void __launch_bounds__(256) k_tma(const uint8_t *src, float *out,
                                             uint32_t nset) {
    __shared__ __align__(128) float lut[32];
    __shared__ __align__(128) uint32_t tile[NTHREAD * PITCH];
    __shared__ __align__(8) unsigned long long bar;
    const uint32_t per_set = NWORD / KW;      // 10 CTAs per frame set
    const uint32_t set = blockIdx.x / per_set;
    const uint32_t k0 = (blockIdx.x - set * per_set) * KW;
    if (set >= nset) return;
    const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_s));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    build_lut(lut);                            // includes __syncthreads
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                     :: "r"(bar_s), "r"(NTHREAD * KW * 4) : "memory");
        for (int s = 0; s < NTHREAD; ++s) {
            const uint8_t *g = src + ((size_t)set * NTHREAD + s) * FRAME + 32
                + 4 * k0;
            uint32_t d = (uint32_t)__cvta_generic_to_shared(tile + s * PITCH);
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes"
                " [%0], [%1], %2, [%3];"
                :: "r"(d), "l"(g), "r"(KW * 4), "r"(bar_s) : "memory");
        }
    }
    asm volatile(
        "{\n .reg .pred p;\n WAIT_%=:\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
        " @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n"
        :: "r"(bar_s) : "memory");
    for (uint32_t item = threadIdx.x; item < KW * 4; item += 256) {
        uint32_t g = item & 3u, k = item >> 2;
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) w[j] = tile[(g * 4 + j) * PITCH + k];
        size_t lw = (size_t)set * NWORD + k0 + k;
        emit16(lut, w, out + (lw * 16) * NTHREAD + g * 4);
    }
}

int main(int argc, char **argv) {
    double gib = argc > 1 ? atof(argv[1]) : 1.0;
    uint32_t nset = (uint32_t)(gib * (1 << 30) / (NTHREAD * FRAME));
    size_t nbytes = (size_t)nset * NTHREAD * FRAME;
    size_t nout = (size_t)nset * SPF * NTHREAD;
    uint8_t *src; float *a, *b;
    cudaMalloc(&src, nbytes); cudaMalloc(&a, nout * 4); cudaMalloc(&b, nout * 4);
    std::vector<uint8_t> h(nbytes);
    uint32_t x = 12345;
    for (auto &v : h) { x = x * 1664525u + 1013904223u; v = x >> 24; }
    cudaMemcpy(src, h.data(), nbytes, cudaMemcpyHostToDevice);
    uint32_t nitems = nset * NWORD * 4;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best[2] = {1e9f, 1e9f};
    for (int rep = 0; rep < 8; ++rep) {
        cudaEventRecord(e0);
        k_direct<<<(nitems + 255) / 256, 256>>>(src, a, nitems);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best[0]) best[0] = ms;
        cudaEventRecord(e0);
        k_tma<<<nset * (NWORD / KW), 256>>>(src, b, nset);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); if (ms < best[1]) best[1] = ms;
    }
    cudaError_t err = cudaGetLastError();
    std::vector<float> ha(1 << 20), hb(1 << 20);
    cudaMemcpy(ha.data(), a + nout - (1 << 20), 4 << 20, cudaMemcpyDeviceToHost);
    cudaMemcpy(hb.data(), b + nout - (1 << 20), 4 << 20, cudaMemcpyDeviceToHost);
    bool same = ha == hb;
    double tot = (double)nbytes + (double)nout * 4;
    printf("C2 decode, %.2f GiB packed (%s, outputs %s)\n", gib,
           cudaGetErrorString(err), same ? "identical" : "DIFFER");
    printf("  direct coalesced LDG.32          : %7.1f GB/s (%.3f ms)\n",
           tot / best[0] / 1e6, best[0]);
    printf("  TMA bulk copy -> smem -> decode  : %7.1f GB/s (%.3f ms)\n",
           tot / best[1] / 1e6, best[1]);
    return 0;
}
