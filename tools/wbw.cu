// Write-bandwidth micro-benchmark for B200: how fast can a kernel stream
// float4 stores to HBM, and does the store flavour / shape matter?
// Development tool (not part of the library).  nvcc -O3 -arch=sm_100a.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
    printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

enum { ST_DEFAULT, ST_CS, ST_NOALLOC, ST_WT, ST_EVICT_FIRST };

template <int KIND>
__device__ __forceinline__ void st4(float4 *p, float4 v) {
    if (KIND == ST_DEFAULT) *p = v;
    else if (KIND == ST_CS) __stcs(p, v);
    else if (KIND == ST_WT) __stwt(p, v);
    else if (KIND == ST_NOALLOC)
        asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                     :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    else {
        uint64_t pol;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;"
                     :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
    }
}

// grid-stride, consecutive float4 per lane (512 B per warp instruction)
template <int KIND, int UNROLL>
__global__ void __launch_bounds__(256) k_fill(float4 *out, size_t n4) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    float4 v = make_float4(1.f, 2.f, 3.f, (float)threadIdx.x);
    for (; i + (UNROLL - 1) * stride < n4; i += UNROLL * stride) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) st4<KIND>(out + i + u * stride, v);
    }
    for (; i < n4; i += stride) st4<KIND>(out + i, v);
}

// the decode kernel's store shape: each warp instruction writes 8 segments of
// 64 B, 1 KiB apart; a lane writes 16 consecutive rows of 64 B.
template <int KIND>
__global__ void __launch_bounds__(256) k_fill_rows(float4 *out, size_t nitems) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    float4 v = make_float4(1.f, 2.f, 3.f, (float)threadIdx.x);
    for (size_t item = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
         item < nitems; item += stride) {
        size_t lw = item >> 2, g = item & 3;
        float4 *dst = out + lw * 64 + g;         // 16 rows x 4 float4 per lw
#pragma unroll
        for (int i = 0; i < 16; ++i) st4<KIND>(dst + i * 4, v);
    }
}

// block-contiguous: each CTA owns a contiguous 32 KiB tile per iteration
template <int KIND>
__global__ void __launch_bounds__(256) k_fill_tiles(float4 *out, size_t ntiles) {
    float4 v = make_float4(1.f, 2.f, 3.f, (float)threadIdx.x);
    for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        float4 *dst = out + t * 2048 + threadIdx.x;
#pragma unroll
        for (int i = 0; i < 8; ++i) st4<KIND>(dst + i * 256, v);
    }
}

// TMA bulk store from shared memory: one elected thread per CTA issues
// cp.async.bulk.global.shared::cta of a TILE-byte tile.
template <int TILE>
__global__ void __launch_bounds__(128) k_fill_tma(char *out, size_t ntiles) {
    extern __shared__ __align__(128) char smem[];
    for (int i = threadIdx.x; i < TILE / 16; i += blockDim.x)
        reinterpret_cast<float4 *>(smem)[i] = make_float4(1.f, 2.f, 3.f, 4.f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
        for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         :: "l"(out + t * TILE), "r"(s), "r"(TILE) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

template <typename F>
static float time_ms(F launch, int reps = 7) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 2; ++i) launch();
    std::vector<float> t;
    for (int i = 0; i < reps; ++i) {
        cudaEventRecord(a); launch(); cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); t.push_back(ms);
    }
    std::sort(t.begin(), t.end());
    return t[t.size() / 2];
}

int main(int argc, char **argv) {
    size_t nbytes = (size_t)8 << 30;
    char *buf;
    CK(cudaMalloc(&buf, nbytes));
    float4 *out = (float4 *)buf;
    size_t n4 = nbytes / 16;
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("SMs %d, buffer %.1f GiB\n", sms, nbytes / 1073741824.0);
    float ms = time_ms([&] { cudaMemsetAsync(buf, 1, nbytes); });
    printf("%-44s %8.1f GB/s\n", "cudaMemset", nbytes / ms / 1e6);
#define RUN(name, expr) do { float m = time_ms([&] { expr; }); \
    cudaError_t e = cudaGetLastError(); \
    printf("%-44s %8.1f GB/s %s\n", name, nbytes / m / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e)); } while (0)
    for (int cps : {2, 4, 8, 16}) {
        int grid = sms * cps;
        char nm[96];
        snprintf(nm, 96, "fill default u1 grid=%dxSM", cps); RUN(nm, (k_fill<ST_DEFAULT, 1><<<grid, 256>>>(out, n4)));
        snprintf(nm, 96, "fill default u4 grid=%dxSM", cps); RUN(nm, (k_fill<ST_DEFAULT, 4><<<grid, 256>>>(out, n4)));
        snprintf(nm, 96, "fill .cs u4 grid=%dxSM", cps); RUN(nm, (k_fill<ST_CS, 4><<<grid, 256>>>(out, n4)));
        snprintf(nm, 96, "fill .wt u4 grid=%dxSM", cps); RUN(nm, (k_fill<ST_WT, 4><<<grid, 256>>>(out, n4)));
        snprintf(nm, 96, "fill no_allocate u4 grid=%dxSM", cps); RUN(nm, (k_fill<ST_NOALLOC, 4><<<grid, 256>>>(out, n4)));
        snprintf(nm, 96, "fill evict_first u4 grid=%dxSM", cps); RUN(nm, (k_fill<ST_EVICT_FIRST, 4><<<grid, 256>>>(out, n4)));
        snprintf(nm, 96, "rows(8x64B) default grid=%dxSM", cps); RUN(nm, (k_fill_rows<ST_DEFAULT><<<grid, 256>>>(out, n4 / 16)));
        snprintf(nm, 96, "rows(8x64B) .cs grid=%dxSM", cps); RUN(nm, (k_fill_rows<ST_CS><<<grid, 256>>>(out, n4 / 16)));
        snprintf(nm, 96, "tiles 32KiB default grid=%dxSM", cps); RUN(nm, (k_fill_tiles<ST_DEFAULT><<<grid, 256>>>(out, n4 / 2048)));
        snprintf(nm, 96, "tiles 32KiB .cs grid=%dxSM", cps); RUN(nm, (k_fill_tiles<ST_CS><<<grid, 256>>>(out, n4 / 2048)));
    }
    RUN("fill default u1 full grid", (k_fill<ST_DEFAULT, 1><<<(unsigned)(n4 / 256), 256>>>(out, n4)));
    RUN("fill .cs u1 full grid", (k_fill<ST_CS, 1><<<(unsigned)(n4 / 256), 256>>>(out, n4)));
    cudaFuncSetAttribute(k_fill_tma<16384>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    cudaFuncSetAttribute(k_fill_tma<32768>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    cudaFuncSetAttribute(k_fill_tma<65536>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int cps : {1, 2, 4}) {
        char nm[96];
        snprintf(nm, 96, "TMA bulk store 16KiB grid=%dxSM", cps); RUN(nm, (k_fill_tma<16384><<<sms * cps, 128, 16384>>>(buf, nbytes / 16384)));
        snprintf(nm, 96, "TMA bulk store 32KiB grid=%dxSM", cps); RUN(nm, (k_fill_tma<32768><<<sms * cps, 128, 32768>>>(buf, nbytes / 32768)));
        snprintf(nm, 96, "TMA bulk store 64KiB grid=%dxSM", cps); RUN(nm, (k_fill_tma<65536><<<sms * cps, 128, 65536>>>(buf, nbytes / 65536)));
    }
    // read+write copy for reference
    return 0;
}
