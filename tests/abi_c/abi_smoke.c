/* Plain-C client of libbaseband_b200.so: no Python, no torch.  Decodes a
 * small 2-bit VDIF-like frame set through the C ABI and checks every value
 * against a straightforward C restatement, then encodes it back.
 *   gcc -O1 -I include tests/abi_c/abi_smoke.c -ldl -o abi_smoke
 *   ./abi_smoke baseband_b200/libbaseband_b200.so                         */
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "baseband_b200.h"

#define LOAD(name) \
    __typeof__(name) *p_##name = (__typeof__(name) *)dlsym(lib, #name); \
    if (!p_##name) { fprintf(stderr, "missing symbol %s\n", #name); return 2; }

int main(int argc, char **argv) {
    void *lib = dlopen(argc > 1 ? argv[1] : "libbaseband_b200.so", RTLD_NOW);
    if (!lib) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
    LOAD(bb_abi_version) LOAD(bb_last_error) LOAD(bb_device_count)
    LOAD(bb_set_device) LOAD(bb_malloc) LOAD(bb_free) LOAD(bb_memcpy_h2d)
    LOAD(bb_memcpy_d2h) LOAD(bb_stream_create) LOAD(bb_stream_synchronize)
    LOAD(bb_stream_destroy) LOAD(bb_decode_bitfield) LOAD(bb_encode_bitfield)
    LOAD(bb_vdif_scan) LOAD(bb_memset)
    if (p_bb_abi_version() != BB_ABI_VERSION) return 3;
    if (p_bb_device_count() < 1) { fprintf(stderr, "no CUDA device\n"); return 4; }
    if (p_bb_set_device(0) != BB_OK) return 5;

    enum { NSET = 3, NTHREAD = 8, PAYLOAD = 512, HDR = 32,
           FRAME = PAYLOAD + HDR, NFRAME = NSET * NTHREAD,
           SPF = PAYLOAD * 4, NSAMPLE = NSET * SPF };
    const float levels[4] = {-3.316505f, -1.f, 1.f, 3.316505f};
    uint8_t *raw = malloc((size_t)NFRAME * FRAME);
    uint32_t x = 2463534242u;
    for (size_t i = 0; i < (size_t)NFRAME * FRAME; ++i) {
        x ^= x << 13; x ^= x >> 17; x ^= x << 5;
        raw[i] = (uint8_t)(x >> 11);
    }
    /* headers: thread ids in reversed order, frame 5 flagged invalid */
    for (int f = 0; f < NFRAME; ++f) {
        uint32_t w[8] = {0};
        int set = f / NTHREAD, tid = NTHREAD - 1 - f % NTHREAD;
        w[0] = 100u | (f == 5 ? 0x80000000u : 0u);
        w[1] = (uint32_t)set;
        w[2] = (1u << 29) | (FRAME / 8);
        w[3] = (1u << 26) | ((uint32_t)tid << 16);
        memcpy(raw + (size_t)f * FRAME, w, 32);
    }
    int32_t slot_host[1024];
    for (int i = 0; i < 1024; ++i) slot_host[i] = i < NTHREAD ? i : -1;

    void *d_raw, *d_slot, *d_fields, *d_uo, *d_bad, *d_out, *d_back, *stream;
    if (p_bb_malloc(&d_raw, (int64_t)NFRAME * FRAME) || p_bb_malloc(&d_slot, 4096)
        || p_bb_malloc(&d_fields, 4 * BB_VDIF_NFIELD * NFRAME)
        || p_bb_malloc(&d_uo, 8 * NFRAME) || p_bb_malloc(&d_bad, 4)
        || p_bb_malloc(&d_out, 4 * (int64_t)NSAMPLE * NTHREAD)
        || p_bb_malloc(&d_back, (int64_t)NFRAME * FRAME)
        || p_bb_stream_create(&stream)) {
        fprintf(stderr, "alloc: %s\n", p_bb_last_error()); return 6;
    }
    p_bb_memcpy_h2d(d_raw, raw, (int64_t)NFRAME * FRAME, stream);
    p_bb_memcpy_h2d(d_slot, slot_host, 4096, stream);
    p_bb_memset(d_bad, 0, 4, stream);
    p_bb_memset(d_back, 0, (int64_t)NFRAME * FRAME, stream);
    int rc = p_bb_vdif_scan(d_raw, NULL, FRAME, NFRAME, HDR, NTHREAD, NTHREAD,
                            d_slot, d_fields, d_uo, d_bad, 0, 0, 0, 0, stream);
    if (rc) { fprintf(stderr, "scan: %s\n", p_bb_last_error()); return 7; }
    rc = p_bb_decode_bitfield(d_raw, d_uo, NSET, NTHREAD, PAYLOAD, 2, 1, 0,
                              BB_CODEC_LEVELS, levels, -9.f, 0, NSAMPLE,
                              d_out, stream);
    if (rc) { fprintf(stderr, "decode: %s\n", p_bb_last_error()); return 8; }
    rc = p_bb_encode_bitfield(d_out, BB_F32, d_back, d_uo, NSET, NTHREAD,
                              PAYLOAD, 2, 1, BB_QUANT_OFFSET_BINARY, stream);
    if (rc) { fprintf(stderr, "encode: %s\n", p_bb_last_error()); return 9; }
    float *out = malloc(4 * (size_t)NSAMPLE * NTHREAD);
    uint8_t *back = malloc((size_t)NFRAME * FRAME);
    int32_t bad = -1;
    p_bb_memcpy_d2h(out, d_out, 4 * (int64_t)NSAMPLE * NTHREAD, stream);
    p_bb_memcpy_d2h(back, d_back, (int64_t)NFRAME * FRAME, stream);
    p_bb_memcpy_d2h(&bad, d_bad, 4, stream);
    p_bb_stream_synchronize(stream);
    if (bad != 0) { fprintf(stderr, "inconsistent frames: %d\n", bad); return 10; }

    long errors = 0;
    for (int f = 0; f < NFRAME; ++f) {
        int set = f / NTHREAD, tid = NTHREAD - 1 - f % NTHREAD;
        const uint8_t *pl = raw + (size_t)f * FRAME + HDR;
        for (int s = 0; s < SPF; ++s) {
            float want = f == 5 ? -9.f
                : levels[(pl[s / 4] >> (2 * (s % 4))) & 3];
            float got = out[((size_t)set * SPF + s) * NTHREAD + tid];
            errors += memcmp(&want, &got, 4) != 0;
        }
        if (f != 5)
            errors += memcmp(back + (size_t)f * FRAME + HDR, pl, PAYLOAD) != 0;
    }
    p_bb_free(d_raw); p_bb_free(d_slot); p_bb_free(d_fields); p_bb_free(d_uo);
    p_bb_free(d_bad); p_bb_free(d_out); p_bb_free(d_back);
    p_bb_stream_destroy(stream);
    printf("abi_smoke: %d frames, %d samples/thread, %ld mismatches\n",
           NFRAME, NSAMPLE, errors);
    return errors ? 1 : 0;
}
