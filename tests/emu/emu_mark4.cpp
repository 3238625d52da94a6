// CPU emulation of the Mark 4 kernels — TEST INFRASTRUCTURE ONLY (see
// emu_bitfield.cpp).
#include <string>
#include <vector>
#include "../../baseband_b200/csrc/bb_bitfield.cuh"
#include "../../baseband_b200/csrc/bb_mark4_plan.h"
#include "../../include/baseband_b200.h"

using namespace bb;

static std::string g_err4;

static void run_dec(const std::vector<M4Launch> &launches) {
    for (const M4Launch &l : launches) {
        alignas(16) float lut[DecodeLut<2>::kFloats];
        for (int i = 0; i < DecodeLut<2>::kFloats; ++i) {
            int entry = i >> 1, which = i & 1;
            int code = which ? (entry >> 2) : (entry & 3);
            lut[i] = l.g.levels[2 * (code & 1) + (code >> 1)];
        }
        if (l.mode == M4_WARP) {
            for (uint32_t chunk = 0; chunk < l.g.nitems / 32; ++chunk) {
                uint32_t w[32];
                bool ok[32];
                for (uint32_t lane = 0; lane < 32; ++lane)
                    ok[lane] = m4w_load(l.g, chunk, lane, w[lane]);
                const bool interior = m4w_interior(l.g, chunk);
                for (uint32_t q = 0; q < 128; ++q) {
                    const M4Lane lc = m4w_lane(l.g, l.g.pos, q & 31u);
                    uint32_t src = m4w_src_lane(l.g, lc, q);
                    if (interior && l.g.std4) {
                        const uint32_t ss = m4s_src_lane(l.g, q);
                        m4s_emit_fast(l.g, lut, m4w_chunk_out(l.g, chunk), q,
                                      m4s_row(l.g, q), m4_reorder32(w[ss]),
                                      ok[ss]);
                    } else if (interior)
                        m4w_emit_fast(l.g, lc, lut, m4w_chunk_out(l.g, chunk),
                                      q, w[src], ok[src]);
                    else
                        m4w_emit(l.g, lc, l.g.levels, chunk, q, w[src],
                                 ok[src]);
                }
            }
            continue;
        }
        for (uint32_t item = 0; item < l.g.nitems; ++item) {
            if (l.mode == M4_FAST) m4_dec_fast(l.g, lut, item);
            else if (l.mode == M4_GENERIC_VEC) m4_dec_generic<true>(l.g, item);
            else m4_dec_generic<false>(l.g, item);
        }
    }
}

template <typename T>
static void run_enc(const std::vector<M4Launch> &launches) {
    const QuantConsts<T> c = make_quant_consts<T>();
    for (const M4Launch &l : launches)
        for (uint32_t item = 0; item < l.g.nitems; ++item) {
            if (l.mode == M4_FAST) m4_enc_fast<T>(l.g, c, item);
            else if (l.mode == M4_HALF) m4_enc_half<T>(l.g, l.g.pos, c, item);
            else m4_enc_generic<T>(l.g, c, item);
        }
}

extern "C" {

const char *emu_mark4_error(void) { return g_err4.c_str(); }

int bb_mark4_decode(const void *src, const int64_t *unit_offset,
                    int64_t nframe, int32_t nchan, int32_t fanout, int32_t ft,
                    const float *levels_host, float fill_value,
                    int64_t sample_start, int64_t nsample, float *out,
                    void *stream) {
    std::vector<M4Launch> l;
    if (!plan_m4_frames(false, src, unit_offset, nframe, nchan, fanout, ft,
                        levels_host, fill_value, sample_start, nsample, out,
                        nullptr, l, g_err4))
        return BB_ERR_ARGUMENT;
    run_dec(l);
    return 0;
}

int bb_mark4_encode(const void *in, int32_t in_dtype, void *dst,
                    const int64_t *unit_offset, int64_t nframe, int32_t nchan,
                    int32_t fanout, int32_t ft, void *stream) {
    std::vector<M4Launch> l;
    if (!plan_m4_frames(true, dst, unit_offset, nframe, nchan, fanout, ft,
                        nullptr, 0.f, 0, nframe * 20000ll * fanout, nullptr,
                        in, l, g_err4))
        return BB_ERR_ARGUMENT;
    if (in_dtype == BB_F32) run_enc<float>(l); else run_enc<double>(l);
    return 0;
}

int bb_mark4_decode_words(const void *words, int64_t nword, int32_t nchan,
                          int32_t fanout, int32_t ft, const float *levels_host,
                          float *out, void *stream) {
    std::vector<M4Launch> l;
    if (!plan_m4_words(false, words, nword, nchan, fanout, ft, levels_host,
                       out, nullptr, l, g_err4))
        return BB_ERR_ARGUMENT;
    run_dec(l);
    return 0;
}

int bb_mark4_encode_words(const void *in, int32_t in_dtype, void *words,
                          int64_t nword, int32_t nchan, int32_t fanout,
                          int32_t ft, void *stream) {
    std::vector<M4Launch> l;
    if (!plan_m4_words(true, words, nword, nchan, fanout, ft, nullptr,
                       nullptr, in, l, g_err4))
        return BB_ERR_ARGUMENT;
    if (in_dtype == BB_F32) run_enc<float>(l); else run_enc<double>(l);
    return 0;
}

}  // extern "C"
