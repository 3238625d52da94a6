// CPU emulation of the int8 transpose kernels — TEST INFRASTRUCTURE ONLY.
// The two per-thread phases of each CTA are run as two loops.
#include <vector>
#include "../../baseband_b200/csrc/bb_int8.cuh"
#include "../../include/baseband_b200.h"

using namespace bb;

static bool geom(I8Geom &g, int64_t nunit, int64_t nrow, int64_t ncol, int ib,
                 uint64_t &nblocks, bool fast = false) {
    if (ib != 1 && ib != 2) return false;
    g.nunit = (uint32_t)nunit; g.nrow = (uint32_t)nrow; g.ncol = (uint32_t)ncol;
    g.ib = ib;
    const int rows = fast ? kF8Rows : kI8Rows;
    g.tiles_r = (uint32_t)((nrow + rows - 1) / rows);
    uint32_t tc = (fast ? kF8Words * 4 : kI8RowBytes) / ib;
    g.tiles_c = (uint32_t)((ncol + tc - 1) / tc);
    g.group = 16 < g.tiles_c ? 16 : g.tiles_c;
    if (g.group < 1) g.group = 1;
    nblocks = (uint64_t)nunit * g.tiles_r * g.tiles_c;
    return true;
}

extern "C" {

int bb_decode_int8_transposed(const void *src, const int64_t *unit_offset,
                              int64_t nunit, int64_t nrow, int64_t ncol,
                              int32_t item_nbytes, const int64_t *col_begin,
                              const int64_t *col_end, const int64_t *out_col0,
                              float *out, void *stream) {
    I8Geom g;
    uint64_t nblocks;
    const bool fast = nrow % 2 == 0
        && (reinterpret_cast<uintptr_t>(out) & 15u) == 0;
    if (!geom(g, nunit, nrow, ncol, item_nbytes, nblocks, fast))
        return BB_ERR_ARGUMENT;
    g.src = (const uint8_t *)src;
    g.unit_offset = (const long long *)unit_offset;
    g.col_begin = (const long long *)col_begin;
    g.col_end = (const long long *)col_end;
    g.out_col0 = (const long long *)out_col0;
    g.out = out; g.in = nullptr;
    if (fast) {
        std::vector<uint32_t> tile64(kF8SmemWords);
        for (uint64_t b = 0; b < nblocks; ++b) {
            for (uint32_t t = 0; t < kF8Threads; ++t)
                f8_dec_load(g, tile64.data(), (uint32_t)b, t);
            for (uint32_t t = 0; t < kF8Threads; ++t)
                f8_dec_store(g, tile64.data(), (uint32_t)b, t);
        }
        return 0;
    }
    alignas(16) uint8_t tile[kI8SmemBytes];
    for (uint64_t b = 0; b < nblocks; ++b) {
        for (uint32_t t = 0; t < kI8Threads; ++t) i8_dec_load(g, tile, (uint32_t)b, t);
        for (uint32_t t = 0; t < kI8Threads; ++t) i8_dec_store(g, tile, (uint32_t)b, t);
    }
    return 0;
}

int bb_encode_int8_transposed(const void *in, int32_t in_dtype, void *dst,
                              const int64_t *unit_offset, int64_t nunit,
                              int64_t nrow, int64_t ncol, int32_t item_nbytes,
                              void *stream) {
    I8Geom g;
    uint64_t nblocks;
    const bool fast = nrow % 2 == 0
        && (reinterpret_cast<uintptr_t>(in) & 15u) == 0;
    if (!geom(g, nunit, nrow, ncol, item_nbytes, nblocks, fast))
        return BB_ERR_ARGUMENT;
    g.src = (const uint8_t *)dst;
    g.unit_offset = (const long long *)unit_offset;
    g.col_begin = g.col_end = g.out_col0 = nullptr;
    g.out = nullptr; g.in = in;
    if (fast) {
        std::vector<uint32_t> tile64(kF8SmemWords);
        for (uint64_t b = 0; b < nblocks; ++b) {
            for (uint32_t t = 0; t < kF8Threads; ++t) {
                if (in_dtype == BB_F32)
                    f8_enc_load<float>(g, tile64.data(), (uint32_t)b, t);
                else
                    f8_enc_load<double>(g, tile64.data(), (uint32_t)b, t);
            }
            for (uint32_t t = 0; t < kF8Threads; ++t)
                f8_enc_store(g, tile64.data(), (uint32_t)b, t);
        }
        return 0;
    }
    alignas(16) uint8_t tile[kI8SmemBytes];
    for (uint64_t b = 0; b < nblocks; ++b) {
        for (uint32_t t = 0; t < kI8Threads; ++t) {
            if (in_dtype == BB_F32) i8_enc_load<float>(g, tile, (uint32_t)b, t);
            else i8_enc_load<double>(g, tile, (uint32_t)b, t);
        }
        for (uint32_t t = 0; t < kI8Threads; ++t) i8_enc_store(g, tile, (uint32_t)b, t);
    }
    return 0;
}

int bb_decode_int8_timefirst(const void *src, const int64_t *unit_offset,
                             int64_t nunit, int64_t nsample, int32_t nchan,
                             int32_t npol, int32_t item_nbytes,
                             const int64_t *t_begin, const int64_t *t_end,
                             const int64_t *out_t0, float *out, void *stream) {
    TFGeom g;
    if (tf_fill_geom(g, nunit, nsample, nchan, npol, item_nbytes,
                     (reinterpret_cast<uintptr_t>(out) & 15u) == 0))
        return BB_ERR_ARGUMENT;
    g.src = (const uint8_t *)src;
    g.unit_offset = (const long long *)unit_offset;
    g.t_begin = (const long long *)t_begin;
    g.t_end = (const long long *)t_end;
    g.out_t0 = (const long long *)out_t0;
    g.out = out; g.in = nullptr;
    for (uint32_t u = 0; u < g.nunit; ++u)
        for (uint32_t i = 0; i < g.items_per_unit; ++i) tf_decode(g, u, i);
    return 0;
}

int bb_encode_int8_timefirst(const void *in, int32_t in_dtype, void *dst,
                             const int64_t *unit_offset, int64_t nunit,
                             int64_t nsample, int32_t nchan, int32_t npol,
                             int32_t item_nbytes, void *stream) {
    TFGeom g;
    if (tf_fill_geom(g, nunit, nsample, nchan, npol, item_nbytes,
                     (reinterpret_cast<uintptr_t>(in) & 15u) == 0))
        return BB_ERR_ARGUMENT;
    g.src = (const uint8_t *)dst;
    g.unit_offset = (const long long *)unit_offset;
    g.t_begin = g.t_end = g.out_t0 = nullptr;
    g.out = nullptr; g.in = in;
    for (uint32_t u = 0; u < g.nunit; ++u)
        for (uint32_t i = 0; i < g.items_per_unit; ++i) {
            if (in_dtype == BB_F32) tf_encode<float>(g, u, i);
            else tf_encode<double>(g, u, i);
        }
    return 0;
}

}  // extern "C"
