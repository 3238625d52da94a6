// int8 <-> float32 with a 2-D transpose through a shared-memory tile.
//
// GUPPI channels-first payloads are stored [chan][time][pol][re,im] and decode
// to (time, pol, chan) (baseband/guppi/payload.py:90-110); MKBF DADA heaps are
// stored [pol][chan][256 times][re,im] and decode to (time, pol, chan)
// (baseband/dada/payload.py:54-89).  Both are
//     in  unit: [nrow][ncol] items of IB bytes (1 real, 2 complex)
//     out     : [column][nrow] items -> IB floats each
// A CTA moves one TR x TC tile: rows are read with coalesced 4-byte loads
// into shared memory (row pitch padded by one word -> conflict-free column
// reads), then written so that consecutive lanes cover consecutive rows of
// one output column = contiguous floats in HBM.
//
// Written as two per-thread phases separated by a CTA barrier so the CPU
// emulation can run them as two loops.
#pragma once
#include "bb_common.cuh"
#include "bb_quant.cuh"

namespace bb {

constexpr int kI8Rows = 32;            // tile rows (channels)
constexpr int kI8RowBytes = 128;       // tile row width in bytes
constexpr int kI8Pitch = kI8RowBytes + 4;
constexpr int kI8Threads = 256;
constexpr int kI8SmemBytes = kI8Rows * kI8Pitch;

struct I8Geom {
    const uint8_t *src;              // decode: packed input; encode: output
    const long long *unit_offset;    // [nunit]
    const long long *col_begin, *col_end, *out_col0;   // decode only
    float *out;
    const void *in;                  // encode input (float or double)
    uint32_t nunit, nrow, ncol, ib;  // ib = item bytes (1 or 2)
    uint32_t tiles_r, tiles_c;       // tiles per unit
};

struct I8Tile { uint32_t unit, r0, j0; };

BB_HD I8Tile i8_tile(const I8Geom &p, uint32_t block) {
    I8Tile t;
    uint32_t per_unit = p.tiles_r * p.tiles_c;
    t.unit = block / per_unit;
    uint32_t rem = block - t.unit * per_unit;
    uint32_t tr = rem / p.tiles_c;
    t.r0 = tr * kI8Rows;
    t.j0 = (rem - tr * p.tiles_c) * (kI8RowBytes / p.ib);
    return t;
}

// Phase 1 of decode: packed rows -> shared tile.
BB_HD void i8_dec_load(const I8Geom &p, uint8_t *smem, uint32_t block,
                       uint32_t tid) {
    I8Tile t = i8_tile(p, block);
    const long long off = p.unit_offset[t.unit];
    if (off < 0) return;
    if ((long long)t.j0 >= p.col_end[t.unit]
        || (long long)t.j0 + kI8RowBytes / p.ib <= p.col_begin[t.unit]) return;
    const uint32_t lane = tid & 31u, warp = tid >> 5;
    const size_t rowbytes = (size_t)p.ncol * p.ib;
    const size_t b0 = (size_t)t.j0 * p.ib + 4 * lane;     // byte within row
    for (uint32_t r = warp; r < kI8Rows; r += kI8Threads / 32) {
        if (t.r0 + r >= p.nrow) break;
        const uint8_t *row = p.src + off + (size_t)(t.r0 + r) * rowbytes;
        uint32_t w = 0;
        if (b0 + 4 <= rowbytes
            && (reinterpret_cast<uintptr_t>(row + b0) & 3u) == 0) {
            w = *reinterpret_cast<const uint32_t *>(row + b0);
        } else {
            for (int k = 0; k < 4; ++k)
                if (b0 + k < rowbytes) w |= (uint32_t)row[b0 + k] << (8 * k);
        }
        *reinterpret_cast<uint32_t *>(smem + r * kI8Pitch + 4 * lane) = w;
    }
}

// Phase 2 of decode: shared tile -> transposed float output.
BB_HD void i8_dec_store(const I8Geom &p, const uint8_t *smem, uint32_t block,
                        uint32_t tid) {
    I8Tile t = i8_tile(p, block);
    const long long off = p.unit_offset[t.unit];
    if (off < 0) return;
    const uint32_t tc = kI8RowBytes / p.ib;               // tile columns
    const uint32_t tr = p.nrow - t.r0 < (uint32_t)kI8Rows ? p.nrow - t.r0
                                                          : kI8Rows;
    const long long cb = p.col_begin[t.unit], ce = p.col_end[t.unit];
    const long long oc0 = p.out_col0[t.unit];
    const uint32_t total = tc * tr;
    for (uint32_t idx = tid; idx < total; idx += kI8Threads) {
        uint32_t c = tr == kI8Rows ? idx >> 5 : idx / tr;
        uint32_t r = tr == kI8Rows ? idx & 31u : idx - c * tr;
        long long j = (long long)t.j0 + c;
        if (j < cb || j >= ce || j >= (long long)p.ncol) continue;
        size_t o = (size_t)(oc0 + j - cb) * p.nrow + t.r0 + r;
        const int8_t *s = reinterpret_cast<const int8_t *>(
            smem + r * kI8Pitch + c * p.ib);
        if (p.ib == 2) {
            *reinterpret_cast<F2 *>(p.out + 2 * o) = F2{(float)s[0], (float)s[1]};
        } else {
            p.out[o] = (float)s[0];
        }
    }
}

// Encode phase 1: float input (column-major items) -> int8 in the tile.
template <typename T>
BB_HD void i8_enc_load(const I8Geom &p, uint8_t *smem, uint32_t block,
                       uint32_t tid) {
    I8Tile t = i8_tile(p, block);
    if (p.unit_offset[t.unit] < 0) return;
    const uint32_t tc = kI8RowBytes / p.ib;
    const uint32_t tr = p.nrow - t.r0 < (uint32_t)kI8Rows ? p.nrow - t.r0
                                                          : kI8Rows;
    const T *in = reinterpret_cast<const T *>(p.in);
    const uint32_t total = tc * tr;
    for (uint32_t idx = tid; idx < total; idx += kI8Threads) {
        uint32_t c = tr == kI8Rows ? idx >> 5 : idx / tr;
        uint32_t r = tr == kI8Rows ? idx & 31u : idx - c * tr;
        size_t j = (size_t)t.j0 + c;
        if (j >= p.ncol) continue;
        size_t o = ((size_t)t.unit * p.ncol + j) * p.nrow + t.r0 + r;
        uint8_t *s = smem + r * kI8Pitch + c * p.ib;
        for (uint32_t k = 0; k < p.ib; ++k)
            s[k] = (uint8_t)quant_sint<T, 8>(in[o * p.ib + k]);
    }
}

// Encode phase 2: tile rows -> packed output.
BB_HD void i8_enc_store(const I8Geom &p, const uint8_t *smem, uint32_t block,
                        uint32_t tid) {
    I8Tile t = i8_tile(p, block);
    const long long off = p.unit_offset[t.unit];
    if (off < 0) return;
    const uint32_t lane = tid & 31u, warp = tid >> 5;
    const size_t rowbytes = (size_t)p.ncol * p.ib;
    const size_t b0 = (size_t)t.j0 * p.ib + 4 * lane;
    uint8_t *dst = const_cast<uint8_t *>(p.src);
    for (uint32_t r = warp; r < kI8Rows; r += kI8Threads / 32) {
        if (t.r0 + r >= p.nrow) break;
        uint8_t *row = dst + off + (size_t)(t.r0 + r) * rowbytes;
        uint32_t w = *reinterpret_cast<const uint32_t *>(
            smem + r * kI8Pitch + 4 * lane);
        if (b0 + 4 <= rowbytes
            && (reinterpret_cast<uintptr_t>(row + b0) & 3u) == 0) {
            *reinterpret_cast<uint32_t *>(row + b0) = w;
        } else {
            for (int k = 0; k < 4; ++k)
                if (b0 + k < rowbytes) row[b0 + k] = (uint8_t)(w >> (8 * k));
        }
    }
}

}  // namespace bb
