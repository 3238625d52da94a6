// Frame location and frame indexing on the device (sm_100a): the replacement
// for the reference's byte-wise sync search (`locate_frames`,
// baseband/base/base.py:181-335), its `RawOffsets` bookkeeping
// (baseband/base/offsets.py:6-126) and the frame-by-frame repair of streams
// with missing / duplicated / re-ordered frames (baseband/vdif/base.py:536-755,
// baseband/base/base.py:1083-1219).  A chunk of the file is searched for every
// place the (masked) sync pattern of the format occurs with another one a
// frame later; the headers found are turned into (frame index, thread slot)
// and their byte offsets scattered into the stream's frame table -- first
// occurrence wins, frames flagged invalid end up as missing.  Integer only.
#include "bb_runtime.cuh"

namespace bb {

constexpr int kIndexBlock = 256;
constexpr int kMaxPattern = 256;          // bytes (Mark 4, 64 tracks: 256)

struct Pattern {
    uint8_t pat[kMaxPattern];
    uint8_t mask[kMaxPattern];
    int n;                                // bytes used
};

__device__ __forceinline__ bool match_at(const uint8_t *src, long long nbytes,
                                         long long at, const Pattern &p) {
    if (at < 0 || at + p.n > nbytes) return false;
    for (int i = 0; i < p.n; ++i)
        if ((src[at + i] ^ p.pat[i]) & p.mask[i]) return false;
    return true;
}

// One thread per byte position.  `first` = index of the first pattern byte
// with a non-zero mask: the cheap rejection test.
__global__ void __launch_bounds__(kIndexBlock)
k_locate_frames(const uint8_t *src, long long nbytes, long long own_stop,
                const Pattern p, int first, long long pattern_offset,
                long long frame_nbytes, int check, int at_eof,
                long long base, long long *locations, int max_out,
                int *count, long long *unverified, int max_unverified,
                int *count_unverified) {
    const long long loc = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (loc >= own_stop) return;
    const long long at = loc + pattern_offset;
    if (at + p.n > nbytes) return;
    if ((src[at + first] ^ p.pat[first]) & p.mask[first]) return;
    if (!match_at(src, nbytes, at, p)) return;
    // the frame must fit (in the file when the region ends it, else in what
    // the next chunk will see as well)
    if (frame_nbytes > 0) {
        if (at_eof && loc + frame_nbytes > nbytes) return;
        if (check) {
            const long long c = at + (long long)check * frame_nbytes;
            // a check position inside the data must hold the pattern too
            if (c >= 0 && c + p.n <= nbytes) {
                if (!match_at(src, nbytes, c, p)) {
                    // a header that is not followed by another one: kept
                    // aside, the caller may still want it (the last frame
                    // before trailing damage)
                    if (unverified) {
                        const int u = atomicAdd(count_unverified, 1);
                        if (u < max_unverified) unverified[u] = base + loc;
                    }
                    return;
                }
            } else if (!at_eof && c + p.n > nbytes) {
                return;                   // cannot be verified in this chunk
            }
        }
    }
    const int slot = atomicAdd(count, 1);
    if (slot < max_out) locations[slot] = base + loc;
}

__device__ __forceinline__ uint32_t ldw_any(const uint8_t *p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16)
        | ((uint32_t)p[3] << 24);
}

// Table entries while it is being built: 2 * offset + invalid, so that an
// atomicMin keeps the first occurrence in the file together with its flag.
constexpr unsigned long long kEmpty = ~0ull;

// VDIF: (seconds, frame_nr, thread_id) -> (set index, slot)
// (`_get_index`, baseband/vdif/base.py:386-390).
__global__ void __launch_bounds__(kIndexBlock)
k_vdif_index(const uint8_t *src, long long base, const long long *locations,
             const int *count, int max_loc, const int *thread_slot,
             int nthread, int seconds0, int frame_nr0, int fps, int thread0,
             long long nset_max, unsigned long long *table, int *stats) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = *count < max_loc ? *count : max_loc;
    if (i >= n) return;
    const long long off = locations[i];
    const uint8_t *h = src + (off - base);
    const uint32_t w0 = ldw_any(h), w1 = ldw_any(h + 4), w3 = ldw_any(h + 12);
    const long long seconds = w0 & 0x3fffffffu;
    const unsigned invalid = w0 >> 31;
    const long long frame_nr = w1 & 0xffffffu;
    const int tid = (w3 >> 16) & 0x3ff;
    const long long index = (seconds - seconds0) * fps + frame_nr - frame_nr0;
    const int slot = thread_slot[tid];
    if (slot < 0 || slot >= nthread) return;        // thread not selected
    if (index < 0 || index >= nset_max) {
        atomicAdd(stats + 1, 1);                    // outside the table
        return;
    }
    atomicMin(table + index * nthread + slot,
              2ull * (unsigned long long)off + invalid);
    const int capped = (int)(index < 0x7fffffff ? index : 0x7fffffff);
    atomicMax(stats, capped);
    // the stream ends with the last good frame of the first header's thread
    // (`_last_header`, baseband/vdif/base.py:493-519)
    if (tid == thread0) atomicMax(stats + 3, capped);
}

__device__ __forceinline__ int bcd_digits(uint32_t v, int ndigit) {
    int out = 0, scale = 1;
    for (int d = 0; d < ndigit; ++d, scale *= 10) {
        const int nib = (v >> (4 * d)) & 0xf;
        if (nib > 9) return -1;
        out += nib * scale;
    }
    return out;
}

// Mark 5B: (jday, seconds, frame_nr) -> frame index
// (baseband/mark5b/base.py:206-213; jday wraps every 1000 days).
__global__ void __launch_bounds__(kIndexBlock)
k_mark5b_index(const uint8_t *src, long long base, const long long *locations,
               const int *count, int max_loc, int jday0, int seconds0,
               int frame_nr0, int fps, long long nset_max,
               unsigned long long *table, int *stats) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = *count < max_loc ? *count : max_loc;
    if (i >= n) return;
    const long long off = locations[i];
    const uint8_t *h = src + (off - base);
    const uint32_t w1 = ldw_any(h + 4), w2 = ldw_any(h + 8);
    const int jday = bcd_digits(w2 >> 20, 3), seconds = bcd_digits(w2, 5);
    if (jday < 0 || seconds < 0) {
        atomicAdd(stats + 2, 1);                    // invalid BCD time code
        return;
    }
    const long long dday = (((long long)jday - jday0 + 1500) % 1000) - 500;
    const long long index = ((long long)seconds - seconds0 + 86400ll * dday)
        * fps + (long long)(w1 & 0x7fffu) - frame_nr0;
    if (index < 0 || index >= nset_max) {
        atomicAdd(stats + 1, 1);
        return;
    }
    atomicMin(table + index, 2ull * (unsigned long long)off);
    atomicMax(stats, (int)(index < 0x7fffffff ? index : 0x7fffffff));
}

// Mark 4: one warp per header found.  The time code of one track is gathered
// bit by bit (a header word of a track is spread over 32 steps of the
// ntrack-bit stream words, baseband/mark4/header.py:47-63): step
// 32 * w + lane, bit `track`, ballot, bit reversal.  Unit year, day of year,
// h, m, s and ms (baseband/mark4/header.py:223-262; the ms digit stands for
// quarters: + (ms % 5) / 4 ms) give the time in 0.25 ms ticks relative to the
// first header, which must be a whole number of frame periods.
__global__ void __launch_bounds__(kIndexBlock)
k_mark4_index(const uint8_t *src, long long base, const long long *locations,
              const int *count, int max_loc, int wordbytes, int track,
              int year0, int yday0, int days_year0, int days_prev_year,
              long long tick0, long long tick_step, int check_crc,
              long long nset_max, unsigned long long *table, int *stats) {
    const int lane = threadIdx.x & 31;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n = *count < max_loc ? *count : max_loc;
    if (i >= n) return;                               // warp-uniform
    const long long off = locations[i];
    const uint8_t *h = src + (off - base);
    uint32_t w5[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const uint8_t b = h[(size_t)(32 * k + lane) * wordbytes
                            + (track >> 3)];
        w5[k] = __brev(__ballot_sync(0xffffffffu, (b >> (track & 7)) & 1));
    }
    if (lane) return;
    if (check_crc) {
        // CRC-12 (x^12 + x^11 + x^3 + x^2 + x + 1, baseband/mark4/header.py:
        // 25-27) over the 160 header bits of the track, first step first: a
        // header that ends in its CRC leaves no remainder.  The all-ones sync
        // also matches a few bytes either side of the true frame start when
        // the neighbouring header bits are ones; those candidates fail here.
        uint32_t rem = 0u;
#pragma unroll 1
        for (int k = 0; k < 5; ++k)
            for (int b = 31; b >= 0; --b) {
                const uint32_t top = (rem >> 11) & 1u;
                rem = ((rem << 1) | ((w5[k] >> b) & 1u)) & 0xfffu;
                if (top) rem ^= 0x80fu;
            }
        if (rem) {
            atomicAdd(stats + 2, 1);
            return;
        }
    }
    const uint32_t w[2] = {w5[3], w5[4]};
    const int y = bcd_digits(w[0] >> 28, 1), doy = bcd_digits(w[0] >> 16, 3),
        hh = bcd_digits(w[0] >> 8, 2), mm = bcd_digits(w[0], 2),
        ss = bcd_digits(w[1] >> 24, 2), ms = bcd_digits(w[1] >> 12, 3);
    if (y < 0 || doy < 1 || doy > 366 || hh < 0 || hh > 23 || mm < 0
        || mm > 59 || ss < 0 || ss > 60 || ms < 0 || ms % 5 == 4) {
        atomicAdd(stats + 2, 1);                      // not a time code
        return;
    }
    const int dy = ((y - year0 % 10 + 15) % 10) - 5;
    long long ddays;
    if (dy == 0) ddays = doy - yday0;
    else if (dy == 1) ddays = doy + days_year0 - yday0;
    else if (dy == -1) ddays = doy - days_prev_year - yday0;
    else { atomicAdd(stats + 1, 1); return; }
    const long long ticks = ddays * (86400ll * 4000ll)
        + ((long long)hh * 3600 + mm * 60 + ss) * 4000ll + ms * 4 + ms % 5
        - tick0;
    if (ticks % tick_step) {
        atomicAdd(stats + 2, 1);                      // off the frame grid
        return;
    }
    const long long index = ticks / tick_step;
    if (index < 0 || index >= nset_max) {
        atomicAdd(stats + 1, 1);
        return;
    }
    atomicMin(table + index, 2ull * (unsigned long long)off);
    atomicMax(stats, (int)(index < 0x7fffffff ? index : 0x7fffffff));
}

__global__ void __launch_bounds__(kIndexBlock)
k_index_fill(unsigned long long *table, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) table[i] = kEmpty;
}

// Encoded entries -> byte offsets of the frames, -1 where no (valid) frame.
__global__ void __launch_bounds__(kIndexBlock)
k_index_finalize(const unsigned long long *table, long long n,
                 long long *offsets) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long v = table[i];
    offsets[i] = (v == kEmpty || (v & 1ull)) ? -1 : (long long)(v >> 1);
}

static unsigned blocks_for(long long n) {
    return (unsigned)((n + kIndexBlock - 1) / kIndexBlock);
}

}  // namespace bb

using namespace bb;

extern "C" int bb_locate_frames(
    const void *src, int64_t nbytes, int64_t own_stop, const uint8_t *pattern,
    const uint8_t *mask, int32_t pattern_nbytes, int64_t pattern_offset,
    int64_t frame_nbytes, int32_t check, int32_t at_eof, int64_t base,
    int64_t *locations, int32_t max_locations, int32_t *count,
    int64_t *unverified, int32_t max_unverified, int32_t *count_unverified,
    void *stream) {
    if (!src || !pattern || !locations || !count
        || (unverified && !count_unverified))
        return set_error(BB_ERR_ARGUMENT, "null pointer");
    if (pattern_nbytes < 1 || pattern_nbytes > kMaxPattern)
        return set_error(BB_ERR_ARGUMENT, "pattern must be 1..%d bytes",
                         kMaxPattern);
    if (nbytes < 0 || own_stop < 0 || frame_nbytes < 0 || max_locations < 0
        || pattern_offset < 0)
        return set_error(BB_ERR_ARGUMENT, "negative size");
    Pattern p;
    p.n = pattern_nbytes;
    int first = -1;
    for (int i = 0; i < pattern_nbytes; ++i) {
        p.pat[i] = pattern[i];
        p.mask[i] = mask ? mask[i] : 0xff;
        if (first < 0 && p.mask[i]) first = i;
    }
    if (first < 0) return set_error(BB_ERR_ARGUMENT, "mask is all zero");
    if (own_stop > nbytes) own_stop = nbytes;
    if (own_stop == 0) return BB_OK;
    k_locate_frames<<<blocks_for(own_stop), kIndexBlock, 0,
                      as_stream(stream)>>>(
        (const uint8_t *)src, nbytes, own_stop, p, first, pattern_offset,
        frame_nbytes, check, at_eof, base, (long long *)locations,
        max_locations, count, (long long *)unverified, max_unverified,
        count_unverified);
    BB_CHECK_LAUNCH("bb_locate_frames");
    return BB_OK;
}

extern "C" int bb_index_table_init(uint64_t *table, int64_t nentry,
                                   void *stream) {
    if (!table || nentry < 0) return set_error(BB_ERR_ARGUMENT, "bad table");
    if (nentry == 0) return BB_OK;
    k_index_fill<<<blocks_for(nentry), kIndexBlock, 0, as_stream(stream)>>>(
        (unsigned long long *)table, nentry);
    BB_CHECK_LAUNCH("bb_index_table_init");
    return BB_OK;
}

extern "C" int bb_vdif_index(
    const void *src, int64_t base, const int64_t *locations,
    const int32_t *count, int32_t max_locations, const int32_t *thread_slot,
    int32_t nthread, int32_t seconds0, int32_t frame_nr0,
    int32_t frames_per_second, int32_t thread0, int64_t nset_max,
    uint64_t *table, int32_t *stats, void *stream) {
    if (!src || !locations || !count || !thread_slot || !table || !stats)
        return set_error(BB_ERR_ARGUMENT, "null pointer");
    if (nthread < 1 || frames_per_second < 1 || nset_max < 0)
        return set_error(BB_ERR_ARGUMENT, "bad geometry");
    if (max_locations == 0) return BB_OK;
    k_vdif_index<<<blocks_for(max_locations), kIndexBlock, 0,
                   as_stream(stream)>>>(
        (const uint8_t *)src, base, (const long long *)locations, count,
        max_locations, thread_slot, nthread, seconds0, frame_nr0,
        frames_per_second, thread0, nset_max, (unsigned long long *)table,
        stats);
    BB_CHECK_LAUNCH("bb_vdif_index");
    return BB_OK;
}

extern "C" int bb_mark5b_index(
    const void *src, int64_t base, const int64_t *locations,
    const int32_t *count, int32_t max_locations, int32_t jday0,
    int32_t seconds0, int32_t frame_nr0, int32_t frames_per_second,
    int64_t nset_max, uint64_t *table, int32_t *stats, void *stream) {
    if (!src || !locations || !count || !table || !stats)
        return set_error(BB_ERR_ARGUMENT, "null pointer");
    if (frames_per_second < 1 || nset_max < 0)
        return set_error(BB_ERR_ARGUMENT, "bad geometry");
    if (max_locations == 0) return BB_OK;
    k_mark5b_index<<<blocks_for(max_locations), kIndexBlock, 0,
                     as_stream(stream)>>>(
        (const uint8_t *)src, base, (const long long *)locations, count,
        max_locations, jday0, seconds0, frame_nr0, frames_per_second,
        nset_max, (unsigned long long *)table, stats);
    BB_CHECK_LAUNCH("bb_mark5b_index");
    return BB_OK;
}

extern "C" int bb_mark4_index(
    const void *src, int64_t base, const int64_t *locations,
    const int32_t *count, int32_t max_locations, int32_t ntrack,
    int32_t track, int32_t year0, int32_t yday0, int32_t days_year0,
    int32_t days_prev_year, int64_t tick0, int64_t tick_step,
    int32_t check_crc, int64_t nset_max, uint64_t *table, int32_t *stats,
    void *stream) {
    if (!src || !locations || !count || !table || !stats)
        return set_error(BB_ERR_ARGUMENT, "null pointer");
    if ((ntrack != 16 && ntrack != 32 && ntrack != 64) || track < 0
        || track >= ntrack || tick_step < 1 || nset_max < 0)
        return set_error(BB_ERR_ARGUMENT, "bad geometry");
    if (max_locations == 0) return BB_OK;
    k_mark4_index<<<blocks_for((long long)max_locations * 32), kIndexBlock, 0,
                    as_stream(stream)>>>(
        (const uint8_t *)src, base, (const long long *)locations, count,
        max_locations, ntrack / 8, track, year0, yday0, days_year0,
        days_prev_year, tick0, tick_step, check_crc, nset_max,
        (unsigned long long *)table, stats);
    BB_CHECK_LAUNCH("bb_mark4_index");
    return BB_OK;
}

extern "C" int bb_index_table_finish(const uint64_t *table, int64_t nentry,
                                     int64_t *offsets, void *stream) {
    if (!table || !offsets || nentry < 0)
        return set_error(BB_ERR_ARGUMENT, "bad table");
    if (nentry == 0) return BB_OK;
    k_index_finalize<<<blocks_for(nentry), kIndexBlock, 0,
                       as_stream(stream)>>>(
        (const unsigned long long *)table, nentry, (long long *)offsets);
    BB_CHECK_LAUNCH("bb_index_table_finish");
    return BB_OK;
}
