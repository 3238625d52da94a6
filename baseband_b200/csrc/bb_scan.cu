// Batched header parse / validity kernels (sm_100a): the producers of the
// unit-offset tables the decode kernels consume.  Tiny traffic (headers
// only, plus a short-circuited payload check for Mark 5B), integer only.
#include "bb_runtime.cuh"

namespace bb {

constexpr int kScanBlock = 256;
constexpr long long kMissing = -2;     // slot never written by the scan

__device__ __forceinline__ uint32_t ldw(const uint8_t *p) {
    return *reinterpret_cast<const uint32_t *>(p);
}

__device__ __forceinline__ uint32_t bits(uint32_t w, int lo, int n) {
    return (w >> lo) & ((n >= 32) ? 0xffffffffu : ((1u << n) - 1u));
}

__global__ void k_fill_i64(long long *p, long long v, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ---------------------------------------------------------------- VDIF
// Field table: baseband/vdif/header.py:529-542 (+ edv :557-559).
__global__ void __launch_bounds__(kScanBlock)
k_vdif_scan(const uint8_t *src, const long long *frame_offset,
            long long frame_stride, long long nframe, int header_nbytes,
            int frames_per_set, int nthread, const int *thread_slot,
            int *fields, long long *unit_offset, int *n_bad,
            long long index0, int seconds0, int frame_nr0, int fps) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nframe) return;
    long long off = frame_offset ? frame_offset[i] : i * frame_stride;
    const uint8_t *h = src + off;
    uint32_t w[8];
#pragma unroll
    for (int k = 0; k < 8; ++k)
        w[k] = (k < 4 || header_nbytes >= 32) ? ldw(h + 4 * k) : 0u;
    const uint32_t invalid = bits(w[0], 31, 1), seconds = bits(w[0], 0, 30);
    const uint32_t frame_nr = bits(w[1], 0, 24), tid = bits(w[3], 16, 10);
    if (fields) {
        int *f = fields + i;
        f[BB_VDIF_INVALID * nframe] = invalid;
        f[BB_VDIF_LEGACY * nframe] = bits(w[0], 30, 1);
        f[BB_VDIF_SECONDS * nframe] = seconds;
        f[BB_VDIF_REF_EPOCH * nframe] = bits(w[1], 24, 6);
        f[BB_VDIF_FRAME_NR * nframe] = frame_nr;
        f[BB_VDIF_VERSION * nframe] = bits(w[2], 29, 3);
        f[BB_VDIF_LG2_NCHAN * nframe] = bits(w[2], 24, 5);
        f[BB_VDIF_FRAME_LENGTH * nframe] = bits(w[2], 0, 24);
        f[BB_VDIF_COMPLEX * nframe] = bits(w[3], 31, 1);
        f[BB_VDIF_BITS_PER_SAMPLE * nframe] = bits(w[3], 26, 5);
        f[BB_VDIF_THREAD_ID * nframe] = tid;
        f[BB_VDIF_STATION_ID * nframe] = bits(w[3], 0, 16);
        f[BB_VDIF_EDV * nframe] = bits(w[4], 24, 8);
        f[BB_VDIF_WORD4 * nframe] = (int)w[4];
        f[BB_VDIF_WORD5 * nframe] = (int)w[5];
        f[BB_VDIF_WORD6 * nframe] = (int)w[6];
        f[BB_VDIF_WORD7 * nframe] = (int)w[7];
    }
    if (!unit_offset) return;
    long long set = i / frames_per_set;
    if ((set + 1) * (long long)frames_per_set > nframe) return;  // partial set
    // All frames of a set share frame_nr with its first frame; `seconds` is
    // deliberately not compared: some recorders get it wrong in part of the
    // threads (baseband/vdif/frame.py:207-216, sample_vlbi.vdif).
    long long i0 = set * frames_per_set;
    if (i != i0) {
        long long off0 = frame_offset ? frame_offset[i0] : i0 * frame_stride;
        uint32_t b = ldw(src + off0 + 4);
        if (bits(b, 0, 24) != frame_nr) atomicAdd(n_bad, 1);
    } else if (fps > 0) {
        // the set must carry the frame index its position implies
        // (baseband/vdif/base.py:386-390, on the first frame of the set)
        const long long index = ((long long)seconds - seconds0) * fps
            + (long long)frame_nr - frame_nr0;
        if (index != index0 + set) atomicAdd(n_bad, 1);
    }
    int slot = thread_slot[tid];
    if (slot < 0 || slot >= nthread) return;      // thread not selected
    long long value = invalid ? -1 : off + header_nbytes;
    unsigned long long prev = atomicExch(
        reinterpret_cast<unsigned long long *>(unit_offset + set * nthread
                                               + slot),
        (unsigned long long)value);
    if ((long long)prev != kMissing) atomicAdd(n_bad, 1);   // duplicate id
}

__global__ void k_count_missing(long long *unit_offset, long long n,
                                int *n_bad) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && unit_offset[i] == kMissing) {
        unit_offset[i] = -1;
        atomicAdd(n_bad, 1);
    }
}

// ---------------------------------------------------------------- Mark 5B
__device__ __forceinline__ int bcd(uint32_t v, int ndigit) {
    int out = 0, scale = 1;
    for (int d = 0; d < ndigit; ++d) {
        uint32_t nib = (v >> (4 * d)) & 0xfu;
        if (nib > 9) return -1;                   // base/utils.py:27-31
        out += nib * scale;
        scale *= 10;
    }
    return out;
}

// One lane per frame for the header; frames whose first three payload words
// equal the fill pattern (baseband/mark5b/frame.py:62-72 short-circuit) are
// then checked in full by the whole warp, coalesced.
__global__ void __launch_bounds__(kScanBlock)
k_mark5b_scan(const uint8_t *src, const long long *frame_offset,
              long long frame_stride, long long nframe, int *fields,
              long long *unit_offset, int *n_bad, long long index0,
              int jday0, int seconds0, int frame_nr0, int fps) {
    const uint32_t kFill = 0x11223344u;
    const int lane = threadIdx.x & 31;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool live = i < nframe;
    long long off = 0;
    bool candidate = false;
    uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
    if (live) {
        off = frame_offset ? frame_offset[i] : i * frame_stride;
        if (off < 0) {
            // a frame the stream's index knows to be absent
            if (unit_offset) unit_offset[i] = -1;
            if (fields)
                for (int k = 0; k < BB_M5B_NFIELD; ++k)
                    fields[k * nframe + i] = 0;
            live = false;
        }
    }
    if (live) {
        const uint8_t *h = src + off;
        w0 = ldw(h); w1 = ldw(h + 4); w2 = ldw(h + 8); w3 = ldw(h + 12);
        candidate = ldw(h + 16) == kFill && ldw(h + 20) == kFill
            && ldw(h + 24) == kFill;
    }
    bool valid = !candidate;
    unsigned todo = __ballot_sync(0xffffffffu, candidate);
    while (todo) {
        int b = __ffs(todo) - 1;
        todo &= todo - 1;
        long long boff = __shfl_sync(0xffffffffu, off, b);
        const uint8_t *pl = src + boff + 16;
        // no early exit: the loads stay independent, so a batch of them is
        // in flight at once (a true fill frame costs ~10 latencies, not 78)
        uint32_t acc = 0u;
#pragma unroll 8
        for (int k = 3 + lane; k < 2500; k += 32)
            acc |= ldw(pl + 4 * k) ^ kFill;
        bool any = __any_sync(0xffffffffu, acc != 0u);
        if (lane == b) valid = any;
    }
    if (!live) return;
    const int bj = bits(w2, 20, 12), bs = bits(w2, 0, 20),
        bf = bits(w3, 16, 16);
    if (n_bad && fps > 0) {
        // frame index from the header time must equal the position
        // (baseband/mark5b/base.py:206-213; jday wraps every 1000 days)
        const int jday = bcd(bj, 3), seconds = bcd(bs, 5);
        const long long dday = (((long long)jday - jday0 + 1500) % 1000) - 500;
        const long long index = ((long long)seconds - seconds0
                                 + 86400ll * dday) * fps
            + (long long)bits(w1, 0, 15) - frame_nr0;
        if (w0 != 0xABADDEEDu || jday < 0 || seconds < 0
            || index != index0 + i)
            atomicAdd(n_bad, 1);
    }
    if (fields) {
        int *f = fields + i;
        f[BB_M5B_SYNC * nframe] = (int)w0;
        f[BB_M5B_USER * nframe] = bits(w1, 16, 16);
        f[BB_M5B_INTERNAL_TVG * nframe] = bits(w1, 15, 1);
        f[BB_M5B_FRAME_NR * nframe] = bits(w1, 0, 15);
        f[BB_M5B_BCD_JDAY * nframe] = bj;
        f[BB_M5B_BCD_SECONDS * nframe] = bs;
        f[BB_M5B_BCD_FRACTION * nframe] = bf;
        f[BB_M5B_CRC * nframe] = bits(w3, 0, 16);
        f[BB_M5B_JDAY * nframe] = bcd(bj, 3);
        f[BB_M5B_SECONDS * nframe] = bcd(bs, 5);
        int frac = bcd(bf, 4);
        // "unrounded" to the 156250 ns grid: mark5b/header.py:223-225
        f[BB_M5B_FRACTION_NS * nframe] = frac < 0 ? -1
            : 156250 * ((frac * 100000 + 156249) / 156250);
        f[BB_M5B_VALID * nframe] = valid ? 1 : 0;
    }
    if (unit_offset) unit_offset[i] = valid ? off + 16 : -1;
}

// ---------------------------------------------------------------- Mark 4
// One warp per frame.  Header = 160 steps of an ntrack-bit word; word w, bit
// b of a track's header is step 32*w + 31 - b (baseband/mark4/header.py:47-63)
// so a warp ballot over 32 consecutive steps, bit-reversed, is one header
// word.  Error flags are bits 15..12 of word 1 = steps 48..51; the frame is
// valid iff no track has one set (baseband/mark4/frame.py:78-87).
// BCD of v < 1000 (three digits).
__device__ __forceinline__ uint32_t bcd3(uint32_t v) {
    return ((v / 100u) << 8) | (((v / 10u) % 10u) << 4) | (v % 10u);
}

// Time-code words 3 and 4 (CRC bits zero) of a Mark 4 header at `ticks`
// quarter milliseconds after 00:00 of MJD mjd0
// (baseband/mark4/header.py:223-262, :509-533): unit year, day of year,
// hour, minute | second, millisecond.
__device__ void mark4_time_words(int mjd0, long long ticks, uint32_t &w3,
                                 uint32_t &w4) {
    const long long kDay = 86400ll * 4000ll;
    const long long day = ticks / kDay;
    const uint32_t tick = (uint32_t)(ticks - day * kDay);
    // civil date from the day count (days since 1970-01-01 = MJD - 40587)
    const long long z = (long long)mjd0 + day - 40587 + 719468;
    const long long era = (z >= 0 ? z : z - 146096) / 146097;
    const uint32_t doe = (uint32_t)(z - era * 146097);
    const uint32_t yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
    const uint32_t doy = doe - (365 * yoe + yoe / 4 - yoe / 100);  // from 1 Mar
    long long year = (long long)yoe + era * 400;
    uint32_t yday;
    if (doy >= 306) {                    // January, February of the next year
        year += 1;
        yday = doy - 306 + 1;
    } else {
        const bool leap = (year % 4 == 0 && year % 100 != 0) || year % 400 == 0;
        yday = doy + 59 + (leap ? 1 : 0) + 1;
    }
    const uint32_t sec = tick / 4000u, ms = (tick % 4000u) / 4u;
    const uint32_t hour = sec / 3600u, minute = sec / 60u % 60u,
        second = sec % 60u;
    w3 = ((uint32_t)(year % 10) << 28) | (bcd3(yday) << 16)
        | (bcd3(hour) << 8) | bcd3(minute);
    w4 = (bcd3(second) << 24) | (bcd3(ms) << 12);
}

template <typename W>
__global__ void __launch_bounds__(kScanBlock)
k_mark4_scan(const uint8_t *src, const long long *frame_offset,
             long long frame_stride, long long nframe, int track,
             uint32_t *words5, long long *unit_offset, int *n_bad,
             long long index0, int mjd0, long long tick0,
             long long tick_step) {
    const int lane = threadIdx.x & 31;
    long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= nframe) return;                     // warp-uniform
    long long off = frame_offset ? frame_offset[i] : i * frame_stride;
    if (off < 0) {                 // absent according to the stream's index
        if (lane == 0) {
            if (unit_offset) unit_offset[i] = -1;
            if (words5)
                for (int w = 0; w < 5; ++w) words5[i * 5 + w] = 0u;
        }
        return;                                   // warp-uniform
    }
    const W *st = reinterpret_cast<const W *>(src + off);
    bool bad = false;
    uint32_t got3 = 0u, got4 = 0u;
#pragma unroll
    for (int w = 0; w < 5; ++w) {
        W v = st[32 * w + lane];
        unsigned m = __brev(__ballot_sync(0xffffffffu, (v >> track) & 1));
        if (words5 && lane == 0) words5[i * 5 + w] = m;
        if (w == 3) got3 = m;
        if (w == 4) got4 = m;
        if (w == 1 && lane >= 16 && lane < 20 && v != 0) bad = true;
    }
    bad = __any_sync(0xffffffffu, bad);
    if (n_bad && tick_step > 0 && lane == 0) {
        // the time code of the chosen track must be the one the position of
        // the frame implies (what the writer would generate; CRC bits aside)
        uint32_t w3, w4;
        mark4_time_words(mjd0, tick0 + tick_step * (index0 + i), w3, w4);
        if (got3 != w3 || (got4 & 0xfffff000u) != w4) atomicAdd(n_bad, 1);
    }
    if (unit_offset && lane == 0)
        unit_offset[i] = bad ? -1 : off + (long long)sizeof(W) * 160;
}


// ---------------------------------------------------------------- writers
// Frame assembly for the stream writers: one warp per frame copies the
// frame's header (built on the host, a few bytes per frame) to the head of
// the frame, writes the unit-offset entries the encode kernel takes, and --
// where the format marks invalid frames by a payload pattern (Mark 5B,
// baseband/mark5b/frame.py:126-133) -- fills the payload of those frames and
// takes them out of the encode (-1).
__global__ void __launch_bounds__(kScanBlock)
k_frames_assemble(uint8_t *dst, long long nframe, long long frame_stride,
                  int header_nbytes, const uint8_t *headers,
                  const uint8_t *valid, uint32_t fill_word,
                  long long payload_nbytes, int units_per_frame,
                  long long unit_stride, long long *unit_offset) {
    const int lane = threadIdx.x & 31;
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= nframe) return;                     // warp-uniform
    uint8_t *frame = dst + i * frame_stride;
    const uint8_t *h = headers + i * (long long)header_nbytes;
    if ((header_nbytes & 3) == 0
        && ((reinterpret_cast<uintptr_t>(frame)
             | reinterpret_cast<uintptr_t>(h)) & 3) == 0) {
        const uint32_t *hs = reinterpret_cast<const uint32_t *>(h);
        uint32_t *hd = reinterpret_cast<uint32_t *>(frame);
        for (int k = lane; k < header_nbytes / 4; k += 32) hd[k] = hs[k];
    } else {
        for (int k = lane; k < header_nbytes; k += 32) frame[k] = h[k];
    }
    const bool ok = !valid || valid[i];
    if (unit_offset)
        for (int u = lane; u < units_per_frame; u += 32)
            unit_offset[i * units_per_frame + u] = ok
                ? i * frame_stride + header_nbytes + u * unit_stride : -1;
    if (!ok) {
        uint32_t *pl = reinterpret_cast<uint32_t *>(frame + header_nbytes);
        for (long long k = lane; k < payload_nbytes / 4; k += 32)
            pl[k] = fill_word;
    }
}

}  // namespace bb

using namespace bb;

extern "C" int bb_vdif_scan(
    const void *src, const int64_t *frame_offset, int64_t frame_stride,
    int64_t nframe, int32_t header_nbytes, int32_t frames_per_set,
    int32_t nthread, const int32_t *thread_slot, int32_t *fields,
    int64_t *unit_offset, int32_t *n_inconsistent, int64_t index0,
    int32_t seconds0, int32_t frame_nr0, int32_t frames_per_second,
    void *stream) {
    if (!src) return set_error(BB_ERR_ARGUMENT, "null src");
    if (header_nbytes != 16 && header_nbytes != 32)
        return set_error(BB_ERR_ARGUMENT, "header_nbytes must be 16 or 32");
    if (unit_offset && (!thread_slot || !n_inconsistent || frames_per_set < 1
                        || nthread < 1))
        return set_error(BB_ERR_ARGUMENT,
                         "thread_slot, n_inconsistent, frames_per_set and "
                         "nthread are needed to build unit offsets");
    if (!aligned(src, 4) || (frame_stride & 3))
        return set_error(BB_ERR_ALIGNMENT, "frames must be 4-byte aligned");
    if (nframe <= 0) return BB_OK;
    cudaStream_t s = as_stream(stream);
    long long nunit = unit_offset ? (nframe / frames_per_set) * nthread : 0;
    if (nunit) {
        k_fill_i64<<<(unsigned)((nunit + 255) / 256), 256, 0, s>>>(
            (long long *)unit_offset, kMissing, nunit);
        BB_CHECK_LAUNCH("bb_vdif_scan fill");
    }
    k_vdif_scan<<<(unsigned)((nframe + kScanBlock - 1) / kScanBlock),
                  kScanBlock, 0, s>>>(
        (const uint8_t *)src, (const long long *)frame_offset, frame_stride,
        nframe, header_nbytes, frames_per_set, nthread, thread_slot, fields,
        (long long *)unit_offset, n_inconsistent, index0, seconds0, frame_nr0,
        frames_per_second);
    BB_CHECK_LAUNCH("bb_vdif_scan");
    if (nunit) {
        k_count_missing<<<(unsigned)((nunit + 255) / 256), 256, 0, s>>>(
            (long long *)unit_offset, nunit, n_inconsistent);
        BB_CHECK_LAUNCH("bb_vdif_scan count");
    }
    return BB_OK;
}

extern "C" int bb_mark5b_scan(
    const void *src, const int64_t *frame_offset, int64_t frame_stride,
    int64_t nframe, int32_t *fields, int64_t *unit_offset,
    int32_t *n_inconsistent, int64_t index0, int32_t jday0, int32_t seconds0,
    int32_t frame_nr0, int32_t frames_per_second, void *stream) {
    if (!src) return set_error(BB_ERR_ARGUMENT, "null src");
    if (!aligned(src, 4) || (frame_stride & 3))
        return set_error(BB_ERR_ALIGNMENT, "frames must be 4-byte aligned");
    if (nframe <= 0) return BB_OK;
    k_mark5b_scan<<<(unsigned)((nframe + kScanBlock - 1) / kScanBlock),
                    kScanBlock, 0, as_stream(stream)>>>(
        (const uint8_t *)src, (const long long *)frame_offset, frame_stride,
        nframe, fields, (long long *)unit_offset, n_inconsistent, index0,
        jday0, seconds0, frame_nr0, frames_per_second);
    BB_CHECK_LAUNCH("bb_mark5b_scan");
    return BB_OK;
}

extern "C" int bb_mark4_scan(
    const void *src, const int64_t *frame_offset, int64_t frame_stride,
    int64_t nframe, int32_t ntrack, int32_t track, uint32_t *words5,
    int64_t *unit_offset, int32_t *n_inconsistent, int64_t index0,
    int32_t mjd0, int64_t tick0, int64_t tick_step, void *stream) {
    if (!src) return set_error(BB_ERR_ARGUMENT, "null src");
    if (ntrack != 16 && ntrack != 32 && ntrack != 64)
        return set_error(BB_ERR_UNSUPPORTED, "ntrack must be 16, 32 or 64");
    if (track < 0 || track >= ntrack)
        return set_error(BB_ERR_ARGUMENT, "track out of range");
    if (!aligned(src, ntrack / 8) || (frame_stride % (ntrack / 8)))
        return set_error(BB_ERR_ALIGNMENT, "frames must be word aligned");
    if (nframe <= 0) return BB_OK;
    unsigned grid = (unsigned)((nframe * 32 + kScanBlock - 1) / kScanBlock);
    cudaStream_t s = as_stream(stream);
    const uint8_t *p = (const uint8_t *)src;
    const long long *fo = (const long long *)frame_offset;
    if (ntrack == 64)
        k_mark4_scan<unsigned long long><<<grid, kScanBlock, 0, s>>>(
            p, fo, frame_stride, nframe, track, words5, (long long *)unit_offset,
            n_inconsistent, index0, mjd0, tick0, tick_step);
    else if (ntrack == 32)
        k_mark4_scan<uint32_t><<<grid, kScanBlock, 0, s>>>(
            p, fo, frame_stride, nframe, track, words5, (long long *)unit_offset,
            n_inconsistent, index0, mjd0, tick0, tick_step);
    else
        k_mark4_scan<uint16_t><<<grid, kScanBlock, 0, s>>>(
            p, fo, frame_stride, nframe, track, words5, (long long *)unit_offset,
            n_inconsistent, index0, mjd0, tick0, tick_step);
    BB_CHECK_LAUNCH("bb_mark4_scan");
    return BB_OK;
}

extern "C" int bb_frames_assemble(
    void *dst, int64_t nframe, int64_t frame_stride, int32_t header_nbytes,
    const void *headers, const uint8_t *valid, uint32_t fill_word,
    int64_t payload_nbytes, int32_t units_per_frame, int64_t unit_stride,
    int64_t *unit_offset, void *stream) {
    if (!dst || (!headers && header_nbytes > 0))
        return set_error(BB_ERR_ARGUMENT, "null dst or headers");
    if (nframe < 0 || header_nbytes < 0 || units_per_frame < 1
        || frame_stride < header_nbytes)
        return set_error(BB_ERR_ARGUMENT, "bad frame geometry");
    if (valid && (!aligned(dst, 4) || (frame_stride & 3) || (header_nbytes & 3)
                  || (payload_nbytes & 3)))
        return set_error(BB_ERR_ALIGNMENT,
                         "payload fill needs 4-byte aligned payloads");
    if (nframe == 0) return BB_OK;
    const long long nthr = (long long)nframe * 32;
    k_frames_assemble<<<(unsigned)((nthr + kScanBlock - 1) / kScanBlock),
                        kScanBlock, 0, as_stream(stream)>>>(
        (uint8_t *)dst, nframe, frame_stride, header_nbytes,
        (const uint8_t *)headers, valid, fill_word, payload_nbytes,
        units_per_frame, unit_stride, (long long *)unit_offset);
    BB_CHECK_LAUNCH("bb_frames_assemble");
    return BB_OK;
}
