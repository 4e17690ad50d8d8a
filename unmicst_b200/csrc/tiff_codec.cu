// Host-only helpers of the TIFF page reader (unmicst_b200/tiffio.py): the LZW decoder that
// tifffile / imagecodecs provide to the reference's skio.imread / tifffile.imread calls
// (UnMicst1-5.py:794-797).  Deflate goes through Python's zlib; uncompressed data needs nothing.
#include <stdint.h>
#include <string.h>

#include "../../include/unmicst_b200.h"

// TIFF 6.0 LZW: MSB-first codes of 9..12 bits, ClearCode 256, EndOfInformation 257, code width grows one
// code early (at 511, 1023, 2047 entries).  Returns the number of bytes written, or a negative UMX_E* code.
extern "C" int64_t umx_tiff_lzw_decode(const uint8_t* src, int64_t src_bytes, uint8_t* dst, int64_t dst_capacity) {
    if (!src || !dst || src_bytes < 0 || dst_capacity < 0) return UMX_EINVAL;
    enum { kClear = 256, kEoi = 257, kFirst = 258, kMax = 4096 };
    static_assert(kMax == 1 << 12, "12-bit codes");
    uint16_t prefix[kMax];
    uint8_t suffix[kMax], first[kMax];
    uint16_t length[kMax];
    for (int i = 0; i < 256; ++i) { prefix[i] = 0; suffix[i] = (uint8_t)i; first[i] = (uint8_t)i; length[i] = 1; }
    int next = kFirst, width = 9, prev = -1;
    uint32_t bits = 0; int nbits = 0;
    int64_t ip = 0, op = 0;
    for (;;) {
        while (nbits < width) {
            if (ip >= src_bytes) return op;              // ran out of input without EOI: what was decoded stands
            bits = (bits << 8) | src[ip++]; nbits += 8;
        }
        const int code = (int)((bits >> (nbits - width)) & ((1u << width) - 1));
        nbits -= width;
        if (code == kEoi) return op;
        if (code == kClear) { next = kFirst; width = 9; prev = -1; continue; }
        if (prev < 0) {                                   // first code after a clear: a literal
            if (code > 255) return UMX_EINVAL;
            if (op >= dst_capacity) return op;
            dst[op++] = (uint8_t)code; prev = code;
            continue;
        }
        int len; uint8_t f;
        if (code < next) { len = length[code]; f = first[code]; }
        else if (code == next) { len = length[prev] + 1; f = first[prev]; }
        else return UMX_EINVAL;                           // corrupt stream
        // emit string(code) (or string(prev) + first(prev)) backwards
        int64_t end = op + len;
        const int64_t room = dst_capacity - op;
        if (code < next) {
            int c = code; int64_t q = end;
            while (q > op) { --q; if (q - op < room) dst[q] = suffix[c]; c = prefix[c]; }
        } else {
            if (len - 1 < room) dst[end - 1] = f;
            int c = prev; int64_t q = end - 1;
            while (q > op) { --q; if (q - op < room) dst[q] = suffix[c]; c = prefix[c]; }
        }
        if (next < kMax) {
            prefix[next] = (uint16_t)prev; suffix[next] = f; first[next] = first[prev]; length[next] = (uint16_t)(length[prev] + 1);
            ++next;
            if (next == (1 << width) - 1 && width < 12) ++width;      // "early change"
        }
        prev = code;
        op = end < dst_capacity ? end : dst_capacity;
        if (op >= dst_capacity) return op;
    }
}
