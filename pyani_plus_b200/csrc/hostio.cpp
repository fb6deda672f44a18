// hostio.cpp -- host-side text I/O of libpanib200.so for the drop-in path: FASTA text -> base-stream form, and
// the decimal text of a sketch for sourmash's signature JSON.  No CUDA in this file.
//
// Why native: at 1,000+ genomes the GPU needs tens of milliseconds for everything it does, and what the user
// waits for is the host walking 5 MB of text per genome twice (identity scan, then sketching) and printing
// 5,000 64-bit integers per genome.  Both routines release the GIL under ctypes, so
// the Python side runs them on a thread pool.
#include <emmintrin.h>  // SSE2: part of the x86-64 baseline, no run-time dispatch needed
#include <stdint.h>
#include <string.h>

#include "../../include/panib200.h"

namespace panib {
void set_error(const char *fmt, ...);
}
using panib::set_error;

// ------------------------------------------------------------------------------------------------
// FASTA text -> base-stream form.  Restates pyani_plus/utils.py:40-90 (fasta_bytes_iterator): a record
// starts at a line whose FIRST byte is '>', lines are split at '\n' only, anything before the first
// record is skipped, and the sequence is every byte of the following lines with space, tab, CR and LF
// removed.  Records are written back to back with ONE 'N' between them (the invalid separator of the
// base stream).
//
// A sequence line is 60-80 bytes of bases and, at most, a trailing CR, so the line is tested 16 bytes at
// a time for the three characters to drop and copied whole when there is none; only lines that do hold
// one take the byte loop.
// ------------------------------------------------------------------------------------------------
namespace {

struct FastaScan {
    int64_t n_records = 0, total_bases = 0, title_off = -1, title_len = 0, written = 0;
};

inline bool is_dropped(uint8_t c) { return c == ' ' || c == '\t' || c == '\r'; }

// number of bytes of [p, p + len) that are space, tab or CR
inline int64_t count_dropped(const uint8_t *p, int64_t len) {
    const __m128i sp = _mm_set1_epi8(' '), tb = _mm_set1_epi8('\t'), cr = _mm_set1_epi8('\r');
    int64_t i = 0, hits = 0;
    for (; i + 16 <= len; i += 16) {
        const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i *>(p + i));
        const __m128i m = _mm_or_si128(_mm_or_si128(_mm_cmpeq_epi8(v, sp), _mm_cmpeq_epi8(v, tb)), _mm_cmpeq_epi8(v, cr));
        hits += __builtin_popcount((unsigned)_mm_movemask_epi8(m));
    }
    for (; i < len; i++) hits += is_dropped(p[i]) ? 1 : 0;
    return hits;
}

FastaScan fasta_walk(const uint8_t *text, int64_t n, uint8_t *dst, int64_t cap, bool *overflow) {
    FastaScan r;
    int64_t pos = 0;
    bool in_record = false;
    while (pos < n) {
        const uint8_t *nl = static_cast<const uint8_t *>(memchr(text + pos, '\n', (size_t)(n - pos)));
        const int64_t end = nl ? (int64_t)(nl - text) : n;  // line = [pos, end), newline excluded
        if (end > pos && text[pos] == '>') {
            if (r.n_records == 0) {  // title of the first record, right-stripped like bytes.rstrip()
                int64_t te = end;
                while (te > pos + 1 && (text[te - 1] == ' ' || text[te - 1] == '\t' || text[te - 1] == '\r' ||
                                        text[te - 1] == '\n' || text[te - 1] == '\v' || text[te - 1] == '\f'))
                    te--;
                r.title_off = pos + 1;
                r.title_len = te - (pos + 1);
            }
            if (r.n_records > 0) {  // separator between records
                if (dst) { if (r.written < cap) dst[r.written] = 'N'; else *overflow = true; }
                r.written++;
            }
            r.n_records++;
            in_record = true;
        } else if (in_record && end > pos) {
            const uint8_t *line = text + pos;
            const int64_t len = end - pos;
            const int64_t kept = len - count_dropped(line, len);
            if (dst) {
                if (r.written + kept > cap) {
                    *overflow = true;
                } else if (kept == len) {
                    memcpy(dst + r.written, line, (size_t)len);
                } else {
                    uint8_t *out = dst + r.written;
                    for (int64_t i = 0; i < len; i++)
                        if (!is_dropped(line[i])) *out++ = line[i];
                }
            }
            r.written += kept;
            r.total_bases += kept;
        }
        pos = end + 1;
    }
    return r;
}

}  // namespace

// dst == NULL only measures.  Returns the stream-form length (bases + separators), or PANIB_E_ARG if dst is
// too small (n bytes are always enough).  out4[0..3] = n_records, total_bases, title_off, title_len.
extern "C" int64_t panib_fasta_to_stream(const uint8_t *text, int64_t n, uint8_t *dst, int64_t dst_cap,
                                         int64_t *out4) {
    if (!text || n < 0) {
        set_error("panib_fasta_to_stream: bad arguments");
        return PANIB_E_ARG;
    }
    bool overflow = false;
    const FastaScan r = fasta_walk(text, n, dst, dst_cap, &overflow);
    if (overflow) {
        set_error("panib_fasta_to_stream: destination too small (%lld needed)", (long long)r.written);
        return PANIB_E_ARG;
    }
    if (out4) {
        out4[0] = r.n_records;
        out4[1] = r.total_bases;
        out4[2] = r.title_off;
        out4[3] = r.title_len;
    }
    return r.written;
}

// ------------------------------------------------------------------------------------------------
// Decimal text of n unsigned 64-bit integers, `sep` between them (sep == 0: nothing between them).
// The "mins" array of a sourmash signature is the first form, the input of its md5sum the second
// (reference format: tests/fixtures/*/intermediates/sourmash/*.sig; pinned by
// tests/snakemake/test_sourmash_workflow.py:43-67).  Two digits per division, written backwards into a
// 20-byte scratch.  Returns the number of bytes written, or PANIB_E_ARG when dst_cap is too small
// (21 * n bytes are always enough).
// ------------------------------------------------------------------------------------------------
extern "C" int64_t panib_format_u64(const uint64_t *values, int64_t n, int sep, char *dst, int64_t dst_cap) {
    static const char pairs[201] =
        "00010203040506070809101112131415161718192021222324252627282930313233343536373839"
        "40414243444546474849505152535455565758596061626364656667686970717273747576777879"
        "8081828384858687888990919293949596979899";
    if (n < 0 || (n > 0 && (!values || !dst))) {
        set_error("panib_format_u64: bad arguments");
        return PANIB_E_ARG;
    }
    int64_t w = 0;
    char buf[20];
    for (int64_t i = 0; i < n; i++) {
        uint64_t v = values[i];
        int p = 20;
        while (v >= 100) {
            const unsigned d = (unsigned)(v % 100);
            v /= 100;
            p -= 2;
            memcpy(buf + p, pairs + 2 * d, 2);
        }
        if (v >= 10) {
            p -= 2;
            memcpy(buf + p, pairs + 2 * v, 2);
        } else {
            buf[--p] = (char)('0' + v);
        }
        const int len = 20 - p;
        if (w + len + 1 > dst_cap) {
            set_error("panib_format_u64: destination too small");
            return PANIB_E_ARG;
        }
        if (sep && i) dst[w++] = (char)sep;
        memcpy(dst + w, buf + p, (size_t)len);
        w += len;
    }
    return w;
}
