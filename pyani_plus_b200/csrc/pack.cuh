// pack.cuh -- ASCII base -> 2-bit code + validity, shared by the CUDA pack kernel and the host emulation.
// sourmash upper-cases the sequence and skips every k-mer window holding a byte outside ACGT
// (SURVEY.md 8c conventions 2-3; pinned by the reference's tests/test_coverage.py:163-174).
#pragma once
#include <stdint.h>

#include "kmer_hash.cuh"

namespace panib {

// code: A=0 C=1 G=2 T=3 (order-preserving, so integer compares of packed k-mers are lexicographic);
// returns true when the byte is a valid base.
PANIB_HD bool base_code(uint8_t c, uint32_t &code) {
    const uint32_t u = c & 0xDFu;  // clear bit 5: a-z -> A-Z (only letters can land on A/C/G/T)
    code = ((u >> 1) & 3u) ^ ((u >> 2) & 1u);
    return u == 'A' || u == 'C' || u == 'G' || u == 'T';
}

// 16 ASCII bytes (as four little-endian words) -> packed word; invalid bits returned in low 16 bits.
PANIB_HD uint32_t pack16(const uint32_t *w4, uint32_t &invalid16) {
    uint32_t packed = 0, inv = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const uint8_t c = (uint8_t)(w4[i >> 2] >> (8 * (i & 3)));
        uint32_t code;
        const bool ok = base_code(c, code);
        packed |= (ok ? code : 0u) << (2 * i);
        inv |= (ok ? 0u : 1u) << i;
    }
    invalid16 = inv;
    return packed;
}

}  // namespace panib
