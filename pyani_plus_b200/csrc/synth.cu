// synth.cu -- synthetic genome family of BASELINE.json configs 2-5, generated on the
// device straight into the ASCII base-stream layout.  Integer-only, counter-based recipe; the CPU
// oracle (oracle/panib_oracle.c: oracle_synth_genome) states the same recipe independently and
// tests/test_gpu_parity.py checks the two byte for byte.
#include "common.cuh"
#include "kmer_hash.cuh"

namespace panib {

__host__ __device__ inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

// identity p_g = 0.80 + 0.19 u_g, u_g = 53 random bits / 2^53; returns (1 - p_g) * 2^53
__host__ __device__ inline uint64_t synth_threshold(uint64_t seed, uint64_t g) {
    const uint64_t u = splitmix64(splitmix64(seed ^ 0x5851F42D4C957F2DULL) + g) >> 11;
    const uint64_t two53 = 1ULL << 53;
    // floor(u * 19 / 100) without 128-bit arithmetic: u < 2^53, u*19 < 2^58
    return two53 / 5 - (u * 19ULL) / 100ULL;
}

__global__ void __launch_bounds__(256)
synth_ascii_kernel(uint64_t seed, int64_t g0, int64_t n_genomes, int64_t length, int64_t tiles_per_genome,
                   uint8_t *__restrict__ out) {
    const int64_t per_genome = tiles_per_genome * kTileBases;
    const int64_t total = n_genomes * per_genome;
    const uint64_t ka = splitmix64(seed + 0xFFFFFFFFFFFFULL);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int64_t gi = t / per_genome, pos = t % per_genome;
        uint8_t c = 'N';
        if (pos < length) {
            const uint64_t g = (uint64_t)(g0 + gi);
            const uint64_t thr = synth_threshold(seed, g);
            const uint64_t kg = splitmix64(seed + g);
            uint32_t code = (uint32_t)(splitmix64(ka ^ (uint64_t)pos) & 3u);
            const uint64_t r = splitmix64(kg ^ (uint64_t)pos);
            if ((r >> 11) < thr) code = (code + 1u + (uint32_t)((r & 0x7FFu) % 3u)) & 3u;
            c = (uint8_t)(0x54474341u >> (8 * code));
        }
        out[t] = c;
    }
}

}  // namespace panib

using namespace panib;

extern "C" int panib_synth_ascii(uint64_t seed, int64_t g0, int64_t n_genomes, int64_t length, uint8_t *d_ascii,
                                 void *stream) {
    if (n_genomes <= 0 || length < 0) return PANIB_OK;
    const int64_t tiles_per_genome = length / kTileBases + 1;
    synth_ascii_kernel<<<148 * 16, 256, 0, (cudaStream_t)stream>>>(seed, g0, n_genomes, length, tiles_per_genome,
                                                                   d_ascii);
    return check_launch("synth_ascii_kernel");
}
