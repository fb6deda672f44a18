// api.cu -- library-level entry points of libpanib200.so (version, errors, planning helpers).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

#define PANIB_STR_(x) #x
#define PANIB_STR(x) PANIB_STR_(x)

namespace panib {

static thread_local char t_error[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_error, sizeof t_error, fmt, ap);
    va_end(ap);
}

}  // namespace panib

using namespace panib;

extern "C" const char *panib_version(void) {
    return PANIB_VERSION_STRING " sm_100a (CUDA " PANIB_STR(CUDART_VERSION) ")";
}

extern "C" int panib_last_error(char *buf, size_t n) {
    const size_t len = strlen(t_error);
    if (buf && n) {
        strncpy(buf, t_error, n - 1);
        buf[n - 1] = 0;
    }
    return (int)len;
}

extern "C" int panib_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" uint64_t panib_launch_count(void) { return g_launches.load(); }

// sourmash max_hash_for_scaled: 0 -> 0, 1 -> u64::MAX, else (u64::MAX as f64 / scaled as f64) as u64,
// where u64::MAX as f64 == 2^64.  Pinned by the "max_hash" values of the reference's fixture .sig files.
extern "C" uint64_t panib_max_hash(uint64_t scaled) {
    if (scaled == 0) return 0;
    if (scaled == 1) return UINT64_MAX;
    return (uint64_t)(18446744073709551616.0 / (double)scaled);
}

// Bucket plan of one genome: expected survivors = n_kmers / scaled; buckets are sized so that the
// expected load of a PANIB_BUCKET_SLOTS-slot bucket is <= 1/(2*slack) ... i.e. half full at slack = 1.
extern "C" int panib_plan_buckets(int64_t n_kmers, uint64_t scaled, double slack, int32_t *nb, uint64_t *bmul) {
    if (n_kmers < 0 || scaled == 0 || !(slack >= 1.0) || !nb || !bmul) {
        set_error("panib_plan_buckets: bad arguments");
        return PANIB_E_ARG;
    }
    const double expected = (double)n_kmers / (double)scaled;
    double want = expected * 2.0 * slack / (double)kBucketSlots;
    int64_t n = (int64_t)want + 1;
    if (n > (1 << 24)) n = 1 << 24;
    const unsigned __int128 universe = (unsigned __int128)panib_max_hash(scaled) + 1;
    if ((unsigned __int128)n > universe) n = (int64_t)universe;
    if (n < 1) n = 1;
    *nb = (int32_t)n;
    // bucket(h) = floor(h * bmul / 2^64) with bmul = floor(2^64 * n / universe) <= 2^64 * n / universe,
    // so bucket(h) <= h * n / universe < n for h < universe; monotone in h.
    unsigned __int128 m = (((unsigned __int128)n) << 64) / universe;
    if (m > (unsigned __int128)UINT64_MAX) m = UINT64_MAX;  // only when n == universe (degenerate)
    *bmul = (uint64_t)m;
    return PANIB_OK;
}

// (panib_fasta_to_stream and panib_format_u64, the host-side text routines, live in hostio.cpp)
