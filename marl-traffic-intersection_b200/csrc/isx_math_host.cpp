// isx_math_host.cpp — host build of isx_math.cuh behind a C ABI, used ONLY to prove on the CPU that
// the restated libm functions agree bit-for-bit with this machine's glibc (tests/test_math_host.py).
// Built as libisx_math_host.so with g++ -O2 -ffp-contract=off (no -march): see build.py.
#include "isx_math.cuh"

#include <atomic>
#include <cmath>
#include <thread>
#include <vector>

namespace {
inline bool same(float a, float b) {
    uint32_t x = isx::f2u(a), y = isx::f2u(b);
    if (x == y) return true;
    return (a != a) && (b != b);   // any NaN matches any NaN
}

template <class F>
uint64_t sweep(uint32_t lo, uint32_t hi, int nthreads, uint32_t* first_bad, F f) {
    std::atomic<uint64_t> bad{0};
    std::atomic<uint32_t> fb{0xffffffffu};
    std::vector<std::thread> th;
    const uint64_t total = (uint64_t)hi - lo + 1;
    for (int t = 0; t < nthreads; ++t) {
        th.emplace_back([&, t]() {
            uint64_t a = lo + total * t / nthreads, b = lo + total * (t + 1) / nthreads;
            uint64_t nb = 0;
            for (uint64_t u = a; u < b; ++u) {
                if (!f(isx::u2f((uint32_t)u))) {
                    ++nb;
                    uint32_t cur = fb.load();
                    while ((uint32_t)u < cur && !fb.compare_exchange_weak(cur, (uint32_t)u)) {}
                }
            }
            bad += nb;
        });
    }
    for (auto& x : th) x.join();
    if (first_bad) *first_bad = fb.load();
    return bad.load();
}

inline uint64_t splitmix(uint64_t& s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
}  // namespace

extern "C" {

void isxm_sincosf(const float* x, int n, float* s, float* c) { for (int i = 0; i < n; ++i) isx::sincosf_(x[i], s + i, c + i); }
void isxm_tanf(const float* x, int n, float* o) { for (int i = 0; i < n; ++i) o[i] = isx::tanf_(x[i]); }
void isxm_atan2f(const float* y, const float* x, int n, float* o) { for (int i = 0; i < n; ++i) o[i] = isx::atan2f_(y[i], x[i]); }
void isxm_hypotf(const float* y, const float* x, int n, float* o) { for (int i = 0; i < n; ++i) o[i] = isx::hypotf_(y[i], x[i]); }
void isxm_wrap(const float* x, int n, float* o) { for (int i = 0; i < n; ++i) o[i] = isx::wrap_angle(x[i]); }

// Sweeps over the float bit patterns [lo, hi] (inclusive); return the mismatch count vs libm.
uint64_t isxm_sweep_sincosf(uint32_t lo, uint32_t hi, int nthreads, uint32_t* first_bad) {
    return sweep(lo, hi, nthreads, first_bad, [](float x) {
        float s, c, s2, c2;
        isx::sincosf_(x, &s, &c);
        sincosf(x, &s2, &c2);
        return same(s, s2) && same(c, c2);
    });
}
// glibc's separate sinf/cosf entry points must agree with its sincosf (the compiler may emit either)
uint64_t isxm_sweep_sinf_cosf(uint32_t lo, uint32_t hi, int nthreads, uint32_t* first_bad) {
    return sweep(lo, hi, nthreads, first_bad, [](float x) {
        float s, c;
        isx::sincosf_(x, &s, &c);
        volatile float vx = x;
        return same(s, sinf(vx)) && same(c, cosf(vx));
    });
}
// fast exact fmod path vs libm fmodf, as used by wrap_angle (b = 2*pi_f) and with a second modulus
uint64_t isxm_sweep_fmod(uint32_t lo, uint32_t hi, int nthreads, uint32_t* first_bad) {
    return sweep(lo, hi, nthreads, first_bad, [](float x) {
        volatile float vx = x;
        const float t = isx::TWO_PI_F;
        return same(isx::fmodf_(x, t), fmodf(vx, t)) && same(isx::fmodf_(x, -t), fmodf(vx, -t)) && same(isx::fmodf_(x, 1.5f), fmodf(vx, 1.5f));
    });
}
uint64_t isxm_sweep_tanf(uint32_t lo, uint32_t hi, int nthreads, uint32_t* first_bad) {
    return sweep(lo, hi, nthreads, first_bad, [](float x) { return same(isx::tanf_(x), tanf(x)); });
}
uint64_t isxm_sweep_atanf(uint32_t lo, uint32_t hi, int nthreads, uint32_t* first_bad) {
    return sweep(lo, hi, nthreads, first_bad, [](float x) { return same(isx::atanf_(x), atanf(x)) && same(isx::atan2f_(x, 1.0f), atan2f(x, 1.0f)); });
}
// x paired with a fixed second operand (bit pattern `other`): atan2f(x, o), atan2f(o, x), hypotf(x, o)
uint64_t isxm_sweep_pair(uint32_t lo, uint32_t hi, uint32_t other, int nthreads, uint32_t* first_bad) {
    const float o = isx::u2f(other);
    return sweep(lo, hi, nthreads, first_bad, [o](float x) {
        return same(isx::atan2f_(x, o), atan2f(x, o)) && same(isx::atan2f_(o, x), atan2f(o, x)) &&
               same(isx::hypotf_(x, o), hypotf(x, o));
    });
}
// n random pairs; mode 0: arbitrary bit patterns, mode 1: "scene-like" magnitudes |v| < 1000
uint64_t isxm_random_pairs(uint64_t seed, uint64_t n, int mode, int nthreads, float* bad_y, float* bad_x) {
    std::atomic<uint64_t> bad{0};
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) {
        th.emplace_back([&, t]() {
            uint64_t s = seed * 1315423911ull + (uint64_t)t * 0x9E3779B97F4A7C15ull;
            uint64_t nb = 0;
            for (uint64_t i = t; i < n; i += nthreads) {
                uint64_t r = splitmix(s);
                float y, x;
                if (mode == 0) { y = isx::u2f((uint32_t)r); x = isx::u2f((uint32_t)(r >> 32)); }
                else {
                    y = ((float)(int32_t)(uint32_t)r) * (1000.0f / 2147483648.0f);
                    x = ((float)(int32_t)(uint32_t)(r >> 32)) * (1000.0f / 2147483648.0f);
                }
                bool ok = same(isx::atan2f_(y, x), atan2f(y, x)) && same(isx::hypotf_(y, x), hypotf(y, x));
                if (!ok) { if (nb == 0 && bad_y) { *bad_y = y; *bad_x = x; } ++nb; }
            }
            bad += nb;
        });
    }
    for (auto& x : th) x.join();
    return bad.load();
}

}  // extern "C"
