// isx_host_expand.cpp — host side of the host-buffer step (isx_step_host / isx_step_pinned): rebuilds the float32 obs rows
// [agents][127] in HOST memory from the compact record the device ships over PCIe.
//
// Why: an obs row is 508 bytes, but only 31 of its floats are free-form.  The lidar part is exactly
// float(4k) * (1/250) of the u8 hit index k the beam kernel stores anyway (k = 0: no hit -> 250 * (1/250)), and the columns
// behind the last beam are always 0 (IntersectionEnv.cpp:424,510-514).  So the device->host copy carries 32 floats (31
// features + an alive flag) and R bytes per agent — 200 B instead of 508 B at 72 beams — and the rows are completed here,
// bit-identical by construction: the same two IEEE operations (int -> float conversion, one multiplication by the same
// rounded constant) the kernel applies.  This is data-format work of the transport, not simulation: nothing here steps an env.
//
// The rows of a batch are contiguous, so a worker's slice is one dense float stream.  It is produced 8 rows (4064 bytes =
// 127 x 32 B) at a time in an L1-resident staging block and written with non-temporal 32-byte stores: no read-for-ownership
// of the 266 MB destination (65,536 envs x 8 agents) and no cache pollution.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <immintrin.h>

#include "isx_host_expand.h"

namespace isx {

namespace {

constexpr int OBS = 127, FEAT = 31, REC = 32;

struct Lut {
    float v[256];
    Lut() {
        const float inv = 1.0f / 250.0f;                       // the kernel's (1.0f / LIDAR_MAX_DIST)
        for (int k = 0; k < 256; ++k) v[k] = (float)(4 * k) * inv;
        v[0] = 250.0f * inv;                                   // no hit: LIDAR_MAX_DIST * (1 / LIDAR_MAX_DIST)
    }
};
const Lut g_lut;

// one row into `row` (127 floats, any alignment)
inline void row_scalar(const float* rec, const uint8_t* hits, int R, float* row) {
    if (rec[FEAT] != 0.0f) {
        memcpy(row, rec, sizeof(float) * FEAT);
        for (int i = 0; i < R; ++i) row[FEAT + i] = g_lut.v[hits[i]];
        for (int i = FEAT + R; i < OBS; ++i) row[i] = 0.0f;
    } else {
        memset(row, 0, sizeof(float) * OBS);                   // dead ego: all-zero row (IntersectionEnv.cpp:426-429)
    }
}

void rows_scalar(const float* rec, const uint8_t* hits, int R, float* dst, size_t n) {
    for (size_t a = 0; a < n; ++a) row_scalar(rec + a * REC, hits + a * (size_t)R, R, dst + a * OBS);
}

__attribute__((target("avx2"))) inline void row_avx2(const float* rec, const uint8_t* hits, int R, float* row) {
    if (rec[FEAT] == 0.0f) { memset(row, 0, sizeof(float) * OBS); return; }
    _mm256_storeu_ps(row, _mm256_loadu_ps(rec));
    _mm256_storeu_ps(row + 8, _mm256_loadu_ps(rec + 8));
    _mm256_storeu_ps(row + 16, _mm256_loadu_ps(rec + 16));
    _mm256_storeu_ps(row + 23, _mm256_loadu_ps(rec + 23));    // floats 23..30 (overlaps 23, stops before the flag)
    const __m256 inv = _mm256_set1_ps(1.0f / 250.0f);
    const __m256 none = _mm256_set1_ps(g_lut.v[0]);
    const __m256i zero = _mm256_setzero_si256();
    int i = 0;
    for (; i + 8 <= R; i += 8) {
        const __m256i k = _mm256_cvtepu8_epi32(_mm_loadl_epi64(reinterpret_cast<const __m128i*>(hits + i)));
        const __m256 f = _mm256_mul_ps(_mm256_cvtepi32_ps(_mm256_slli_epi32(k, 2)), inv);      // float(4k) * (1/250)
        const __m256 isz = _mm256_castsi256_ps(_mm256_cmpeq_epi32(k, zero));
        _mm256_storeu_ps(row + FEAT + i, _mm256_blendv_ps(f, none, isz));
    }
    for (; i < R; ++i) row[FEAT + i] = g_lut.v[hits[i]];
    for (i = FEAT + R; i < OBS; ++i) row[i] = 0.0f;
}

// n rows; when `stream` the destination is 32-byte aligned and n is a multiple of 8: staged blocks + non-temporal stores
__attribute__((target("avx2"))) void rows_avx2(const float* rec, const uint8_t* hits, int R, float* dst, size_t n, bool stream) {
    if (!stream) {
        for (size_t a = 0; a < n; ++a) row_avx2(rec + a * REC, hits + a * (size_t)R, R, dst + a * OBS);
        return;
    }
    alignas(32) float stage[8 * OBS];
    for (size_t a = 0; a < n; a += 8) {
        for (int j = 0; j < 8; ++j) row_avx2(rec + (a + j) * REC, hits + (a + j) * (size_t)R, R, stage + j * OBS);
        float* out = dst + a * OBS;
        for (int c = 0; c < OBS; ++c) _mm256_stream_ps(out + 8 * c, _mm256_load_ps(stage + 8 * c));
    }
    _mm_sfence();
}

bool have_avx2() {
    static const bool v = __builtin_cpu_supports("avx2");
    return v;
}
bool use_streaming_stores() {      // ISX_EXPAND_NT=0: plain cached stores (tuning aid; default on)
    static const bool v = [] { const char* e = getenv("ISX_EXPAND_NT"); return !(e && e[0] == '0'); }();
    return v;
}

}  // namespace

void expand_obs_rows(const float* rec, const uint8_t* hits, int R, float* dst, size_t n) {
    if (n == 0) return;
    if (!have_avx2()) { rows_scalar(rec, hits, R, dst, n); return; }
    if (!use_streaming_stores()) { rows_avx2(rec, hits, R, dst, n, false); return; }
    // head: rows until the destination is 32-byte aligned (a row is 508 B, so alignment recurs every 8 rows)
    size_t head = 0;
    while (head < n && (reinterpret_cast<uintptr_t>(dst + head * OBS) & 31u) != 0 && head < 8) ++head;
    if ((reinterpret_cast<uintptr_t>(dst + head * OBS) & 31u) != 0) { rows_avx2(rec, hits, R, dst, n, false); return; }   // never aligns (odd base)
    rows_avx2(rec, hits, R, dst, head, false);
    const size_t body = (n - head) & ~(size_t)7;
    rows_avx2(rec + head * REC, hits + head * (size_t)R, R, dst + head * OBS, body, true);
    const size_t done = head + body;
    rows_avx2(rec + done * REC, hits + done * (size_t)R, R, dst + done * OBS, n - done, false);
}

const float* expand_lidar_lut() { return g_lut.v; }

}  // namespace isx
