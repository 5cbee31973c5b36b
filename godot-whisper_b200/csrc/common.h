// Shared host-side helpers: logging through the whisper_log_set callback, f16 <-> f32 conversion, wall clock.
#pragma once

#include "../../include/whisper_b200.h"

#include <chrono>
#include <cstdarg>
#include <cstdint>
#include <cstring>

namespace wb200 {

// Global log sink (whisper.cpp:865-871 keeps the same thing in g_state).
void log_set(ggml_log_callback cb, void * user_data);
void log_msg(ggml_log_level level, const char * fmt, ...) __attribute__((format(printf, 2, 3)));

#define WB_LOG_ERROR(...) ::wb200::log_msg(GGML_LOG_LEVEL_ERROR, __VA_ARGS__)
#define WB_LOG_WARN(...)  ::wb200::log_msg(GGML_LOG_LEVEL_WARN,  __VA_ARGS__)
#define WB_LOG_INFO(...)  ::wb200::log_msg(GGML_LOG_LEVEL_INFO,  __VA_ARGS__)

inline int64_t time_us() {
    using namespace std::chrono;
    return duration_cast<microseconds>(steady_clock::now().time_since_epoch()).count();
}

// IEEE binary16 <-> binary32, round-to-nearest-even — the same mapping F16C (_cvtss_sh / _cvtsh_ss) and
// CUDA's __float2half_rn implement, i.e. the conversion the reference applies before every mat-mul (ggml.c:9841-9857).
inline float f16_to_f32(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp  = (h >> 10) & 0x1Fu;
    uint32_t mant = h & 0x3FFu;
    uint32_t bits;
    if (exp == 0) {
        if (mant == 0) {
            bits = sign;
        } else {  // subnormal: normalise
            int e = -1;
            do { ++e; mant <<= 1; } while ((mant & 0x400u) == 0);
            mant &= 0x3FFu;
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | (mant << 13);
        }
    } else if (exp == 31) {
        bits = sign | 0x7F800000u | (mant << 13);
    } else {
        bits = sign | ((exp + 112u) << 23) | (mant << 13);
    }
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

inline uint16_t f32_to_f16(float f) {
    uint32_t x;
    memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    x &= 0x7FFFFFFFu;
    if (x >= 0x7F800000u) {  // inf / nan
        return (uint16_t)(sign | 0x7C00u | ((x > 0x7F800000u) ? (0x200u | ((x >> 13) & 0x3FFu)) : 0u));
    }
    if (x >= 0x477FF000u) {  // rounds to >= 65520 -> inf
        return (uint16_t)(sign | 0x7C00u);
    }
    if (x < 0x38800000u) {  // subnormal half or zero
        if (x < 0x33000000u) return (uint16_t) sign;  // < 2^-25 -> 0 (2^-25 exactly ties to even = 0)
        const int e = (int)(x >> 23);                 // biased exponent, 102..112
        uint32_t m = (x & 0x7FFFFFu) | 0x800000u;     // 24-bit significand
        const int shift = 126 - e;                    // 14..24
        const uint32_t half = m >> shift;
        const uint32_t rem  = m & ((1u << shift) - 1u);
        const uint32_t mid  = 1u << (shift - 1);
        uint32_t r = half;
        if (rem > mid || (rem == mid && (half & 1u))) r++;
        return (uint16_t)(sign | r);
    }
    // normal
    uint32_t e = (x >> 23) - 112u;
    uint32_t m = x & 0x7FFFFFu;
    uint32_t r = (e << 10) | (m >> 13);
    const uint32_t rem = m & 0x1FFFu;
    if (rem > 0x1000u || (rem == 0x1000u && (r & 1u))) r++;
    return (uint16_t)(sign | r);
}

inline float round_f16(float f) { return f16_to_f32(f32_to_f16(f)); }

}  // namespace wb200
