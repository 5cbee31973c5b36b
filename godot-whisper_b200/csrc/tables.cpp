// The two 65 536-entry f16 activation tables the reference evaluates GELU and the softmax exponential through
// (/root/reference/thirdparty/whisper.cpp/ggml.c:2218-2236 builds them, :1416-1423 and :11181-11183 use them).
// Built on the host with the same libm calls and uploaded to HBM once per context; device kernels index them with
// the f16 bit pattern of their argument, so GELU and exp are bit-identical to the reference for identical inputs.
#include "tables.h"
#include "common.h"

#include <cmath>

namespace wb200 {

void build_f16_tables(uint16_t * gelu, uint16_t * exp_tbl) {
    const float GELU_COEF_A    = 0.044715f;
    const float SQRT_2_OVER_PI = 0.79788456080286535587989211986876f;
    for (int i = 0; i < 65536; ++i) {
        const float f = f16_to_f32((uint16_t) i);
        // ggml.c:1404-1406.  The inner polynomial is written as the fused multiply-add gcc contracts it to when the
        // reference is built for a CPU with FMA (x86-64-v3 and up, i.e. oracle/_ref); without contraction exactly one
        // of the 65 536 entries (0xBFFF) differs by one f16 ulp.
        const float g = 0.5f * f * (1.0f + tanhf(SQRT_2_OVER_PI * f * fmaf(GELU_COEF_A * f, f, 1.0f)));
        gelu[i]    = f32_to_f16(g);
        exp_tbl[i] = f32_to_f16(expf(f));
    }
}

}  // namespace wb200

extern "C" WHISPER_B200_API void whisper_b200_f16_tables(uint16_t * gelu, uint16_t * exp_tbl) {
    wb200::build_f16_tables(gelu, exp_tbl);
}
