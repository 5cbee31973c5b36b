#pragma once
#include <cstdint>
namespace wb200 {
// gelu[i] = f16(gelu_tanh_f32(f32(half(i)))), exp_tbl[i] = f16(expf(f32(half(i))))   (65 536 entries each)
void build_f16_tables(uint16_t * gelu, uint16_t * exp_tbl);
}
