// device_math.cuh -- small device helpers shared by the kernels.
#pragma once

#include <cstdint>

namespace pb
{
    // The weight of detector pixel (s, t), with exactly the reference's float operations (round-to-nearest
    // multiply/add/sqrt/divide, no FMA contraction), so the result is bit-identical to the OpenMP backend.
    __device__ __forceinline__ float pixel_weight(uint32_t s, uint32_t t, float h_min, float v_min, float d_sd,
                                                  float l_px_row, float l_px_col)
    {
        const float h_s = __fadd_rn(__fadd_rn(l_px_row / 2.f, __fmul_rn(static_cast<float>(s), l_px_row)), h_min);
        const float v_t = __fadd_rn(__fadd_rn(l_px_col / 2.f, __fmul_rn(static_cast<float>(t), l_px_col)), v_min);
        const float sum = __fadd_rn(__fadd_rn(__fmul_rn(d_sd, d_sd), __fmul_rn(h_s, h_s)), __fmul_rn(v_t, v_t));
        return __fdiv_rn(d_sd, __fsqrt_rn(sum));
    }

    __device__ __forceinline__ float2 cmul(float2 a, float2 b)
    {
        return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
    }

    // a * conj(b)
    __device__ __forceinline__ float2 cmul_conj(float2 a, float2 b)
    {
        return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
    }
}
