// shenc.cu -- real spherical-harmonics direction encoding, degree <= 4 (adjacent row f-1 of SURVEY.md section 8).
//
// Replaces shencoder/src/shencoder.cu:27-123 (kernel_sh) and :358-382 (kernel_sh_backward) for the degrees
// LAENeRF uses (4 -> 16 channels in the NeRF colour net, 3 -> 9 in the style offset net).  One thread per sample;
// a warp reads 384 contiguous bytes of directions and writes degree^2 contiguous values per sample through
// 16-byte vectors.  Output dtype fp32 (the reference's, custom_fwd(cast_inputs=float32)) or fp16 so that the result
// can feed the colour net without a cast pass.
#include "common.cuh"
#include "sh_core.cuh"

namespace lnrf {

template <typename T>
__device__ __forceinline__ void store_row(T* out, const float* v, uint32_t n);
template <>
__device__ __forceinline__ void store_row<float>(float* out, const float* v, uint32_t n) {
    if (n == 16) {
#pragma unroll
        for (int i = 0; i < 4; i++) __stcs(reinterpret_cast<float4*>(out) + i, make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
    } else {
        for (uint32_t i = 0; i < n; i++) out[i] = v[i];
    }
}
template <>
__device__ __forceinline__ void store_row<__half>(__half* out, const float* v, uint32_t n) {
    if (n == 16) {
#pragma unroll
        for (int i = 0; i < 2; i++) {
            union { uint4 u; __half2 h[4]; } p;
#pragma unroll
            for (int j = 0; j < 4; j++) p.h[j] = __floats2half2_rn(v[8 * i + 2 * j], v[8 * i + 2 * j + 1]);
            __stcs(reinterpret_cast<uint4*>(out) + i, p.u);
        }
    } else {
        for (uint32_t i = 0; i < n; i++) out[i] = __float2half_rn(v[i]);
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
k_sh_forward(const float* __restrict__ inputs, T* __restrict__ outputs, uint32_t B, uint32_t degree, T* __restrict__ dy_dx) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint32_t C2 = degree * degree;
    const float x = inputs[(size_t)b * 3], y = inputs[(size_t)b * 3 + 1], z = inputs[(size_t)b * 3 + 2];
    float v[16];
    sh_basis(x, y, z, degree, v);
    store_row<T>(outputs + (size_t)b * C2, v, C2);
    if (dy_dx) {  // layout [B, 3, C2] (shencoder.cu:125-128)
        float dx[16], dy[16], dz[16];
        sh_jacobian(x, y, z, degree, dx, dy, dz);
        T* d = dy_dx + (size_t)b * 3 * C2;
        store_row<T>(d, dx, C2);
        store_row<T>(d + C2, dy, C2);
        store_row<T>(d + 2 * C2, dz, C2);
    }
}

// shencoder.cu:358-382: grad_inputs[b, d] += sum_c grad[b, c] * dy_dx[b, d, c]
__global__ void __launch_bounds__(256)
k_sh_backward(const float* __restrict__ grad, uint32_t B, uint32_t degree, const float* __restrict__ dy_dx,
              float* __restrict__ grad_inputs) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * 3) return;
    const uint32_t b = t / 3, d = t - b * 3, C2 = degree * degree;
    const float* g = grad + (size_t)b * C2;
    const float* j = dy_dx + ((size_t)b * 3 + d) * C2;
    float acc = grad_inputs[t];
    for (uint32_t c = 0; c < C2; c++) acc += g[c] * j[c];
    grad_inputs[t] = acc;
}

}  // namespace lnrf

using namespace lnrf;

extern "C" {

int lnrf_sh_encode_forward(const float* inputs, void* outputs, uint32_t B, uint32_t degree, void* dy_dx, lnrf_dtype out_dtype,
                           lnrf_stream_t stream) {
    LNRF_REQUIRE(degree >= 1 && degree <= 8, "SH encoder only supports degree in [1, 8]");  // sphere_harmonics.py:70
    if (degree > 4) {
        set_error("sh_encode_forward: degree %u not built (this library implements degrees 1..4, the ones LAENeRF uses)", degree);
        return LNRF_ERR_UNSUPPORTED;
    }
    if (B == 0) return LNRF_OK;
    LNRF_REQUIRE(inputs && outputs, "sh_encode_forward: null pointer");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (out_dtype == LNRF_F32)
        k_sh_forward<float><<<div_up(B, 256u), 256, 0, st>>>(inputs, (float*)outputs, B, degree, (float*)dy_dx);
    else
        k_sh_forward<__half><<<div_up(B, 256u), 256, 0, st>>>(inputs, (__half*)outputs, B, degree, (__half*)dy_dx);
    LNRF_LAUNCH_CHECK("sh_encode_forward");
    return LNRF_OK;
}

int lnrf_sh_encode_backward(const float* grad, uint32_t B, uint32_t degree, const float* dy_dx, float* grad_inputs,
                            lnrf_stream_t stream) {
    LNRF_REQUIRE(degree >= 1 && degree <= 4, "sh_encode_backward: degree %u not built (1..4)", degree);
    if (B == 0) return LNRF_OK;
    LNRF_REQUIRE(grad && dy_dx && grad_inputs, "sh_encode_backward: null pointer");
    k_sh_backward<<<div_up(B * 3u, 256u), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(grad, B, degree, dy_dx, grad_inputs);
    LNRF_LAUNCH_CHECK("sh_encode_backward");
    return LNRF_OK;
}

}  // extern "C"
