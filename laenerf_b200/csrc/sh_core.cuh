// sh_core.cuh -- real spherical-harmonics basis (degree <= 4) and its Jacobian, shared by shenc.cu and the fused
// NeRF network kernel (nerfnet.cu).  Follows shencoder/src/shencoder.cu:27-123 of the reference.
#pragma once
#include "common.cuh"

namespace lnrf {

// normalisation constants of the real SH basis (closed forms in the comments)
constexpr float kSH0 = 0.28209479177387814f;   // 1 / (2 sqrt(pi))
constexpr float kSH1 = 0.48860251190291987f;   // sqrt(3) / (2 sqrt(pi))
constexpr float kSH2a = 1.0925484305920792f;   // sqrt(15) / (2 sqrt(pi))
constexpr float kSH2b = 0.94617469575755997f;  // 3 sqrt(5) / (4 sqrt(pi))
constexpr float kSH2c = 0.31539156525251999f;  // sqrt(5) / (4 sqrt(pi))
constexpr float kSH2d = 0.54627421529603959f;  // sqrt(15) / (4 sqrt(pi))
constexpr float kSH3a = 0.59004358992664352f;  // sqrt(70) / (8 sqrt(pi))
constexpr float kSH3b = 2.8906114426405538f;   // sqrt(105) / (2 sqrt(pi))
constexpr float kSH3c = 0.45704579946446572f;  // sqrt(42) / (8 sqrt(pi))
constexpr float kSH3d = 0.3731763325901154f;   // sqrt(7) / (4 sqrt(pi))
constexpr float kSH3e = 1.4453057213202769f;   // sqrt(105) / (4 sqrt(pi))

__device__ __forceinline__ void sh_basis(float x, float y, float z, uint32_t degree, float* v) {
    const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
    v[0] = kSH0;
    if (degree <= 1) return;
    v[1] = -kSH1 * y; v[2] = kSH1 * z; v[3] = -kSH1 * x;
    if (degree <= 2) return;
    v[4] = kSH2a * xy; v[5] = -kSH2a * yz; v[6] = kSH2b * z2 - kSH2c; v[7] = -kSH2a * xz; v[8] = kSH2d * x2 - kSH2d * y2;
    if (degree <= 3) return;
    v[9] = kSH3a * y * (-3.0f * x2 + y2); v[10] = kSH3b * xy * z; v[11] = kSH3c * y * (1.0f - 5.0f * z2);
    v[12] = kSH3d * z * (5.0f * z2 - 3.0f); v[13] = kSH3c * x * (1.0f - 5.0f * z2); v[14] = kSH3e * z * (x2 - y2);
    v[15] = kSH3a * x * (-x2 + 3.0f * y2);
}

// Jacobian rows d/dx, d/dy, d/dz of the basis above
__device__ __forceinline__ void sh_jacobian(float x, float y, float z, uint32_t degree, float* dx, float* dy, float* dz) {
    const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
    for (int i = 0; i < 16; i++) dx[i] = dy[i] = dz[i] = 0.f;
    if (degree <= 1) return;
    dy[1] = -kSH1; dz[2] = kSH1; dx[3] = -kSH1;
    if (degree <= 2) return;
    dx[4] = kSH2a * y; dy[4] = kSH2a * x;
    dy[5] = -kSH2a * z; dz[5] = -kSH2a * y;
    dz[6] = 2.0f * kSH2b * z;
    dx[7] = -kSH2a * z; dz[7] = -kSH2a * x;
    dx[8] = 2.0f * kSH2d * x; dy[8] = -2.0f * kSH2d * y;
    if (degree <= 3) return;
    dx[9] = -6.0f * kSH3a * xy; dy[9] = kSH3a * (-3.0f * x2 + 3.0f * y2);
    dx[10] = kSH3b * yz; dy[10] = kSH3b * xz; dz[10] = kSH3b * xy;
    dy[11] = kSH3c * (1.0f - 5.0f * z2); dz[11] = -10.0f * kSH3c * yz;
    dz[12] = kSH3d * (15.0f * z2 - 3.0f);
    dx[13] = kSH3c * (1.0f - 5.0f * z2); dz[13] = -10.0f * kSH3c * xz;
    dx[14] = 2.0f * kSH3e * xz; dy[14] = -2.0f * kSH3e * yz; dz[14] = kSH3e * (x2 - y2);
    dx[15] = kSH3a * (-3.0f * x2 + 3.0f * y2); dy[15] = 6.0f * kSH3a * xy;
}

}  // namespace lnrf
