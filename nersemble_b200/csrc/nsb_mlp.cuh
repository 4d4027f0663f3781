// Small register-resident GEMM helpers shared by the forward and backward field kernels.
#pragma once
#include "nsb_common.cuh"

namespace nsb {

constexpr int kFieldPackedU4 = 256 + 128 + 256 + 512 + 128;  // base0, base1, head0, head1, head2 (uint4 units)
constexpr int kFeatStride = 40;  // halfs per feature row in smem (bank-conflict-free fragment loads)

// generic small GEMM with weights resident in shared memory (fragment order [kt][pair][lane])
template <int KT, int NP>
__device__ __forceinline__ void smem_gemm(float (&acc)[2 * NP][4], const uint32_t (&a)[KT][4], const uint4 *w, int lane) {
#pragma unroll
    for (int i = 0; i < 2 * NP; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[i][k] = 0.f;
#pragma unroll
    for (int kt = 0; kt < KT; ++kt) {
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            uint4 b = w[(kt * NP + p) * 32 + lane];
            mma16816(acc[2 * p], a[kt], b.x, b.y);
            mma16816(acc[2 * p + 1], a[kt], b.z, b.w);
        }
    }
}

template <int NT>
__device__ __forceinline__ void relu_pack_nobias(const float (&acc)[NT][4], uint32_t (&nxt)[NT / 2][4]) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        nxt[nt / 2][(nt & 1) * 2 + 0] = pack_h2(fmaxf(acc[nt][0], 0.f), fmaxf(acc[nt][1], 0.f));
        nxt[nt / 2][(nt & 1) * 2 + 1] = pack_h2(fmaxf(acc[nt][2], 0.f), fmaxf(acc[nt][3], 0.f));
    }
}

}  // namespace nsb
