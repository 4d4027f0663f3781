// Fused per-sample field evaluation for sm_100a:
//   windowed posenc + warp code -> SE(3) deformation MLP (tensor cores) -> warp ->
//   32-member hash-ensemble gather + time-latent blend (HBM-bound) -> density MLP -> colour MLP.
//
// Replaces (reference, relative to /root/reference/src/nersemble/nerfstudio/):
//   field_components/deformation_field.py:77-107,148-166   SE3WarpingField / compute_offsets
//   field_components/windowed_nerf_encoding.py:33-74       WindowedNeRFEncoding
//   field_components/hash_ensemble.py:93-158                HashEnsemble.forward (8 tcnn grids + einsum)
//   fields/nersemble_nerfacto_field.py:250-301,303-383      get_density / get_outputs (tcnn MLPs)
//
// Work decomposition: see the comment block above field_kernel_ws (persistent, one CTA per SM, warp-specialised:
// tensor warps run the deformation MLP and the density/colour MLPs on mma.sync with the deformation weights
// streamed through a cp.async.bulk ring; gather warps do nothing but the 128-byte-line hash gather).
#include <algorithm>
#include <cstdlib>
#include <cstring>

#define NSB_NO_SPIN_GUARD 1   // see nsb_common.cuh mbar_wait
#include "nsb_common.cuh"
#include "nsb_gather.cuh"
#include "nsb_march.cuh"
#include "nsb_mlp.cuh"
#include "nsb_tc.cuh"

namespace nsb {

constexpr int kSlabBytes = 2048;  // one k-tile (16) x 64 output columns, fragment order
constexpr int kChunkSlabs = 4;
constexpr int kChunkBytes = kSlabBytes * kChunkSlabs;
constexpr int kStages = 4;
constexpr int kBiasFloats = 6 * 128 + 8;

struct FieldArgs {
    nsb_field_params P;
    nsb_field_opts O;
    nsb_samples S;
    nsb_field_out out;
    float aabb_size[3];
};

__device__ __forceinline__ void zero_acc(float (&acc)[8][4]) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[i][k] = 0.f;
}

// -------------------------------------------------------------------------------------------
// SE(3) exponential map applied to p (util/pytorch3d.py:107-191 + deformation_field.py:95-107)
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ void cross3(const float a[3], const float b[3], float o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ void se3_apply(const float p[3], const float r[3], const float v[3], float out[3]) {
    float n2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
    float th = sqrtf(fmaxf(n2, 1e-4f));
    float s, c;
    sincosf(th, &s, &c);
    float f1 = s / th, f2 = (1.0f - c) / (th * th), f3 = (th - s) / (th * th * th);
    float rp[3], rrp[3], rv[3], rrv[3];
    cross3(r, p, rp); cross3(r, rp, rrp); cross3(r, v, rv); cross3(r, rv, rrv);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float o = (p[k] + f1 * rp[k] + f2 * rrp[k]) + (v[k] + f2 * rv[k] + f3 * rrv[k]);
        out[k] = isnan(o) ? p[k] : o;
    }
}

// ===========================================================================================
// v2: warp-specialised persistent kernel, 1 CTA per SM (profiles/r1: with every warp doing every stage the
// gather is latency-bound on the number of gathering warps and cannot overlap the deformation phase).
//   4 TENSOR warps (120 regs): deformation MLP of tile i+1 (32 rows each, B fragments reused by two
//     m-tiles), then density+colour MLPs of tile i.  Tensor warp 0 lane 0 also refills the weight ring.
//   24 GATHER warps (64 regs): nothing but the hash-ensemble gather; rows of a tile are claimed
//     dynamically; each warp keeps 2 m-tiles (2 x 2 LDG.256) in flight.
// Hand-off through shared memory, double-buffered by tile parity, two mbarrier arrays:
//   xs_full[b]   tensor -> gather : warped positions of tile i are in sm.xs[b]
//   feat_full[b] gather -> tensor : blended features of tile i are in sm.feat[b]
// The reverse (buffer-free) edges are implied by program order: tensor warps issue D(i+2) only after
// F(i), which waited for feat_full(i), i.e. for every gather warp to be done with tile i.
// ===========================================================================================
#ifndef NSB_WS_MT
#define NSB_WS_MT 1
#endif
constexpr int kMT = NSB_WS_MT;                    // m-tiles (16 rows) per tensor warp
constexpr int kTRows = 16 * kMT;                  // rows per tensor warp
constexpr int kTensorWarps = NSB_TILE / kTRows;   // 8 (kMT=1) or 4 (kMT=2)
constexpr int kGatherWarps = 28 - kTensorWarps;   // 20 or 24: 28 warps = 7 warpgroups
constexpr int kThreadsWS = (kTensorWarps + kGatherWarps) * 32;
// Register budget: 896 threads are launched with 72 registers each (64512 of the SM's 65536).
// setmaxnreg can only move registers WITHIN the CTA's allocation (inc blocks until a dec released
// enough -- an inc that exceeds the pool hangs the kernel), so
//   kTensorWarps*32*kTensorRegs + kGatherWarps*32*kGatherRegs <= 64512.
#ifndef NSB_GATHER_REGS
#define NSB_GATHER_REGS 64
#endif
#ifndef NSB_TENSOR_REGS
#define NSB_TENSOR_REGS (NSB_WS_MT == 1 ? 80 : 120)
#endif
constexpr int kGatherRegs = NSB_GATHER_REGS;
constexpr int kTensorRegs = NSB_TENSOR_REGS;
static_assert(kTensorWarps * 32 * kTensorRegs + kGatherWarps * 32 * kGatherRegs <= kThreadsWS * 72, "register pool");
constexpr int kLaunchBoundWS = kThreadsWS;
// Slab layout of deform_packed_tb: layers 0 and 4 without their 128 warp-code columns (those enter as the
// per-timestep bias deform_code_bias[t][0|1][128]).
constexpr int kT_L0 = 0;                  // 2 x 3
constexpr int kT_L1 = kT_L0 + 6;          // 2 x 8
constexpr int kT_L2 = kT_L1 + 16;
constexpr int kT_L3 = kT_L2 + 16;
constexpr int kT_L4 = kT_L3 + 16;         // 2 x 11
constexpr int kT_L5 = kT_L4 + 22;         // 2 x 8
constexpr int kT_HEADS = kT_L5 + 16;      // 92
constexpr int kTbNumSlabs = kT_HEADS + 2; // 94
constexpr int kTbNumChunks = (kTbNumSlabs + kChunkSlabs - 1) / kChunkSlabs;  // 24
static_assert(kTbNumChunks % (2 * kStages) == 0 && kT_HEADS % kChunkSlabs == 0, "ring parity / heads chunk");
static_assert(kTensorWarps % 4 == 0 && kGatherWarps % 4 == 0, "setmaxnreg works on warpgroups");

struct alignas(16) TensorScratch {
    float pos[kTRows][4];          // world position xyz, w = timestep (int bits)
    float dirsel[2][kTRows][4];    // [tile parity] ray direction xyz, w = in-box selector
    uint4 enc[kMT][3][32];         // posenc A fragments [m-tile][k-tile][lane]
    uint4 act[2][kMT][8][32];      // [ping-pong][m-tile][k-tile][lane] hidden activations (lane-private)
};

struct alignas(128) SmemWS {
    uint8_t ring[kStages][kChunkBytes];
    uint4 field_w[kFieldPackedU4];
    float bias[kBiasFloats];
    uint64_t full[kStages];
    uint64_t empty[kStages];
    uint64_t xs_full[2];
    uint64_t feat_full[2];
    int tile_ctr[2];
    int64_t n_dyn;                          // fused render kernel: the packed sample count (known on the device only)
    uint64_t sampler_done;                  // fused render kernel: mbarrier, the tensor warps' sampler phase -> the gather warps
    alignas(16) float xs[2][NSB_TILE][4];  // normalised warped position (0 outside the box), w = timestep bits
    alignas(16) __half feat[2][NSB_TILE * kFeatStride];
    TensorScratch ts[kTensorWarps];
    uint2 blend_b[kGatherWarps][4 * 32];   // per gather warp: the current timestep's B fragments (lane-private columns)
    uint4 cv_stage[kGatherWarps][32];      // per gather warp: the current sample's corner values (training forward)
};

struct RingRefill {
    bool active;           // tensor warp 0, lane 0
    uint32_t gbase;        // global chunk index of this tile's chunk 0
    uint32_t total;        // chunks this CTA will consume in total
    const uint8_t *src;
    __device__ __forceinline__ void issue(SmemWS &sm, uint32_t gt) const {
        if (active && gt < total) {
            const uint32_t s = gt % kStages, kf = gt / kStages, c = gt % kTbNumChunks;
            mbar_wait<20>(&sm.empty[s], (kf & 1) ^ 1);   // every tensor warp released the previous use
            const uint32_t bytes = (c == kTbNumChunks - 1) ? (kTbNumSlabs - c * kChunkSlabs) * kSlabBytes : kChunkBytes;
            mbar_expect_tx(&sm.full[s], bytes);
            bulk_g2s(sm.ring[s], src + (size_t)c * kChunkBytes, bytes, &sm.full[s]);
        }
    }
};

// one N-half of a deformation layer for TWO m-tiles (32 rows): each B fragment pair feeds 4 HMMAs
template <class AFn>
__device__ __forceinline__ void ring_gemm2(float (&acc)[kMT][8][4], const int j0, const int KT, AFn &&afn, SmemWS &sm,
                                           const RingRefill &rf, int lane) {
    static_assert(kChunkSlabs == 4 && kStages == 4 && kTbNumChunks % (2 * kStages) == 0, "ring index arithmetic");
#pragma unroll 2
    for (int kt = 0; kt < KT; ++kt) {
        const int j = j0 + kt;
        const int chunk = j >> 2, stage = chunk & 3;
        if ((j & 3) == 0) {
            rf.issue(sm, rf.gbase + chunk + (kStages - 1));   // refill three chunks ahead
            __syncwarp();
            mbar_wait<20>(&sm.full[stage], (chunk >> 2) & 1);
        }
        uint32_t a[kMT][4];
#pragma unroll
        for (int m = 0; m < kMT; ++m) afn(m, kt, a[m]);
        const uint4 *slab = reinterpret_cast<const uint4 *>(&sm.ring[stage][(j & 3) * kSlabBytes]);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const uint4 b = slab[p * 32 + lane];
#pragma unroll
            for (int m = 0; m < kMT; ++m) {
                mma16816(acc[m][2 * p], a[m], b.x, b.y);
                mma16816(acc[m][2 * p + 1], a[m], b.z, b.w);
            }
        }
        if ((j & 3) == 3) {
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.empty[stage]);
        }
    }
}

// bias + ReLU + pack one N-half straight into the ping-pong activation buffer.
//   CODE = false: `sbias` is the layer's shared bias row (smem), one load serves both row halves.
//   CODE = true : layers 0 / 4 -- `gbias` is the per-timestep code-bias table (global, L1/L2 resident: 24 x 1 KB) and
//                 off[m][0|1] the float offset of the bias row of the m-tile's rows g / g+8.  Offsets instead of
//                 pointers: four 64-bit row pointers per m-tile cost 8 registers for the whole MLP.
template <int HALF, bool CODE>
__device__ __forceinline__ void relu_store2(const float (&acc)[kMT][8][4], uint4 (*dst)[8][32], const float *sbias,
                                            const float *gbias, const uint32_t (&off)[kMT][2], int q, int lane) {
#pragma unroll
    for (int m = 0; m < kMT; ++m) {
#pragma unroll
        for (int kt = 0; kt < 4; ++kt) {
            uint32_t r[4];
#pragma unroll
            for (int o = 0; o < 2; ++o) {
                const int nt = 2 * kt + o, col = HALF * 64 + nt * 8 + 2 * q;
                float2 b0, b1;
                if (CODE) {
                    b0 = __ldg(reinterpret_cast<const float2 *>(gbias + off[m][0] + col));
                    b1 = __ldg(reinterpret_cast<const float2 *>(gbias + off[m][1] + col));
                } else {
                    b0 = *reinterpret_cast<const float2 *>(sbias + col);
                    b1 = b0;
                }
                r[o * 2 + 0] = pack_h2(fmaxf(acc[m][nt][0] + b0.x, 0.f), fmaxf(acc[m][nt][1] + b0.y, 0.f));
                r[o * 2 + 1] = pack_h2(fmaxf(acc[m][nt][2] + b1.x, 0.f), fmaxf(acc[m][nt][3] + b1.y, 0.f));
            }
            dst[m][HALF * 4 + kt][lane] = make_uint4(r[0], r[1], r[2], r[3]);
        }
    }
}

__device__ __forceinline__ void zero_acc2(float (&acc)[kMT][8][4]) {
#pragma unroll
    for (int m = 0; m < kMT; ++m) zero_acc(acc[m]);
}

// density + colour MLPs for one m-tile (16 rows); feat rows / dirsel rows are this m-tile's row 0
template <bool HEAD>
__device__ __forceinline__ void field_mlp_tile(const FieldArgs &A, const uint4 *field_w, const __half *feat,
                                               const float (*dirsel)[4], int64_t row0, int64_t n, int lane) {
    const int g = lane >> 2, q = lane & 3;
    uint32_t fa[2][4];
#pragma unroll
    for (int kt = 0; kt < 2; ++kt) {
        fa[kt][0] = *reinterpret_cast<const uint32_t *>(&feat[g * kFeatStride + kt * 16 + 2 * q]);
        fa[kt][1] = *reinterpret_cast<const uint32_t *>(&feat[(g + 8) * kFeatStride + kt * 16 + 2 * q]);
        fa[kt][2] = *reinterpret_cast<const uint32_t *>(&feat[g * kFeatStride + kt * 16 + 2 * q + 8]);
        fa[kt][3] = *reinterpret_cast<const uint32_t *>(&feat[(g + 8) * kFeatStride + kt * 16 + 2 * q + 8]);
    }
    float b0acc[8][4];
    smem_gemm<2, 4>(b0acc, fa, field_w, lane);
    uint32_t h1[4][4];
    relu_pack_nobias<8>(b0acc, h1);
    float b1acc[2][4];
    smem_gemm<4, 1>(b1acc, h1, field_w + 256, lane);
    const float sel0 = dirsel[g][3], sel1 = dirsel[g + 8][3];
    if (q == 0 && A.out.sigma) {
        if (row0 + g < n) A.out.sigma[row0 + g] = expf(b1acc[0][0]) * sel0;
        if (row0 + g + 8 < n) A.out.sigma[row0 + g + 8] = expf(b1acc[0][2]) * sel1;
    }
    if (!HEAD) return;
    uint32_t ha[2][4];
    {
        float g00 = b1acc[0][0], g02 = b1acc[0][2];
        if (q == 0) { g00 = 1.0f; g02 = 1.0f; }
        ha[0][0] = pack_h2(g00, b1acc[0][1]);
        ha[0][1] = pack_h2(g02, b1acc[0][3]);
        ha[0][2] = pack_h2(b1acc[1][0], b1acc[1][1]);
        ha[0][3] = pack_h2(b1acc[1][2], b1acc[1][3]);
        float e[2][2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = g + 8 * h;
            const float d0 = (dirsel[r][0] + 1.0f) / 2.0f, d1 = (dirsel[r][1] + 1.0f) / 2.0f,
                        d2 = (dirsel[r][2] + 1.0f) / 2.0f;
            e[h][0] = q == 0 ? d0 : (q == 1 ? d2 : 1.0f);
            e[h][1] = q == 0 ? d1 : 1.0f;
        }
        ha[1][0] = pack_h2(e[0][0], e[0][1]);
        ha[1][1] = pack_h2(e[1][0], e[1][1]);
        ha[1][2] = pack_h2(1.0f, 1.0f);
        ha[1][3] = pack_h2(1.0f, 1.0f);
    }
    float c0acc[8][4];
    smem_gemm<2, 4>(c0acc, ha, field_w + 384, lane);
    uint32_t c1in[4][4];
    relu_pack_nobias<8>(c0acc, c1in);
    float c1acc[8][4];
    smem_gemm<4, 4>(c1acc, c1in, field_w + 640, lane);
    uint32_t c2in[4][4];
    relu_pack_nobias<8>(c1acc, c2in);
    float c2acc[2][4];
    smem_gemm<4, 1>(c2acc, c2in, field_w + 1152, lane);
    if (A.out.rgb) {
        const int64_t sa = row0 + g, sb = row0 + g + 8;
        if (q == 0) {
            if (sa < n) { A.out.rgb[3 * sa + 0] = 1.f / (1.f + expf(-c2acc[0][0])); A.out.rgb[3 * sa + 1] = 1.f / (1.f + expf(-c2acc[0][1])); }
            if (sb < n) { A.out.rgb[3 * sb + 0] = 1.f / (1.f + expf(-c2acc[0][2])); A.out.rgb[3 * sb + 1] = 1.f / (1.f + expf(-c2acc[0][3])); }
        } else if (q == 1) {
            if (sa < n) A.out.rgb[3 * sa + 2] = 1.f / (1.f + expf(-c2acc[0][0]));
            if (sb < n) A.out.rgb[3 * sb + 2] = 1.f / (1.f + expf(-c2acc[0][2]));
        }
    }
}

// SAVE: training instantiation, additionally stores the warped positions and the deformation activations the
// backward kernels read; compiled out of the inference instantiation (the stores cost ~40 B of spills there).
template <bool DEFORM, bool FIELD, bool HEAD, bool SAVE>
__global__ void __launch_bounds__(kLaunchBoundWS, 1) field_kernel_ws(const __grid_constant__ FieldArgs A) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    SmemWS &sm = *reinterpret_cast<SmemWS *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr bool FEAT_GIVEN = false;
    // NOTE: nothing is computed ahead of the role split on purpose.  Values shared by both roles get registers that
    // suit the 88-register tensor role, and the 64-register gather role then pays for them with spills inside its
    // sample loop (measured: 2.59 -> 2.77 ms).  Each role derives its loop bounds itself.

#define NSB_N_SAMPLES A.S.n_samples
#include "nsb_field_setup.inc"
    if (warp >= kTensorWarps) {
        // =============================== GATHER warps ===============================
        if (kGatherRegs != 72) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kGatherRegs));
#include "nsb_field_gather_role.inc"
        return;
    }

    // =============================== TENSOR warps ===============================
    if (kTensorRegs != 72) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kTensorRegs));
#include "nsb_field_tensor_role.inc"
#undef NSB_N_SAMPLES
}

// Training forward over the KEPT samples when the density pre-pass already gathered their features (SURVEY 8(f)-1:
// "reuse the pre-pass evaluation"): deformation MLP (activations saved for the backward) + density / colour MLPs from
// nsb_samples.given_feat; the gather warps exit at once -- no table line is read a second time.
template <bool DEFORM>
__global__ void __launch_bounds__(kLaunchBoundWS, 1) field_kernel_ws_given(const __grid_constant__ FieldArgs A) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    SmemWS &sm = *reinterpret_cast<SmemWS *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr bool FIELD = true, HEAD = true, SAVE = true, FEAT_GIVEN = true;
#define NSB_N_SAMPLES A.S.n_samples
#include "nsb_field_setup.inc"
    if (warp >= kTensorWarps) {
        if (kGatherRegs != 72) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kGatherRegs));
#include "nsb_field_gather_role.inc"
        return;
    }
    if (kTensorRegs != 72) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kTensorRegs));
#include "nsb_field_tensor_role.inc"
#undef NSB_N_SAMPLES
}

// ===========================================================================================
// tcgen05 variant of the inference kernels: the deformation MLP on the 5th-generation tensor cores (TMEM accumulator,
// one issuing thread, weights read from shared memory once per 128-row tile) -- nsb_field_tensor_role_tc.inc.
// The gather role is the unchanged include; only the tensor role and the shared-memory plan differ.
// ===========================================================================================
#ifndef NSB_TC_PAIR
#define NSB_TC_PAIR 0       // 1: two tiles per pass over the weights (nsb_field_tensor_role_tc2.inc); 0: one tile (..._tc.inc)
#endif
#ifndef NSB_FRAME_GATHER_WARPS
#define NSB_FRAME_GATHER_WARPS 8     // measured (2^20 samples): 20 warps 1.25 ms, 16: 1.24, 12: 1.21, 8: 1.14
#endif
constexpr int kFrameGatherWarps = NSB_FRAME_GATHER_WARPS;   // frame-table gather: how many of the gather warps do work
#if NSB_TC_PAIR
constexpr int kTcStages = 3, kTcBlocksPerTile = 14, kTcTmemCols = 256, kTcTiles = 2;
#else
constexpr int kTcStages = 4, kTcBlocksPerTile = 14, kTcTmemCols = 128, kTcTiles = 1;
#endif
#ifdef NSB_TC_PROF
__device__ unsigned long long g_tc_prof[8], g_tc_prof_k[8];
#define KPROF(i) if (threadIdx.x == 0 && blockIdx.x == 3) g_tc_prof_k[i] = (unsigned long long)clock64();
#else
#define KPROF(i)
#endif
#ifndef NSB_TC_SPLIT
#define NSB_TC_SPLIT 1      // deformation (tcgen05) and density / colour MLPs (mma.sync) on separate warp groups
#endif
constexpr size_t kTcPackedBytes = 12 * 16384 + 2 * 2048;

struct alignas(1024) SmemTC {
    uint8_t wring[kTcStages][16384];        // weight blocks [128 n x 64 k] (heads: [16 x 64]) in UMMA core-matrix order
#if NSB_TC_PAIR
    uint8_t a_enc[2][16384];                // posenc A operand [128 rows x 64 k] of the pair's two tiles
    uint8_t act[2][32768];                  // hidden activations A operand [128 rows x 128 k] of the two tiles
#else
    uint8_t a_enc[16384];                   // posenc A operand [128 rows x 64 k]
    uint8_t act[32768];                     // hidden activations A operand [128 rows x 128 k]
#endif
    uint4 field_w[kFieldPackedU4];
    alignas(16) float bias[kBiasFloats];
#if NSB_TC_PAIR
    uint64_t full[kTcStages], empty[kTcStages], acc_bar[2], f_done[2];
#else
    uint64_t full[kTcStages], empty[kTcStages], acc_bar, f_done[2];
#endif
    uint64_t xs_full[2], feat_full[2];
    uint32_t tmem_base;
    int tile_ctr[2];
    int64_t n_dyn;
    uint64_t sampler_done;
    alignas(16) float xs[2][NSB_TILE][4];
    alignas(16) __half feat[2][NSB_TILE * kFeatStride];
    alignas(16) float dirsel[2][NSB_TILE][4];
    uint2 blend_b[kGatherWarps][4 * 32];
    uint4 cv_stage[kGatherWarps][1];        // SAVE is never instantiated with the tc role; the gather include names it
};
static_assert(sizeof(SmemTC) <= 227 * 1024, "shared memory plan");

// setup shared by the tc kernels (textual for the same reason as the other role bodies)
#define NSB_TC_SETUP()                                                                                              \
    {                                                                                                               \
        const uint4 *src = reinterpret_cast<const uint4 *>(A.P.field_packed);                                      \
        for (int i = tid; i < kFieldPackedU4; i += kThreadsWS) sm.field_w[i] = __ldg(src + i);                     \
        for (int i = tid; i < kBiasFloats; i += kThreadsWS) sm.bias[i] = __ldg(A.P.deform_bias + i);               \
        for (int i = tid; i < (int)(sizeof(sm.a_enc) / 16); i += kThreadsWS) reinterpret_cast<uint4 *>(&sm.a_enc)[i] = make_uint4(0u, 0u, 0u, 0u); \
        if (tid == 0) {                                                                                             \
            for (int s = 0; s < kTcStages; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }         \
            for (int t = 0; t < kTcTiles; ++t) mbar_init(reinterpret_cast<uint64_t *>(&sm.acc_bar) + t, 1);            \
            mbar_init(&sm.f_done[0], 4); mbar_init(&sm.f_done[1], 4);                                               \
            for (int b = 0; b < 2; ++b) {                                                                           \
                mbar_init(&sm.xs_full[b], 4);                                                                       \
                mbar_init(&sm.feat_full[b], kGatherWarps);                                                          \
                sm.tile_ctr[b] = 0;                                                                                 \
            }                                                                                                       \
            mbar_fence_init();                                                                                      \
        }                                                                                                           \
        if (warp == 0) tc::tmem_alloc(&sm.tmem_base, kTcTmemCols);                                                  \
        tc::fence_before_sync();                                                                                    \
        __syncthreads();                                                                                            \
        tc::fence_after_sync();                                                                                     \
    }

// FRAME: the gather role reads the member-blended frame table (nsb_field_gather_role_frame.inc)
template <bool HEAD, bool FRAME>
__global__ void __launch_bounds__(kLaunchBoundWS, 1) field_kernel_tc(const __grid_constant__ FieldArgs A) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    SmemTC &sm = *reinterpret_cast<SmemTC *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr bool FIELD = true, SAVE = false, FEAT_GIVEN = false;
#define NSB_N_SAMPLES A.S.n_samples
    NSB_TC_SETUP()
    if (warp >= kTensorWarps) {
        if (kGatherRegs != 72) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kGatherRegs));
        if constexpr (FRAME) {
#include "nsb_field_gather_role_frame.inc"
        } else {
#include "nsb_field_gather_role.inc"
        }
        return;
    }
    if (kTensorRegs != 72) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kTensorRegs));
#if NSB_TC_PAIR
#include "nsb_field_tensor_role_tc2.inc"
#else
#include "nsb_field_tensor_role_tc.inc"
#endif
#undef NSB_N_SAMPLES
}

// field_kernel_ws with the sample count read from DEVICE memory (nsb_samples.n_samples_dev; the sync-free training
// sampler: the host never learns how many candidates the march produced).  grid = number of SMs; A.S.n_samples is the
// capacity of the sample arrays.  Same role bodies.
template <bool DEFORM, bool FIELD, bool HEAD, bool SAVE>
__global__ void __launch_bounds__(kLaunchBoundWS, 1) field_kernel_ws_dyn(const __grid_constant__ FieldArgs A) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    SmemWS &sm = *reinterpret_cast<SmemWS *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr bool FEAT_GIVEN = false;
    if (tid == 0) sm.n_dyn = min(*A.S.n_samples_dev, A.S.n_samples);      // visible after the setup's __syncthreads
#define NSB_N_SAMPLES (*reinterpret_cast<const volatile int64_t *>(&sm.n_dyn))
#include "nsb_field_setup.inc"
    if (warp >= kTensorWarps) {
        if (kGatherRegs != 72) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kGatherRegs));
#include "nsb_field_gather_role.inc"
        return;
    }
    if (kTensorRegs != 72) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kTensorRegs));
#include "nsb_field_tensor_role.inc"
#undef NSB_N_SAMPLES
}

// ===========================================================================================
// render_kernel_ws: sampler -> field -> composite in ONE launch (north star: "fused into one kernel").
// A persistent cooperative kernel, one CTA per SM, whose phases are separated by grid-wide barriers:
//   S  sampler.  Fixed stride: one warp per ray (march_fixed_warp).  Occupancy grid (nerfacc traverse_grids): count per ray
//      (one thread per ray) | barrier | per-CTA chunk sums | barrier | exclusive scan -> packed_info, total | barrier |
//      fill.  The packed sample count stays on the device (shared-memory word n_dyn): no host synchronisation.
//   F  the field phase = the body of field_kernel_ws, textually (same setup / gather role / tensor role includes).
//   C  compositing by the tensor warps (composite_ray, one warp per ray) | barrier | global depth clip.
// Per-sample sigma / rgb / offsets cross from F to C through the caller's workspace (32 B per sample: L2-resident
// at 2^20 samples, 0.2 % of the gather traffic).  Every phase runs the device code of the stand-alone kernels, so the
// results are bit-identical to nsb_march_* + nsb_field_forward + nsb_composite_forward.
// ===========================================================================================
struct RenderKArgs {
    FieldArgs F;              // S.* and out.* point into the caller's workspace
    nsb_march_args M;         // occupancy sampler (counts / offsets / outputs in the workspace)
    nsb_composite_args C;
    int32_t sampler, n_per_ray;
    float near_plane;
    int64_t capacity;
    nsb_render_ws_header *hdr;
    int64_t *partials;        // [gridDim.x] chunk sums of the ray-count scan
    int64_t *packed_info;     // [n_rays][2] out (C.packed_info is the same buffer, const)
};

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Grid-wide barrier of the tensor warps (the only warps that run the sampler / compositing phases): one leader thread
// per CTA on a global arrival counter (cooperative launch: all CTAs are resident).  `target` = arrivals expected so
// far = (barrier index + 1) * gridDim.x; the counter is zeroed by the host before the launch.
__device__ __forceinline__ void phase_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kTensorWarps * 32) : "memory"); }
__device__ __forceinline__ void grid_barrier(uint32_t *ctr, const uint32_t target) {
    phase_sync();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
        while (ld_acquire_u32(ctr) < target) __nanosleep(40);
        __threadfence();
    }
    phase_sync();
}

__device__ __forceinline__ int64_t phase_sum_i64(int64_t v, int64_t *smem_warp /* [kTensorWarps] */, const int tid) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    phase_sync();
    if ((tid & 31) == 0) smem_warp[tid >> 5] = v;
    phase_sync();
    int64_t t = 0;
    for (int w = 0; w < kTensorWarps; ++w) t += smem_warp[w];
    return t;
}

// Where the extra phases run.  The 64-register gather role is allocated erratically by ptxas: with the sampler phase
// inline, or as a call ahead of the role split, it went from 1 to 37-82 spill instructions (tools/spill_report.py;
// a spill in the sample loop costs more than everything this kernel saves).  So BOTH extra phases run on the TENSOR
// warps, as calls after the role split: the gather warps' code is the text of field_kernel_ws plus one named barrier
// on which they wait for the sampler (and the sample count in shared memory) before their first tile.
constexpr int kPhaseThreads = kTensorWarps * 32;

#ifndef NSB_RK_SAMPLER_ATTR
#define NSB_RK_SAMPLER_ATTR __forceinline__
#endif
#ifndef NSB_RK_COMPOSITE_ATTR
#define NSB_RK_COMPOSITE_ATTR __forceinline__
#endif
// SAMPLER: 0 fixed-stride march fused; 1 occupancy march fused (count | scan | fill); 2 samples GIVEN: a preceding
// launch (march_occ_coop_kernel, nsb_render.cu) filled the packed arrays and left the count in the workspace header.
// the fixed-stride march of this CTA's rays as a real call (tensor warps, after the role split)
__device__ __noinline__ void fixed_march_rays(const RenderKArgs &K, const int warp, const int lane) {
    const int64_t R = K.C.n_rays;
    for (int64_t r = (int64_t)blockIdx.x * kTensorWarps + warp; r < R; r += (int64_t)gridDim.x * kTensorWarps) {
        const float t0 = march_fixed_t0(K.F.S.origins, K.F.S.directions, K.F.P.aabb, r, K.near_plane);
        march_fixed_warp(t0, r, K.n_per_ray, K.M.step, K.M.t_starts, K.M.t_ends, K.M.ray_indices, lane);
        if (lane == 0) { K.packed_info[2 * r] = r * K.n_per_ray; K.packed_info[2 * r + 1] = K.n_per_ray; }
    }
}

template <int SAMPLER, class SM = SmemWS>
__device__ NSB_RK_SAMPLER_ATTR void render_sampler_phase(const RenderKArgs &K) {
    constexpr bool OCC = SAMPLER == 1;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    SM &sm = *reinterpret_cast<SM *>(smem_raw);
    const FieldArgs &A = K.F;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;      // tid < kPhaseThreads
    uint32_t *const bar = &K.hdr->barrier;
    const int64_t R = K.C.n_rays;
    if (blockIdx.x == 0 && tid == 0) {
        K.hdr->depth_range[0] = 0xffffffffu;
        K.hdr->depth_range[1] = 0u;
        if (SAMPLER != 2 && !(SAMPLER == 4 && K.sampler != 0)) K.hdr->status = 0;
    }
    if constexpr (SAMPLER == 4) {
        // run-time choice between the fixed march and given samples in ONE binary (the tcgen05 render kernel): two
        // instantiations were two draws of ptxas' allocation of the gather role, and the fixed-march draw ran 15 % slower
        if (K.sampler == 0) {
            fixed_march_rays(K, warp, lane);
            if (tid == 0) {
                sm.n_dyn = R * K.n_per_ray;
                if (blockIdx.x == 0) K.hdr->n_total = R * K.n_per_ray;
            }
        } else if (tid == 0) {
            sm.n_dyn = min(__ldcg(&K.hdr->n_total), K.capacity);
        }
        grid_barrier(bar, 1u * gridDim.x);
    } else if constexpr (SAMPLER == 2) {
        if (tid == 0) sm.n_dyn = min(__ldcg(&K.hdr->n_total), K.capacity);
        grid_barrier(bar, 1u * gridDim.x);          // depth_range initialised before any CTA composites
    } else if constexpr (!OCC) {
        for (int64_t r = (int64_t)blockIdx.x * kTensorWarps + warp; r < R; r += (int64_t)gridDim.x * kTensorWarps) {
            const float t0 = march_fixed_t0(A.S.origins, A.S.directions, A.P.aabb, r, K.near_plane);
            march_fixed_warp(t0, r, K.n_per_ray, K.M.step, K.M.t_starts, K.M.t_ends, K.M.ray_indices, lane);
            if (lane == 0) { K.packed_info[2 * r] = r * K.n_per_ray; K.packed_info[2 * r + 1] = K.n_per_ray; }
        }
        if (tid == 0) {
            sm.n_dyn = R * K.n_per_ray;     // the host sized the workspace for exactly this
            if (blockIdx.x == 0) K.hdr->n_total = R * K.n_per_ray;
        }
        grid_barrier(bar, 1u * gridDim.x);
    } else {
        // S1: samples per ray
        for (int64_t r = (int64_t)blockIdx.x * kPhaseThreads + tid; r < R; r += (int64_t)gridDim.x * kPhaseThreads)
            K.M.counts[r] = march_occ_ray<false, 1>(K.M, r, 0, 0);
        grid_barrier(bar, 1u * gridDim.x);
        // S2: exclusive scan of the counts: CTA b owns the contiguous chunk [b * chunk, (b + 1) * chunk)
        int64_t *red = reinterpret_cast<int64_t *>(sm.ring);          // scratch: the weight ring is not live yet
        const int64_t chunk = (R + gridDim.x - 1) / gridDim.x;
        const int64_t r0 = min(R, (int64_t)blockIdx.x * chunk), r1 = min(R, r0 + chunk);
        {
            int64_t v = 0;
            for (int64_t r = r0 + tid; r < r1; r += kPhaseThreads) v += __ldcg(K.M.counts + r);
            const int64_t tot = phase_sum_i64(v, red, tid);
            if (tid == 0) K.partials[blockIdx.x] = tot;
        }
        grid_barrier(bar, 2u * gridDim.x);
        {
            int64_t before = 0, total = 0;
            for (int b = tid; b < (int)gridDim.x; b += kPhaseThreads) {
                const int64_t p = __ldcg(K.partials + b);
                total += p;
                if (b < (int)blockIdx.x) before += p;
            }
            total = phase_sum_i64(total, red, tid);
            before = phase_sum_i64(before, red, tid);
            if (tid == 0) {
                sm.n_dyn = min(total, K.capacity);
                if (blockIdx.x == 0) { K.hdr->n_total = total; if (total > K.capacity) K.hdr->status = 1; }
            }
            // slabs of kPhaseThreads rays: warp scan + warp totals in shared memory + running carry
            int64_t carry = before;
            for (int64_t s0 = r0; s0 < r1; s0 += kPhaseThreads) {
                const int64_t r = s0 + tid;
                const int64_t c = r < r1 ? (int64_t)__ldcg(K.M.counts + r) : 0;
                int64_t inc = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int64_t nb = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += nb;
                }
                phase_sync();
                if (lane == 31) red[warp] = inc;
                phase_sync();
                int64_t wbase = 0, slab = 0;
                for (int w = 0; w < kTensorWarps; ++w) {
                    const int64_t t = red[w];
                    if (w < warp) wbase += t;
                    slab += t;
                }
                if (r < r1) {   // beyond the caller's capacity (status = 1) rays are truncated: nothing downstream reads out of bounds
                    const int64_t st = min(carry + wbase + inc - c, K.capacity);
                    K.packed_info[2 * r] = st;
                    K.packed_info[2 * r + 1] = min(c, K.capacity - st);
                }
                carry += slab;
            }
        }
        grid_barrier(bar, 3u * gridDim.x);
        // S3: fill (packed_info of a ray may come from another CTA: L2 reads)
        for (int64_t r = (int64_t)blockIdx.x * kPhaseThreads + tid; r < R; r += (int64_t)gridDim.x * kPhaseThreads)
            march_occ_ray<true, 1>(K.M, r, __ldcg(K.C.packed_info + 2 * r), K.capacity);
        grid_barrier(bar, 4u * gridDim.x);
    }
}

template <int SAMPLER>
__device__ NSB_RK_COMPOSITE_ATTR void render_composite_phase(const RenderKArgs &K) {
    constexpr uint32_t kBarS = SAMPLER == 1 ? 4u : 1u;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t *const bar = &K.hdr->barrier;
    const int64_t R = K.C.n_rays;
    grid_barrier(bar, (kBarS + 1u) * gridDim.x);
    for (int64_t ray = (int64_t)blockIdx.x * kTensorWarps + warp; ray < R; ray += (int64_t)gridDim.x * kTensorWarps)
        composite_ray(K.C, ray, lane);
    grid_barrier(bar, (kBarS + 2u) * gridDim.x);
    {   // DepthRenderer('expected'): clip to the global [min, max] of the sample midpoints
        const uint32_t lo_u = __ldcg(&K.hdr->depth_range[0]), hi_u = __ldcg(&K.hdr->depth_range[1]);
        // no sample in the whole batch: the reference inserts one fake sample with t_start = t_end = 1 on ray 0
        // (nersemble_volumetric_sampler.py:110-114), whose midpoint clips every ray's depth (0) to 1
        const float lo = lo_u != 0xffffffffu ? ordered_to_float(lo_u) : 1.0f, hi = lo_u != 0xffffffffu ? ordered_to_float(hi_u) : 1.0f;
        for (int64_t r = (int64_t)blockIdx.x * kPhaseThreads + tid; r < R; r += (int64_t)gridDim.x * kPhaseThreads)
            K.C.out_depth[r] = fminf(fmaxf(K.C.out_depth[r], lo), hi);
    }
}

template <bool DEFORM, int SAMPLER>
__global__ void __launch_bounds__(kLaunchBoundWS, 1) render_kernel_ws(const __grid_constant__ RenderKArgs K) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    SmemWS &sm = *reinterpret_cast<SmemWS *>(smem_raw);
    const FieldArgs &A = K.F;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr bool FIELD = true, HEAD = true, SAVE = false, FEAT_GIVEN = false;

#define NSB_N_SAMPLES (*reinterpret_cast<const volatile int64_t *>(&sm.n_dyn))
    if (tid == 0) mbar_init(&sm.sampler_done, 1);                          // fenced + published by the setup below
#include "nsb_field_setup.inc"
    if (warp >= kTensorWarps) {
        if (kGatherRegs != 72) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kGatherRegs));
        mbar_wait<200>(&sm.sampler_done, 0);                               // the sampler is done, sm.n_dyn is set
#include "nsb_field_gather_role.inc"
        return;
    }
    if (kTensorRegs != 72) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kTensorRegs));
    render_sampler_phase<SAMPLER>(K);                                      // S (ends with a grid barrier)
    // hand-off through an mbarrier rather than a named barrier shared by the two roles: compute-sanitizer synccheck
    // reports a barrier that warps reach from two different instructions as divergent
    if (tid == 0) mbar_arrive(&sm.sampler_done);
    {
#include "nsb_field_tensor_role.inc"
    }
#undef NSB_N_SAMPLES
    render_composite_phase<SAMPLER>(K);                                    // C
}

// render_kernel_ws with the deformation MLP on tcgen05 / TMEM (nsb_field_tensor_role_tc.inc); SAMPLER 0 or 2
template <int SAMPLER, bool FRAME>
__global__ void __launch_bounds__(kLaunchBoundWS, 1) render_kernel_tc(const __grid_constant__ RenderKArgs K) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    SmemTC &sm = *reinterpret_cast<SmemTC *>(smem_raw);
    const FieldArgs &A = K.F;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr bool FIELD = true, HEAD = true, SAVE = false, FEAT_GIVEN = false;
#define NSB_N_SAMPLES (*reinterpret_cast<const volatile int64_t *>(&sm.n_dyn))
    KPROF(0)
    if (tid == 0) mbar_init(&sm.sampler_done, 1);
    NSB_TC_SETUP()
    KPROF(1)
    if (warp >= kTensorWarps) {
        if (kGatherRegs != 72) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kGatherRegs));
        mbar_wait<200>(&sm.sampler_done, 0);
        if constexpr (FRAME) {
#include "nsb_field_gather_role_frame.inc"
        } else {
#include "nsb_field_gather_role.inc"
        }
        return;
    }
    if (kTensorRegs != 72) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kTensorRegs));
    render_sampler_phase<SAMPLER, SmemTC>(K);
    KPROF(2)
    if (tid == 0) mbar_arrive(&sm.sampler_done);
    {
#if NSB_TC_PAIR
#include "nsb_field_tensor_role_tc2.inc"
#else
#include "nsb_field_tensor_role_tc.inc"
#endif
    }
#undef NSB_N_SAMPLES
    KPROF(3)
    render_composite_phase<SAMPLER>(K);
    KPROF(4)
}

// -------------------------------------------------------------------------------------------
// stand-alone HashEnsemble.forward (component API): one warp per sample, grid-stride
// -------------------------------------------------------------------------------------------
struct HashArgs {
    nsb_field_params P;
    nsb_field_opts O;
    const float *x;
    const float *codes;
    int64_t n;
    void *out;
    int out_is_half;
};

__global__ void __launch_bounds__(256) hash_blend_kernel(const __grid_constant__ HashArgs H) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t s = warp_global; s < H.n; s += n_warps) {
        const float x = H.x[3 * s + 0], y = H.x[3 * s + 1], z = H.x[3 * s + 2];
#if NSB_GATHER_MMA
        const BlendB Bf = make_blend_b(H.O, H.codes + s * NSB_MEMBERS, lane);
        const float val = gather_blend_mma(H.P, x, y, z, Bf, lane);
#else
        const int mg = lane & 7;
        float4 c = __ldg(reinterpret_cast<const float4 *>(H.codes + s * NSB_MEMBERS) + mg);
        float cw[4];
        cw[0] = fmaf(c.x, H.O.cw_scale[4 * mg + 0], H.O.cw_bias[4 * mg + 0]);
        cw[1] = fmaf(c.y, H.O.cw_scale[4 * mg + 1], H.O.cw_bias[4 * mg + 1]);
        cw[2] = fmaf(c.z, H.O.cw_scale[4 * mg + 2], H.O.cw_bias[4 * mg + 2]);
        cw[3] = fmaf(c.w, H.O.cw_scale[4 * mg + 3], H.O.cw_bias[4 * mg + 3]);
        const float val = gather_blend<2>(H.P, x, y, z, cw, lane);
#endif
        if (H.out_is_half)
            reinterpret_cast<__half *>(H.out)[s * 32 + lane] = __float2half_rn(val);
        else
            reinterpret_cast<float *>(H.out)[s * 32 + lane] = val;
    }
}

// -------------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------------
static int g_num_sms = 0;
static int num_sms() {
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

template <bool D, bool F, bool H, bool SV>
static int launch_field_ws_(const FieldArgs &A, cudaStream_t st) {
    const size_t smem = sizeof(SmemWS);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(field_kernel_ws<D, F, H, SV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(field_kernel_ws): %s", cudaGetErrorString(e));
            return 1;
        }
        configured = true;
    }
    const int64_t n_tiles = (A.S.n_samples + NSB_TILE - 1) / NSB_TILE;
    const int grid = (int)std::min<int64_t>(n_tiles, (int64_t)num_sms());
    field_kernel_ws<D, F, H, SV><<<grid, kThreadsWS, smem, st>>>(A);
    return check_launch("field_kernel_ws");
}
// device-side sample count: instantiated for the field evaluations of the training sampler (density pre-pass; SAVE for the
// pre-pass-reuse variant) -- each instantiation costs ~15 s of ptxas
template <bool D, bool F, bool H, bool SV>
static int launch_field_ws_dyn_(const FieldArgs &A, cudaStream_t st) {
    const size_t smem = sizeof(SmemWS);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(field_kernel_ws_dyn<D, F, H, SV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(field_kernel_ws_dyn): %s", cudaGetErrorString(e)); return 1; }
        configured = true;
    }
    const int64_t n_tiles = (A.S.n_samples + NSB_TILE - 1) / NSB_TILE;
    const int grid = (int)std::min<int64_t>(n_tiles, (int64_t)num_sms());
    field_kernel_ws_dyn<D, F, H, SV><<<grid, kThreadsWS, smem, st>>>(A);
    return check_launch("field_kernel_ws_dyn");
}
template <bool D, bool F, bool H>
static int launch_field_ws(const FieldArgs &A, cudaStream_t st) {
    const bool save = A.out.xs || A.out.deform_acts || A.out.deform_enc || A.out.corner_vals;
    return save ? launch_field_ws_<D, F, H, true>(A, st) : launch_field_ws_<D, F, H, false>(A, st);
}

template <bool H, bool FR>
static int launch_field_tc_(const FieldArgs &A, cudaStream_t st) {
    const size_t smem = sizeof(SmemTC);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(field_kernel_tc<H, FR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(field_kernel_tc): %s", cudaGetErrorString(e)); return 1; }
        configured = true;
    }
    const int64_t n_tiles = (A.S.n_samples + NSB_TILE - 1) / NSB_TILE;
    const int grid = (int)std::min<int64_t>(n_tiles, (int64_t)num_sms());
    field_kernel_tc<H, FR><<<grid, kThreadsWS, smem, st>>>(A);
    return check_launch("field_kernel_tc");
}
template <bool H>
static int launch_field_tc(const FieldArgs &A, cudaStream_t st) {
    return A.P.frame_table ? launch_field_tc_<H, true>(A, st) : launch_field_tc_<H, false>(A, st);
}

template <bool D>
static int launch_field_given(const FieldArgs &A, cudaStream_t st) {
    const size_t smem = sizeof(SmemWS);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(field_kernel_ws_given<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(field_kernel_ws_given): %s", cudaGetErrorString(e)); return 1; }
        configured = true;
    }
    const int64_t n_tiles = (A.S.n_samples + NSB_TILE - 1) / NSB_TILE;
    const int grid = (int)std::min<int64_t>(n_tiles, (int64_t)num_sms());
    field_kernel_ws_given<D><<<grid, kThreadsWS, smem, st>>>(A);
    return check_launch("field_kernel_ws_given");
}

template <bool D, bool F, bool H>
static int launch_field(const FieldArgs &A, cudaStream_t st) {
    if (A.S.given_feat) {
        if (F && H && !A.S.n_samples_dev) return launch_field_given<D>(A, st);
        set_error("nsb_field_forward: given_feat needs the full evaluation (rgb) and a host-side sample count");
        return 1;
    }
    if (A.S.n_samples_dev) {
        // always the SAVE instantiation (its stores are skipped at run time when the pointers are NULL): ptxas gives its
        // gather role 0 spill instructions, the non-SAVE density instantiation 37 (tools/spill_report.py)
        if (F && !H) return launch_field_ws_dyn_<D, true, false, true>(A, st);
        set_error("nsb_field_forward: n_samples_dev is supported for the density evaluation (no rgb) only");
        return 1;
    }
    if (D && F && A.P.deform_packed_umma && !A.S.sample_code_bias && !(A.out.xs || A.out.deform_acts || A.out.deform_enc || A.out.corner_vals || A.out.feat))
        return launch_field_tc<H>(A, st);       // inference with the deformation MLP on tcgen05 / TMEM
    return launch_field_ws<D, F, H>(A, st);
}

}  // namespace nsb

using namespace nsb;

extern "C" size_t nsb_deform_packed_bytes(void) { return (size_t)kTbNumSlabs * kSlabBytes; }
extern "C" size_t nsb_field_packed_bytes(void) { return kFieldPackedU4 * sizeof(uint4); }
extern "C" size_t nsb_deform_packed_umma_bytes(void) { return kTcPackedBytes; }
#ifdef NSB_TC_PROF
extern "C" int nsb_debug_tc_prof(unsigned long long *out8) {
    return (int)cudaMemcpyFromSymbol(out8, nsb::g_tc_prof, sizeof(unsigned long long) * 8);
}
extern "C" int nsb_debug_tc_prof_kernel(unsigned long long *out8) {      // clock64 at the phase boundaries of render_kernel_tc
    return (int)cudaMemcpyFromSymbol(out8, nsb::g_tc_prof_k, sizeof(unsigned long long) * 8);
}
#endif

extern "C" int nsb_field_forward(const nsb_field_params *params, const nsb_field_opts *opts, const nsb_samples *samples,
                                 const nsb_field_out *out, void *stream) {
    if (!params || !opts || !samples || !out) { set_error("nsb_field_forward: null argument"); return 1; }
    if (samples->n_samples <= 0) return 0;
    if (params->levels.n_levels != NSB_MAX_LEVELS) { set_error("nsb_field_forward: n_levels must be 16"); return 1; }
    if (!samples->origins && !samples->positions) { set_error("nsb_field_forward: no sample source"); return 1; }
    if (samples->origins && (!samples->directions || !samples->t_starts || !samples->t_ends || !samples->ray_indices)) {
        set_error("nsb_field_forward: ray-based samples need directions, t_starts, t_ends, ray_indices");
        return 1;
    }
    const bool deform = opts->use_deformation != 0;
    const bool need_field = out->sigma || out->rgb || out->feat || out->xs;
    const bool head = opts->compute_rgb != 0 && out->rgb != nullptr;
    if (!need_field && !(deform && out->offsets)) { set_error("nsb_field_forward: no outputs requested"); return 1; }
    if (need_field && (!params->tables || !params->field_packed)) { set_error("nsb_field_forward: tables/field_packed missing"); return 1; }
    if (need_field && !params->blend_codes && !samples->sample_blend_codes) { set_error("nsb_field_forward: no blend codes"); return 1; }
    if (deform && (!params->deform_packed_tb || !params->deform_bias || (!params->deform_code_bias && !samples->sample_code_bias))) {
        set_error("nsb_field_forward: deformation parameters missing (deform_packed_tb, deform_bias, code bias)");
        return 1;
    }
    if (deform && samples->sample_code_bias && (samples->n_samples > (int64_t(1) << 24) || (need_field && !samples->sample_blend_codes))) {
        set_error("nsb_field_forward: per-sample code bias needs <= 2^24 samples and per-sample blend codes");
        return 1;
    }
    if (params->n_timesteps < 1) { set_error("nsb_field_forward: n_timesteps < 1"); return 1; }
    FieldArgs A;
    A.P = *params; A.O = *opts; A.S = *samples; A.out = *out;
    for (int k = 0; k < 3; ++k) A.aabb_size[k] = params->aabb[3 + k] - params->aabb[k];
    cudaStream_t st = (cudaStream_t)stream;
    if (deform) {
        if (!need_field) return launch_field<true, false, false>(A, st);
        return head ? launch_field<true, true, true>(A, st) : launch_field<true, true, false>(A, st);
    }
    return head ? launch_field<false, true, true>(A, st) : launch_field<false, true, false>(A, st);
}

extern "C" int nsb_hash_blend_forward(const nsb_field_params *params, const nsb_field_opts *opts, const float *x,
                                      const float *codes, int64_t n, void *out, int32_t out_is_half, void *stream) {
    if (!params || !opts || !x || !codes || !out) { set_error("nsb_hash_blend_forward: null argument"); return 1; }
    if (n <= 0) return 0;
    if (params->levels.n_levels != NSB_MAX_LEVELS) { set_error("nsb_hash_blend_forward: n_levels must be 16"); return 1; }
    HashArgs H;
    H.P = *params; H.O = *opts; H.x = x; H.codes = codes; H.n = n; H.out = out; H.out_is_half = out_is_half;
    const int64_t warps_needed = n;
    const int blocks = (int)std::min<int64_t>((warps_needed + 7) / 8, (int64_t)num_sms() * 8);
    hash_blend_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(H);
    return check_launch("hash_blend_kernel");
}

// -------------------------------------------------------------------------------------------
// nsb_blend_tables: frame table = the 32 members of every table entry blended with ONE timestep's weights
// (cw[m] = code[m] * cw_scale[m] + cw_bias[m], hash_ensemble.py:119-139), float2 per entry.  8 lanes per entry
// (16 B = 4 members each), fp32 accumulation; one streaming pass over the tables (HBM-bound: 1 GB at ~6 TB/s).
// -------------------------------------------------------------------------------------------
namespace nsb {
__global__ void __launch_bounds__(256) blend_tables_kernel(const uint4 *__restrict__ tables, const float *__restrict__ code,
                                                           const __grid_constant__ nsb_field_opts O, float2 *__restrict__ out,
                                                           int64_t n_entries) {
    const int lane = threadIdx.x & 31, sub = lane & 7;
    float cw[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) cw[i] = fmaf(__ldg(code + 4 * sub + i), O.cw_scale[4 * sub + i], O.cw_bias[4 * sub + i]);
    const int64_t stride = (int64_t)gridDim.x * (blockDim.x >> 3);
    for (int64_t e = (int64_t)blockIdx.x * (blockDim.x >> 3) + (threadIdx.x >> 3); e < ((n_entries + 3) & ~(int64_t)3); e += stride) {
        float f0 = 0.f, f1 = 0.f;
        if (e < n_entries) {
            const uint4 v = __ldcs(tables + e * 8 + sub);          // streamed once: evict-first
            const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 f = unpack_h2(u[i]);
                f0 = fmaf(cw[i], f.x, f0); f1 = fmaf(cw[i], f.y, f1);
            }
        }
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) { f0 += __shfl_xor_sync(0xffffffffu, f0, o); f1 += __shfl_xor_sync(0xffffffffu, f1, o); }
        if (sub == 0 && e < n_entries) out[e] = make_float2(f0, f1);
    }
}
}  // namespace nsb

extern "C" int nsb_blend_tables(const nsb_field_params *params, const nsb_field_opts *opts, int32_t timestep, int64_t n_entries,
                                void *out, void *stream) {
    if (!params || !opts || !out || !params->tables || !params->blend_codes) { set_error("nsb_blend_tables: null argument"); return 1; }
    if (timestep < 0 || timestep >= params->n_timesteps) { set_error("nsb_blend_tables: timestep out of range"); return 1; }
    if (n_entries <= 0) return 0;
    const int blocks = (int)std::min<int64_t>((n_entries + 31) / 32, (int64_t)num_sms() * 8);
    blend_tables_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4 *>(params->tables),
                                                                  params->blend_codes + (size_t)timestep * NSB_MEMBERS, *opts,
                                                                  reinterpret_cast<float2 *>(out), n_entries);
    return check_launch("blend_tables_kernel");
}

// -------------------------------------------------------------------------------------------
// nsb_render_forward: host side of render_kernel_ws
// -------------------------------------------------------------------------------------------
namespace nsb {
constexpr size_t kRenderHdrBytes = 64, kRenderPartials = 1024;
static_assert(sizeof(nsb_render_ws_header) == kRenderHdrBytes, "workspace header layout");

template <int SAMPLER, bool FR>
static int launch_render_tc_(const RenderKArgs &K, cudaStream_t st) {
    const size_t smem = sizeof(SmemTC);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(render_kernel_tc<SAMPLER, FR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(render_kernel_tc): %s", cudaGetErrorString(e)); return 1; }
        configured = true;
    }
    const int grid = num_sms();
    if ((size_t)grid > kRenderPartials) { set_error("nsb_render_forward: more SMs than scan partials"); return 1; }
    cudaError_t e = cudaMemsetAsync(&K.hdr->barrier, 0, sizeof(uint32_t), st);
    if (e != cudaSuccess) { set_error("nsb_render_forward: memset: %s", cudaGetErrorString(e)); return 2; }
    void *kargs[] = {const_cast<RenderKArgs *>(&K)};
    e = cudaLaunchCooperativeKernel(reinterpret_cast<const void *>(render_kernel_tc<SAMPLER, FR>), dim3(grid), dim3(kThreadsWS), kargs, smem, st);
    if (e != cudaSuccess) { set_error("render_kernel_tc: %s", cudaGetErrorString(e)); return 2; }
    return check_launch("render_kernel_tc");
}
template <int SAMPLER>
static int launch_render_tc(const RenderKArgs &K, cudaStream_t st) {
    return K.F.P.frame_table ? launch_render_tc_<SAMPLER, true>(K, st) : launch_render_tc_<SAMPLER, false>(K, st);
}

template <bool D, int SAMPLER>
static int launch_render(const RenderKArgs &K, cudaStream_t st) {
    if constexpr (D && SAMPLER != 1)
        if (K.F.P.deform_packed_umma)       // every instantiation is its own draw of ptxas' allocation: measured (tools/
            // render_time.py, 2^20 samples) <0> 2.35 ms, <4> (run-time sampler choice, fixed march as a call) 2.13 fixed /
            // 2.18 given, <2> 2.02 given -> the fixed march runs the <4> binary, given samples the <2> binary
            return SAMPLER == 0 ? launch_render_tc<4>(K, st) : launch_render_tc<2>(K, st);
    const size_t smem = sizeof(SmemWS);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(render_kernel_ws<D, SAMPLER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(render_kernel_ws): %s", cudaGetErrorString(e)); return 1; }
        configured = true;
    }
    const int grid = num_sms();     // one CTA per SM, all resident: the grid barriers rely on it (cooperative launch checks)
    if ((size_t)grid > kRenderPartials) { set_error("nsb_render_forward: more SMs than scan partials"); return 1; }
    cudaError_t e = cudaMemsetAsync(&K.hdr->barrier, 0, sizeof(uint32_t), st);
    if (e != cudaSuccess) { set_error("nsb_render_forward: memset: %s", cudaGetErrorString(e)); return 2; }
    void *kargs[] = {const_cast<RenderKArgs *>(&K)};
    e = cudaLaunchCooperativeKernel(reinterpret_cast<const void *>(render_kernel_ws<D, SAMPLER>), dim3(grid), dim3(kThreadsWS),
                                    kargs, smem, st);
    if (e != cudaSuccess) { set_error("render_kernel_ws: %s", cudaGetErrorString(e)); return 2; }
    return check_launch("render_kernel_ws");
}
}  // namespace nsb

extern "C" size_t nsb_render_workspace_bytes(int64_t n_rays) {
    return kRenderHdrBytes + kRenderPartials * sizeof(int64_t) + (size_t)std::max<int64_t>(n_rays, 1) * sizeof(int32_t) + 64;
}

extern "C" int nsb_render_forward(const nsb_field_params *params, const nsb_field_opts *opts, const nsb_render_args *ra, void *stream) {
    if (!params || !opts || !ra) { set_error("nsb_render_forward: null argument"); return 1; }
    if (ra->n_rays <= 0) return 0;
    if (params->levels.n_levels != NSB_MAX_LEVELS) { set_error("nsb_render_forward: n_levels must be 16"); return 1; }
    const bool deform = opts->use_deformation != 0;
    if (!ra->origins || !ra->directions || !ra->t_starts || !ra->t_ends || !ra->ray_indices || !ra->sigma || !ra->rgb ||
        !ra->packed_info || !ra->out_rgb || !ra->out_acc || !ra->out_depth || !ra->workspace || (deform && !ra->offsets)) {
        set_error("nsb_render_forward: null buffer");
        return 1;
    }
    if (!params->tables || !params->field_packed || !params->blend_codes) { set_error("nsb_render_forward: tables/field_packed/blend_codes missing"); return 1; }
    if (deform && (!params->deform_packed_tb || !params->deform_bias || !params->deform_code_bias)) {
        set_error("nsb_render_forward: deformation parameters missing");
        return 1;
    }
    if (params->n_timesteps < 1 || ra->capacity <= 0) { set_error("nsb_render_forward: n_timesteps < 1 or capacity <= 0"); return 1; }
    if (ra->sampler == 0) {
        if (ra->n_per_ray <= 0 || ra->capacity < ra->n_rays * (int64_t)ra->n_per_ray) { set_error("nsb_render_forward: capacity < n_rays * n_per_ray"); return 1; }
    } else if (ra->sampler == 1 || ra->sampler == 3) {
        if (!ra->near_planes || !ra->far_planes || !ra->binaries || !ra->aabbs) { set_error("nsb_render_forward: occupancy sampler arguments"); return 1; }
        if (ra->levels < 1 || ra->levels > 8 || (ra->sampler == 3 && ra->levels != 1)) {
            set_error("nsb_render_forward: levels must be in [1,8] (1 for the single-launch variant)");
            return 1;
        }
    } else { set_error("nsb_render_forward: unknown sampler"); return 1; }
    uint8_t *ws = reinterpret_cast<uint8_t *>(ra->workspace);
    RenderKArgs K;
    memset(&K, 0, sizeof(K));
    K.hdr = reinterpret_cast<nsb_render_ws_header *>(ws);
    K.partials = reinterpret_cast<int64_t *>(ws + kRenderHdrBytes);
    int32_t *counts = reinterpret_cast<int32_t *>(ws + kRenderHdrBytes + kRenderPartials * sizeof(int64_t));
    K.F.P = *params; K.F.O = *opts;
    K.F.O.compute_rgb = 1;
    for (int k = 0; k < 3; ++k) K.F.aabb_size[k] = params->aabb[3 + k] - params->aabb[k];
    K.F.S.n_samples = 0;                 // device-side: sm.n_dyn
    K.F.S.origins = ra->origins; K.F.S.directions = ra->directions; K.F.S.ray_times = ra->ray_times;
    K.F.S.t_starts = ra->t_starts; K.F.S.t_ends = ra->t_ends; K.F.S.ray_indices = ra->ray_indices;
    K.F.out.sigma = ra->sigma; K.F.out.rgb = ra->rgb; K.F.out.offsets = deform ? ra->offsets : nullptr;
    K.M.n_rays = ra->n_rays; K.M.origins = ra->origins; K.M.directions = ra->directions;
    K.M.near_planes = ra->near_planes; K.M.far_planes = ra->far_planes; K.M.binaries = ra->binaries; K.M.aabbs = ra->aabbs;
    K.M.levels = ra->levels; K.M.res = ra->res; K.M.step = ra->step; K.M.cone_angle = ra->cone_angle;
    K.M.counts = counts; K.M.offsets = nullptr; K.M.t_starts = ra->t_starts; K.M.t_ends = ra->t_ends; K.M.ray_indices = ra->ray_indices;
    K.C.n_rays = ra->n_rays; K.C.n_samples = ra->capacity; K.C.packed_info = ra->packed_info; K.packed_info = ra->packed_info;
    K.C.t_starts = ra->t_starts; K.C.t_ends = ra->t_ends; K.C.sigma = ra->sigma; K.C.rgb = ra->rgb;
    K.C.offsets = deform ? ra->offsets : nullptr; K.C.training = ra->training;
    K.C.out_rgb = ra->out_rgb; K.C.out_acc = ra->out_acc; K.C.out_depth = ra->out_depth;
    K.C.out_deform = deform ? ra->out_deform : nullptr; K.C.out_weights = ra->weights;
    K.C.workspace = K.hdr->depth_range;
    K.sampler = ra->sampler; K.n_per_ray = ra->n_per_ray; K.near_plane = ra->near_plane; K.capacity = ra->capacity;
    cudaStream_t st = (cudaStream_t)stream;
    if (ra->sampler == 3) return deform ? launch_render<true, 1>(K, st) : launch_render<false, 1>(K, st);
    if (ra->sampler == 1) {
        // occupancy march as its own small cooperative launch (count | scan | fill, the count stays on the device), then
        // field + composite in one: with the marcher inside render_kernel_ws ptxas spills 66-82 instructions in the gather
        // role of the field phase (tools/spill_report.py); sampler == 3 selects that single-launch variant.
        const int rc = launch_march_occ_coop(K.M, K.packed_info, K.hdr, K.partials, K.capacity, ra->march_scratch, st);
        if (rc) return rc;
        return deform ? launch_render<true, 2>(K, st) : launch_render<false, 2>(K, st);
    }
    return deform ? launch_render<true, 0>(K, st) : launch_render<false, 0>(K, st);
}
