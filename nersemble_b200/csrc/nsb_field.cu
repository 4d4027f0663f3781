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

#define NSB_NO_SPIN_GUARD 1   // see nsb_common.cuh mbar_wait
#include "nsb_common.cuh"
#include "nsb_gather.cuh"
#include "nsb_mlp.cuh"

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
constexpr int kGatherRegs = 64;
constexpr int kTensorRegs = kMT == 1 ? 80 : 120;
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
    // NOTE: nothing is computed ahead of the role split on purpose.  Values shared by both roles get registers that
    // suit the 88-register tensor role, and the 64-register gather role then pays for them with spills inside its
    // sample loop (measured: 2.59 -> 2.77 ms).  Each role derives its loop bounds itself.

    if (FIELD) {
        const uint4 *src = reinterpret_cast<const uint4 *>(A.P.field_packed);
        for (int i = tid; i < kFieldPackedU4; i += kThreadsWS) sm.field_w[i] = __ldg(src + i);
    }
    if (DEFORM)
        for (int i = tid; i < kBiasFloats; i += kThreadsWS) sm.bias[i] = __ldg(A.P.deform_bias + i);
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.empty[s], kTensorWarps);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&sm.xs_full[b], kTensorWarps);
            mbar_init(&sm.feat_full[b], kGatherWarps);
            sm.tile_ctr[b] = 0;
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp >= kTensorWarps) {
        // =============================== GATHER warps ===============================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kGatherRegs));
        if (!FIELD) return;
        const int g = lane >> 2, q = lane & 3;
        const uint8_t *tab = reinterpret_cast<const uint8_t *>(A.P.tables) + q * 32;
        // 32-bit loop state: the gather role runs in 64 registers (n_tiles < 2^31 for any n_samples < 2^38)
        uint32_t bid, nct;
        asm volatile("mov.u32 %0, %%ctaid.x;" : "=r"(bid));     // volatile: not CSE'd with the tensor role's copy
        asm volatile("mov.u32 %0, %%nctaid.x;" : "=r"(nct));
        const int n_tiles32 = (int)((A.S.n_samples + NSB_TILE - 1) / NSB_TILE);
        uint2 *bslot = sm.blend_b[warp - kTensorWarps];   // lane-private columns: no synchronisation needed
        const BlendBSmem Bf{bslot + lane};
        int cur_ts = -1;
        for (int tile = (int)bid, it = 0; tile < n_tiles32; tile += (int)nct, ++it) {
            const int b = it & 1;
            const int rows_valid = (int)min((int64_t)NSB_TILE, A.S.n_samples - (int64_t)tile * NSB_TILE);
            mbar_wait<20>(&sm.xs_full[b], (uint32_t)(it >> 1) & 1);
            int row = 0, nrow = 0;
            if (lane == 0) { row = atomicAdd(&sm.tile_ctr[b], 1); nrow = atomicAdd(&sm.tile_ctr[b], 1); }
            row = __shfl_sync(0xffffffffu, row, 0);
            nrow = __shfl_sync(0xffffffffu, nrow, 0);
            GatherTile Ga;
            QuadIdx Q;
            Q.entry = 0; Q.w = 0.f;
            float4 xs = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < rows_valid) {
                xs = *reinterpret_cast<const float4 *>(sm.xs[b][row]);
                Q = quad_compute<0>(A.P, xs.x, xs.y, xs.z, lane);
                gather_issue_q<0, 0>(A.P, tab, Q, g, Ga);
            }
            while (row < rows_valid) {
                // blend-weight fragments: rebuilt only when the timestep changes (a tile is 128 consecutive samples,
                // i.e. one or two rays = one or two timesteps; the code row is a global load through an L1 that the
                // table lines flood, at the head of the sample's dependent chain)
                if (A.S.sample_blend_codes) {
                    park_blend_b(make_blend_b(A.O, A.S.sample_blend_codes + ((int64_t)tile * NSB_TILE + row) * NSB_MEMBERS, lane),
                                 bslot, lane);
                } else if (__float_as_int(xs.w) != cur_ts) {
                    cur_ts = __float_as_int(xs.w);
                    park_blend_b(make_blend_b(A.O, A.P.blend_codes + (size_t)cur_ts * NSB_MEMBERS, lane), bslot, lane);
                }
                const bool has_next = nrow < rows_valid;
                __half *feat_row = &sm.feat[b][row * kFeatStride];
                float4 nx = xs;
                // training: corner values are staged in this warp's 512 B of shared memory (one 32-bit address live across
                // the sample instead of a 64-bit global pointer) and leave as ONE coalesced 16 B-per-lane store
                __half2 *cv_row = SAVE ? reinterpret_cast<__half2 *>(sm.cv_stage[warp - kTensorWarps]) : nullptr;
                gather_sample_quad<SAVE>(A.P, tab, xs.x, xs.y, xs.z,
                                         has_next ? reinterpret_cast<const float4 *>(sm.xs[b][nrow]) : nullptr, nx, Bf, Ga, Q,
                                         feat_row, cv_row, lane);
                if (SAVE && A.out.corner_vals) {
                    __syncwarp();
                    reinterpret_cast<uint4 *>(A.out.corner_vals)[((int64_t)tile * NSB_TILE + row) * 32 + lane] =
                        sm.cv_stage[warp - kTensorWarps][lane];
                    __syncwarp();
                }
                if (A.out.feat) {   // feature output / training: the row goes to global too
                    __syncwarp();
                    reinterpret_cast<__half *>(A.out.feat)[((int64_t)tile * NSB_TILE + row) * 32 + lane] = feat_row[lane];
                }
                row = nrow;
                xs = nx;
                if (has_next) {
                    int t = 0;
                    if (lane == 0) t = atomicAdd(&sm.tile_ctr[b], 1);
                    nrow = __shfl_sync(0xffffffffu, t, 0);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.feat_full[b]);
        }
        return;
    }

    // =============================== TENSOR warps ===============================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kTensorRegs));
    const int64_t n = A.S.n_samples;
    const int64_t n_tiles = (n + NSB_TILE - 1) / NSB_TILE;
    const int64_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    TensorScratch &ts = sm.ts[warp];
    const int g = lane >> 2, q = lane & 3;
    const float amin0 = A.P.aabb[0], amin1 = A.P.aabb[1], amin2 = A.P.aabb[2];
    RingRefill rf;
    rf.active = DEFORM && warp == 0 && lane == 0;
    rf.total = (uint32_t)(my_tiles * kTbNumChunks);
    rf.src = reinterpret_cast<const uint8_t *>(A.P.deform_packed_tb);
    rf.gbase = 0;
    if (DEFORM) {
        for (uint32_t c = 0; c < kStages - 1; ++c) rf.issue(sm, c);
        __syncwarp();
    }

    // D(i): per-row inputs + deformation of tile iteration `it` -> sm.xs[it&1]
    auto deform_stage = [&](int64_t it) {
        const int64_t tile = blockIdx.x + it * gridDim.x;
        const int b = (int)(it & 1);
        const int64_t row0 = tile * NSB_TILE + warp * kTRows;
        __syncwarp();
        if (lane < kTRows) {   // one row per lane
            const int64_t s = row0 + lane;
            float px = 0.f, py = 0.f, pz = 0.f, ddx = 0.f, ddy = 0.f, ddz = 0.f, tt = 0.f;
            if (s < n) {
                if (A.S.origins != nullptr) {
                    const int ri = A.S.ray_indices[s];
                    const float mid = __fadd_rn(A.S.t_starts[s], A.S.t_ends[s]);
                    ddx = A.S.directions[3 * (int64_t)ri + 0];
                    ddy = A.S.directions[3 * (int64_t)ri + 1];
                    ddz = A.S.directions[3 * (int64_t)ri + 2];
                    px = __fadd_rn(A.S.origins[3 * (int64_t)ri + 0], __fmul_rn(__fmul_rn(ddx, mid), 0.5f));
                    py = __fadd_rn(A.S.origins[3 * (int64_t)ri + 1], __fmul_rn(__fmul_rn(ddy, mid), 0.5f));
                    pz = __fadd_rn(A.S.origins[3 * (int64_t)ri + 2], __fmul_rn(__fmul_rn(ddz, mid), 0.5f));
                    if (A.S.ray_times) tt = A.S.ray_times[ri];
                } else {
                    px = A.S.positions[3 * s + 0]; py = A.S.positions[3 * s + 1]; pz = A.S.positions[3 * s + 2];
                    ddx = ddy = ddz = 1.0f;
                    if (A.S.sample_directions) {
                        ddx = A.S.sample_directions[3 * s + 0]; ddy = A.S.sample_directions[3 * s + 1];
                        ddz = A.S.sample_directions[3 * s + 2];
                    }
                    if (A.S.sample_times) tt = A.S.sample_times[s];
                }
            }
            int tsi = __float2int_rn(__fmul_rn(tt, (float)(A.P.n_timesteps - 1)));
            tsi = min(max(tsi, 0), A.P.n_timesteps - 1);
            // component API (per-sample warp codes): the code-bias rows are indexed by SAMPLE instead of by timestep
            // (n <= 2^24 and per-sample blend codes whenever the field is evaluated: checked on the host)
            if (A.S.sample_code_bias) tsi = (int)min(s, n - 1);
            ts.pos[lane][0] = px; ts.pos[lane][1] = py; ts.pos[lane][2] = pz; ts.pos[lane][3] = __int_as_float(tsi);
            ts.dirsel[b][lane][0] = ddx; ts.dirsel[b][lane][1] = ddy; ts.dirsel[b][lane][2] = ddz;
        }
        __syncwarp();
        float wx = 0.f, wy = 0.f, wz = 0.f;   // warped world position of row `lane`
        if (DEFORM) {
            rf.gbase = (uint32_t)(it * kTbNumChunks);
            float pn[kMT][2][3];   // [m-tile][row g / g+8][xyz]
#pragma unroll
            for (int m = 0; m < kMT; ++m)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int r = m * 16 + g + 8 * h;
                    pn[m][h][0] = __fdiv_rn(__fsub_rn(ts.pos[r][0], amin0), A.aabb_size[0]);
                    pn[m][h][1] = __fdiv_rn(__fsub_rn(ts.pos[r][1], amin1), A.aabb_size[1]);
                    pn[m][h][2] = __fdiv_rn(__fsub_rn(ts.pos[r][2], amin2), A.aabb_size[2]);
                }
#pragma unroll
            for (int m = 0; m < kMT; ++m)
#pragma unroll
                for (int kt = 0; kt < 3; ++kt) {
                    uint32_t w4[4];
#pragma unroll
                    for (int hi = 0; hi < 2; ++hi) {
                        const int i = kt * 8 + hi * 4 + q;
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            float e0 = 0.f, e1 = 0.f;
                            if (i < 21) {
                                const int d = i / 7, j = i - d * 7;
                                const float pd = d == 0 ? pn[m][h][0] : (d == 1 ? pn[m][h][1] : pn[m][h][2]);   // no local-memory indexing
                                const float arg = (6.283185307179586f * pd) * (float)(1 << j);
                                const float wj = A.O.pe_window[j];
                                e0 = wj * sinf(arg);
                                e1 = wj * sinf(arg + 1.5707963267948966f);
                            } else if (i == 21) {
                                e0 = 6.283185307179586f * pn[m][h][0];
                                e1 = 6.283185307179586f * pn[m][h][1];
                            } else if (i == 22) {
                                e0 = 6.283185307179586f * pn[m][h][2];
                            }
                            w4[hi * 2 + h] = pack_h2(e0, e1);
                        }
                    }
                    ts.enc[m][kt][lane] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
                    if (SAVE && kMT == 1 && A.out.deform_enc)
                        reinterpret_cast<uint4 *>(A.out.deform_enc)[((tile * kTensorWarps + warp) * 3 + kt) * 32 + lane] =
                            make_uint4(w4[0], w4[1], w4[2], w4[3]);
                }
            // layers 0/4 read the per-timestep code bias: float offset of the bias row of rows g / g+8
            uint32_t cb[kMT][2];
#pragma unroll
            for (int m = 0; m < kMT; ++m)
#pragma unroll
                for (int h = 0; h < 2; ++h) cb[m][h] = (uint32_t)__float_as_int(ts.pos[m * 16 + g + 8 * h][3]) * 256u;
            const float *const gcb = A.S.sample_code_bias ? A.S.sample_code_bias : A.P.deform_code_bias;
            auto in_a = [&](int m, int kt, uint32_t(&a)[4]) {
                const uint4 v = ts.enc[m][kt][lane];
                a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
            };
            float acc[kMT][8][4];
            // layer l reads act[src] (or the input), writes act[dst]
            auto hid = [&](int src) {
                return [&, src](int m, int kt, uint32_t(&a)[4]) {
                    const uint4 v = ts.act[src][m][kt][lane];
                    a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
                };
            };
            auto hid1 = hid(1), hid0 = hid(0);
            auto skip_a = [&](int m, int kt, uint32_t(&a)[4]) {
                if (kt < 8) hid1(m, kt, a); else in_a(m, kt - 8, a);
            };
            auto save_act = [&](int layer, int buf) {   // training: keep the layer output for the backward pass
                if (SAVE && kMT == 1 && A.out.deform_acts) {
                    uint4 *dst = reinterpret_cast<uint4 *>(A.out.deform_acts) + (((size_t)tile * kTensorWarps + warp) * 6 + layer) * 256;
#pragma unroll 2      // fully unrolled = 32 transient registers in the 80-register tensor role
                    for (int kt = 0; kt < 8; ++kt) dst[kt * 32 + lane] = ts.act[buf][0][kt][lane];
                }
            };
            // layer 0: posenc (48) -> act[0]; the 128 warp-code columns are in the bias
            zero_acc2(acc); ring_gemm2(acc, kT_L0, 3, in_a, sm, rf, lane); relu_store2<0, true>(acc, ts.act[0], nullptr, gcb, cb, q, lane);
            zero_acc2(acc); ring_gemm2(acc, kT_L0 + 3, 3, in_a, sm, rf, lane); relu_store2<1, true>(acc, ts.act[0], nullptr, gcb, cb, q, lane);
            save_act(0, 0);
            // layer 1: act[0] -> act[1]
            zero_acc2(acc); ring_gemm2(acc, kT_L1, 8, hid0, sm, rf, lane); relu_store2<0, false>(acc, ts.act[1], sm.bias + 1 * 128, nullptr, cb, q, lane);
            zero_acc2(acc); ring_gemm2(acc, kT_L1 + 8, 8, hid0, sm, rf, lane); relu_store2<1, false>(acc, ts.act[1], sm.bias + 1 * 128, nullptr, cb, q, lane);
            save_act(1, 1);
            // layer 2: act[1] -> act[0]
            zero_acc2(acc); ring_gemm2(acc, kT_L2, 8, hid1, sm, rf, lane); relu_store2<0, false>(acc, ts.act[0], sm.bias + 2 * 128, nullptr, cb, q, lane);
            zero_acc2(acc); ring_gemm2(acc, kT_L2 + 8, 8, hid1, sm, rf, lane); relu_store2<1, false>(acc, ts.act[0], sm.bias + 2 * 128, nullptr, cb, q, lane);
            save_act(2, 0);
            // layer 3: act[0] -> act[1]
            zero_acc2(acc); ring_gemm2(acc, kT_L3, 8, hid0, sm, rf, lane); relu_store2<0, false>(acc, ts.act[1], sm.bias + 3 * 128, nullptr, cb, q, lane);
            zero_acc2(acc); ring_gemm2(acc, kT_L3 + 8, 8, hid0, sm, rf, lane); relu_store2<1, false>(acc, ts.act[1], sm.bias + 3 * 128, nullptr, cb, q, lane);
            save_act(3, 1);
            // layer 4 (skip): [act[1] | posenc] -> act[0]
            zero_acc2(acc); ring_gemm2(acc, kT_L4, 11, skip_a, sm, rf, lane); relu_store2<0, true>(acc, ts.act[0], nullptr, gcb + 128, cb, q, lane);
            zero_acc2(acc); ring_gemm2(acc, kT_L4 + 11, 11, skip_a, sm, rf, lane); relu_store2<1, true>(acc, ts.act[0], nullptr, gcb + 128, cb, q, lane);
            save_act(4, 0);
            // layer 5: act[0] -> act[1]
            zero_acc2(acc); ring_gemm2(acc, kT_L5, 8, hid0, sm, rf, lane); relu_store2<0, false>(acc, ts.act[1], sm.bias + 5 * 128, nullptr, cb, q, lane);
            zero_acc2(acc); ring_gemm2(acc, kT_L5 + 8, 8, hid0, sm, rf, lane); relu_store2<1, false>(acc, ts.act[1], sm.bias + 5 * 128, nullptr, cb, q, lane);
            save_act(5, 1);
            // heads (last chunk: slabs 92,93)
            float hacc[kMT][2][4];
#pragma unroll
            for (int m = 0; m < kMT; ++m)
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int k = 0; k < 4; ++k) hacc[m][i][k] = 0.f;
            {
                constexpr int chunk = kT_HEADS / kChunkSlabs, stage = chunk % kStages;
                rf.issue(sm, rf.gbase + chunk + (kStages - 1));
                __syncwarp();
                mbar_wait<20>(&sm.full[stage], (chunk / kStages) & 1);
                const uint4 *hw = reinterpret_cast<const uint4 *>(&sm.ring[stage][0]);
#pragma unroll
                for (int kt = 0; kt < 8; ++kt) {
                    const uint4 bw = hw[kt * 32 + lane];
#pragma unroll
                    for (int m = 0; m < kMT; ++m) {
                        uint32_t am[4];
                        hid1(m, kt, am);
                        mma16816(hacc[m][0], am, bw.x, bw.y);
                        mma16816(hacc[m][1], am, bw.z, bw.w);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.empty[stage]);
            }
            const float hb0 = sm.bias[6 * 128 + 2 * q], hb1 = sm.bias[6 * 128 + 2 * q + 1];
#pragma unroll
            for (int m = 0; m < kMT; ++m) {
                const float c0 = hacc[m][0][0] + hb0, c1 = hacc[m][0][1] + hb1, c2 = hacc[m][0][2] + hb0, c3 = hacc[m][0][3] + hb1;
                const int rsel = q & 1;
                float vr[6];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int srcl = (lane & ~3) | k;
                    const float a0 = __shfl_sync(0xffffffffu, c0, srcl), a1 = __shfl_sync(0xffffffffu, c1, srcl);
                    const float b0 = __shfl_sync(0xffffffffu, c2, srcl), b1 = __shfl_sync(0xffffffffu, c3, srcl);
                    vr[2 * k] = rsel ? b0 : a0;
                    vr[2 * k + 1] = rsel ? b1 : a1;
                }
                const int r = m * 16 + g + 8 * rsel;
                // recomputed (same ops => same bits as pn above): keeping pn live across the MLP costs 6 registers
                const float p[3] = {__fdiv_rn(__fsub_rn(ts.pos[r][0], amin0), A.aabb_size[0]),
                                    __fdiv_rn(__fsub_rn(ts.pos[r][1], amin1), A.aabb_size[1]),
                                    __fdiv_rn(__fsub_rn(ts.pos[r][2], amin2), A.aabb_size[2])};
                const float v[3] = {vr[0], vr[1], vr[2]}, rr[3] = {vr[3], vr[4], vr[5]};
                float pw[3];
                se3_apply(p, rr, v, pw);
                __syncwarp();   // lanes q = 2, 3 read the rows lanes q = 0, 1 update below
                if (q < 2) {
                    const float o0 = pw[0] - p[0], o1 = pw[1] - p[1], o2 = pw[2] - p[2];
                    const int64_t s = row0 + r;
                    if (A.out.offsets && s < n) {
                        A.out.offsets[3 * s + 0] = o0; A.out.offsets[3 * s + 1] = o1; A.out.offsets[3 * s + 2] = o2;
                    }
                    // reuse pos[] as the warped world position (normalised offsets added to world coords: ref quirk)
                    ts.pos[r][0] += o0; ts.pos[r][1] += o1; ts.pos[r][2] += o2;
                }
            }
            __syncwarp();
        }
        if (FIELD && lane < kTRows) {
            wx = ts.pos[lane][0]; wy = ts.pos[lane][1]; wz = ts.pos[lane][2];
            float x = __fdiv_rn(__fsub_rn(wx, amin0), A.aabb_size[0]);
            float y = __fdiv_rn(__fsub_rn(wy, amin1), A.aabb_size[1]);
            float z = __fdiv_rn(__fsub_rn(wz, amin2), A.aabb_size[2]);
            const bool sel = (x > 0.f) && (x < 1.f) && (y > 0.f) && (y < 1.f) && (z > 0.f) && (z < 1.f);
            float4 o4 = make_float4(sel ? x : 0.f, sel ? y : 0.f, sel ? z : 0.f, ts.pos[lane][3]);
            *reinterpret_cast<float4 *>(sm.xs[b][warp * kTRows + lane]) = o4;
            if (SAVE && A.out.xs && row0 + lane < n)
                *reinterpret_cast<float4 *>(A.out.xs + 4 * (row0 + lane)) = make_float4(o4.x, o4.y, o4.z, sel ? 1.f : 0.f);
            ts.dirsel[b][lane][3] = sel ? 1.f : 0.f;
            if (warp == 0 && lane == 0) sm.tile_ctr[b] = 0;
        }
        if (FIELD) {
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.xs_full[b]);
        }
    };

    // it = -1 is the pipeline prologue: ONE inlined copy of deform_stage (two copies doubled the kernel's code size)
#pragma unroll 1
    for (int64_t it = -1; it < my_tiles; ++it) {
        if (it + 1 < my_tiles) deform_stage(it + 1);
        if (FIELD && it >= 0) {
            const int64_t tile = blockIdx.x + it * gridDim.x;
            const int b = (int)(it & 1);
            mbar_wait<100>(&sm.feat_full[b], (uint32_t)(it >> 1) & 1);
#pragma unroll 1
            for (int m = 0; m < kMT; ++m)
                field_mlp_tile<HEAD>(A, sm.field_w, &sm.feat[b][(warp * kTRows + m * 16) * kFeatStride],
                                     &ts.dirsel[b][m * 16], tile * NSB_TILE + warp * kTRows + m * 16, n, lane);
        }
    }
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
template <bool D, bool F, bool H>
static int launch_field_ws(const FieldArgs &A, cudaStream_t st) {
    const bool save = A.out.xs || A.out.deform_acts || A.out.deform_enc || A.out.corner_vals;
    return save ? launch_field_ws_<D, F, H, true>(A, st) : launch_field_ws_<D, F, H, false>(A, st);
}

template <bool D, bool F, bool H>
static int launch_field(const FieldArgs &A, cudaStream_t st) { return launch_field_ws<D, F, H>(A, st); }

}  // namespace nsb

using namespace nsb;

extern "C" size_t nsb_deform_packed_bytes(void) { return (size_t)kTbNumSlabs * kSlabBytes; }
extern "C" size_t nsb_field_packed_bytes(void) { return kFieldPackedU4 * sizeof(uint4); }

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
