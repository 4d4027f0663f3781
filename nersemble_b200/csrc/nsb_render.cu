// Ray marching, visibility filter and alpha compositing for sm_100a.
//
// Replaces (reference, relative to /root/reference/src/nersemble/nerfstudio/):
//   nerfacc OccGridEstimator.sampling -> traverse_grids   (model_components/nersemble_volumetric_sampler.py:95-108)
//   nerfacc render_visibility_from_density                (training pre-pass of sampling())
//   nerfacc pack_info / render_weight_from_density + RGB/Depth/Accumulation renderers
//                                                         (models/nersemble_instant_ngp.py:325-343)
//   DeformationRenderer                                   (model_components/nersemble_deformation_renderer.py:10-29)
//
// The marcher keeps nerfacc's arithmetic order in float32 WITHOUT fused multiply-add
// (__fmul_rn/__fadd_rn) so that sample positions are bit-exact to the CPU oracle
// (oracle/tp/nerfacc_cpu.py): t advances by repeated float32 addition, exactly like
// `t_last = t_last + dt` in nerfacc's traverse_grids kernel.
#include <algorithm>

#include "nsb_common.cuh"

namespace nsb {

// ---------------------------------------------------------------------------------------------
// fixed-stride marcher (BASELINE configs 1/2): n_per_ray steps from max(t_enter, near)
// ---------------------------------------------------------------------------------------------
__global__ void march_fixed_kernel(const float *__restrict__ origins, const float *__restrict__ directions,
                                   int64_t n_rays, const float *__restrict__ aabb, int n_per_ray, float step,
                                   float near_plane, float *__restrict__ t_starts, float *__restrict__ t_ends,
                                   int32_t *__restrict__ ray_indices, int64_t *__restrict__ packed_info) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    const float o[3] = {origins[3 * r], origins[3 * r + 1], origins[3 * r + 2]};
    const float d[3] = {directions[3 * r], directions[3 * r + 1], directions[3 * r + 2]};
    float tmin, tmax;
    const bool hit = ray_aabb(o, d, aabb, tmin, tmax);
    float t = hit ? fmaxf(tmin, near_plane) : near_plane;
    const int64_t base = r * n_per_ray;
    for (int k = 0; k < n_per_ray; ++k) {
        t_starts[base + k] = t;
        t = __fadd_rn(t, step);
        t_ends[base + k] = t;
        ray_indices[base + k] = (int32_t)r;
    }
    if (packed_info) {
        packed_info[2 * r] = base;
        packed_info[2 * r + 1] = n_per_ray;
    }
}

// ---------------------------------------------------------------------------------------------
// occupancy-grid marcher: nerfacc csrc/grid.cu traverse_grids (one thread per ray), two passes
// ---------------------------------------------------------------------------------------------
struct MarchArgs {
    nsb_march_args a;
};

__device__ __forceinline__ float calc_dt(float t, float cone_angle, float dt_min, float dt_max) {
    return fminf(fmaxf(__fmul_rn(t, cone_angle), dt_min), dt_max);
}

template <bool FILL>
__global__ void __launch_bounds__(128) march_occ_kernel(const __grid_constant__ MarchArgs M) {
    const nsb_march_args &a = M.a;
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.n_rays) return;
    const float o[3] = {a.origins[3 * r], a.origins[3 * r + 1], a.origins[3 * r + 2]};
    const float d[3] = {a.directions[3 * r], a.directions[3 * r + 1], a.directions[3 * r + 2]};
    const float near_plane = a.near_planes[r], far_plane = a.far_planes[r];
    const float step = a.step, cone = a.cone_angle;
    if (!(isfinite(o[0]) && isfinite(o[1]) && isfinite(o[2]) && isfinite(d[0]) && isfinite(d[1]) && isfinite(d[2])) ||
        (d[0] == 0.f && d[1] == 0.f && d[2] == 0.f)) {
        if (!FILL) a.counts[r] = 0;   // degenerate ray: no samples (never hang)
        return;
    }
    const int res = a.res, levels = a.levels;
    const float eps = 1e-6f;
    float inv_d[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) inv_d[k] = __fdiv_rn(1.0f, d[k]);

    // sorted aabb intersections over levels (tiny insertion sort, stable like torch.sort on ties)
    constexpr int kMaxLevels = 8;
    float tv[2 * kMaxLevels];
    int ti[2 * kMaxLevels];
    bool hits[kMaxLevels];
    for (int lv = 0; lv < levels; ++lv) {
        float t0, t1;
        bool h = ray_aabb(o, d, a.aabbs + 6 * lv, t0, t1);
        hits[lv] = h;
        tv[lv] = h ? t0 : INFINITY;
        tv[levels + lv] = h ? t1 : INFINITY;
        ti[lv] = lv;
        ti[levels + lv] = levels + lv;
    }
    for (int i = 1; i < 2 * levels; ++i) {
        float v = tv[i]; int id = ti[i]; int j = i - 1;
        while (j >= 0 && tv[j] > v) { tv[j + 1] = tv[j]; ti[j + 1] = ti[j]; --j; }
        tv[j + 1] = v; ti[j + 1] = id;
    }

    int64_t out = FILL ? a.offsets[r] : 0;
    int32_t count = 0;
    float t_last = near_plane;
    bool continuous = false;
    const float resf = (float)res;
    for (int i = 0; i < 2 * levels - 1; ++i) {
        const int level = ti[i] % levels;
        if (!hits[level]) continue;
        const float this_tmin = fmaxf(tv[i], near_plane);
        const float this_tmax = fminf(tv[i + 1], far_plane);
        if (!(this_tmin < this_tmax)) continue;
        if (!continuous) {
            if (step <= 0.0f) {
                t_last = this_tmin;
            } else {
                while (true) {
                    const float dt = calc_dt(t_last, cone, step, 1e10f);
                    if (__fadd_rn(t_last, __fmul_rn(dt, 0.5f)) >= this_tmin) break;
                    t_last = __fadd_rn(t_last, dt);
                }
            }
        }
        const float *ab = a.aabbs + 6 * level;
        float tdist[3], delta[3];
        int cur[3], fin[3], stp[3];
        const float ts_eps = __fadd_rn(this_tmin, eps), te_eps = __fsub_rn(this_tmax, eps);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float ext = __fsub_rn(ab[3 + k], ab[k]);
            const float voxel = __fdiv_rn(ext, resf);
            const float rs = __fadd_rn(o[k], __fmul_rn(d[k], ts_eps));
            const float re = __fadd_rn(o[k], __fmul_rn(d[k], te_eps));
            int c = (int)__fmul_rn(__fdiv_rn(__fsub_rn(rs, ab[k]), ext), resf);
            int f = (int)__fmul_rn(__fdiv_rn(__fsub_rn(re, ab[k]), ext), resf);
            c = min(max(c, 0), res - 1);
            f = min(max(f, 0), res - 1);
            cur[k] = c; fin[k] = f;
            const int start_index = c + (d[k] > 0.0f ? 1 : 0);
            const float tmax_k =
                __fadd_rn(__fmul_rn(__fadd_rn(ab[k], __fsub_rn(__fmul_rn((float)start_index, voxel), rs)), inv_d[k]), this_tmin);
            const float sf = d[k] == 0.0f ? 0.0f : (d[k] > 0.0f ? 1.0f : -1.0f);
            tdist[k] = d[k] == 0.0f ? this_tmax : tmax_k;
            delta[k] = d[k] == 0.0f ? this_tmax : __fmul_rn(__fmul_rn(voxel, inv_d[k]), sf);
            stp[k] = (int)sf;
        }
        const int ovf[3] = {fin[0] + stp[0], fin[1] + stp[1], fin[2] + stp[2]};
        for (int guard = 0; guard < 3 * res + 3; ++guard) {
            const float t_trav = fminf(fminf(tdist[0], fminf(tdist[1], tdist[2])), this_tmax);
            const size_t cell = ((size_t)level * res + cur[0]) * res * res + (size_t)cur[1] * res + cur[2];
            if (!a.binaries[cell]) {
                if (step <= 0.0f) {
                    t_last = t_trav;
                } else {
                    while (true) {
                        const float dt = calc_dt(t_last, cone, step, 1e10f);
                        if (__fadd_rn(t_last, __fmul_rn(dt, 0.5f)) >= t_trav) break;
                        t_last = __fadd_rn(t_last, dt);
                    }
                }
                continuous = false;
            } else {
                while (true) {
                    float t_next;
                    if (step <= 0.0f) {
                        t_next = t_trav;
                    } else {
                        const float dt = calc_dt(t_last, cone, step, 1e10f);
                        if (__fadd_rn(t_last, __fmul_rn(dt, 0.5f)) >= t_trav) break;
                        t_next = __fadd_rn(t_last, dt);
                    }
                    if (FILL) {
                        a.t_starts[out] = t_last;
                        a.t_ends[out] = t_next;
                        a.ray_indices[out] = (int32_t)r;
                        ++out;
                    }
                    ++count;
                    continuous = true;
                    t_last = t_next;
                    if (t_next >= t_trav) break;
                }
            }
            int ax;
            if (tdist[0] < tdist[1] && tdist[0] < tdist[2]) ax = 0;
            else if (tdist[1] < tdist[2]) ax = 1;
            else ax = 2;
            // (dynamic register-array indexing avoided)
            bool done;
            if (ax == 0) { cur[0] += stp[0]; tdist[0] = __fadd_rn(tdist[0], delta[0]); done = cur[0] == ovf[0]; }
            else if (ax == 1) { cur[1] += stp[1]; tdist[1] = __fadd_rn(tdist[1], delta[1]); done = cur[1] == ovf[1]; }
            else { cur[2] += stp[2]; tdist[2] = __fadd_rn(tdist[2], delta[2]); done = cur[2] == ovf[2]; }
            if (done) break;
        }
    }
    if (!FILL) a.counts[r] = count;
}

// ---------------------------------------------------------------------------------------------
// visibility mask (nerfacc render_visibility_from_density): one warp per ray
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_incl_scan(float v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += n;
    }
    return v;
}

__global__ void __launch_bounds__(256) visibility_kernel(const int64_t *__restrict__ packed_info, int64_t n_rays,
                                                         const float *__restrict__ t_starts,
                                                         const float *__restrict__ t_ends,
                                                         const float *__restrict__ sigma, float early_stop_eps,
                                                         float alpha_thre, uint8_t *__restrict__ mask,
                                                         int32_t *__restrict__ kept) {
    const int lane = threadIdx.x & 31;
    const int64_t ray = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (ray >= n_rays) return;
    const int64_t start = packed_info[2 * ray], cnt = packed_info[2 * ray + 1];
    float carry = 0.f;
    int32_t k = 0;
    for (int64_t b = 0; b < cnt; b += 32) {
        const int64_t i = b + lane;
        float sd = 0.f;
        if (i < cnt) sd = sigma[start + i] * (t_ends[start + i] - t_starts[start + i]);
        const float incl = warp_incl_scan(sd, lane);
        const float excl = carry + (incl - sd);
        const float T = expf(-excl);
        const float alpha = 1.0f - expf(-sd);
        bool vis = T >= early_stop_eps;
        if (alpha_thre > 0.0f) vis = vis && (alpha >= alpha_thre);
        if (i < cnt) mask[start + i] = vis ? 1 : 0;
        k += __popc(__ballot_sync(0xffffffffu, vis && i < cnt));
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0 && kept) kept[ray] = k;
}

// ---------------------------------------------------------------------------------------------
// compositing: one warp per ray
// ---------------------------------------------------------------------------------------------
struct CompArgs {
    nsb_composite_args a;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(256) composite_kernel(const __grid_constant__ CompArgs C) {
    const nsb_composite_args &a = C.a;
    const int lane = threadIdx.x & 31;
    const int64_t ray = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (ray >= a.n_rays) return;
    const int64_t start = a.packed_info[2 * ray], cnt = a.packed_info[2 * ray + 1];
    float carry = 0.f;
    float acc = 0.f, cr = 0.f, cg = 0.f, cb = 0.f, dep = 0.f, d0 = 0.f, d1 = 0.f, d2 = 0.f;
    float mn = INFINITY, mx = -INFINITY;
    for (int64_t b = 0; b < cnt; b += 32) {
        const int64_t i = b + lane;
        const bool ok = i < cnt;
        const int64_t s = start + (ok ? i : 0);
        const float ts = ok ? a.t_starts[s] : 0.f, te = ok ? a.t_ends[s] : 0.f;
        const float sd = ok ? a.sigma[s] * (te - ts) : 0.f;
        // render_weight_from_density: T = exp(-exclusive_sum(sigma*dt)), alpha = 1 - exp(-sigma*dt)
        const float incl = warp_incl_scan(sd, lane);
        const float excl = carry + (incl - sd);
        const float w = ok ? expf(-excl) * (1.0f - expf(-sd)) : 0.f;
        carry += __shfl_sync(0xffffffffu, incl, 31);
        if (ok) {
            if (a.out_weights) a.out_weights[s] = w;
            float r = a.rgb[3 * s], g = a.rgb[3 * s + 1], bl = a.rgb[3 * s + 2];
            if (!a.training) {  // RGBRenderer eval path: nan_to_num before compositing
                r = isnan(r) ? 0.f : (isinf(r) ? (r > 0 ? 3.4028234663852886e38f : -3.4028234663852886e38f) : r);
                g = isnan(g) ? 0.f : (isinf(g) ? (g > 0 ? 3.4028234663852886e38f : -3.4028234663852886e38f) : g);
                bl = isnan(bl) ? 0.f : (isinf(bl) ? (bl > 0 ? 3.4028234663852886e38f : -3.4028234663852886e38f) : bl);
            }
            const float mid = (ts + te) / 2.0f;
            acc += w; cr += w * r; cg += w * g; cb += w * bl; dep += w * mid;
            mn = fminf(mn, mid); mx = fmaxf(mx, mid);
            if (a.offsets) { d0 += w * a.offsets[3 * s]; d1 += w * a.offsets[3 * s + 1]; d2 += w * a.offsets[3 * s + 2]; }
        }
    }
    acc = warp_sum(acc); cr = warp_sum(cr); cg = warp_sum(cg); cb = warp_sum(cb); dep = warp_sum(dep);
    if (a.offsets) { d0 = warp_sum(d0); d1 = warp_sum(d1); d2 = warp_sum(d2); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lane == 0) {
        // white background: rgb + 1*(1-acc); eval clamps to [0,1]
        float r = cr + (1.0f - acc), g = cg + (1.0f - acc), bl = cb + (1.0f - acc);
        if (!a.training) { r = fminf(fmaxf(r, 0.f), 1.f); g = fminf(fmaxf(g, 0.f), 1.f); bl = fminf(fmaxf(bl, 0.f), 1.f); }
        a.out_rgb[3 * ray] = r; a.out_rgb[3 * ray + 1] = g; a.out_rgb[3 * ray + 2] = bl;
        a.out_acc[ray] = acc;
        a.out_depth[ray] = dep / (acc + 1e-10f);   // clipped by depth_clip_kernel
        if (a.out_deform) { a.out_deform[3 * ray] = d0; a.out_deform[3 * ray + 1] = d1; a.out_deform[3 * ray + 2] = d2; }
        if (cnt > 0) {
            atomicMin(&a.workspace[0], float_to_ordered(mn));
            atomicMax(&a.workspace[1], float_to_ordered(mx));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// compositing backward (training mode): one warp per ray, two passes over the ray's samples.
//   w_i = T_i a_i, T_i = exp(-sum_{j<i} sd_j), a_i = 1-exp(-sd_i), sd = sigma*dt
//   dL/dsd_i = G_i T_i (1-a_i) - sum_{j>i} G_j w_j,  G_i = dL/dw_i (all consumers of w_i)
// ---------------------------------------------------------------------------------------------
struct CompBwdArgs {
    nsb_composite_bwd_args a;
};

__global__ void __launch_bounds__(256) composite_bwd_kernel(const __grid_constant__ CompBwdArgs C) {
    const nsb_composite_bwd_args &a = C.a;
    const int lane = threadIdx.x & 31;
    const int64_t ray = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (ray >= a.n_rays) return;
    const int64_t start = a.packed_info[2 * ray], cnt = a.packed_info[2 * ray + 1];
    if (cnt == 0) return;
    const float gr = a.d_out_rgb[3 * ray], gg = a.d_out_rgb[3 * ray + 1], gb = a.d_out_rgb[3 * ray + 2];
    const float gacc = a.d_out_acc ? a.d_out_acc[ray] : 0.f;
    float gdep = a.d_out_depth ? a.d_out_depth[ray] : 0.f;
    // pass 1: forward recompute of acc, N = sum w*mid, and of total = sum_j G_j w_j (needs acc, N first -> split)
    float carry = 0.f, acc = 0.f, N = 0.f;
    for (int64_t b = 0; b < cnt; b += 32) {
        const int64_t i = b + lane;
        const bool ok = i < cnt;
        const int64_t s = start + (ok ? i : 0);
        const float ts = ok ? a.t_starts[s] : 0.f, te = ok ? a.t_ends[s] : 0.f;
        const float sd = ok ? a.sigma[s] * (te - ts) : 0.f;
        const float incl = warp_incl_scan(sd, lane);
        const float w = ok ? expf(-(carry + (incl - sd))) * (1.0f - expf(-sd)) : 0.f;
        carry += __shfl_sync(0xffffffffu, incl, 31);
        acc += w;
        N += w * ((ts + te) / 2.0f);
    }
    acc = warp_sum(acc);
    N = warp_sum(N);
    const float inv = 1.0f / (acc + 1e-10f);
    const float draw = N * inv;
    if (a.workspace[0] != 0xffffffffu) {   // torch.clip backward: gradient only inside [min, max]
        const float lo = ordered_to_float(a.workspace[0]), hi = ordered_to_float(a.workspace[1]);
        if (!(draw >= lo && draw <= hi)) gdep = 0.f;
    }
    const float gbg = -(gr + gg + gb);     // white background: rgb + (1 - acc)
    // pass 2a: total = sum_j G_j w_j
    float total = 0.f;
    carry = 0.f;
    for (int64_t b = 0; b < cnt; b += 32) {
        const int64_t i = b + lane;
        const bool ok = i < cnt;
        const int64_t s = start + (ok ? i : 0);
        const float ts = ok ? a.t_starts[s] : 0.f, te = ok ? a.t_ends[s] : 0.f;
        const float sd = ok ? a.sigma[s] * (te - ts) : 0.f;
        const float incl = warp_incl_scan(sd, lane);
        const float w = ok ? expf(-(carry + (incl - sd))) * (1.0f - expf(-sd)) : 0.f;
        carry += __shfl_sync(0xffffffffu, incl, 31);
        if (ok) {
            const float G = (a.d_weights ? a.d_weights[s] : 0.f) + gr * a.rgb[3 * s] + gg * a.rgb[3 * s + 1] +
                            gb * a.rgb[3 * s + 2] + gbg + gacc + gdep * (((ts + te) / 2.0f) - draw) * inv;
            total += G * w;
        }
    }
    total = warp_sum(total);
    // pass 2b: gradients
    carry = 0.f;
    float pcarry = 0.f;
    for (int64_t b = 0; b < cnt; b += 32) {
        const int64_t i = b + lane;
        const bool ok = i < cnt;
        const int64_t s = start + (ok ? i : 0);
        const float ts = ok ? a.t_starts[s] : 0.f, te = ok ? a.t_ends[s] : 0.f;
        const float dt = te - ts;
        const float sd = ok ? a.sigma[s] * dt : 0.f;
        const float incl = warp_incl_scan(sd, lane);
        const float T = expf(-(carry + (incl - sd)));
        const float e = expf(-sd);
        const float w = ok ? T * (1.0f - e) : 0.f;
        carry += __shfl_sync(0xffffffffu, incl, 31);
        float G = 0.f;
        if (ok)
            G = (a.d_weights ? a.d_weights[s] : 0.f) + gr * a.rgb[3 * s] + gg * a.rgb[3 * s + 1] + gb * a.rgb[3 * s + 2] +
                gbg + gacc + gdep * (((ts + te) / 2.0f) - draw) * inv;
        const float gw = G * w;
        const float pin = warp_incl_scan(gw, lane);          // inclusive prefix of G_j w_j
        const float suffix = total - (pcarry + pin);          // sum_{j>i} G_j w_j
        pcarry += __shfl_sync(0xffffffffu, pin, 31);
        if (ok) {
            a.d_sigma[s] = (G * T * e - suffix) * dt;
            a.d_rgb[3 * s] = gr * w; a.d_rgb[3 * s + 1] = gg * w; a.d_rgb[3 * s + 2] = gb * w;
        }
    }
}

__global__ void depth_init_kernel(uint32_t *ws) {
    ws[0] = 0xffffffffu;
    ws[1] = 0u;
}

// DepthRenderer('expected'): depth = clip(depth, steps.min(), steps.max()) over ALL samples of the batch
__global__ void depth_clip_kernel(float *depth, int64_t n_rays, const uint32_t *ws) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    if (ws[0] == 0xffffffffu) return;  // no samples at all
    const float lo = ordered_to_float(ws[0]), hi = ordered_to_float(ws[1]);
    depth[r] = fminf(fmaxf(depth[r], lo), hi);
}

// ---------------------------------------------------------------------------------------------
// occupancy-grid EMA update (nerfacc OccGridEstimator._update): see nsb.h.  occ values are >= 0 (densities), so the
// int ordering of their bit patterns is the float ordering and atomicMax(int) resolves duplicate cells.
// ---------------------------------------------------------------------------------------------
__global__ void occ_decay_kernel(const float *__restrict__ occs, const int64_t *__restrict__ ids, int64_t n, float decay,
                                 float *__restrict__ tmp) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) tmp[ids[i]] = occs[ids[i]] * decay;      // duplicates store the same value
}
__global__ void occ_max_kernel(const int64_t *__restrict__ ids, const float *__restrict__ occ_new, int64_t n,
                               float *__restrict__ tmp) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = occ_new[i];
    if (v >= 0.f) atomicMax(reinterpret_cast<int *>(tmp + ids[i]), __float_as_int(v));
    else if (isnan(v)) tmp[ids[i]] = v;                  // torch.maximum propagates NaN
}
__global__ void occ_commit_kernel(float *__restrict__ occs, const int64_t *__restrict__ ids, int64_t n,
                                  const float *__restrict__ tmp) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) occs[ids[i]] = tmp[ids[i]];
}
__global__ void __launch_bounds__(256) occ_mean_kernel(const float *__restrict__ occs, int64_t n_cells, double *__restrict__ acc) {
    double s = 0.0, c = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_cells; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = occs[i];
        if (v >= 0.f) { s += (double)v; c += 1.0; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); c += __shfl_xor_sync(0xffffffffu, c, o); }
    if ((threadIdx.x & 31) == 0 && c > 0.0) { atomicAdd(acc, s); atomicAdd(acc + 1, c); }
}
__global__ void occ_binarise_kernel(const float *__restrict__ occs, int64_t n_cells, const double *__restrict__ acc,
                                    float occ_thre, uint8_t *__restrict__ binaries) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cells) return;
    const float mean = (float)(acc[0] / acc[1]);        // 0/0 = NaN like torch.mean of an empty selection
    const float thre = fminf(mean, occ_thre);            // torch.clamp(mean, max=occ_thre); NaN mean -> occ_thre
    binaries[i] = occs[i] > (isnan(mean) ? occ_thre : thre) ? 1 : 0;
}

}  // namespace nsb

using namespace nsb;

extern "C" int nsb_march_fixed(const float *origins, const float *directions, int64_t n_rays, const float *aabb6,
                               int32_t n_per_ray, float step, float near_plane, float *t_starts, float *t_ends,
                               int32_t *ray_indices, int64_t *packed_info, void *stream) {
    if (!origins || !directions || !aabb6 || !t_starts || !t_ends || !ray_indices) {
        set_error("nsb_march_fixed: null argument");
        return 1;
    }
    if (n_rays <= 0 || n_per_ray <= 0) return 0;
    const int threads = 128;
    const int blocks = (int)((n_rays + threads - 1) / threads);
    march_fixed_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(origins, directions, n_rays, aabb6, n_per_ray, step,
                                                                   near_plane, t_starts, t_ends, ray_indices, packed_info);
    return check_launch("march_fixed_kernel");
}

extern "C" int nsb_march_occupancy(const nsb_march_args *args, void *stream) {
    if (!args || !args->origins || !args->directions || !args->near_planes || !args->far_planes || !args->binaries ||
        !args->aabbs) {
        set_error("nsb_march_occupancy: null argument");
        return 1;
    }
    if (args->levels < 1 || args->levels > 8) { set_error("nsb_march_occupancy: levels must be in [1,8]"); return 1; }
    if (args->n_rays <= 0) return 0;
    MarchArgs M; M.a = *args;
    const int threads = 128;
    const int blocks = (int)((args->n_rays + threads - 1) / threads);
    if (args->t_starts == nullptr) {
        if (!args->counts) { set_error("nsb_march_occupancy: pass 1 needs counts"); return 1; }
        march_occ_kernel<false><<<blocks, threads, 0, (cudaStream_t)stream>>>(M);
    } else {
        if (!args->offsets || !args->t_ends || !args->ray_indices) { set_error("nsb_march_occupancy: pass 2 needs offsets/t_ends/ray_indices"); return 1; }
        march_occ_kernel<true><<<blocks, threads, 0, (cudaStream_t)stream>>>(M);
    }
    return check_launch("march_occ_kernel");
}

extern "C" int nsb_visibility_mask(const int64_t *packed_info, int64_t n_rays, const float *t_starts,
                                   const float *t_ends, const float *sigma, float early_stop_eps, float alpha_thre,
                                   uint8_t *mask, int32_t *kept_counts, void *stream) {
    if (!packed_info || !t_starts || !t_ends || !sigma || !mask) { set_error("nsb_visibility_mask: null argument"); return 1; }
    if (n_rays <= 0) return 0;
    const int blocks = (int)((n_rays + 7) / 8);
    visibility_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(packed_info, n_rays, t_starts, t_ends, sigma,
                                                              early_stop_eps, alpha_thre, mask, kept_counts);
    return check_launch("visibility_kernel");
}

extern "C" int nsb_composite_backward(const nsb_composite_bwd_args *args, void *stream) {
    if (!args || !args->packed_info || !args->t_starts || !args->t_ends || !args->sigma || !args->rgb || !args->d_out_rgb ||
        !args->workspace || !args->d_sigma || !args->d_rgb) {
        set_error("nsb_composite_backward: null argument");
        return 1;
    }
    if (args->n_rays <= 0) return 0;
    CompBwdArgs C; C.a = *args;
    const int blocks = (int)((args->n_rays + 7) / 8);
    composite_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(C);
    return check_launch("composite_bwd_kernel");
}

extern "C" int nsb_composite_forward(const nsb_composite_args *args, void *stream) {
    if (!args || !args->packed_info || !args->t_starts || !args->t_ends || !args->sigma || !args->rgb || !args->out_rgb ||
        !args->out_acc || !args->out_depth || !args->workspace) {
        set_error("nsb_composite_forward: null argument");
        return 1;
    }
    if (args->n_rays <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    CompArgs C; C.a = *args;
    depth_init_kernel<<<1, 1, 0, st>>>(args->workspace);
    const int blocks = (int)((args->n_rays + 7) / 8);
    composite_kernel<<<blocks, 256, 0, st>>>(C);
    depth_clip_kernel<<<(int)((args->n_rays + 255) / 256), 256, 0, st>>>(args->out_depth, args->n_rays, args->workspace);
    return check_launch("composite_kernel");
}

extern "C" size_t nsb_occ_update_scratch_bytes(int64_t n_cells) { return (size_t)n_cells * sizeof(float) + 2 * sizeof(double) + 8; }

extern "C" int nsb_occ_update(float *occs, uint8_t *binaries, int64_t n_cells, const int64_t *cell_ids, const float *occ_new,
                              int64_t n, float ema_decay, float occ_thre, void *scratch, void *stream) {
    if (!occs || !binaries || !scratch || n_cells <= 0 || (n > 0 && (!cell_ids || !occ_new))) {
        set_error("nsb_occ_update: null argument");
        return 1;
    }
    cudaStream_t st = (cudaStream_t)stream;
    double *acc = reinterpret_cast<double *>(scratch);                                   // 8-byte aligned head
    float *tmp = reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(scratch) + 2 * sizeof(double));
    const int T = 256;
    if (n > 0) {
        const int blocks = (int)((n + T - 1) / T);
        occ_decay_kernel<<<blocks, T, 0, st>>>(occs, cell_ids, n, ema_decay, tmp);
        occ_max_kernel<<<blocks, T, 0, st>>>(cell_ids, occ_new, n, tmp);
        occ_commit_kernel<<<blocks, T, 0, st>>>(occs, cell_ids, n, tmp);
    }
    cudaMemsetAsync(acc, 0, 2 * sizeof(double), st);
    occ_mean_kernel<<<296, 256, 0, st>>>(occs, n_cells, acc);
    occ_binarise_kernel<<<(int)((n_cells + T - 1) / T), T, 0, st>>>(occs, n_cells, acc, occ_thre, binaries);
    return check_launch("nsb_occ_update");
}
