// gsb_blend.cu -- per-16x16-tile front-to-back alpha blending.
// Replaces render.comp:30-99 (dispatch src/Renderer.cpp:654-677).  The reference makes every
// pixel thread gather idx + uv + conic + colour from global memory for every Gaussian of the
// tile's run (render.comp:62-65,87; README.md:87 lists staging as a TODO).  Here one CTA owns one
// tile and
//   * stages the run in batches of 256 compact 48-B records into shared memory once;
//   * while staging, each thread classifies its Gaussian against the eight 8x4-pixel blocks of
//     the tile (one block per warp): a block whose best-case exponent is below the shader's own
//     alpha < 1/255 cut (with an fp32 error margin) can never contribute, so that warp never
//     touches the record (bit-identical result: those pairs hit `continue` in render.comp:78);
//   * each warp compacts the batch with one ballot per 32 records and walks only its survivors,
//     reading each record as a shared-memory broadcast; the walk is branch-free per lane (the
//     shader's `continue`s and `break` are predicates on the four state updates).
// The per-pixel `break` (render.comp:83-85) becomes a per-lane done flag + warp / block votes.
//
// EXACT mode: -fmad=false, ops in render.comp's order, exp = the fixed IEEE sequence below
// (bit-identical to oracle exp-mode 1).  FAST mode: explicit FMA + ex2.approx.
#include "gsb_cull.cuh"
#include "gsb_internal.cuh"

namespace gsb {

namespace {

constexpr int BLEND_THREADS = 256;
constexpr unsigned FULL = 0xffffffffu;

// Bit-defined exp for x in [-87, 0]; mirrors gso_exp_shared() in oracle/gs_oracle.c op for op: Cody-Waite
// reduction, degree-5 Horner with the constant term 1, 2^n applied through the exponent bits (13 instructions).
__device__ __forceinline__ float exp_shared(float x) {
    x = fmaxf(x, -87.0f);
    const float t = __fmul_rn(x, 1.44269504088896341f);
    const float tm = __fadd_rn(t, 12582912.0f);  // low mantissa bits = rint(t) in two's complement
    const float n = __fsub_rn(tm, 12582912.0f);
    float r = __fmaf_rn(n, -0.693359375f, x);
    r = __fmaf_rn(n, 2.12194440e-4f, r);
    float p = __fmaf_rn(8.290082216262817e-3f, r, 4.1899293661117554e-2f);
    p = __fmaf_rn(p, r, 1.6667647659778595e-1f);
    p = __fmaf_rn(p, r, 4.9999138712882996e-1f);
    p = __fmaf_rn(p, r, 9.999997019767761e-1f);
    p = __fmaf_rn(p, r, 1.0f);
    return __uint_as_float(__float_as_uint(p) + (__float_as_uint(tm) << 23));
}

#if GSB_BLEND_TDONE
// same sequence without the clamp: only evaluated results with power in [cut, 0] are used (identity there)
__device__ __forceinline__ float exp_shared_inrange(float x) {
    const float t = __fmul_rn(x, 1.44269504088896341f);
    const float tm = __fadd_rn(t, 12582912.0f);
    const float n = __fsub_rn(tm, 12582912.0f);
    float r = __fmaf_rn(n, -0.693359375f, x);
    r = __fmaf_rn(n, 2.12194440e-4f, r);
    float p = __fmaf_rn(8.290082216262817e-3f, r, 4.1899293661117554e-2f);
    p = __fmaf_rn(p, r, 1.6667647659778595e-1f);
    p = __fmaf_rn(p, r, 4.9999138712882996e-1f);
    p = __fmaf_rn(p, r, 9.999997019767761e-1f);
    p = __fmaf_rn(p, r, 1.0f);
    return __uint_as_float(__float_as_uint(p) + (__float_as_uint(tm) << 23));
}
#endif

// shared-memory loads by 32-bit shared-window address (one LDS each, immediate offsets, no generic-pointer arithmetic)
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float2 lds_f2(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

__device__ __forceinline__ uint32_t unorm8(float v) {
    v = fminf(fmaxf(v, 0.0f), 1.0f);  // NaN -> 0
    return __float2uint_rn(v * 255.0f);
}

// Bit w set <=> warp w's 8x4 pixel block may receive a contribution from this Gaussian (gsb_cull.cuh).
__device__ __forceinline__ uint32_t block_mask(float ux, float uy, float A, float B, float C, float cut, float tile_x0, float tile_y0) {
    if (!(A > 0.0f) || !(C > 0.0f)) return 0xffu;  // not positive definite / NaN: never cull
    const float inv_a = __frcp_rn(A), inv_c = __frcp_rn(C);
    uint32_t mask = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) {
        const float x0 = tile_x0 + (float)((w & 1) * 8), y0 = tile_y0 + (float)((w >> 1) * 4);
        if (rect_may_contribute(ux, uy, A, B, C, inv_a, inv_c, x0, y0, 8.0f, 4.0f, cut)) mask |= 1u << w;
    }
    return mask;
}

// one staged record (48 B): three 16-B slots so the inner loop addresses it with one byte offset
struct __align__(16) StagedRec {
    float4 r0;  // uv.x uv.y -conic.x/2 -conic.y      (exact sign / power-of-two scalings: render.comp:66 becomes
    float4 r1;  // -conic.z/2 power_cut opacity r       ((-A/2 dx) dx + (-C/2 dy) dy) + ((-B) dx) dy, bit for bit)
    float4 r2;  // g b - -
};

#ifndef GSB_BLEND_CHECK
#define GSB_BLEND_CHECK 8  // records walked between two "is the whole warp done" votes (measured: 8 -> 0.673 ms, 16 -> 0.685, 32 -> 0.733)
#endif
#ifndef GSB_BLEND_TDONE
#define GSB_BLEND_TDONE 0  // 1: a finished pixel is T == 0 (no separate flag in the walk); needs GSB_BLEND_PREDICATED
#endif
#ifndef GSB_BLEND_PREDICATED
#define GSB_BLEND_PREDICATED 1
#endif
#ifndef GSB_BLEND_MIN_BLOCKS
#define GSB_BLEND_MIN_BLOCKS 5  // <= 51 registers, 5 CTAs per SM (measured: 0.873 ms; 4 CTAs 0.886, 6 CTAs 0.891, 80 registers 0.972)
#endif
template <int MODE>
__global__ void __launch_bounds__(BLEND_THREADS, GSB_BLEND_MIN_BLOCKS) k_blend(const __grid_constant__ BlendParams P) {
    __shared__ StagedRec s_rec[BLEND_THREADS];
    __shared__ uint32_t s_mask[BLEND_THREADS];
    __shared__ uint16_t s_list[BLEND_THREADS / 32][BLEND_THREADS];  // per warp: byte offsets of the records it must visit
    __shared__ uint32_t s_used;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tx = blockIdx.x % P.tiles_x;
    const uint32_t ty = P.tile_row_begin + blockIdx.x / P.tiles_x;
    uint2 range = P.ranges[ty * P.tiles_x + tx];  // render.comp:43-44; stored as (start, ~end), empty = all ones
    range.y = ~range.y;
    // warp w owns the 8x4 pixel block at (8 * (w & 1), 4 * (w >> 1)) of the tile
    const uint32_t px = tx * GSB_TILE + (warp & 1) * 8 + (lane & 7);
    const uint32_t py = ty * GSB_TILE + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = px < P.width && py < P.height;  // :37-39
    const float fx = (float)px, fy = (float)py;
    const float tile_x0 = (float)(tx * GSB_TILE), tile_y0 = (float)(ty * GSB_TILE);
    if (tid == 0) s_used = 0;
    __syncthreads();

    float T = 1.0f, c0 = 0.0f, c1 = 0.0f, c2 = 0.0f;
#if GSB_BLEND_TDONE
    if (!inside) T = 0.0f;
#define BLEND_DONE (T == 0.0f)
#else
    bool done = !inside;
#define BLEND_DONE done
#endif
    uint32_t used = 0;
    const uint32_t rec_sh = (uint32_t)__cvta_generic_to_shared(&s_rec[0]);  // < 64 KiB: fits the u16 list entries

    for (uint32_t base = range.x; base < range.y; base += BLEND_THREADS) {
        const uint32_t cnt = min((uint32_t)BLEND_THREADS, range.y - base);
        if ((uint32_t)tid < cnt) {
            const uint32_t cid = __ldg(P.vals + base + tid);
            const float4* rec = P.recs + (size_t)cid * 3;
            const float4 a = __ldg(rec), b = __ldg(rec + 1);
            const float cut = power_cut(b.y);
            s_rec[tid].r0 = make_float4(a.x, a.y, -0.5f * a.z, -a.w);
            s_rec[tid].r1 = make_float4(-0.5f * b.x, cut, b.y, b.z);
            s_rec[tid].r2 = make_float4(b.w, __ldg(reinterpret_cast<const float*>(rec + 2)), 0.f, 0.f);
            s_mask[tid] = block_mask(a.x, a.y, a.z, a.w, b.x, cut, tile_x0, tile_y0);
        }
        __syncthreads();
        if (!__all_sync(FULL, BLEND_DONE)) {
#if GSB_BLEND_TDONE
            const bool was_done = BLEND_DONE;
#endif
            // compact this warp's survivors of the batch into a list of shared-memory addresses (one ballot per 32 records)
            uint32_t n = 0;
            for (uint32_t c = 0; c < cnt; c += 32) {
                const bool mine = (c + lane < cnt) && ((s_mask[c + lane] >> warp) & 1u);
                const unsigned bits = __ballot_sync(FULL, mine);
                if (mine) s_list[warp][n + __popc(bits & ((1u << lane) - 1u))] = (uint16_t)(rec_sh + (c + lane) * sizeof(StagedRec));
                n += __popc(bits);
            }
            __syncwarp();
            // The walk is branch-free per lane: the shader's `continue`s (render.comp:68-70, :78-80) and `break` (:83-85)
            // become the predicate `ok`; only the every-16 "whole warp done" test is a (warp-uniform) branch.
            const uint32_t list_sh = (uint32_t)__cvta_generic_to_shared(&s_list[warp][0]);
            uint32_t fin_k = 0xffffffffu;
            for (uint32_t k0 = 0; k0 < n; k0 += GSB_BLEND_CHECK) {
                if (__all_sync(FULL, BLEND_DONE)) break;
                const uint32_t k1 = min(n, k0 + (uint32_t)GSB_BLEND_CHECK);
                for (uint32_t k = k0; k < k1; k++) {
                    const uint32_t addr = lds_u16(list_sh + 2u * k);
                    const float4 a = lds_f4(addr);
                    const float4 b = lds_f4(addr + 16u);
                    const float2 gb = lds_f2(addr + 32u);
                    const float dx = a.x - fx, dy = a.y - fy;  // :64
                    float power, alpha;
                    if (MODE == GSB_MODE_EXACT) {
                        power = ((a.z * dx) * dx + (b.x * dy) * dy) + (a.w * dx) * dy;  // :66 (pre-scaled conic)
#if GSB_BLEND_TDONE
                        alpha = fminf(0.99f, b.z * exp_shared_inrange(power));             // :77
#else
                        alpha = fminf(0.99f, b.z * exp_shared(power));                     // :77
#endif
                    } else {
                        power = fmaf(a.z * dx, dx, fmaf(b.x * dy, dy, (a.w * dx) * dy));
                        alpha = fminf(0.99f, b.z * __expf(power));
                    }
#if GSB_BLEND_TDONE
                    // A finished pixel carries T == 0 (a live one has T >= 1e-4): its test_T is 0, so it "finishes" again at
                    // every record it would touch and never accumulates; fin_k keeps the FIRST such record.  A NaN power
                    // fails both comparisons and is skipped, as in the oracle (its exp clamps NaN to exp(-87)).
                    bool ok = (power <= 0.0f && power >= b.y) && !(alpha < 1.0f / 255.0f);  // :68-70, :78-80
                    const float test_T = T * (1.0f - alpha);                               // :82
                    const bool fin = ok && test_T < 0.0001f;                               // :83-85
                    fin_k = fin ? min(fin_k, k) : fin_k;
                    ok = ok && !fin;
#else
                    // :68-70 and, below the Gaussian's cut, alpha < 1/255 (:78); a NaN power passes like in the shader
                    bool ok = !done && !(power > 0.0f || power < b.y) && !(alpha < 1.0f / 255.0f);  // :78-80
                    const float test_T = T * (1.0f - alpha);  // :82
                    if (ok && test_T < 0.0001f) {             // :83-85
                        done = true;
                        fin_k = k;
                        ok = false;
                    }
#endif
#if GSB_BLEND_PREDICATED
                    // select form: the products are computed unconditionally, only the four state updates are predicated
                    if (MODE == GSB_MODE_EXACT) {
                        const float n0 = c0 + (b.w * alpha) * T, n1 = c1 + (gb.x * alpha) * T, n2 = c2 + (gb.y * alpha) * T;  // :87
                        c0 = ok ? n0 : c0;
                        c1 = ok ? n1 : c1;
                        c2 = ok ? n2 : c2;
                    } else {
                        const float w = alpha * T;
                        c0 = ok ? fmaf(b.w, w, c0) : c0;
                        c1 = ok ? fmaf(gb.x, w, c1) : c1;
                        c2 = ok ? fmaf(gb.y, w, c2) : c2;
                    }
#if GSB_BLEND_TDONE
                    T = fin ? 0.0f : (ok ? test_T : T);  // :88, and the break
#else
                    T = ok ? test_T : T;  // :88
#endif
#else
                    if (ok) {
                        if (MODE == GSB_MODE_EXACT) {
                            c0 = c0 + (b.w * alpha) * T;  // :87
                            c1 = c1 + (gb.x * alpha) * T;
                            c2 = c2 + (gb.y * alpha) * T;
                        } else {
                            const float w = alpha * T;
                            c0 = fmaf(b.w, w, c0);
                            c1 = fmaf(gb.x, w, c1);
                            c2 = fmaf(gb.y, w, c2);
                        }
                        T = test_T;  // :88
                    }
#endif
                }
            }
#if GSB_BLEND_TDONE
            if (!was_done && fin_k != 0xffffffffu) used = base - range.x + (s_list[warp][fin_k] - rec_sh) / (uint32_t)sizeof(StagedRec) + 1;
            else if (!BLEND_DONE) used = base - range.x + cnt;
#else
            if (fin_k != 0xffffffffu) used = base - range.x + (s_list[warp][fin_k] - rec_sh) / (uint32_t)sizeof(StagedRec) + 1;
            else if (!done) used = base - range.x + cnt;
#endif
        }
        if (__syncthreads_and(BLEND_DONE)) break;
    }

    if (inside) {
        atomicMax(&s_used, used);
        const uint32_t row = py - P.tile_row_begin * GSB_TILE;
        unsigned char* dst = static_cast<unsigned char*>(P.out) + (size_t)row * P.row_pitch_bytes;
        if (P.format == GSB_FORMAT_RGBA32F) {
            reinterpret_cast<float4*>(dst)[px] = make_float4(c0, c1, c2, 1.0f);  // :98 vec4(c, 1)
        } else {
            const uint32_t r = unorm8(c0), g = unorm8(c1), b = unorm8(c2);
            const uint32_t v = (P.format == GSB_FORMAT_BGRA8) ? (b | (g << 8) | (r << 16) | 0xff000000u)
                                                              : (r | (g << 8) | (b << 16) | 0xff000000u);
            reinterpret_cast<uint32_t*>(dst)[px] = v;
        }
    }
    __syncthreads();
    if (tid == 0 && s_used) atomicAdd(&P.ctl->blend_consumed, (unsigned long long)s_used);
}

}  // namespace

cudaError_t launch_blend(const BlendParams& p, cudaStream_t s) {
    const uint32_t rows = p.tile_row_end - p.tile_row_begin;
    const uint32_t blocks = rows * p.tiles_x;
    if (blocks == 0) return cudaSuccess;
    if (p.mode == GSB_MODE_EXACT) k_blend<GSB_MODE_EXACT><<<blocks, BLEND_THREADS, 0, s>>>(p);
    else k_blend<GSB_MODE_FAST><<<blocks, BLEND_THREADS, 0, s>>>(p);
    return cudaGetLastError();
}

}  // namespace gsb
