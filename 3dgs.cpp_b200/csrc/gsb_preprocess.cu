// gsb_preprocess.cu -- ONE kernel for the reference's first three stages:
//   preprocess.comp:115-182   (project, cull, EWA cov2d, conic, radius, tile AABB, degree-3 SH colour)
//   prefix_sum.comp:32-58 x (log2 N + 1) dispatches (Renderer.cpp:497-526)  -> single-pass decoupled look-back scan
//   preprocess_sort.comp:31-60 (emit (tile<<32 | depth) keys + payload at the scan offset)
// and it removes the mid-frame fence + host read of M (Renderer.cpp:391,538): M stays in HBM.
//
// B200 mapping: 256 Gaussians per CTA, one per thread; coalesced LDG.128 of the SoA position /
// covariance arrays; the 192-B SH record is fetched by cull survivors only; survivors are compacted
// (dense 48-B blend records => dense, L2-friendly gathers in the blend); key emission is a
// block-cooperative expansion so every 8-B key / 4-B payload store is coalesced, in exactly the
// reference's order (Gaussian-major, x-outer, y-inner).
//
// Arithmetic: compiled with -fmad=false; every fp32 operation is a single IEEE op in the order
// of the GLSL source, so results are value-identical to oracle/gs_oracle.c (bit-exact parity).
#include "gsb_internal.cuh"

namespace gsb {

namespace {

constexpr int PRE_THREADS = 256;
constexpr unsigned FULL = 0xffffffffu;

// look-back status word: [63:62] flag, [61:32] survivors, [31:0] tile instances (saturating)
constexpr unsigned long long ST_AGG = 1ull << 62;
constexpr unsigned long long ST_PREFIX = 2ull << 62;
constexpr unsigned long long ST_FLAGS = 3ull << 62;

__device__ __forceinline__ unsigned long long st_pack(unsigned long long flag, uint32_t surv, unsigned long long tiles) {
    if (tiles > 0xffffffffull) tiles = 0xffffffffull;
    return flag | ((unsigned long long)surv << 32) | tiles;
}
__device__ __forceinline__ unsigned long long ld_status(const unsigned long long* p) {
    return *reinterpret_cast<const volatile unsigned long long*>(p);
}
__device__ __forceinline__ void st_status(unsigned long long* p, unsigned long long v) {
    *reinterpret_cast<volatile unsigned long long*>(p) = v;
}

// common.glsl:16-33
__device__ constexpr float SH_C0 = 0.28209479177387814f;
__device__ constexpr float SH_C1 = 0.4886025119029199f;
__device__ constexpr float SH_C2_0 = 1.0925484305920792f, SH_C2_1 = -1.0925484305920792f,
                           SH_C2_2 = 0.31539156525252005f, SH_C2_3 = -1.0925484305920792f,
                           SH_C2_4 = 0.5462742152960396f;
__device__ constexpr float SH_C3_0 = -0.5900435899266435f, SH_C3_1 = 2.890611442640554f,
                           SH_C3_2 = -0.4570457994644658f, SH_C3_3 = 0.3731763325901154f,
                           SH_C3_4 = -0.4570457994644658f, SH_C3_5 = 1.445305721320277f,
                           SH_C3_6 = -0.5900435899266435f;

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// preprocess.comp:73-108 compute_sh(); sh = 48 floats RGB-interleaved, as 12 float4.
__device__ __forceinline__ void compute_sh(const float4* __restrict__ sh4, float px, float py, float pz,
                                           const float* cam, float& r, float& g, float& b) {
    float f[48];
#pragma unroll
    for (int k = 0; k < 12; k++) {
        const float4 t = __ldg(sh4 + k);
        f[4 * k + 0] = t.x;
        f[4 * k + 1] = t.y;
        f[4 * k + 2] = t.z;
        f[4 * k + 3] = t.w;
    }
    const float dx = px - cam[0], dy = py - cam[1], dz = pz - cam[2];
    const float len = sqrtf((dx * dx + dy * dy) + dz * dz);
    const float x = dx / len, y = dy / len, z = dz / len;
    const float xx = x * x, yy = y * y;
    const float w6 = ((2.0f * z) * z - xx) - yy;
    const float w8 = xx - yy;
    const float w9 = (3.0f * x) * x - yy;
    const float w11 = ((4.0f * z) * z - xx) - yy;
    const float w12 = ((2.0f * z) * z - (3.0f * x) * x) - (3.0f * y) * y;
    const float w15 = xx - (3.0f * y) * y;
    float c[3];
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
#define SH(k) f[(k) * 3 + ch]
        float a = SH_C0 * SH(0);
        a = a - (SH_C1 * SH(1)) * y;
        a = a + (SH_C1 * SH(2)) * z;
        a = a - (SH_C1 * SH(3)) * x;
        a = a + ((SH_C2_0 * SH(4)) * x) * y;
        a = a + ((SH_C2_1 * SH(5)) * y) * z;
        a = a + (SH_C2_2 * SH(6)) * w6;
        a = a + ((SH_C2_3 * SH(7)) * z) * x;
        a = a + (SH_C2_4 * SH(8)) * w8;
        a = a + ((SH_C3_0 * SH(9)) * w9) * y;
        a = a + (((SH_C3_1 * SH(10)) * x) * y) * z;
        a = a + ((SH_C3_2 * SH(11)) * w11) * y;
        a = a + ((SH_C3_3 * SH(12)) * z) * w12;
        a = a + ((SH_C3_4 * SH(13)) * x) * w11;
        a = a + ((SH_C3_5 * SH(14)) * w8) * z;
        a = a + ((SH_C3_6 * SH(15)) * x) * w15;
        c[ch] = a + 0.5f;
#undef SH
    }
    r = c[0] < 0.0f ? 0.0f : c[0];  // :102-104 only the red channel is clamped
    g = c[1];
    b = c[2];
}

template <bool DEBUG>
__global__ void __launch_bounds__(PRE_THREADS) k_preprocess(const __grid_constant__ PreprocessParams P) {
    __shared__ uint32_t s_chunk;
    __shared__ uint32_t s_wsurv[PRE_THREADS / 32], s_wnt[PRE_THREADS / 32];
    __shared__ uint32_t s_base_surv;
    __shared__ unsigned long long s_base_tiles;
    __shared__ uint32_t s_off[PRE_THREADS + 1];
    __shared__ uint4 s_info[PRE_THREADS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_chunk = atomicAdd(&P.ctl->pre_ticket, 1u);
    __syncthreads();
    const uint32_t chunk = s_chunk;
    const uint32_t num_chunks = (P.n + PRE_THREADS - 1) / PRE_THREADS;
    const uint32_t i = chunk * PRE_THREADS + tid;

    const gsb_uniforms& U = P.ubo;
    const int W = (int)U.width, H = (int)U.height;
    const int tiles_x = (int)((U.width + GSB_TILE - 1) / GSB_TILE);
    const int tiles_y = (int)((U.height + GSB_TILE - 1) / GSB_TILE);

    bool surv = false;
    uint32_t nt = 0;
    float uvx = 0.f, uvy = 0.f, conx = 0.f, cony = 0.f, conz = 0.f, opac = 0.f, depth = 0.f, radii = 0.f;
    float px = 0.f, py = 0.f, pz = 0.f;
    int bx0 = 0, by0 = 0, bx1 = 0, by1 = 0;

    if (i < P.n) {
        const float4 po = __ldg(P.pos_op + i);
        px = po.x;
        py = po.y;
        pz = po.z;
        opac = po.w;
        const float* pm = U.proj_mat;
        const float* vm = U.view_mat;
        // :130-134  mat4 * vec4(p, 1): ((m0*x + m1*y) + m2*z) + m3*1
        const float hx = ((pm[0] * px + pm[4] * py) + pm[8] * pz) + pm[12];
        const float hy = ((pm[1] * px + pm[5] * py) + pm[9] * pz) + pm[13];
        const float hw = ((pm[3] * px + pm[7] * py) + pm[11] * pz) + pm[15];
        const float p_w = 1.0f / hw;
        const float ndcx = hx * p_w, ndcy = hy * p_w;
        const float vx = ((vm[0] * px + vm[4] * py) + vm[8] * pz) + vm[12];
        const float vy = ((vm[1] * px + vm[5] * py) + vm[9] * pz) + vm[13];
        const float vz = ((vm[2] * px + vm[6] * py) + vm[10] * pz) + vm[14];
        if (!(vz <= 0.2f)) {  // :135 (NaN is not culled by the shader's test either)
            // get_projection_jacobian_approx :34-50
            const float limx = 1.3f * U.tan_fovx, limy = 1.3f * U.tan_fovy;
            const float txtz = vx / vz, tytz = vy / vz;
            const float tx = fminf(limx, fmaxf(-limx, txtz)) * vz;
            const float ty = fminf(limy, fmaxf(-limy, tytz)) * vz;
            const float focal_x = (float)U.width / (2.0f * U.tan_fovx);
            const float focal_y = (float)U.height / (2.0f * U.tan_fovy);
            const float ja = focal_x / vz, jb = focal_y / vz;
            const float g0 = -(focal_x * tx) / (vz * vz), g1 = -(focal_y * ty) / (vz * vz);
            // T = transpose(mat3(view)) * J  (:55,:61): T[0][r] = V[r][0]*ja + V[r][2]*g0, T[1][r] = V[r][1]*jb + V[r][2]*g1
            float T0[3], T1[3];
#pragma unroll
            for (int r = 0; r < 3; r++) {
                T0[r] = vm[r * 4 + 0] * ja + vm[r * 4 + 2] * g0;
                T1[r] = vm[r * 4 + 1] * jb + vm[r * 4 + 2] * g1;
            }
            const float4 ca = __ldg(P.cov_a + i);
            const float2 cb = __ldg(P.cov_b + i);
            // Sigma columns (:56-60): S[0]=(c0,c1,c2) S[1]=(c1,c3,c4) S[2]=(c2,c4,c5)
            const float S[3][3] = {{ca.x, ca.y, ca.z}, {ca.y, ca.w, cb.x}, {ca.z, cb.x, cb.y}};
            // tmp = transpose(T) * Sigma: tmp[k][r] = (T_r[0]*S[k][0] + T_r[1]*S[k][1]) + T_r[2]*S[k][2]
            float tm0[3], tm1[3];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                tm0[k] = (T0[0] * S[k][0] + T0[1] * S[k][1]) + T0[2] * S[k][2];
                tm1[k] = (T1[0] * S[k][0] + T1[1] * S[k][1]) + T1[2] * S[k][2];
            }
            // cov2d = tmp * T (:62): c[c][r] = (tmp[0][r]*T_c[0] + tmp[1][r]*T_c[1]) + tmp[2][r]*T_c[2]
            const float c00 = (tm0[0] * T0[0] + tm0[1] * T0[1]) + tm0[2] * T0[2];
            const float c01 = (tm1[0] * T0[0] + tm1[1] * T0[1]) + tm1[2] * T0[2];  // [0][1]: col 0, row 1
            const float c10 = (tm0[0] * T1[0] + tm0[1] * T1[1]) + tm0[2] * T1[2];  // [1][0]
            const float c11 = (tm1[0] * T1[0] + tm1[1] * T1[1]) + tm1[2] * T1[2];
            const float m00 = c00 + 0.3f, m01 = c01, m10 = c10, m11 = c11 + 0.3f;  // :63-65
            const float det = m00 * m11 - m10 * m01;                               // :138
            if (!(det <= 0.0f)) {                                                  // :139-141
                const float ood = 1.0f / det;
                conx = m11 * ood;
                cony = -m01 * ood;
                conz = m00 * ood;  // :142-143
                const float mid = 0.5f * (m00 + m11);
                const float sq = sqrtf(fmaxf(0.1f, mid * mid - det));
                const float lambda = fmaxf(mid + sq, mid - sq);
                radii = ceilf(3.0f * sqrtf(lambda));                     // :146-151
                uvx = ((ndcx + 1.0f) * (float)W - 1.0f) * 0.5f;          // :157 ndc2Pix
                uvy = ((ndcy + 1.0f) * (float)H - 1.0f) * 0.5f;
                // :159-164; cvt.rzi.s32.f32 saturates where GLSL int() is undefined
                bx0 = clampi(__float2int_rz((uvx - radii) / 16.0f), 0, tiles_x);
                by0 = clampi(__float2int_rz((uvy - radii) / 16.0f), 0, tiles_y);
                bx1 = clampi(__float2int_rz((((uvx + radii) + 16.0f) - 1.0f) / 16.0f), 0, tiles_x);  // "+ TILE_WIDTH - 1"
                by1 = clampi(__float2int_rz((((uvy + radii) + 16.0f) - 1.0f) / 16.0f), 0, tiles_y);
                // multi-GPU band clip (identity for the whole frame)
                if ((uint32_t)by0 < P.tile_row_begin) by0 = (int)P.tile_row_begin;
                if ((uint32_t)by1 > P.tile_row_end) by1 = (int)P.tile_row_end;
                if (by1 < by0) by1 = by0;
                nt = (uint32_t)(bx1 - bx0) * (uint32_t)(by1 - by0);  // :168
                surv = nt != 0;                                      // :169-171
                depth = vz;
            }
        }
    }

    // ---- block scan of (survivor, tiles) ----
    const unsigned surv_mask = __ballot_sync(FULL, surv);
    const uint32_t surv_rank_w = __popc(surv_mask & ((1u << lane) - 1u));
    uint32_t nt_incl = nt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(FULL, nt_incl, o);
        if (lane >= o) nt_incl += t;
    }
    if (lane == 31) {
        s_wsurv[warp] = __popc(surv_mask);
        s_wnt[warp] = nt_incl;
    }
    __syncthreads();
    uint32_t surv_before = 0, nt_before = 0, blk_surv = 0, blk_nt = 0;
#pragma unroll
    for (int w = 0; w < PRE_THREADS / 32; w++) {
        const uint32_t a = s_wsurv[w], b = s_wnt[w];
        if (w < warp) {
            surv_before += a;
            nt_before += b;
        }
        blk_surv += a;
        blk_nt += b;
    }
    const uint32_t local_rank = surv_before + surv_rank_w;      // compact slot within the chunk
    const uint32_t local_off = nt_before + (nt_incl - nt);      // exclusive tile offset within the chunk

    // publish this chunk's aggregate as early as possible
    if (tid == 0) st_status(P.status + chunk, st_pack(chunk == 0 ? ST_PREFIX : ST_AGG, blk_surv, blk_nt));

    // ---- SH colour of survivors (overlaps the look-back of other chunks) ----
    float colr = 0.f, colg = 0.f, colb = 0.f;
    if (surv) compute_sh(reinterpret_cast<const float4*>(P.sh) + (size_t)i * 12, px, py, pz, U.camera_position, colr, colg, colb);

    // ---- decoupled look-back by warp 0 ----
    if (warp == 0) {
        uint32_t ex_s = 0;
        unsigned long long ex_t = 0;
        if (chunk != 0) {
            int look = (int)chunk - 1;
            while (true) {
                const int idx = look - lane;
                unsigned long long st = ST_PREFIX;  // virtual predecessor of chunk 0
                if (idx >= 0) {
                    st = ld_status(P.status + idx);
                    while ((st & ST_FLAGS) == 0) st = ld_status(P.status + idx);
                }
                const unsigned pm = __ballot_sync(FULL, (st & ST_FLAGS) == ST_PREFIX);
                const int first = pm ? (__ffs(pm) - 1) : 32;
                uint32_t cs = (lane <= first) ? (uint32_t)((st >> 32) & 0x3fffffffu) : 0u;
                unsigned long long ct = (lane <= first) ? (st & 0xffffffffull) : 0ull;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    cs += __shfl_xor_sync(FULL, cs, o);
                    ct += __shfl_xor_sync(FULL, ct, o);
                }
                ex_s += cs;
                ex_t += ct;
                if (pm) break;
                look -= 32;
            }
            if (lane == 0) st_status(P.status + chunk, st_pack(ST_PREFIX, ex_s + blk_surv, ex_t + blk_nt));
        }
        if (lane == 0) {
            s_base_surv = ex_s;
            s_base_tiles = ex_t;
            if (chunk == num_chunks - 1) {  // global totals
                const unsigned long long total = ex_t + blk_nt;
                P.ctl->instances_total = total;
                P.ctl->num_instances = total > P.capacity ? P.capacity : (uint32_t)total;
                P.ctl->overflow = total > P.capacity ? 1u : 0u;
                P.ctl->num_visible = ex_s + blk_surv;
            }
        }
    }
    s_off[tid] = local_off;
    if (tid == 0) s_off[PRE_THREADS] = blk_nt;
    __syncthreads();
    const uint32_t base_surv = s_base_surv;
    const unsigned long long base_tiles = s_base_tiles;

    // ---- compacted blend record + emission descriptors ----
    const uint32_t cid = base_surv + local_rank;
    if (surv) {
        float4* rec = P.recs + (size_t)cid * 3;
        rec[0] = make_float4(uvx, uvy, conx, cony);
        rec[1] = make_float4(conz, opac, colr, colg);
        rec[2] = make_float4(colb, depth, radii, __uint_as_float(i));
        s_info[tid] = make_uint4((uint32_t)bx0 | ((uint32_t)by0 << 16), (uint32_t)(by1 - by0), __float_as_uint(depth), cid);
    }
    if (DEBUG && i < P.n) {
        P.dbg_tiles[i] = nt;
        const unsigned long long incl = base_tiles + local_off + nt;
        P.dbg_scan[i] = (uint32_t)incl;
        P.dbg_aabb[i] = surv ? make_uint4(bx0, by0, bx1, by1) : make_uint4(0, 0, 0, 0);
    }
    __syncthreads();

    // ---- block-cooperative key emission, preprocess_sort.comp:43-58 order ----
    const uint32_t tileX = (uint32_t)tiles_x;
    for (uint32_t j = tid; j < blk_nt; j += PRE_THREADS) {
        int lo = 0, hi = PRE_THREADS - 1;  // smallest g with s_off[g + 1] > j
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (s_off[mid + 1] <= j) lo = mid + 1;
            else hi = mid;
        }
        const uint4 inf = s_info[lo];
        const uint32_t k = j - s_off[lo];
        const uint32_t h = inf.y;
        const uint32_t x = (inf.x & 0xffffu) + k / h;  // x outer (:47)
        const uint32_t y = (inf.x >> 16) + k % h;      // y inner (:48)
        const unsigned long long slot = base_tiles + j;
        if (slot < P.capacity) {
            P.keys[slot] = ((unsigned long long)(x + y * tileX) << 32) | inf.z;  // :49-54
            P.vals[slot] = inf.w;                                               // compact id (orig idx in rec[2].w)
        }
    }
}

}  // namespace

cudaError_t launch_preprocess(const PreprocessParams& p, bool debug, cudaStream_t s) {
    if (p.n == 0) return cudaSuccess;
    const unsigned blocks = (p.n + PRE_THREADS - 1) / PRE_THREADS;
    if (debug) k_preprocess<true><<<blocks, PRE_THREADS, 0, s>>>(p);
    else k_preprocess<false><<<blocks, PRE_THREADS, 0, s>>>(p);
    return cudaGetLastError();
}

}  // namespace gsb
