// gsb_sort.cu -- device-wide Onesweep LSD radix sort of (u64 key, u32 payload) pairs.
// Replaces the reference's 8 x (sort/hist.comp:69-94 + sort/sort.comp:99-213) dispatch loop
// (src/Renderer.cpp:598-629).  Same contract: stable, ascending by key, 8-bit digits; but
//   * ONE histogram kernel reads the keys once and produces the digit histograms of all passes,
//   * each pass is ONE kernel: per-tile ranking + chained-scan decoupled look-back over tiles
//     (no O(workgroups^2) histogram re-reads as in sort.comp:112) + shared-memory staged,
//     coalesced scatter of key and payload,
//   * only P = ceil(key_bits / 8) passes run (the reference always runs 8; bits >= 32+log2(T)
//     are zero), and M is read from device memory (no host round trip).
// HBM traffic: M * (8 + 24 P) bytes.  No tensor cores: integer/byte work.
#include "gsb_internal.cuh"

namespace gsb {

namespace {

constexpr int SORT_IPT = 16;  // keys per thread
#ifndef GSB_SORT_THREADS
#define GSB_SORT_THREADS 256  // 256 threads x 16 = 4096-pair tiles, two CTAs per SM (phases of the two CTAs overlap)
#endif
constexpr int SORT_THREADS = GSB_SORT_THREADS;
constexpr int SORT_CTAS_PER_SM = SORT_THREADS == 512 ? 1 : 2;
constexpr int SORT_TILE = SORT_THREADS * SORT_IPT;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int RADIX = 256;
constexpr unsigned FULL = 0xffffffffu;

constexpr int HIST_THREADS = 512;
constexpr int HIST_IPT = 8;
constexpr int HIST_TILE = HIST_THREADS * HIST_IPT;

// look-back word: [63:32] epoch, [31:30] flag, [29:0] count.  Epoch tagging means the status
// array never needs clearing between passes or frames.
constexpr uint32_t LB_AGG = 1u << 30;
constexpr uint32_t LB_PREFIX = 2u << 30;
constexpr uint32_t LB_COUNT = (1u << 30) - 1u;

__device__ __forceinline__ unsigned long long ld_volatile(const unsigned long long* p) {
    return *reinterpret_cast<const volatile unsigned long long*>(p);
}
__device__ __forceinline__ void st_volatile(unsigned long long* p, unsigned long long v) {
    *reinterpret_cast<volatile unsigned long long*>(p) = v;
}

// ------------------------------------------------------------------------------------------
// Histogram of every digit in one read of the keys (replaces hist.comp, run once not 8x).
// ------------------------------------------------------------------------------------------
template <int P>
__global__ void __launch_bounds__(HIST_THREADS) k_sort_hist(const unsigned long long* __restrict__ keys,
                                                            const uint32_t* __restrict__ d_m, Control* ctl) {
    __shared__ uint32_t s_hist[P][RADIX];
    const int tid = threadIdx.x;
    for (int k = tid; k < P * RADIX; k += HIST_THREADS) (&s_hist[0][0])[k] = 0;
    __syncthreads();
    const uint32_t m = *d_m;
    const uint32_t num_tiles = (m + HIST_TILE - 1) / HIST_TILE;
    for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const uint32_t base = tile * HIST_TILE;
        unsigned long long k[HIST_IPT];
#pragma unroll
        for (int it = 0; it < HIST_IPT; it++) {
            const uint32_t idx = base + it * HIST_THREADS + tid;
            k[it] = idx < m ? __ldg(keys + idx) : ~0ull;
        }
#pragma unroll
        for (int it = 0; it < HIST_IPT; it++) {
            const uint32_t idx = base + it * HIST_THREADS + tid;
            const bool valid = idx < m;
#pragma unroll
            for (int p = 0; p < P; p++) {
                const uint32_t d = (uint32_t)(k[it] >> (8 * p)) & 255u;
                if (p >= 3) {
                    // upper digits (depth exponent byte, tile id) are skewed: lanes that share lane 0's digit are
                    // counted with one ballot and a single shared atomic instead of up to 32 serialised ones
                    const uint32_t d0 = __shfl_sync(FULL, d, 0);
                    const unsigned same = __ballot_sync(FULL, valid && d == d0);
                    if ((threadIdx.x & 31) == 0) {
                        if (same) atomicAdd(&s_hist[p][d0], (uint32_t)__popc(same));
                    } else if (valid && d != d0) {
                        atomicAdd(&s_hist[p][d], 1u);
                    }
                } else if (valid) {
                    atomicAdd(&s_hist[p][d], 1u);
                }
            }
        }
    }
    __syncthreads();
    for (int k = tid; k < P * RADIX; k += HIST_THREADS) {
        const uint32_t c = (&s_hist[0][0])[k];
        if (c) atomicAdd(&ctl->hist[0][0] + k, c);
    }
}

// ------------------------------------------------------------------------------------------
// One Onesweep pass (replaces one hist.comp + sort.comp pair).
//
// Persistent kernel, ONE 512-thread CTA per SM, tile = 8192 pairs (64 KB keys + 32 KB payloads).
//  * TMA: the next tile's keys and payloads are fetched by cp.async.bulk (UBLKCP) into the second
//    shared-memory buffer while the current tile is ranked and scattered; completion is an
//    mbarrier transaction count.  No registers are tied up by loads in flight.
//  * ranking: warp-striped, match.any per 8-bit digit, per-warp digit counters in smem (stable).
//  * the tile is permuted IN PLACE in shared memory (raw -> digit-sorted), so the global scatter
//    writes runs of consecutive addresses per digit (coalesced 8-B key / 4-B payload stores).
//  * chained scan: tile aggregate published right after ranking, decoupled look-back per digit
//    with a window of LB_WINDOW predecessors in flight, done after the in-place permutation so
//    the predecessors' latency overlaps local work.
// ------------------------------------------------------------------------------------------
constexpr int LB_WINDOW = 8;

struct PassSmem {
    unsigned long long keys[2][SORT_TILE];  // 2 x 64 KB (double buffer: current / prefetch)
    uint32_t vals[2][SORT_TILE];            // 2 x 32 KB
    uint32_t whist[SORT_WARPS][RADIX];      // per-warp digit counters -> exclusive offsets across warps
    uint32_t bin_start[RADIX];              // exclusive scan of the tile's digit counts
    int32_t out_base[RADIX];                // global index of bin d's first element minus bin_start[d]
    uint32_t gexcl[RADIX];                  // exclusive scan of the global histogram of this pass
    uint32_t warp_sums[SORT_WARPS];
    unsigned long long mbar[2];             // TMA completion barriers, one per buffer
    uint32_t tile[2];                       // ticket held by each buffer
};
static_assert(sizeof(PassSmem) * SORT_CTAS_PER_SM + 1024 * SORT_CTAS_PER_SM <= 227 * 1024, "PassSmem exceeds the 227 KB shared memory of an sm_100 SM");
// digit -> counter slot: XOR swizzle so digits that differ by a multiple of 32 (tile ids of one Gaussian
// are tiles_x apart) do not pile up in one shared-memory bank
__device__ __forceinline__ uint32_t sw(uint32_t d) { return d ^ (d >> 5); }
// Lanes of the warp holding the same 8-bit digit, by 8 ballots.  MATCH.ANY costs time proportional to the
// number of distinct values in the warp (measured: 37% of the pass on random digits); this is constant time.
__device__ __forceinline__ unsigned match_digit8(uint32_t d) {
    unsigned m = FULL;
#pragma unroll
    for (int b = 0; b < 8; b++) {
        const bool bit = (d >> b) & 1u;
        const unsigned vote = __ballot_sync(FULL, bit);
        m &= bit ? vote : ~vote;
    }
    return m;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA bulk copy global -> shared, completion counted in bytes on `bar` (SASS: UBLKCP)
__device__ __forceinline__ void tma_load(void* dst_smem, const void* src_gmem, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// exclusive scan of one value per thread across the block (NW warps)
template <int NW>
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* warp_sums) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    uint32_t before = 0;
#pragma unroll
    for (int w = 0; w < NW; w++) {
        const uint32_t s = warp_sums[w];
        if (w < warp) before += s;
    }
    __syncthreads();
    return before + incl - v;
}

__global__ void __launch_bounds__(SORT_THREADS, SORT_CTAS_PER_SM)
    k_onesweep_pass(const unsigned long long* __restrict__ kin, const uint32_t* __restrict__ vin,
                    unsigned long long* __restrict__ kout, uint32_t* __restrict__ vout,
                    const uint32_t* __restrict__ d_m, Control* ctl, int pass, unsigned long long* status,
                    uint32_t status_tiles, uint32_t epoch) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    PassSmem& S = *reinterpret_cast<PassSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int shift = 8 * pass;
    const uint32_t m = *d_m;
    uint32_t num_tiles = (m + SORT_TILE - 1) / SORT_TILE;
    if (num_tiles > status_tiles) num_tiles = status_tiles;  // host sizes status for the arena capacity
    const unsigned long long epoch_hi = (unsigned long long)epoch << 32;

    auto fetch = [&](int buf) {  // thread 0: take the next ticket and start its TMA loads into `buf`
        const uint32_t t = atomicAdd(&ctl->sort_ticket[pass], 1u);
        S.tile[buf] = t;
        if (t < num_tiles && m - t * SORT_TILE >= (uint32_t)SORT_TILE) {  // full tile: TMA; the ragged last tile is loaded by the threads
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy accesses to `buf` are done (barrier) -> async proxy may write
            mbar_expect_tx(&S.mbar[buf], SORT_TILE * 12);
            tma_load(S.keys[buf], kin + (size_t)t * SORT_TILE, SORT_TILE * 8, &S.mbar[buf]);
            tma_load(S.vals[buf], vin + (size_t)t * SORT_TILE, SORT_TILE * 4, &S.mbar[buf]);
        }
    };

    if (tid == 0) {
        mbar_init(&S.mbar[0], 1);
        mbar_init(&S.mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // exclusive prefix of the global histogram of this digit (thread d <-> bin d)
    {
        const uint32_t c = tid < RADIX ? ctl->hist[pass][tid] : 0u;
        const uint32_t ex = block_excl_scan<SORT_WARPS>(c, S.warp_sums);
        if (tid < RADIX) S.gexcl[tid] = ex;
    }
    if (tid == 0) fetch(0);
    __syncthreads();

    int cur = 0;
    uint32_t parity[2] = {0u, 0u};
    while (true) {
        const uint32_t tile = S.tile[cur];
        if (tile >= num_tiles) break;
        if (tid == 0) fetch(cur ^ 1);  // buffer cur^1 was released by the barrier that ended the previous iteration
        const uint32_t tile_base = tile * SORT_TILE;
        const uint32_t valid = min((uint32_t)SORT_TILE, m - tile_base);
        unsigned long long* sk = S.keys[cur];
        uint32_t* sv = S.vals[cur];

        // zero the per-warp counters (16 x 256 words, 8 per thread)
#pragma unroll
        for (int k = 0; k < SORT_WARPS * RADIX / SORT_THREADS; k++) (&S.whist[0][0])[k * SORT_THREADS + tid] = 0;
        if (valid == (uint32_t)SORT_TILE) {
            mbar_wait(&S.mbar[cur], parity[cur]);
            parity[cur] ^= 1u;
        } else {  // ragged last tile: plain loads, padded with the maximum key so the padding sorts last
#pragma unroll 4
            for (int it = 0; it < SORT_IPT; it++) {
                const uint32_t li = it * SORT_THREADS + tid;
                const bool ok = li < valid;
                sk[li] = ok ? __ldg(kin + tile_base + li) : ~0ull;
                sv[li] = ok ? __ldg(vin + tile_base + li) : 0u;
            }
        }
        __syncthreads();

        // ---- rank inside the warp (stable): warp-striped so (warp, round, lane) order == memory order ----
        const uint32_t wbase = warp * (32 * SORT_IPT) + lane;
        uint16_t rank[SORT_IPT];
        {
            uint32_t dg[SORT_IPT];
            unsigned peers[SORT_IPT];
#pragma unroll
            for (int it = 0; it < SORT_IPT; it++) dg[it] = sw((uint32_t)(sk[wbase + it * 32] >> shift) & 255u);  // 16 LDS in flight
#pragma unroll
            for (int it = 0; it < SORT_IPT; it++) peers[it] = match_digit8(dg[it]);
#pragma unroll
            for (int it = 0; it < SORT_IPT; it++) {  // the counter update is the only serial part
                const int leader = __ffs(peers[it]) - 1;
                uint32_t prev = 0;
                if (lane == leader) {
                    prev = S.whist[warp][dg[it]];
                    S.whist[warp][dg[it]] = prev + __popc(peers[it]);
                }
                prev = __shfl_sync(FULL, prev, leader);
                rank[it] = (uint16_t)(prev + __popc(peers[it] & ((1u << lane) - 1u)));
                __syncwarp();
            }
        }
        __syncthreads();

        // ---- per digit (thread d < 256): exclusive scan across warps, tile count, publish the aggregate ----
        uint32_t count = 0;
        if (tid < RADIX) {
#pragma unroll
            for (int w = 0; w < SORT_WARPS; w++) {
                const uint32_t c = S.whist[w][sw(tid)];
                S.whist[w][sw(tid)] = count;
                count += c;
            }
            // tile 0 has no predecessors: inclusive prefix right away
            st_volatile(status + (size_t)tile * RADIX + tid, epoch_hi | (tile == 0 ? LB_PREFIX : LB_AGG) | count);
        }
        const uint32_t bstart = block_excl_scan<SORT_WARPS>(count, S.warp_sums);  // threads >= 256 contribute 0
        if (tid < RADIX) S.bin_start[sw(tid)] = bstart;
        __syncthreads();

        // ---- permute the tile in place: raw order -> digit-sorted order ----
        unsigned long long key[SORT_IPT];
        uint32_t val[SORT_IPT];
#pragma unroll
        for (int it = 0; it < SORT_IPT; it++) {
            key[it] = sk[wbase + it * 32];
            val[it] = sv[wbase + it * 32];
        }
        __syncthreads();
        {
            uint32_t pos[SORT_IPT];
#pragma unroll
            for (int it = 0; it < SORT_IPT; it++) {  // all lookups first (independent LDS), then all stores
                const uint32_t d = sw((uint32_t)(key[it] >> shift) & 255u);
                pos[it] = S.bin_start[d] + S.whist[warp][d] + rank[it];
            }
#pragma unroll
            for (int it = 0; it < SORT_IPT; it++) {
                sk[pos[it]] = key[it];
                sv[pos[it]] = val[it];
            }
        }

        // ---- decoupled look-back for digit `tid`, LB_WINDOW predecessors in flight ----
        if (tid < RADIX) {
            uint32_t excl = 0;
            int t = (int)tile - 1;
            bool done = tile == 0;
            while (!done) {
                unsigned long long w[LB_WINDOW];
#pragma unroll
                for (int j = 0; j < LB_WINDOW; j++)
                    w[j] = (t - j >= 0) ? ld_volatile(status + (size_t)(t - j) * RADIX + tid) : (epoch_hi | LB_PREFIX);
#pragma unroll
                for (int j = 0; j < LB_WINDOW; j++) {
                    if (done) break;
                    const uint32_t lo = (uint32_t)w[j];
                    if ((w[j] >> 32) != epoch) break;  // not yet published in this pass: poll again from here
                    excl += lo & LB_COUNT;
                    --t;
                    if (lo & LB_PREFIX) done = true;
                }
            }
            if (tile != 0) st_volatile(status + (size_t)tile * RADIX + tid, epoch_hi | LB_PREFIX | (excl + count));
            S.out_base[sw(tid)] = (int32_t)(S.gexcl[tid] + excl) - (int32_t)bstart;
        }
        __syncthreads();

        // ---- coalesced global scatter: consecutive idx of one digit -> consecutive addresses ----
        {
            unsigned long long k[SORT_IPT];
            uint32_t v[SORT_IPT];
            int32_t g[SORT_IPT];
#pragma unroll
            for (int it = 0; it < SORT_IPT; it++) {
                k[it] = sk[it * SORT_THREADS + tid];
                v[it] = sv[it * SORT_THREADS + tid];
            }
#pragma unroll
            for (int it = 0; it < SORT_IPT; it++)
                g[it] = S.out_base[sw((uint32_t)(k[it] >> shift) & 255u)] + (int32_t)(it * SORT_THREADS + tid);
#pragma unroll
            for (int it = 0; it < SORT_IPT; it++) {
                if ((uint32_t)(it * SORT_THREADS + tid) < valid) {
                    kout[g[it]] = k[it];
                    vout[g[it]] = v[it];
                }
            }
        }
        __syncthreads();  // buffer `cur` may now be refilled by TMA
        cur ^= 1;
    }
}

}  // namespace

uint32_t sort_tile_items() { return SORT_TILE; }

cudaError_t launch_sort(const SortParams& p, uint32_t* passes, cudaStream_t s) {
    const uint32_t P = (p.key_bits + 7) / 8;
    *passes = P;
    if (P == 0 || P > 8) return P == 0 ? cudaSuccess : cudaErrorInvalidValue;
    const uint32_t hint = p.m_hint ? p.m_hint : 1;
    // histogram: persistent grid-stride, at most 2 CTAs per SM
    {
        uint32_t blocks = (hint + HIST_TILE - 1) / HIST_TILE;
        const uint32_t cap = (uint32_t)p.num_sms * 2;
        if (blocks > cap) blocks = cap;
        if (blocks == 0) blocks = 1;
        const unsigned long long* k0 = p.keys[0];
        switch (P) {
            case 1: k_sort_hist<1><<<blocks, HIST_THREADS, 0, s>>>(k0, p.d_m, p.ctl); break;
            case 2: k_sort_hist<2><<<blocks, HIST_THREADS, 0, s>>>(k0, p.d_m, p.ctl); break;
            case 3: k_sort_hist<3><<<blocks, HIST_THREADS, 0, s>>>(k0, p.d_m, p.ctl); break;
            case 4: k_sort_hist<4><<<blocks, HIST_THREADS, 0, s>>>(k0, p.d_m, p.ctl); break;
            case 5: k_sort_hist<5><<<blocks, HIST_THREADS, 0, s>>>(k0, p.d_m, p.ctl); break;
            case 6: k_sort_hist<6><<<blocks, HIST_THREADS, 0, s>>>(k0, p.d_m, p.ctl); break;
            case 7: k_sort_hist<7><<<blocks, HIST_THREADS, 0, s>>>(k0, p.d_m, p.ctl); break;
            default: k_sort_hist<8><<<blocks, HIST_THREADS, 0, s>>>(k0, p.d_m, p.ctl); break;
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        if (p.events && (e = cudaEventRecord(p.events[0], s)) != cudaSuccess) return e;
    }
    const size_t smem = sizeof(PassSmem);
    {
        cudaError_t e = cudaFuncSetAttribute(k_onesweep_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    uint32_t blocks = (hint + SORT_TILE - 1) / SORT_TILE;
    const uint32_t cap = (uint32_t)p.num_sms * SORT_CTAS_PER_SM;  // persistent CTAs; ticket loop: any grid size is correct
    if (blocks > cap) blocks = cap;
    if (blocks == 0) blocks = 1;
    for (uint32_t pass = 0; pass < P; pass++) {
        const int src = pass & 1, dst = src ^ 1;
        k_onesweep_pass<<<blocks, SORT_THREADS, smem, s>>>(p.keys[src], p.vals[src], p.keys[dst], p.vals[dst], p.d_m,
                                                           p.ctl, (int)pass, p.status, p.status_tiles,
                                                           p.epoch_base + pass);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        if (p.events && (e = cudaEventRecord(p.events[1 + pass], s)) != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// ------------------------------------------------------------------------------------------
// Tile ranges (replaces fillBuffer(0) + tile_boundary.comp:22-50, Renderer.cpp:633-652).
// ------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) k_tile_ranges(const unsigned long long* __restrict__ keys,
                                                     const uint32_t* __restrict__ d_m, uint2* __restrict__ ranges) {
    const uint32_t m = *d_m;
    uint32_t* r = reinterpret_cast<uint32_t*>(ranges);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
        const uint32_t key = (uint32_t)(__ldg(keys + i) >> 32);
        if (i == 0) {
            r[key * 2] = 0;
        } else {
            const uint32_t prev = (uint32_t)(__ldg(keys + i - 1) >> 32);
            if (key != prev) {
                r[key * 2] = i;
                r[prev * 2 + 1] = i;
            }
        }
        if (i == m - 1) r[key * 2 + 1] = m;
    }
}
}  // namespace

cudaError_t launch_tile_ranges(const unsigned long long* keys, const uint32_t* d_m, uint32_t m_hint, uint2* ranges,
                               uint32_t num_tiles, int num_sms, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(ranges, 0, (size_t)num_tiles * sizeof(uint2), s);
    if (e != cudaSuccess) return e;
    uint32_t blocks = (m_hint + 255) / 256;
    const uint32_t cap = (uint32_t)num_sms * 8;
    if (blocks > cap) blocks = cap;
    if (blocks == 0) blocks = 1;
    k_tile_ranges<<<blocks, 256, 0, s>>>(keys, d_m, ranges);
    return cudaGetLastError();
}

}  // namespace gsb
