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

constexpr int SORT_THREADS = 256;
constexpr int SORT_IPT = 16;                          // keys per thread
constexpr int SORT_TILE = SORT_THREADS * SORT_IPT;   // 4096 keys per tile
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
                    // upper digits (depth exponent byte, tile id) are heavily skewed: aggregate equal
                    // digits inside the warp so a hot bin costs one shared atomic, not 32 serialised ones
                    const unsigned peers = __match_any_sync(FULL, valid ? d : 0xffffffffu);
                    if (valid && (threadIdx.x & 31) == (__ffs(peers) - 1)) atomicAdd(&s_hist[p][d], (uint32_t)__popc(peers));
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
// ------------------------------------------------------------------------------------------
struct PassSmem {
    union {
        unsigned long long keys[SORT_TILE];  // 32 KB
        uint32_t vals[SORT_TILE];
    };
    uint32_t whist[SORT_WARPS][RADIX];  // per-warp digit counters -> exclusive offsets across warps
    uint32_t bin_start[RADIX];          // exclusive scan of the tile's digit counts
    int32_t out_base[RADIX];            // global index of bin d's first element minus bin_start[d]
    uint32_t gexcl[RADIX];              // exclusive scan of the global histogram of this pass
    uint32_t warp_sums[SORT_WARPS];
    uint32_t tile;
};

// exclusive scan of one value per thread across the 256-thread block
__device__ __forceinline__ uint32_t block_excl_scan_256(uint32_t v, uint32_t* warp_sums, uint32_t* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    uint32_t before = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < SORT_WARPS; w++) {
        const uint32_t s = warp_sums[w];
        if (w < warp) before += s;
        tot += s;
    }
    __syncthreads();
    if (total) *total = tot;
    return before + incl - v;
}

__global__ void __launch_bounds__(SORT_THREADS, 2)
    k_onesweep_pass(const unsigned long long* __restrict__ kin, const uint32_t* __restrict__ vin,
                    unsigned long long* __restrict__ kout, uint32_t* __restrict__ vout,
                    const uint32_t* __restrict__ d_m, Control* ctl, int pass, unsigned long long* status,
                    uint32_t status_tiles, uint32_t epoch) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PassSmem& S = *reinterpret_cast<PassSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int shift = 8 * pass;
    const uint32_t m = *d_m;
    uint32_t num_tiles = (m + SORT_TILE - 1) / SORT_TILE;
    if (num_tiles > status_tiles) num_tiles = status_tiles;  // host sizes status for the arena capacity
    const unsigned long long epoch_hi = (unsigned long long)epoch << 32;

    // exclusive prefix of the global histogram of this digit (thread d <-> bin d)
    {
        const uint32_t c = ctl->hist[pass][tid];
        S.gexcl[tid] = block_excl_scan_256(c, S.warp_sums, nullptr);
    }

    while (true) {
        if (tid == 0) S.tile = atomicAdd(&ctl->sort_ticket[pass], 1u);
        // zero the per-warp counters
#pragma unroll
        for (int k = 0; k < SORT_WARPS; k++) S.whist[k][tid] = 0;
        __syncthreads();
        const uint32_t tile = S.tile;
        if (tile >= num_tiles) break;
        const uint32_t tile_base = tile * SORT_TILE;
        const uint32_t valid = min((uint32_t)SORT_TILE, m - tile_base);

        // ---- load: warp-striped so that (warp, round, lane) order == memory order ----
        unsigned long long key[SORT_IPT];
        uint32_t val[SORT_IPT];
#pragma unroll
        for (int it = 0; it < SORT_IPT; it++) {
            const uint32_t li = warp * (32 * SORT_IPT) + it * 32 + lane;
            const bool ok = li < valid;
            key[it] = ok ? __ldg(kin + tile_base + li) : ~0ull;
            val[it] = ok ? __ldg(vin + tile_base + li) : 0u;
        }

        // ---- rank inside the warp (stable): match equal digits, count predecessors ----
        uint32_t rank[SORT_IPT];  // rank within (warp, digit)
#pragma unroll
        for (int it = 0; it < SORT_IPT; it++) {
            const uint32_t d = (uint32_t)(key[it] >> shift) & 255u;
            const unsigned peers = __match_any_sync(FULL, d);
            const int leader = __ffs(peers) - 1;
            uint32_t prev = 0;
            if (lane == leader) {
                prev = S.whist[warp][d];
                S.whist[warp][d] = prev + __popc(peers);
            }
            prev = __shfl_sync(FULL, prev, leader);
            rank[it] = prev + __popc(peers & ((1u << lane) - 1u));
            __syncwarp();
        }
        __syncthreads();

        // ---- per digit (thread d): exclusive scan across warps, tile count ----
        uint32_t count = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) {
            const uint32_t c = S.whist[w][tid];
            S.whist[w][tid] = count;
            count += c;
        }
        // publish the tile aggregate (tile 0 has no predecessors: inclusive prefix right away)
        st_volatile(status + (size_t)tile * RADIX + tid, epoch_hi | (tile == 0 ? LB_PREFIX : LB_AGG) | count);

        const uint32_t bstart = block_excl_scan_256(count, S.warp_sums, nullptr);
        S.bin_start[tid] = bstart;

        // ---- decoupled look-back for digit `tid` ----
        uint32_t excl = 0;
        if (tile != 0) {
            int t = (int)tile - 1;
            while (true) {
                const unsigned long long w = ld_volatile(status + (size_t)t * RADIX + tid);
                if ((w >> 32) != epoch) continue;  // not yet published in this pass
                const uint32_t lo = (uint32_t)w;
                if ((lo & (LB_AGG | LB_PREFIX)) == 0) continue;
                excl += lo & LB_COUNT;
                if (lo & LB_PREFIX) break;
                --t;
            }
            st_volatile(status + (size_t)tile * RADIX + tid, epoch_hi | LB_PREFIX | (excl + count));
        }
        S.out_base[tid] = (int32_t)(S.gexcl[tid] + excl) - (int32_t)bstart;
        __syncthreads();

        // ---- scatter keys into tile-sorted order in shared memory ----
        uint16_t pos[SORT_IPT];
#pragma unroll
        for (int it = 0; it < SORT_IPT; it++) {
            const uint32_t d = (uint32_t)(key[it] >> shift) & 255u;
            const uint32_t p = S.bin_start[d] + S.whist[warp][d] + rank[it];
            pos[it] = (uint16_t)p;
            S.keys[p] = key[it];
        }
        __syncthreads();
        uint8_t dig[SORT_IPT];
#pragma unroll
        for (int it = 0; it < SORT_IPT; it++) {
            const uint32_t idx = it * SORT_THREADS + tid;
            const unsigned long long k = S.keys[idx];
            const uint32_t d = (uint32_t)(k >> shift) & 255u;
            dig[it] = (uint8_t)d;
            if (idx < valid) kout[S.out_base[d] + (int32_t)idx] = k;
        }
        __syncthreads();
#pragma unroll
        for (int it = 0; it < SORT_IPT; it++) S.vals[pos[it]] = val[it];
        __syncthreads();
#pragma unroll
        for (int it = 0; it < SORT_IPT; it++) {
            const uint32_t idx = it * SORT_THREADS + tid;
            if (idx < valid) vout[S.out_base[dig[it]] + (int32_t)idx] = S.vals[idx];
        }
        __syncthreads();
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
    static_assert(sizeof(PassSmem) <= 48 * 1024, "PassSmem must fit the default dynamic shared memory limit");
    uint32_t blocks = (hint + SORT_TILE - 1) / SORT_TILE;
    const uint32_t cap = (uint32_t)p.num_sms * 4;  // ticket loop: any grid size is correct
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
