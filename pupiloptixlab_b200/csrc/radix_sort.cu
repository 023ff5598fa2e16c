// In-tree LSD radix sort of (64-bit key, 32-bit value) pairs for the BVH builders: 8-bit digits, one pass per digit, each pass
// ONE kernel that reads every pair once and writes it once ("onesweep": Adinets & Merrill 2022) — the stage the OptiX build
// keeps inside optixAccelBuild (framework/world/gas_manager.cpp:211).
//
//   k_sort_histogram   one read of the keys: the 256-bin histogram of EVERY digit (the keys do not change between passes)
//   k_sort_scan        exclusive scan of each histogram -> first output position of every digit value, per pass
//   k_sort_pass        per tile of 4096 pairs (256 threads x 16, warp-striped so that ranking order = input order):
//                        1. rank each pair among the pairs of its warp with the same digit (__match_any_sync + a per-warp
//                           digit counter in shared memory): stable
//                        2. per digit: counts of the tile, offsets of the warps
//                        3. chained scan with decoupled look-back over the tiles before this one: thread d publishes the tile's
//                           count of digit d (flag AGGREGATE), walks back over earlier tiles adding their aggregates until it
//                           meets an INCLUSIVE prefix, publishes its own inclusive prefix.  Tiles take their index from an
//                           atomic ticket, so every tile a tile waits for is already running: no deadlock by scheduling
//                        4. pairs to shared memory in digit order, then out: runs of equal digits go to consecutive addresses
//
// The spin of step 3 is bounded (a hung kernel would cost the GPU box): after ~2^26 polls a tile gives up and raises the error
// flag the host checks.
#include "scene.cuh"

namespace pb2 {
namespace {
constexpr int kSortThreads = 256, kSortItems = 16, kSortTile = kSortThreads * kSortItems, kSortWarps = kSortThreads / 32;
constexpr uint32_t kFlagAggregate = 1u << 30, kFlagInclusive = 2u << 30, kValueMask = (1u << 30) - 1u;

__global__ void __launch_bounds__(256) k_sort_histogram(const uint64_t *__restrict__ keys, uint32_t n, int begin_bit, int n_passes, uint32_t *__restrict__ hist) {
    __shared__ uint32_t s_hist[8][256];
    for (int k = threadIdx.x; k < 8 * 256; k += blockDim.x) (&s_hist[0][0])[k] = 0u;
    __syncthreads();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint64_t key = keys[i] >> begin_bit;
        for (int p = 0; p < n_passes; ++p) atomicAdd(&s_hist[p][(uint32_t)(key >> (8 * p)) & 0xffu], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < n_passes * 256; k += blockDim.x) {
        const uint32_t v = (&s_hist[0][0])[k];
        if (v) atomicAdd(&hist[k], v);
    }
}
// one CTA per pass: exclusive scan of its 256 bins, in place
__global__ void __launch_bounds__(256) k_sort_scan(uint32_t *__restrict__ hist) {
    __shared__ uint32_t s_warp[8];
    uint32_t *h = hist + blockIdx.x * 256;
    const uint32_t v = h[threadIdx.x];
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
        if ((int)(threadIdx.x & 31) >= d) incl += up;
    }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t before = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) before += s_warp[w];
    h[threadIdx.x] = before + incl - v;
}

struct SortSmem {
    uint64_t keys[kSortTile];
    uint32_t vals[kSortTile];
    uint32_t warp_hist[kSortWarps][256]; // per warp: running count of each digit while ranking, then the warp's offset inside the digit's run
    uint32_t digit_start[256];           // first shared-memory slot of each digit's run
    uint32_t global_base[256];           // output index of shared-memory slot j with digit d = global_base[d] + j
    uint32_t tile;
};

__global__ void __launch_bounds__(kSortThreads, 2) k_sort_pass(const uint64_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in, uint64_t *__restrict__ keys_out,
                                                                uint32_t *__restrict__ vals_out, uint32_t n, int shift, const uint32_t *__restrict__ digit_base,
                                                                uint32_t *__restrict__ status, uint32_t *__restrict__ ticket, uint32_t *__restrict__ error_flag) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SortSmem &sm = *reinterpret_cast<SortSmem *>(smem_raw);
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) sm.tile = atomicAdd(ticket, 1u);
    for (int k = threadIdx.x; k < kSortWarps * 256; k += kSortThreads) (&sm.warp_hist[0][0])[k] = 0u;
    __syncthreads();
    const uint32_t tile = sm.tile;
    const uint32_t tile_base = tile * (uint32_t)kSortTile;
    const uint32_t n_valid = min((uint32_t)kSortTile, n - tile_base);

    // ---- 1. load (warp-striped) and rank ----
    uint64_t key[kSortItems];
    uint32_t val[kSortItems];
    uint16_t rank[kSortItems];
    const uint32_t seg = tile_base + warp * (kSortItems * 32u) + lane;
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        const uint32_t g = seg + i * 32u;
        key[i] = g < n ? keys_in[g] : ~0ull; // padding ranks behind every real pair of the last tile and is never written
        val[i] = g < n ? vals_in[g] : 0u;
    }
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        const uint32_t d = (uint32_t)(key[i] >> shift) & 0xffu;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const uint32_t before = sm.warp_hist[warp][d]; // every peer reads the same word
        __syncwarp();
        if ((peers & ((1u << lane) - 1u)) == 0u) sm.warp_hist[warp][d] = before + __popc(peers); // the lowest peer bumps the counter
        __syncwarp();
        rank[i] = (uint16_t)(before + __popc(peers & ((1u << lane) - 1u)));
    }
    __syncthreads();

    // ---- 2. per digit: offsets of the warps, count of the tile ----
    const uint32_t d = threadIdx.x; // one digit value per thread
    uint32_t count = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
        const uint32_t c = sm.warp_hist[w][d];
        sm.warp_hist[w][d] = count;
        count += c;
    }
    // ---- 3. publish, look back ----
    uint32_t *my_status = status + (size_t)tile * 256u + d;
    if (tile == 0) {
        atomicExch(my_status, kFlagInclusive | count);
    } else {
        atomicExch(my_status, kFlagAggregate | count);
    }
    // exclusive scan of the counts over the digits (block scan) for the shared-memory layout
    {
        __shared__ uint32_t s_warp[kSortWarps];
        uint32_t incl = count;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, dd);
            if ((int)lane >= dd) incl += up;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t before = 0;
        for (uint32_t w = 0; w < warp; ++w) before += s_warp[w];
        sm.digit_start[d] = before + incl - count;
    }
    uint32_t exclusive = 0;
    if (tile > 0) {
        for (int t = (int)tile - 1; t >= 0; --t) {
            const volatile uint32_t *st = status + (size_t)t * 256u + d;
            uint32_t v = *st;
            for (uint32_t spin = 0; (v >> 30) == 0u; ++spin) {
                if (spin > (1u << 26)) { // a predecessor never published: report instead of hanging the device
                    atomicExch(error_flag, 1u);
                    v = kFlagInclusive;
                    break;
                }
                __nanosleep(40);
                v = *st;
            }
            exclusive += v & kValueMask;
            if (v & kFlagInclusive) break;
        }
        atomicExch(my_status, kFlagInclusive | ((exclusive + count) & kValueMask));
    }
    sm.global_base[d] = digit_base[d] + exclusive - sm.digit_start[d];
    __syncthreads();

    // ---- 4. to shared memory in digit order, then out ----
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        const uint32_t dg = (uint32_t)(key[i] >> shift) & 0xffu;
        const uint32_t pos = sm.digit_start[dg] + sm.warp_hist[warp][dg] + rank[i];
        sm.keys[pos] = key[i];
        sm.vals[pos] = val[i];
    }
    __syncthreads();
    for (uint32_t j = threadIdx.x; j < n_valid; j += kSortThreads) {
        const uint64_t k = sm.keys[j];
        const uint32_t dg = (uint32_t)(k >> shift) & 0xffu;
        const uint32_t out = sm.global_base[dg] + j;
        keys_out[out] = k;
        vals_out[out] = sm.vals[j];
    }
}
}// namespace

// Sorts n (key, value) pairs by the key bits [begin_bit, end_bit).  keys / vals hold the input and are used as scratch; the result
// is in keys_alt / vals_alt or back in keys / vals — the return value says which (true: the alt buffers).
bool radix_sort_pairs(cudaStream_t st, uint64_t *keys, uint64_t *keys_alt, uint32_t *vals, uint32_t *vals_alt, uint32_t n, int begin_bit, int end_bit) {
    const int n_passes = (end_bit - begin_bit + 7) / 8;
    if (n == 0 || n_passes <= 0) return false;
    if (n_passes > 8) throw std::runtime_error("radix_sort_pairs: at most 64 key bits");
    if (n > kValueMask) throw std::runtime_error("radix_sort_pairs: at most 2^30 - 1 pairs");
    const uint32_t tiles = (n + kSortTile - 1) / kSortTile;
    DevBuf<uint32_t> hist((size_t)n_passes * 256), status((size_t)n_passes * tiles * 256), misc(n_passes + 1); // misc: a ticket per pass, the error flag
    hist.zero(st), status.zero(st), misc.zero(st);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    k_sort_histogram<<<(unsigned)std::min<uint64_t>((n + 255) / 256, (uint64_t)sms * 8), 256, 0, st>>>(keys, n, begin_bit, n_passes, hist.ptr);
    PB2_LAUNCH_CHECK();
    k_sort_scan<<<n_passes, 256, 0, st>>>(hist.ptr);
    PB2_LAUNCH_CHECK();
    PB2_CUDA(cudaFuncSetAttribute(k_sort_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SortSmem))); // per device: set every time
    uint64_t *kin = keys, *kout = keys_alt;
    uint32_t *vin = vals, *vout = vals_alt;
    for (int p = 0; p < n_passes; ++p) {
        k_sort_pass<<<tiles, kSortThreads, sizeof(SortSmem), st>>>(kin, vin, kout, vout, n, begin_bit + 8 * p, hist.ptr + p * 256, status.ptr + (size_t)p * tiles * 256,
                                                                misc.ptr + p, misc.ptr + n_passes);
        PB2_LAUNCH_CHECK();
        std::swap(kin, kout), std::swap(vin, vout);
    }
    uint32_t err = 0;
    PB2_CUDA(cudaMemcpyAsync(&err, misc.ptr + n_passes, sizeof err, cudaMemcpyDeviceToHost, st));
    PB2_CUDA(cudaStreamSynchronize(st));
    if (err) throw std::runtime_error("radix_sort_pairs: a tile waited for a predecessor that never published");
    return kin == keys_alt;
}
}// namespace pb2
