// SAH-driven binary BVH builder (builder 2 of pb2_scene_set_builder): parallel locally-ordered clustering
// (after Meister & Bittner, "Parallel Locally-Ordered Clustering for Bounding Volume Hierarchy Construction", 2018).
//
// Replaces the quality side of the OptiX build the reference asks for with OPTIX_BUILD_FLAG_PREFER_FAST_TRACE
// (framework/world/gas_manager.cpp:191, :211-224): LBVH places a split at the highest differing Morton bit whatever the boxes
// look like; here the tree is built bottom-up by merging the pair of clusters whose union has the smallest surface area —
// the greedy step of the surface-area heuristic — with the search for that partner restricted to the `radius` neighbours
// on either side along the Morton curve, which is what makes it parallel:
//
//   clusters = the primitives in Morton order (bvh_build.cu stages 1-2)
//   repeat until one cluster is left:
//     k_ploc_nn       every cluster finds the neighbour within +-radius whose union with it has the least area
//                     (tile + halo of boxes staged in shared memory; ties broken by index so the choice is symmetric)
//     k_ploc_merge    mutual nearest neighbours merge: the lower one becomes the parent node, the upper one is dropped;
//                     per tile, the survivors are counted
//     k_ploc_scan     exclusive scan of the tile counts (one CTA), new cluster count
//     k_ploc_compact  survivors move to the front of the other buffer IN ORDER (the Morton locality the next round's
//                     neighbour search relies on)
//
// Each round removes 30-45 % of the clusters, so ~2.5 n cluster visits in total.  The binary tree comes out in the BinTree
// arrays of bvh_build.cu (left / right: >= 0 internal node, < 0 = ~sorted position; lo / hi boxes; range.y = primitives below)
// and goes through the same collapse to 8-wide nodes, which picks the binary nodes that survive by surface area.  Subtrees
// are NOT contiguous ranges of the Morton order here, which is why the collapse walks small subtrees instead of using ranges.
#include "scene.cuh"
#include <cfloat>

namespace pb2 {
namespace {
constexpr int kTile = 256;
constexpr int kMaxRadius = 16;

__device__ __forceinline__ float union_area(float3 alo, float3 ahi, float3 blo, float3 bhi) {
    const float3 d = fmax3(ahi, bhi) - fmin3(alo, blo);
    return d.x * d.y + d.y * d.z + d.z * d.x;
}

struct Clusters {
    int *id;         // >= 0 internal node, < 0 = ~sorted position
    float4 *lo, *hi; // lo.w = bits(primitives below)
};

__global__ void __launch_bounds__(256) k_ploc_init(uint32_t n, const uint32_t *__restrict__ sorted, const float4 *__restrict__ box_lo,
                                                    const float4 *__restrict__ box_hi, Clusters c) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t p = sorted[i];
        const float4 l = box_lo[p], h = box_hi[p];
        c.id[i] = ~(int)i;
        c.lo[i] = make_float4(l.x, l.y, l.z, __int_as_float(1));
        c.hi[i] = make_float4(h.x, h.y, h.z, 0.f);
    }
}

// nearest neighbour within +-radius by surface area of the union; total order (area, min index, max index): symmetric, so the
// globally best pair of every neighbourhood is mutual and every round merges something
__global__ void __launch_bounds__(kTile) k_ploc_nn(const uint32_t *__restrict__ n_ptr, Clusters c, int *__restrict__ nn, int radius) {
    __shared__ float s_lo[3][kTile + 2 * kMaxRadius], s_hi[3][kTile + 2 * kMaxRadius];
    const int n = (int)*n_ptr;
    if (n <= 1) return;
    const int tiles = (n + kTile - 1) / kTile;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int start = tile * kTile - radius;
        __syncthreads();
        for (int k = threadIdx.x; k < kTile + 2 * radius; k += kTile) {
            const int g = start + k;
            float4 l = make_float4(FLT_MAX, FLT_MAX, FLT_MAX, 0.f), h = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, 0.f);
            if (g >= 0 && g < n) l = c.lo[g], h = c.hi[g];
            s_lo[0][k] = l.x, s_lo[1][k] = l.y, s_lo[2][k] = l.z;
            s_hi[0][k] = h.x, s_hi[1][k] = h.y, s_hi[2][k] = h.z;
        }
        __syncthreads();
        const int i = tile * kTile + threadIdx.x;
        if (i >= n) continue;
        const int me = threadIdx.x + radius;
        const float3 lo = mk3(s_lo[0][me], s_lo[1][me], s_lo[2][me]), hi = mk3(s_hi[0][me], s_hi[1][me], s_hi[2][me]);
        float best = FLT_MAX;
        int best_j = -1;
        for (int d = -radius; d <= radius; ++d) {
            const int j = i + d;
            if (d == 0 || j < 0 || j >= n) continue;
            const int k = me + d;
            const float a = union_area(lo, hi, mk3(s_lo[0][k], s_lo[1][k], s_lo[2][k]), mk3(s_hi[0][k], s_hi[1][k], s_hi[2][k]));
            // d ascends, so among equal areas the first candidate has the smallest j: for j < i that is the smallest min index,
            // for j > i (min index = i for all of them) the smallest max index; a j < i beats a j > i at equal area
            if (a < best) best = a, best_j = j;
        }
        if (best_j < 0) best_j = i > 0 ? i - 1 : i + 1; // boxes with NaNs compare false everywhere: pair by position so the round still merges
        nn[i] = best_j;
    }
}

struct Tree {
    int *left, *right;
    int2 *range;
    float4 *lo, *hi;
};

// mutual nearest neighbours merge; keep[i] = the cluster survives this round; tile_count[tile] = survivors of the tile
__global__ void __launch_bounds__(kTile) k_ploc_merge(const uint32_t *__restrict__ n_ptr, Clusters c, const int *__restrict__ nn, Tree t,
                                                      uint32_t *__restrict__ node_counter, uint8_t *__restrict__ keep, uint32_t *__restrict__ tile_count) {
    __shared__ uint32_t s_warp[kTile / 32];
    const int n = (int)*n_ptr;
    if (n <= 1) return;
    const int tiles = (n + kTile - 1) / kTile;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int i = tile * kTile + threadIdx.x;
        bool survive = false;
        if (i < n) {
            const int j = nn[i];
            const bool mutual = j >= 0 && nn[j] == i;
            survive = !(mutual && i > j);
            if (mutual && i < j) {
                const float4 li = c.lo[i], hi_ = c.hi[i], lj = c.lo[j], hj = c.hi[j];
                const int cnt = __float_as_int(li.w) + __float_as_int(lj.w);
                const int k = (int)atomicAdd(node_counter, 1u);
                t.left[k] = c.id[i], t.right[k] = c.id[j];
                const float4 lo = make_float4(fminf(li.x, lj.x), fminf(li.y, lj.y), fminf(li.z, lj.z), 0.f);
                const float4 hi = make_float4(fmaxf(hi_.x, hj.x), fmaxf(hi_.y, hj.y), fmaxf(hi_.z, hj.z), 0.f);
                t.lo[k] = lo, t.hi[k] = hi;
                t.range[k] = make_int2(0, cnt);
                // in place: nobody else reads cluster i in this kernel (its partner only drops itself)
                c.id[i] = k;
                c.lo[i] = make_float4(lo.x, lo.y, lo.z, __int_as_float(cnt));
                c.hi[i] = hi;
            }
            keep[i] = survive ? 1 : 0;
        }
        const uint32_t m = __ballot_sync(0xffffffffu, survive);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = __popc(m);
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t s = 0;
#pragma unroll
            for (int w = 0; w < kTile / 32; ++w) s += s_warp[w];
            tile_count[tile] = s;
        }
    }
}

// exclusive scan of the tile counts in place (one CTA of 1024 threads), new cluster count -> n_next
__global__ void __launch_bounds__(1024) k_ploc_scan(const uint32_t *__restrict__ n_ptr, uint32_t *__restrict__ tile_count, uint32_t *__restrict__ n_next) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const int n = (int)*n_ptr;
    if (n <= 1) {
        if (threadIdx.x == 0) *n_next = (uint32_t)n;
        return;
    }
    const int tiles = (n + kTile - 1) / kTile;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < tiles; base += 1024) {
        const int k = base + threadIdx.x;
        const uint32_t v = k < tiles ? tile_count[k] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
            if ((int)(threadIdx.x & 31) >= d) incl += up;
        }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = s_warp[threadIdx.x], wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t up = __shfl_up_sync(0xffffffffu, wi, d);
                if ((int)threadIdx.x >= d) wi += up;
            }
            s_warp[threadIdx.x] = wi - w; // exclusive over warps
        }
        __syncthreads();
        const uint32_t carry = s_carry;
        if (k < tiles) tile_count[k] = carry + s_warp[threadIdx.x >> 5] + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + s_warp[31] + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_next = s_carry;
}

// survivors to the front of the other buffer, order kept
__global__ void __launch_bounds__(kTile) k_ploc_compact(const uint32_t *__restrict__ n_ptr, Clusters in, Clusters out, const uint8_t *__restrict__ keep,
                                                        const uint32_t *__restrict__ tile_offset) {
    __shared__ uint32_t s_warp[kTile / 32];
    const int n = (int)*n_ptr;
    if (n <= 1) return;
    const int tiles = (n + kTile - 1) / kTile;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int i = tile * kTile + threadIdx.x;
        const bool k = i < n && keep[i];
        const uint32_t m = __ballot_sync(0xffffffffu, k);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = __popc(m);
        __syncthreads();
        uint32_t before = 0;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) before += s_warp[w];
        if (k) {
            const uint32_t pos = tile_offset[tile] + before + __popc(m & ((1u << (threadIdx.x & 31)) - 1u));
            out.id[pos] = in.id[i], out.lo[pos] = in.lo[i], out.hi[pos] = in.hi[i];
        }
    }
}
}// namespace

// Returns the reference of the root of the binary tree (an internal node index; n == 1 is handled by the caller).
int build_binary_ploc(cudaStream_t st, uint32_t n, const float4 *box_lo, const float4 *box_hi, const uint32_t *sorted, int *left, int *right, int2 *range,
                      float4 *lo, float4 *hi, int radius, uint32_t *rounds_out) {
    radius = std::max(1, std::min(radius, kMaxRadius));
    DevBuf<int> id_a(n), id_b(n), nn(n);
    DevBuf<float4> lo_a(n), hi_a(n), lo_b(n), hi_b(n);
    DevBuf<uint8_t> keep(n);
    const uint32_t tiles = (n + kTile - 1) / kTile;
    DevBuf<uint32_t> tile_count(tiles), counters(3); // [0] nodes allocated, [1] / [2] cluster counts (ping-pong)
    uint32_t init[3] = { 0u, n, 0u };
    PB2_CUDA(cudaMemcpyAsync(counters.ptr, init, sizeof init, cudaMemcpyHostToDevice, st));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    Clusters a{ id_a.ptr, lo_a.ptr, hi_a.ptr }, b{ id_b.ptr, lo_b.ptr, hi_b.ptr };
    k_ploc_init<<<(unsigned)std::min<uint64_t>((n + 255) / 256, (uint64_t)sms * 8), 256, 0, st>>>(n, sorted, box_lo, box_hi, a);
    PB2_LAUNCH_CHECK();
    const Tree t{ left, right, range, lo, hi };
    uint32_t cur_n = n, rounds = 0;
    int cur = 0; // counters[1 + cur] holds the current cluster count
    while (cur_n > 1) {
        const unsigned grid = (unsigned)std::min<uint64_t>((cur_n + kTile - 1) / kTile, (uint64_t)sms * 8);
        const uint32_t *n_ptr = counters.ptr + 1 + cur;
        k_ploc_nn<<<grid, kTile, 0, st>>>(n_ptr, a, nn.ptr, radius);
        k_ploc_merge<<<grid, kTile, 0, st>>>(n_ptr, a, nn.ptr, t, counters.ptr, keep.ptr, tile_count.ptr);
        k_ploc_scan<<<1, 1024, 0, st>>>(n_ptr, tile_count.ptr, counters.ptr + 1 + (cur ^ 1));
        k_ploc_compact<<<grid, kTile, 0, st>>>(n_ptr, a, b, keep.ptr, tile_count.ptr);
        PB2_LAUNCH_CHECK();
        cur ^= 1;
        std::swap(a, b);
        uint32_t next_n = 0;
        PB2_CUDA(cudaMemcpyAsync(&next_n, counters.ptr + 1 + cur, sizeof next_n, cudaMemcpyDeviceToHost, st));
        PB2_CUDA(cudaStreamSynchronize(st));
        if (next_n >= cur_n || next_n == 0) throw std::runtime_error("pb2_bvh_build: clustering made no progress");
        cur_n = next_n;
        ++rounds;
    }
    int root = 0;
    PB2_CUDA(cudaMemcpyAsync(&root, a.id, sizeof root, cudaMemcpyDeviceToHost, st));
    PB2_CUDA(cudaStreamSynchronize(st));
    if (rounds_out) *rounds_out = rounds;
    return root;
}
}// namespace pb2
