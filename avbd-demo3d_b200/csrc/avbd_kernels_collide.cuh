// avbd_kernels_collide.cuh — broadphase (hashed uniform grid + large-body side
// list, half-space sweep with the SAT cull fused in) and narrowphase (manifold
// build with fused feature-id warm start) kernels.
//
// Replaces the reference's O(n^2) pair loop + per-step new/delete of Manifold
// nodes (solver.cpp:262-279, force.cpp:12-69) and Manifold::initialize
// (manifold.cpp:71-175).  Candidate set == reference semantics:
//   {sphere-overlap pairs} U {pairs that already own a manifold} \ {pairs linked by a user force}
// and A is always the higher creation index (the reference's list is newest-first).
#pragma once
#include <cooperative_groups.h>
#include "avbd_body.cuh"
#include "avbd_forces.cuh"
#include "avbd_launch.h"

namespace avbd {
namespace cg = cooperative_groups;


struct GridView {
    float cell;                         // edge length, >= 2.02 * largest small radius
    unsigned tableMask;                 // table size - 1 (power of two)
    unsigned* key; unsigned* keySorted; int* val; int* valSorted;
    int2* cellRange;                    // per bucket: [start, end) in the sorted order
    int2* sortedCell;                   // {packed cell (10 low bits of cx, cy, cz), world}
    float4* sortedPos;                  // pos.xyz, radius
    float4* sortedFrame;                // 6 per body: {ax0, h.x} {ax1, h.y} {ax2, h.z} {n0, raSelf0} {n1, raSelf1} {n2, raSelf2} (ObbFrame) — the fused SAT cull's operands
    float4* bodyFrame;                  // the same six, indexed by BODY (not by sorted position): what np_sat gathers per candidate (per-body sweep path)
    const int* largeList; const int* worldLargeStart;
};

// Bucket of a grid cell.  Cells are grouped in 4x4x4 blocks: the block is hashed, the position inside the block
// fills the low 6 bits, so the 64 cells of a block — and the bodies in them after the sort — stay contiguous and
// the neighbourhood sweep of neighbouring bodies touches neighbouring memory (a plain per-cell hash scatters them).
__device__ __forceinline__ unsigned cell_hash(int x, int y, int z, int w) {
    unsigned h = ((unsigned)(x >> 2) * 73856093u) ^ ((unsigned)(y >> 2) * 19349663u) ^ ((unsigned)(z >> 2) * 83492791u) ^ ((unsigned)w * 2654435761u);
    h ^= h >> 15; h *= 0x2c1b3c6du; h ^= h >> 12;
    return (h << 6) | (unsigned)((x & 3) | ((y & 3) << 2) | ((z & 3) << 4));
}
// Identity of a cell among the cells a sweep can confuse it with: two cells that share a bucket AND these 30 bits are
// >= 1024 cells apart, so the sphere test rejects whatever the mix-up lets through.
__device__ __forceinline__ int pack_cell(int x, int y, int z) { return (x & 1023) | ((y & 1023) << 10) | ((z & 1023) << 20); }
__device__ __forceinline__ float body_radius(float4 size) { return len(xyz(size)) * 0.5f; }   // rigid.cpp:28
__device__ __forceinline__ int3 cell_of(float4 pos, float cell) {
    return make_int3((int)floorf(pos.x / cell), (int)floorf(pos.y / cell), (int)floorf(pos.z / cell));
}

// K1a: bucket key per body.  Large bodies (flag) go to the sentinel bucket past the table.
// Also clears the bucket table and the step's counters (two memset nodes less in the launch chain of a small world's step).
// `zeroA / zeroB`: two more scratch arrays the step wants cleared before the collision kernels run (per-world diagnostics, the manifold
// build's scan tiles), as 4-byte words.
__global__ void bp_cells(BodyView b, GridView g, Counters* cnt, int* zeroA, int nZeroA, int* zeroB, int nZeroB) {
    cudaGridDependencySynchronize();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int stride = gridDim.x * blockDim.x;
    for (unsigned k = (unsigned)i; k <= g.tableMask; k += stride) g.cellRange[k] = make_int2(0, 0);
    for (int k = i; k < nZeroA; k += stride) zeroA[k] = 0;
    for (int k = i; k < nZeroB; k += stride) zeroB[k] = 0;
    if (i < (int)(sizeof(Counters) / sizeof(int))) reinterpret_cast<int*>(cnt)[i] = 0;
    if (i >= b.n) return;
    if (g.bodyFrame) {          // the SAT frames np_sat gathers by body: written here, in body order (coalesced), not after the sort
        const ObbFrame f = make_obb_frame(xyz(b.pose[i].pos), quat(b.pose[i].rot), xyz(b.size[i]));
        float4* o = g.bodyFrame + 6 * (size_t)i;
        o[0] = f4(f.ax[0], f.h.x); o[1] = f4(f.ax[1], f.h.y); o[2] = f4(f.ax[2], f.h.z);
        o[3] = f4(f.n[0], f.raSelf[0]); o[4] = f4(f.n[1], f.raSelf[1]); o[5] = f4(f.n[2], f.raSelf[2]);
    }
    unsigned k = g.tableMask + 1u;
    if (!(b.flags[i] & kLarge)) {
        int3 c = cell_of(b.pose[i].pos, g.cell);
        k = cell_hash(c.x, c.y, c.z, b.worldId[i]) & g.tableMask;
    }
    g.key[i] = k;
    g.val[i] = i;
}

// K1b: bucket boundaries in sorted order + sorted copies for the pair sweep.  cellRange must be zeroed first.
__global__ void bp_cell_bounds(BodyView b, GridView g) {
    cudaGridDependencySynchronize();
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= b.n) return;
    unsigned k = g.keySorted[p];
    int i = g.valSorted[p];
    float4 pos = b.pose[i].pos;
    float r = body_radius(b.size[i]);
    int3 c = cell_of(pos, g.cell);
    g.sortedPos[p] = make_float4(pos.x, pos.y, pos.z, r);
    if (g.sortedFrame) {
        const ObbFrame f = make_obb_frame(xyz(pos), quat(b.pose[i].rot), xyz(b.size[i]));
        float4* o = g.sortedFrame + 6 * (size_t)p;
        o[0] = f4(f.ax[0], f.h.x); o[1] = f4(f.ax[1], f.h.y); o[2] = f4(f.ax[2], f.h.z);
        o[3] = f4(f.n[0], f.raSelf[0]); o[4] = f4(f.n[1], f.raSelf[1]); o[5] = f4(f.n[2], f.raSelf[2]);
    }
    g.sortedCell[p] = make_int2(pack_cell(c.x, c.y, c.z), b.worldId[i]);
    if (k > g.tableMask) return;
    if (p == 0 || g.keySorted[p - 1] != k) g.cellRange[k].x = p;
    if (p == b.n - 1 || g.keySorted[p + 1] != k) g.cellRange[k].y = p + 1;
}

// codeShift > 0: the pair's code (the SAT's winning axis) rides in the key's bits from codeShift up — above the 2 * keyShift bits the
// sort looks at — so the survivors are sorted keys-only (8 bytes per item through every radix pass instead of 12); codes is then unused.
struct PairSink {
    unsigned long long* keys; int* codes; int cap; int keyShift; int* count; Counters* cnt; int overflowBit; int codeShift;
};
__device__ __forceinline__ unsigned long long sink_key(const PairSink& s, unsigned long long key, int code) {
    return s.codeShift > 0 ? key | ((unsigned long long)(unsigned)code << s.codeShift) : key;
}

// Warp-aggregated append: the lanes that reach this call together take one
// atomic for the group and write a contiguous run.
__device__ __forceinline__ void emit_pair(const PairSink& s, unsigned long long key, int code) {
    cg::coalesced_group grp = cg::coalesced_threads();
    int base = 0;
    if (grp.thread_rank() == 0) base = atomicAdd(s.count, (int)grp.size());
    base = grp.shfl(base, 0);
    int idx = base + (int)grp.thread_rank();
    if (idx < s.cap) {
        s.keys[idx] = sink_key(s, key, code);
        if (s.codes) s.codes[idx] = code;
    } else atomicOr(&s.cnt->overflow, s.overflowBit);
}
__device__ __forceinline__ unsigned long long pair_key(int a, int b, int keyShift) {     // a > b
    return ((unsigned long long)(unsigned)a << keyShift) | (unsigned long long)(unsigned)b;
}

// The reference's test, solver.cpp:264-266.  Symmetric bit for bit in its arguments.
__device__ __forceinline__ bool spheres_overlap(float4 pa, float4 pb) {
    V3 dp = xyz(pa) - xyz(pb);
    float r = pa.w + pb.w;
    return dot(dp, dp) <= r * r;
}

__device__ __forceinline__ int find_key(const unsigned long long* keys, int n, unsigned long long k) {
    int lo = 0, hi = n;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (keys[mid] < k) lo = mid + 1; else hi = mid; }
    return (lo < n && keys[lo] == k) ? lo : -1;
}
// First index in the sorted keys[0, n) whose key is not below k, found by the whole warp: every step probes 32 evenly spaced keys
// of the live range, so a list of two million keys takes 5 dependent loads instead of the 21 of a binary search.  k is warp-uniform.
__device__ __forceinline__ int warp_lower_bound(const unsigned long long* keys, int n, unsigned long long k, int lane) {
    int lo = 0, hi = n;
    while (hi - lo > 32) {
        const int stride = (hi - lo + 32) / 33;
        const long long q = (long long)lo + (long long)(lane + 1) * stride - 1;
        const bool below = q < hi && keys[q] < k;
        const int c = __popc(__ballot_sync(0xffffffffu, below));              // sorted: the probes below k are the first c
        const int nlo = c == 0 ? lo : lo + c * stride;
        const long long nhi = c == 32 ? hi : (long long)lo + (long long)(c + 1) * stride - 1;
        lo = nlo; hi = nhi < hi ? (int)nhi : hi;
    }
    const bool below = lo + lane < hi && keys[lo + lane] < k;
    return lo + __popc(__ballot_sync(0xffffffffu, below));
}
// Slot of key k in last step's sorted manifold keys, -1 if it had none.  The new list is sorted too and mostly the same pairs, so the
// warp's keys sit (nearly) side by side in the old list: the warp locates its first key together, each lane then looks at the slot the
// same distance further on — one coalesced load, a hit for every pair when nothing appeared or vanished in between — and only the lanes
// that miss search on, in a window first.  Same result as find_key for every lane.  Called by the whole warp; k = 0 / has = false for idle lanes.
__device__ __forceinline__ int warp_find_keys(const unsigned long long* keys, int n, unsigned long long k, bool has, int lane) {
    if (n <= 0) return -1;
    const unsigned who = __ballot_sync(0xffffffffu, has);
    if (!who) return -1;
    const int lead = __ffs(who) - 1;
    const unsigned long long k0 = __shfl_sync(0xffffffffu, k, lead);
    const int p0 = warp_lower_bound(keys, n, k0, lane);
    if (!has) return -1;
    const int guess = p0 + (lane - lead);
    if (guess < n) {
        const unsigned long long g = keys[guess];
        if (g == k) return guess;
        if (g > k) {                                                       // between the lead's slot and the guess
            int lo = p0, hi = guess;
            while (lo < hi) { int mid = (lo + hi) >> 1; if (keys[mid] < k) lo = mid + 1; else hi = mid; }
            return (lo < n && keys[lo] == k) ? lo : -1;
        }
    }
    int lo = guess < n ? guess + 1 : p0, hi = lo + 64 < n ? lo + 64 : n;      // past the guess: a window, the rest of the list if the window ends below k
    if (guess >= n) { lo = p0; hi = n; }
    else if (hi < n && keys[hi - 1] < k) { lo = hi; hi = n; }
    while (lo < hi) { int mid = (lo + hi) >> 1; if (keys[mid] < k) lo = mid + 1; else hi = mid; }
    return (lo < n && keys[lo] == k) ? lo : -1;
}

// K1c: small-vs-small sphere-overlap pairs (solver.cpp:262-266).  Every unordered pair of cells is swept from ONE side:
// a body looks at its own cell (later sorted positions only) and at the 13 neighbour cells that follow its own in
// (z, y, x) order — half the lookups and sphere tests of a full 27-cell sweep.  16 lanes per body, one neighbour cell
// per lane (14 used): the hash lookup -> bucket range -> bucket walk chain is one cell long instead of fourteen, which
// is what this latency-bound kernel is made of.  Pairs come out unsorted; A (the high half of the key) is the higher
// creation index: the reference's list is newest first, so its earlier-in-list body is the higher id.
// Appends to the pair list are staged per block in shared memory and flushed with ONE global atomic per block: a
// per-warp atomicAdd on the single list counter serialises at its L2 slice (millions of same-address atomics per step
// were most of this kernel's time).
constexpr int kSweepStage = 1024;          // staged keys per block of 16 bodies (about 150 expected on a dense pile)
__global__ void __launch_bounds__(kThreads) bp_sweep(BodyView b, GridView g, PairSink sink) {
    cudaGridDependencySynchronize();
    __shared__ unsigned long long sKeys[kSweepStage];
    __shared__ int sCount, sBase;
    if (threadIdx.x == 0) sCount = 0;
    __syncthreads();
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int p = t >> 4, k = t & 15;
    if (p < b.n && k < 14 && g.keySorted[p] <= g.tableMask) {
        // k = 0: own cell; 1: (+1,0,0); 2..4: (dx,+1,0); 5..13: (dx,dy,+1)
        int dx, dy, dz;
        if (k < 2) { dx = k; dy = 0; dz = 0; }
        else if (k < 5) { dx = k - 3; dy = 1; dz = 0; }
        else { dx = (k - 5) % 3 - 1; dy = (k - 5) / 3 - 1; dz = 1; }
        int i = g.valSorted[p];
        float4 pi = g.sortedPos[p];
        int world = g.sortedCell[p].y;
        int3 ci = cell_of(pi, g.cell);
        int nx = ci.x + dx, ny = ci.y + dy, nz = ci.z + dz;
        int2 r = g.cellRange[cell_hash(nx, ny, nz, world) & g.tableMask];
        int want = pack_cell(nx, ny, nz);
        int q = (k == 0) ? (r.x > p + 1 ? r.x : p + 1) : r.x;
        for (; q < r.y; ++q) {
            int2 cq = g.sortedCell[q];
            if (cq.x != want || cq.y != world) continue;
            if (!spheres_overlap(pi, g.sortedPos[q])) continue;
            int j = g.valSorted[q];
            unsigned long long key = i > j ? pair_key(i, j, sink.keyShift) : pair_key(j, i, sink.keyShift);
            cg::coalesced_group grp = cg::coalesced_threads();
            int base = 0;
            if (grp.thread_rank() == 0) base = atomicAdd(&sCount, (int)grp.size());
            int idx = grp.shfl(base, 0) + (int)grp.thread_rank();
            if (idx < kSweepStage) sKeys[idx] = key;
            else emit_pair(sink, key, 1);                  // stage full (a very dense neighbourhood): straight to the list
        }
    }
    __syncthreads();
    int staged = sCount < kSweepStage ? sCount : kSweepStage;
    if (threadIdx.x == 0 && staged > 0) sBase = atomicAdd(sink.count, staged);
    __syncthreads();
    for (int e = threadIdx.x; e < staged; e += blockDim.x) {
        int idx = sBase + e;
        if (idx < sink.cap) sink.keys[idx] = sKeys[e];
        else atomicOr(&sink.cnt->overflow, sink.overflowBit);
    }
}

// K1c' (the step's sweep): the same pair set as bp_sweep, enumerated per CELL instead of per body, with the SAT cull
// (collision.cpp:420-468) fused in.  A warp owns 32 consecutive bodies of the cell-sorted order and walks the runs of
// bodies that share a cell: 14 lanes look up the 14 neighbour buckets of the run's cell ONCE, the warp then reads the
// candidates of those buckets as contiguous pieces of the sorted arrays (one lane per candidate) and tests each against
// the run's bodies out of registers.  Against the per-body sweep that is ~5x fewer bucket lookups and candidate loads on
// a pile with ~5 bodies per cell.  Sphere hits are parked in a per-warp queue; whenever 32 are waiting they run the 6
// face axes with every lane busy, the pairs still alive wait in a second queue for the 9 edge axes, and the survivors
// {key, winning axis} are staged per block and appended to the candidate list with one global atomic per block — the
// 9 sphere pairs per body of a dense pile never travel through memory.  SAT = false stops at the sphere pairs (stage API).
constexpr int kCellStage = 1536;           // staged survivors per block of 256 bodies (about 800 expected on a dense pile)
constexpr int kCellSub = 8;                // bodies of a run handled per pass over its candidates (bounds the hit queue)
constexpr int kCellQ1 = 32 + kCellSub * 32;
template <bool SAT>
__global__ void __launch_bounds__(kThreads) bp_sweep_cells(BodyView b, GridView g, const unsigned long long* excl, int nExcl, PairSink sink, int bpw) {
    cudaGridDependencySynchronize();
    constexpr int W = kThreads / 32;
    constexpr unsigned kFull = 0xffffffffu;
    __shared__ unsigned long long sKeys[kCellStage];
    __shared__ int sCodes[SAT ? kCellStage : 1];
    __shared__ int2 sQ1[W][kCellQ1];
    __shared__ int2 sQ2[SAT ? W : 1][64];
    __shared__ float sQ2sep[SAT ? W : 1][64];
    __shared__ int sQ2k[SAT ? W : 1][64];
    __shared__ int sStart[W][16], sOff[W][16], sWant[W][16];
    __shared__ int sCount, sBase, sHits;
    if (threadIdx.x == 0) { sCount = 0; sHits = 0; }
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const int base = (blockIdx.x * W + w) * bpw;         // bpw bodies per warp: 32, fewer for a small world (a warp's runs are walked one after another)
    const int p = base + lane;
    bool valid = lane < bpw && p < b.n;
    const unsigned myKey = valid ? g.keySorted[p] : 0xffffffffu;
    valid = valid && myKey <= g.tableMask;                 // small bodies sort first: the valid lanes are a prefix of the warp
    const float4 myPos = valid ? g.sortedPos[p] : make_float4(0.f, 0.f, 0.f, 0.f);
    const int2 myCell = valid ? g.sortedCell[p] : make_int2(-1, -1);
    const int3 ci = cell_of(myPos, g.cell);
    const unsigned upKey = __shfl_up_sync(kFull, myKey, 1);
    const int upCell = __shfl_up_sync(kFull, myCell.x, 1), upWorld = __shfl_up_sync(kFull, myCell.y, 1);
    unsigned startMask = __ballot_sync(kFull, valid && (lane == 0 || upKey != myKey || upCell != myCell.x || upWorld != myCell.y));
    const int nValid = __popc(__ballot_sync(kFull, valid));
    int n1 = 0, n2 = 0, hits = 0;

    auto stage = [&](bool has, unsigned long long key, int code) {
        const unsigned m = __ballot_sync(kFull, has);
        if (!m) return;
        int at = 0;
        if (lane == 0) at = atomicAdd(&sCount, __popc(m));
        at = __shfl_sync(kFull, at, 0) + __popc(m & lt);
        if (has) {
            if (at < kCellStage) { sKeys[at] = key; if (SAT) sCodes[at] = code; }
            else emit_pair(sink, key, code);               // stage full (a very dense neighbourhood): straight to the list
        }
    };
    auto load_frame = [&](int q, bool faces) {
        ObbFrame F;
        const float4 ps = g.sortedPos[q];
        const float4* fr = g.sortedFrame + 6 * (size_t)q;
        const float4 a0 = fr[0], a1 = fr[1], a2 = fr[2];
        F.c = xyz(ps); F.h = mk3(a0.w, a1.w, a2.w); F.ax[0] = xyz(a0); F.ax[1] = xyz(a1); F.ax[2] = xyz(a2);
        if (faces) {
            const float4 n0 = fr[3], n1 = fr[4], n2 = fr[5];
            F.n[0] = xyz(n0); F.n[1] = xyz(n1); F.n[2] = xyz(n2); F.raSelf[0] = n0.w; F.raSelf[1] = n1.w; F.raSelf[2] = n2.w;
        }
        return F;
    };

    int cur = 0, runEnd = 0;                               // lanes [cur, runEnd) of the run being walked
    for (;;) {
        bool final = false;
        if (cur >= runEnd) {
            if (!startMask) final = true;
            else {
                cur = __ffs(startMask) - 1; startMask &= startMask - 1;
                runEnd = startMask ? __ffs(startMask) - 1 : 32;
                if (runEnd > nValid) runEnd = nValid;
            }
        }
        const int subEnd = cur + kCellSub < runEnd ? cur + kCellSub : runEnd;
        int total = 0, world = 0;
        if (!final) {
            const int cx = __shfl_sync(kFull, ci.x, cur), cy = __shfl_sync(kFull, ci.y, cur), cz = __shfl_sync(kFull, ci.z, cur);
            world = __shfl_sync(kFull, myCell.y, cur);
            int len = 0, st = 0, want = 0;
            if (lane < 14) {
                // lane 0: own cell (later sorted positions only); 1: (+1,0,0); 2..4: (dx,+1,0); 5..13: (dx,dy,+1)
                int dx, dy, dz;
                if (lane < 2) { dx = lane; dy = 0; dz = 0; }
                else if (lane < 5) { dx = lane - 3; dy = 1; dz = 0; }
                else { dx = (lane - 5) % 3 - 1; dy = (lane - 5) / 3 - 1; dz = 1; }
                const int nx = cx + dx, ny = cy + dy, nz = cz + dz;
                const int2 r = g.cellRange[cell_hash(nx, ny, nz, world) & g.tableMask];
                want = pack_cell(nx, ny, nz);
                st = (lane == 0 && r.x < base + cur + 1) ? base + cur + 1 : r.x;
                len = r.y > st ? r.y - st : 0;
            }
            int incl = len;
#pragma unroll
            for (int d = 1; d < 16; d <<= 1) { const int up = __shfl_up_sync(kFull, incl, d); if (lane >= d) incl += up; }
            total = __shfl_sync(kFull, incl, 15);
            __syncwarp();                                  // the previous pass is done with the tables
            if (lane < 16) { sStart[w][lane] = st; sOff[w][lane] = lane < 14 ? incl - len : 0x7fffffff; sWant[w][lane] = want; }
            __syncwarp();
        }
        int t0 = 0;
        do {
            if (!final) {
                const int t = t0 + lane;
                const bool act = t < total;
                int k = 0;
                if (act) {                                                 // the piece t falls in: the LAST one starting at or before t (empty pieces share their successor's offset)
                    k = sOff[w][8] <= t ? 8 : 0;
                    k += sOff[w][k + 4] <= t ? 4 : 0; k += sOff[w][k + 2] <= t ? 2 : 0; k += sOff[w][k + 1] <= t ? 1 : 0;
                }
                const int q = act ? sStart[w][k] + (t - sOff[w][k]) : 0;
                bool match = false; float4 pq = make_float4(0.f, 0.f, 0.f, 0.f);
                if (act) {
                    const int2 cq = g.sortedCell[q];
                    pq = g.sortedPos[q];
                    match = cq.x == sWant[w][k] && cq.y == world;
                }
                for (int r = cur; r < subEnd; ++r) {
                    const float4 pr = make_float4(__shfl_sync(kFull, myPos.x, r), __shfl_sync(kFull, myPos.y, r), __shfl_sync(kFull, myPos.z, r), __shfl_sync(kFull, myPos.w, r));
                    const bool hit = match && (k != 0 || q > base + r) && spheres_overlap(pr, pq);
                    const unsigned m = __ballot_sync(kFull, hit);
                    if (hit) sQ1[w][n1 + __popc(m & lt)] = make_int2(base + r, q);
                    n1 += __popc(m);
                }
                __syncwarp();
            }
            // drain the queues: 32 pairs at a time, the rest when the walk is over
            for (;;) {
                if (SAT && (n2 >= 32 || (final && n1 == 0 && n2 > 0))) {
                    const int c = n2 < 32 ? n2 : 32;
                    int code = 0; unsigned long long key = 0ull;
                    if (lane < c) {
                        const int e = n2 - c + lane;
                        const int2 pr = sQ2[SAT ? w : 0][e];
                        const int fk = sQ2k[SAT ? w : 0][e];
                        SatFaces f{fk >= 0, fk >= 0 ? sQ2sep[SAT ? w : 0][e] : -FLT_MAX, fk >= 0 ? fk : 0};
                        const Obb A = frame_obb(load_frame(pr.x, false)), B = frame_obb(load_frame(pr.y, false));
                        code = sat_edges(A, B, f);
                        key = pair_key(g.valSorted[pr.x], g.valSorted[pr.y], sink.keyShift);
                    }
                    __syncwarp();
                    n2 -= c;
                    stage(code != 0, key, code);
                    continue;
                }
                if (n1 >= 32 || (final && n1 > 0)) {
                    const int c = n1 < 32 ? n1 : 32;
                    bool alive = false; SatFaces f{false, 0.0f, 0}; unsigned long long key = 0ull; int hi = 0, lo = 0;
                    if (lane < c) {
                        const int2 pr = sQ1[w][n1 - c + lane];
                        const int i = g.valSorted[pr.x], j = g.valSorted[pr.y];
                        hi = i > j ? pr.x : pr.y; lo = i > j ? pr.y : pr.x;            // A is the higher creation index
                        key = i > j ? pair_key(i, j, sink.keyShift) : pair_key(j, i, sink.keyShift);
                        if (SAT && !(nExcl > 0 && find_key(excl, nExcl, key) >= 0)) {
                            const ObbFrame A = load_frame(hi, true), B = load_frame(lo, true);
                            alive = sat_faces_frames(A, B, f);
                        }
                    }
                    __syncwarp();
                    n1 -= c; hits += c;
                    if (!SAT) { stage(lane < c, key, 1); continue; }
                    const unsigned m = __ballot_sync(kFull, alive);
                    if (alive) {
                        const int e = n2 + __popc(m & lt);
                        sQ2[SAT ? w : 0][e] = make_int2(hi, lo); sQ2sep[SAT ? w : 0][e] = f.sep; sQ2k[SAT ? w : 0][e] = f.valid ? f.k : -1;
                    }
                    n2 += __popc(m);
                    __syncwarp();
                    continue;
                }
                break;
            }
            t0 += 32;
        } while (t0 < total);
        if (final) break;
        cur = subEnd;
    }
    if (lane == 0 && hits) atomicAdd(&sHits, hits);
    __syncthreads();
    const int staged = sCount < kCellStage ? sCount : kCellStage;
    if (threadIdx.x == 0) {
        if (staged > 0) sBase = atomicAdd(sink.count, staged);
        if (SAT && sHits) atomicAdd(&sink.cnt->nSphere, sHits);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < staged; e += blockDim.x) {
        const int idx = sBase + e;
        if (idx < sink.cap) { sink.keys[idx] = SAT ? sink_key(sink, sKeys[e], sCodes[e]) : sKeys[e]; if (SAT && sink.codes) sink.codes[idx] = sCodes[e]; }
        else atomicOr(&sink.cnt->overflow, sink.overflowBit);
    }
}

// K1d + K2 in one launch.
// K1d (thread j < n): every body against the large bodies of its own world.
// K2 (thread m < nOld): manifolds that survived last step persist as candidates whether or not their spheres still overlap
// (solver.cpp:274-279 only deletes on initialize()==false).  Pairs whose spheres DO overlap were emitted by the sweeps; this adds
// the rest (rare), so every candidate appears exactly once.  nOld = 0: the sphere pairs only (stage API).
__global__ void bp_side_pairs(BodyView b, GridView g, ManifoldSet old, int nOld, PairSink sink, int anyLarge) {
    cudaGridDependencySynchronize();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < nOld) {
        int4 h = old.hdr[j];
        if (h.z > 0) {
            float4 pa = b.pose[h.x].pos; pa.w = body_radius(b.size[h.x]);
            float4 pb = b.pose[h.y].pos; pb.w = body_radius(b.size[h.y]);
            if (!spheres_overlap(pa, pb)) emit_pair(sink, old.key[j], 1);
        }
    }
    if (!anyLarge || j >= b.n) return;
    int w = b.worldId[j];
    int t0 = g.worldLargeStart[w], t1 = g.worldLargeStart[w + 1];
    if (t0 == t1) return;
    bool jLarge = (b.flags[j] & kLarge) != 0;
    float4 pj = b.pose[j].pos; pj.w = body_radius(b.size[j]);
    for (int t = t0; t < t1; ++t) {
        int l = g.largeList[t];
        if (l == j) continue;
        if (jLarge && j > l) continue;            // large-large pairs: emitted once, by the lower index
        float4 pl = b.pose[l].pos; pl.w = body_radius(b.size[l]);
        if (spheres_overlap(pj, pl)) emit_pair(sink, j > l ? pair_key(j, l, sink.keyShift) : pair_key(l, j, sink.keyShift), 1);
    }
}

// K3a: SAT cull (collision.cpp:420-468).  Candidates are read in emission order — which follows the cell-sorted body order, so
// neighbouring lanes gather neighbouring bodies.  Survivors {key, winning axis} go to `out`; only they (about a fifth of the candidates
// on a dense pile) are sorted.  The candidate count is read on the device.
// Persistent warps, no block barrier: a warp takes 32 candidates at a time and runs the 6 face axes, one candidate per lane (most
// candidates of a pile die there); the pairs still alive wait in the warp's queue in shared memory, and whenever 32 are waiting the 9
// edge axes run with every lane busy; survivors collect in a second per-warp queue and leave 32 at a time with one atomic on the list
// counter.  (The first form of this kernel compacted per BLOCK between two __syncthreads: ncu showed its warps waiting at the barrier
// 7 cycles per issued instruction with the issue slots 41 % busy — a chain of dependent gathers per phase and four resident blocks.)
// `frames` (6 float4 per body, written by bp_cell_bounds this step): the face axes use each body's normalised axes and own
// projection radii from there instead of re-deriving them per candidate — every body sits in ~10 candidates of a dense pile, and the
// square root + three IEEE divisions per axis were most of the face phase's instructions.  sat_faces_frames == sat_faces bit for bit.
constexpr int kSatQueue = 64;
__global__ void __launch_bounds__(kThreads, 4) np_sat(BodyView b, const float4* __restrict__ frames, const unsigned long long* cand, const int* nCand, int cap, int keyShift,
                                                   const unsigned long long* excl, int nExcl, PairSink out) {
    cudaGridDependencySynchronize();
    constexpr int W = kThreads / 32;
    constexpr unsigned kFull = 0xffffffffu;
    __shared__ unsigned long long sAliveKey[W][kSatQueue], sOutKey[W][kSatQueue];
    __shared__ float sAliveSep[W][kSatQueue];
    __shared__ int sAliveK[W][kSatQueue], sOutCode[W][kSatQueue];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    int n = *nCand; if (n > cap) n = cap;
    auto load_obb = [&](int i) {
        Obb o;
        const float4 ps = b.pose[i].pos;
        const float4* fr = frames + 6 * (size_t)i;
        const float4 a0 = fr[0], a1 = fr[1], a2 = fr[2];
        o.c = xyz(ps); o.h = mk3(a0.w, a1.w, a2.w); o.ax[0] = xyz(a0); o.ax[1] = xyz(a1); o.ax[2] = xyz(a2);
        return o;
    };
    auto load_frame = [&](int i) {
        ObbFrame F;
        const Obb o = load_obb(i);
        const float4* fr = frames + 6 * (size_t)i;
        const float4 n0 = fr[3], n1 = fr[4], n2 = fr[5];
        F.c = o.c; F.h = o.h; F.ax[0] = o.ax[0]; F.ax[1] = o.ax[1]; F.ax[2] = o.ax[2];
        F.n[0] = xyz(n0); F.n[1] = xyz(n1); F.n[2] = xyz(n2); F.raSelf[0] = n0.w; F.raSelf[1] = n1.w; F.raSelf[2] = n2.w;
        return F;
    };
    int nAlive = 0, nOut = 0;                                  // queue fill levels (warp-uniform)
    auto flush_out = [&](int count) {                         // the LAST `count` survivors of the queue leave
        int base = 0;
        if (lane == 0) base = atomicAdd(out.count, count);
        base = __shfl_sync(kFull, base, 0);
        if (lane < count) {
            const int idx = base + lane, e = nOut - count + lane;
            if (idx < out.cap) { out.keys[idx] = sink_key(out, sOutKey[w][e], sOutCode[w][e]); if (out.codes) out.codes[idx] = sOutCode[w][e]; }
            else atomicOr(&out.cnt->overflow, out.overflowBit);
        }
        __syncwarp();
        nOut -= count;
    };
    auto run_edges = [&](int count) {                         // edge axes of the LAST `count` (<= 32) pairs of the alive queue
        int code = 0; unsigned long long k = 0ull;
        if (lane < count) {
            const int e = nAlive - count + lane;
            k = sAliveKey[w][e];
            const int fk = sAliveK[w][e];
            const SatFaces g{fk >= 0, fk >= 0 ? sAliveSep[w][e] : -FLT_MAX, fk >= 0 ? fk : 0};
            const int a = (int)(k >> keyShift), c = (int)(k & ((1ull << keyShift) - 1ull));
            code = sat_edges(load_obb(a), load_obb(c), g);
        }
        __syncwarp();
        nAlive -= count;
        const unsigned vote = __ballot_sync(kFull, code != 0);
        if (code) { const int e = nOut + __popc(vote & lt); sOutKey[w][e] = k; sOutCode[w][e] = code; }
        __syncwarp();
        nOut += __popc(vote);
        if (nOut >= 32) flush_out(32);
    };
    const int stride = gridDim.x * kThreads;
    for (int base = (blockIdx.x * W + w) * 32; base < n; base += stride) {
        const int p = base + lane;
        bool alive = false; SatFaces f{false, 0.0f, 0}; unsigned long long k = 0ull;
        if (p < n) {
            k = cand[p];
            if (!(nExcl > 0 && find_key(excl, nExcl, k) >= 0)) {
                const int a = (int)(k >> keyShift), c = (int)(k & ((1ull << keyShift) - 1ull));
                alive = sat_faces_frames(load_frame(a), load_frame(c), f);
            }
        }
        const unsigned av = __ballot_sync(kFull, alive);
        if (alive) { const int e = nAlive + __popc(av & lt); sAliveKey[w][e] = k; sAliveSep[w][e] = f.sep; sAliveK[w][e] = f.valid ? f.k : -1; }
        __syncwarp();
        nAlive += __popc(av);
        if (nAlive >= 32) run_edges(32);
    }
    if (nAlive > 0) run_edges(nAlive);
    if (nOut > 0) flush_out(nOut);
}

__device__ __forceinline__ ContactState load_contact(const ManifoldSet& ms, int ci) {
    ContactLP q = ms.lp[ci];
    return unpack_contact(ms.cA[ci], ms.cB[ci], ms.cN[ci], q.l, q.p);
}

// K3b: build the manifold of each surviving pair at its final (key-sorted) slot, carrying lambda / penalty / stick
// anchors over from last step's manifold of the same pair, then apply the per-step warm-start decay.  One thread per
// manifold; nothing is held in per-thread arrays (the clipper's polygons and the raw contacts live in shared memory, one
// column per thread).  Contacts are written DENSELY, straight to their final place, in manifold order: once a block's raw
// contacts are parked its contact counts are scanned, and the block learns where its range starts from a single-pass chained scan
// over the blocks (decoupled look-back: every block publishes its total at once, then adds up its predecessors' totals until it
// meets one whose inclusive prefix is already known) — cstart[m] = where manifold m's contacts start.  No staging copy, no
// separate scan or compaction pass, and the placement is deterministic (the graph of a step whose topology did not change is
// re-used, and its visit lists hold contact indices).
constexpr int kBuildThreads = 64;
// tile[b] = (status << 32) | value; status 0 not ready, 1 the block's own total, 2 the inclusive prefix up to and including the block.
// Blocks are dispatched in index order, so a predecessor a block waits on is always resident or done.
// Called by the whole first warp.  Publishing and looking back are separate calls: a block publishes its total as soon as it knows
// it, does the work that does not need its base, and only then adds up its predecessors — which have long published by then.
__device__ __forceinline__ void chained_publish(unsigned long long* tile, int block, int total) {
    if ((threadIdx.x & 31) == 0) atomicExch(&tile[block], ((block == 0 ? 2ull : 1ull) << 32) | (unsigned)total);
}
__device__ __forceinline__ int chained_lookback(unsigned long long* tile, int block, int total) {
    const int lane = threadIdx.x & 31;
    if (block == 0) return 0;
    int exclusive = 0;
    for (int start = block - 1;; start -= 32) {
        const int idx = start - lane;
        unsigned long long v;
        for (;;) {
            v = idx >= 0 ? *(volatile unsigned long long*)&tile[idx] : (2ull << 32);
            if (!__any_sync(0xffffffffu, (v >> 32) == 0ull)) break;
            __nanosleep(100);                                             // a predecessor is still clipping: leave the issue slots to the warps that work
        }
        const unsigned incl = __ballot_sync(0xffffffffu, (v >> 32) == 2ull);
        const int stop = incl ? __ffs(incl) - 1 : 31;                 // nearest predecessor whose inclusive prefix is known
        int val = lane <= stop ? (int)(unsigned)(v & 0xffffffffull) : 0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) val += __shfl_xor_sync(0xffffffffu, val, d);
        exclusive += val;
        if (incl) break;
    }
    if (lane == 0) atomicExch(&tile[block], (2ull << 32) | (unsigned)(exclusive + total));
    return exclusive;
}
__global__ void __launch_bounds__(kBuildThreads) np_build(BodyView b, const unsigned long long* cand, const int* info, int nSurvive,
                                                          int keyShift, ManifoldSet old, int nOld, ManifoldSet out,
                                                          SolveParams prm, Counters* cnt, unsigned long long* tile) {
    cudaGridDependencySynchronize();
    extern __shared__ float sPoly[];
    __shared__ int sWarpTotal[kBuildThreads / 32], sBase;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    const bool inRange = s < nSurvive;
    const unsigned long long pairMask = (1ull << (2 * keyShift)) - 1ull;          // info == nullptr: the SAT code rides above the pair bits
    const unsigned long long kc = inRange ? cand[s] : 0ull;
    unsigned long long k = kc & pairMask;
    const int satCode = inRange ? (info ? info[s] : (int)(kc >> (2 * keyShift))) : 0;
    int a = (int)(k >> keyShift), c = (int)(k & ((1ull << keyShift) - 1ull));
    bool build = inRange;
    if (inRange) {
        out.key[s] = k;
        if (s > 0 && (cand[s - 1] & pairMask) == k) {      // the sweeps emit every pair once; a repeat would double a manifold: keep a dead slot, flag it
            atomicOr(&cnt->overflow, 8);
            build = false;
        }
    }
    BodyPose pa{}, pb{}; float4 sa = make_float4(0.f, 0.f, 0.f, 0.f), sb = sa;
    V3 posA = zero3(), posB = zero3(); Q4 rotA = qid(), rotB = qid();
    int oldN = 0, oldBase = 0, slot = -1;
    int oldFeat[4] = {0, 0, 0, 0};
    // Two passes so the expensive half runs converged: the builder's loops (clip, dedupe) reach their emit at different
    // trip counts in every lane, so finishing a contact inside emit would serialise the lanes; emit only parks the raw contact
    // (10 floats) in the clipper's second polygon buffer, which is free once the clipped polygon sits in the first, and a
    // fixed-trip loop then runs Manifold::initialize's per-contact part for all lanes together, in emission order.
    PolyShared poly{sPoly + threadIdx.x, (int)blockDim.x};
    float* raw = sPoly + threadIdx.x + (size_t)(kMaxPoly * 3) * blockDim.x;            // buffer 1, this thread's column
    const int rs = (int)blockDim.x;
    int n = 0;
    slot = warp_find_keys(old.key, nOld, k, build, lane);
    if (build) {
        pa = b.pose[a]; pb = b.pose[c];
        sa = b.size[a]; sb = b.size[c];
        posA = xyz(pa.pos); posB = xyz(pb.pos); rotA = quat(pa.rot); rotB = quat(pb.rot);
        if (slot >= 0) {
            oldN = old.hdr[slot].z; oldBase = old.cstart[slot];
#pragma unroll
            for (int j = 0; j < 4; ++j) if (j < oldN) oldFeat[j] = f2i(old.lp[oldBase + j].p.w);
        }
        auto emit = [&](int feature, V3 rA, V3 rB, V3 normal) {
            float* q = raw + (size_t)(n * 10) * rs;
            q[0] = i2f(feature); q[rs] = rA.x; q[2 * rs] = rA.y; q[3 * rs] = rA.z; q[4 * rs] = rB.x; q[5 * rs] = rB.y; q[6 * rs] = rB.z;
            q[7 * rs] = normal.x; q[8 * rs] = normal.y; q[9 * rs] = normal.z;
            ++n;
        };
        build_contacts_emit(posA, rotA, xyz(sa), posB, rotB, xyz(sb), satCode, poly, emit);
    }
    // this block's range of the dense contact arrays: warp scan of the counts, block total PUBLISHED at once (the blocks after this one
    // can go on); the block's own base — the sum of its predecessors' totals — is only needed by the stores, so it is looked up after the
    // per-contact work below instead of before it (ncu, 1M-box grid: a seventh of this kernel's instructions were the first warp polling
    // its predecessors while the other three sat at the barrier).
    int incl = n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int up = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += up; }
    if (lane == 31) sWarpTotal[warp] = incl;
    __syncthreads();
    int blockTotal = 0;
    if (warp == 0) {
        int mine = 0;
#pragma unroll
        for (int w2 = 0; w2 < kBuildThreads / 32; ++w2) { int t = sWarpTotal[w2]; if (lane == w2) mine = blockTotal; blockTotal += t; }
        chained_publish(tile, (int)blockIdx.x, blockTotal);
        __syncwarp();
        if (lane < kBuildThreads / 32) sWarpTotal[lane] = mine;
    }
    // Manifold::initialize's per-contact part.  A finished contact is parked in this thread's shared-memory column: its 12 geometry floats
    // in the clipper's first polygon buffer (free since the last emit), lambda / penalty over its own raw record (just consumed).
    float* fin = sPoly + threadIdx.x;                                          // buffer 0, this thread's column
    if (inRange) {
        unsigned used = 0u;
        auto loadOld = [&](int j) { return load_contact(old, oldBase + j); };
        BasisCache cache;
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
            if (cc < n) {
                float* q = raw + (size_t)(cc * 10) * rs;
                ContactState ct = contact_initialize(posA, rotA, posB, rotB, f2i(q[0]), mk3(q[rs], q[2 * rs], q[3 * rs]), mk3(q[4 * rs], q[5 * rs], q[6 * rs]),
                                                     mk3(q[7 * rs], q[8 * rs], q[9 * rs]), oldN, oldFeat, used, loadOld, prm, &cache);
                float* g = fin + (size_t)(cc * 12) * rs;
                g[0] = ct.rA.x; g[rs] = ct.rA.y; g[2 * rs] = ct.rA.z; g[3 * rs] = ct.C0n;
                g[4 * rs] = ct.rB.x; g[5 * rs] = ct.rB.y; g[6 * rs] = ct.rB.z; g[7 * rs] = ct.C0t1;
                g[8 * rs] = ct.n.x; g[9 * rs] = ct.n.y; g[10 * rs] = ct.n.z; g[11 * rs] = ct.C0t2;
                const float4 l = pack_lambda(ct), pq = pack_penalty(ct);
                q[0] = l.x; q[rs] = l.y; q[2 * rs] = l.z; q[3 * rs] = l.w; q[4 * rs] = pq.x; q[5 * rs] = pq.y; q[6 * rs] = pq.z; q[7 * rs] = pq.w;
            }
        }
    }
    __syncthreads();
    if (warp == 0) {
        int base = chained_lookback(tile, (int)blockIdx.x, blockTotal);
        if (lane == 0) { sBase = base; if (blockIdx.x == gridDim.x - 1) cnt->nContacts = base + blockTotal; }
    }
    __syncthreads();
    const int first = sBase + sWarpTotal[warp] + incl - n;
    if (!inRange) return;
#pragma unroll 1
    for (int cc = 0; cc < 4; ++cc) {
        if (cc < n) {
            const float* g = fin + (size_t)(cc * 12) * rs;
            const float* q = raw + (size_t)(cc * 10) * rs;
            const int ci = first + cc;
            out.cA[ci] = make_float4(g[0], g[rs], g[2 * rs], g[3 * rs]);
            out.cB[ci] = make_float4(g[4 * rs], g[5 * rs], g[6 * rs], g[7 * rs]);
            out.cN[ci] = make_float4(g[8 * rs], g[9 * rs], g[10 * rs], g[11 * rs]);
            ContactLP lpq; lpq.l = make_float4(q[0], q[rs], q[2 * rs], q[3 * rs]); lpq.p = make_float4(q[4 * rs], q[5 * rs], q[6 * rs], q[7 * rs]);
            out.lp[ci] = lpq;
            out.cM[ci] = s;
        }
    }
    float mu = build ? sqrtf(sa.w * sb.w) : 0.0f;                                   // manifold.cpp:73
    out.hdr[s] = make_int4(a, c, n, __float_as_int(mu));
    out.cstart[s] = first;
    if (build && (slot != s || oldN != n)) cnt->topoChanged = 1;      // same-value racing stores are fine
}

// Stand-alone narrowphase on caller-supplied pairs (parity harness for
// Manifold::collide, collision.cpp:420).  in: 2 x {size3 pos3 quat4} per pair.
__global__ void np_collide_batch(const float* a10, const float* b10, int n, int* count, int* feats4, float* out36) {
    cudaGridDependencySynchronize();
    extern __shared__ float sPoly[];
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* a = a10 + 10 * i; const float* c = b10 + 10 * i;
    V3 sa = mk3(a[0], a[1], a[2]), pa = mk3(a[3], a[4], a[5]); Q4 qa = qmk(a[6], a[7], a[8], a[9]);
    V3 sb = mk3(c[0], c[1], c[2]), pb = mk3(c[3], c[4], c[5]); Q4 qb = qmk(c[6], c[7], c[8], c[9]);
    int code = sat_test(make_obb(pa, qa, sa), make_obb(pb, qb, sb));
    RawContact rc[4];
    PolyShared poly{sPoly + threadIdx.x, (int)blockDim.x};
    int k = code ? build_contacts(pa, qa, sa, pb, qb, sb, code, rc, poly) : 0;
    count[i] = k;
    for (int j = 0; j < 4; ++j) {
        feats4[4 * i + j] = j < k ? rc[j].feature : 0;
        float* o = out36 + 36 * i + 9 * j;
        if (j < k) { o[0] = rc[j].rA.x; o[1] = rc[j].rA.y; o[2] = rc[j].rA.z; o[3] = rc[j].rB.x; o[4] = rc[j].rB.y; o[5] = rc[j].rB.z;
                     o[6] = rc[j].normal.x; o[7] = rc[j].normal.y; o[8] = rc[j].normal.z; }
        else for (int t = 0; t < 9; ++t) o[t] = 0.0f;
    }
}

} // namespace avbd
