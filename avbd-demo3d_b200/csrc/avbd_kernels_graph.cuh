// avbd_kernels_graph.cuh — body/manifold graph (adjacency ranges, greedy colouring, dense contact and
// contact-visit lists) plus the once-per-step body kernels (predict, velocity recovery, diagnostics, state
// pack/unpack).  Compiled in the -fmad=false translation unit with the collision kernels.
#pragma once
#include "avbd_kernels_collide.cuh"
#include "avbd_launch.h"
#include <cub/block/block_radix_sort.cuh>

namespace avbd {

// ------------------------------------------------------------------ adjacency
// Manifolds are sorted by (A,B): a body's "I am A" manifolds are one contiguous
// run [x,y).  Its "I am B" manifolds are a run [z,w) of bList (manifold ids
// stably sorted by B).  adjRange must be zeroed before these two kernels.
__device__ __forceinline__ unsigned adj_a_one(int m, const int4* hdr, int nM, const int* flags, int nBodies, int4* adjRange) {
    int4 h = hdr[m];
    if (m == 0 || hdr[m - 1].x != h.x) adjRange[h.x].x = m;
    if (m == nM - 1 || hdr[m + 1].x != h.x) adjRange[h.x].y = m + 1;
    return (flags[h.y] & kDynamic) ? (unsigned)h.y : (unsigned)nBodies;          // static B never solves: park at the end
}
__global__ void adj_a_ranges(const int4* hdr, int nM, const int* flags, int nBodies, int4* adjRange, unsigned* bKey, int* bVal) {
    cudaGridDependencySynchronize();
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nM) return;
    bKey[m] = adj_a_one(m, hdr, nM, flags, nBodies, adjRange);
    bVal[m] = m;
}
__device__ __forceinline__ void adj_b_one(int t, const unsigned* bKeySorted, int nM, int nBodies, int4* adjRange) {
    unsigned k = bKeySorted[t];
    if (k >= (unsigned)nBodies) return;
    if (t == 0 || bKeySorted[t - 1] != k) adjRange[k].z = t;
    if (t == nM - 1 || bKeySorted[t + 1] != k) adjRange[k].w = t + 1;
}
__global__ void adj_b_ranges(const unsigned* bKeySorted, int nM, int nBodies, int4* adjRange) {
    cudaGridDependencySynchronize();
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nM) return;
    adj_b_one(t, bKeySorted, nM, nBodies, adjRange);
}

// ------------------------------------------------------------------ body -> manifold entries (CSR)
// One entry per LIVE manifold touching a dynamic body, in pair-key order (the body's "I am A" run, then its "I am B" run): the OTHER
// body, 4 bytes — the colouring's adjacency list and nothing else (the sweeps walk the visit lists).  estart is indexed by BODY
// (n + 1 entries, static bodies own none).
// bodyContacts (optional): the body's live contact count, i.e. its number of visits — the same walk yields it, so visit_count need not
// walk the manifolds again once the colour order exists.
__device__ __forceinline__ void entry_count_one(int t, const int* dynList, const int4* adjRange, const int* bList, const int4* hdr, int* deg, int* bodyContacts) {
    int i = dynList[t];
    int4 rg = adjRange[i];
    int k = 0, c = 0;
    for (int m = rg.x; m < rg.y; ++m) { int z = hdr[m].z; k += z > 0 ? 1 : 0; c += z; }
    for (int q = rg.z; q < rg.w; q += 4) {                     // slots first, then their headers side by side
        int mm[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) mm[u] = q + u < rg.w ? bList[q + u] : -1;
#pragma unroll
        for (int u = 0; u < 4; ++u) { int z = mm[u] >= 0 ? hdr[mm[u]].z : 0; k += z > 0 ? 1 : 0; c += z; }
    }
    deg[i] = k;
    if (bodyContacts) bodyContacts[i] = c;
}
__global__ void entry_count(const int* dynList, int nDyn, const int4* adjRange, const int* bList, const int4* hdr, int* deg, int* bodyContacts) {
    cudaGridDependencySynchronize();
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nDyn) return;
    entry_count_one(t, dynList, adjRange, bList, hdr, deg, bodyContacts);
}
__device__ __forceinline__ void entry_fill_one(int t, const int* dynList, const int4* adjRange, const int* bList, const int4* hdr, const int* estart, int* entries) {
    int i = dynList[t];
    int4 rg = adjRange[i];
    int o = estart[i];
    for (int m = rg.x; m < rg.y; ++m) { int4 h = hdr[m]; if (h.z > 0) entries[o++] = h.y; }
    for (int q = rg.z; q < rg.w; q += 4) {
        int mm[4]; int4 h[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) mm[u] = q + u < rg.w ? bList[q + u] : -1;
#pragma unroll
        for (int u = 0; u < 4; ++u) h[u] = mm[u] >= 0 ? hdr[mm[u]] : make_int4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < 4; ++u) if (h[u].z > 0) entries[o++] = h[u].x;
    }
}
__global__ void entry_fill(const int* dynList, int nDyn, const int4* adjRange, const int* bList, const int4* hdr,
                           const int* estart, int* entries) {
    cudaGridDependencySynchronize();
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nDyn) return;
    entry_fill_one(t, dynList, adjRange, bList, hdr, estart, entries);
}

// ------------------------------------------------------------------ contact-visit lists
// Contacts are stored densely (written in place by np_build), so the dual walks them by index; the primal walks, per dynamic body, a run
// of `visits` — one entry per live contact of every manifold touching the body, rebuilt whenever the topology changes.
// Both run over the COLOUR-SORTED body order (colOrder[k], k = 0..nDyn-1): visitStart[k] is indexed by that position, so the
// visits of the bodies of one colour tile are one contiguous run of `visits` and a tile can be walked one visit per thread.
// Entry: {contact id, other body, (visiting body << 3) | first-visit << 2 | anisotropic-inertia << 1 | body-is-A, friction bits}.
// first-visit: this is the visit of the contact that comes FIRST in a sweep (the other endpoint is static or has a higher
// colour) — the one that applies the previous iteration's deferred dual update (avbd_solve.cu).
// Also lists the dynamic bodies NO contact visits (the sweeps' visit pipeline never meets them; primal_free_bodies solves them):
// `freeList` those no user force touches either, `linkedList` those a joint / spring links to another body (they keep their place
// in the colour order).  The lists' order is irrelevant.
__device__ __forceinline__ void visit_count_one(int t, const int* colOrder, const int4* adjRange, const int* bList, const int4* hdr, int* visitCount,
                                                const ForceView& fv, int* freeList, int* linkedList, Counters* cnt, const int* bodyContacts = nullptr) {
    int i = colOrder[t];
    int k = 0;
    if (bodyContacts) k = bodyContacts[i];                  // counted by entry_count on its walk
    else {
        int4 rg = adjRange[i];
        for (int m = rg.x; m < rg.y; ++m) k += hdr[m].z;
        for (int q = rg.z; q < rg.w; ++q) k += hdr[bList[q]].z;
    }
    visitCount[t] = k;
    if (k == 0) {
        bool linked = fv.adjStart != nullptr && fv.adjStart[i + 1] > fv.adjStart[i];
        if (linked) linkedList[atomicAdd(&cnt->nLinkedFree, 1)] = i;           // rare: one atomic each
        else {
            cg::coalesced_group grp = cg::coalesced_threads();
            int base = 0;
            if (grp.thread_rank() == 0) base = atomicAdd(&cnt->nFree, (int)grp.size());
            freeList[grp.shfl(base, 0) + (int)grp.thread_rank()] = i;
        }
    }
}
__global__ void visit_count(const int* colOrder, int nDyn, const int4* adjRange, const int* bList, const int4* hdr, int* visitCount,
                            ForceView fv, int* freeList, int* linkedList, Counters* cnt) {
    cudaGridDependencySynchronize();
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nDyn) return;
    visit_count_one(t, colOrder, adjRange, bList, hdr, visitCount, fv, freeList, linkedList, cnt);
}
// Four manifolds at a time: their slots (the B run goes through bList), then their headers and contact starts, then the other bodies'
// colours are fetched side by side — one body's walk is otherwise a chain of three dependent gathers per manifold.
__device__ __forceinline__ void visit_fill_one(int t, const int* colOrder, const int4* adjRange, const int* bList, const int4* hdr, const int* cstart,
                                               const int* visitStart, const BodyAux* aux, const int* colour, int4* visits) {
    int i = colOrder[t];
    int4 rg = adjRange[i];
    int o = visitStart[t];
    float4 I = aux[i].inert;
    int idx = (i << 3) | ((I.x == I.y && I.y == I.z) ? 0 : 2);
    int mine = colour[i];
#pragma unroll
    for (int side = 0; side < 2; ++side) {                     // 0: the body is A (slots rg.x .. rg.y), 1: it is B (bList[rg.z .. rg.w])
        const int q0 = side ? rg.z : rg.x, q1 = side ? rg.w : rg.y;
        for (int q = q0; q < q1; q += 4) {
            int mm[4], c0[4], co[4]; int4 h[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) mm[u] = q + u < q1 ? (side ? bList[q + u] : q + u) : -1;
#pragma unroll
            for (int u = 0; u < 4; ++u) { h[u] = mm[u] >= 0 ? hdr[mm[u]] : make_int4(0, 0, 0, 0); c0[u] = mm[u] >= 0 ? cstart[mm[u]] : 0; }
#pragma unroll
            for (int u = 0; u < 4; ++u) co[u] = mm[u] >= 0 ? colour[side ? h[u].x : h[u].y] : 0;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (mm[u] < 0) continue;
                const int other = side ? h[u].x : h[u].y;
                const int tag = idx | (side ? 0 : 1) | ((co[u] < 0 || mine < co[u]) ? 4 : 0);     // first visit: the other endpoint is static or of a higher colour
                for (int c = 0; c < h[u].z; ++c) visits[o++] = make_int4(c0[u] + c, other, tag, h[u].w);
            }
        }
    }
}
// colVisit != nullptr: block 0 also writes each colour's visit range (colour_visit_bounds): the ranges and the visit starts are complete
// before this launch.
__global__ void visit_fill(const int* colOrder, int nDyn, const int4* adjRange, const int* bList, const int4* hdr, const int* cstart,
                           const int* visitStart, const BodyAux* aux, const int* colour, int4* visits,
                           const int2* colourRange, const Counters* cnt, int2* colVisit) {
    cudaGridDependencySynchronize();
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (colVisit != nullptr && t < 64) {
        int2 r = t < cnt->nColours ? colourRange[t] : make_int2(0, 0);
        colVisit[t] = r.y > r.x ? make_int2(visitStart[r.x], visitStart[r.y]) : make_int2(0, 0);
    }
    if (t >= nDyn) return;
    visit_fill_one(t, colOrder, adjRange, bList, hdr, cstart, visitStart, aux, colour, visits);
}

// Once per step (the narrowphase rewrites every contact): copy each visit's contact geometry into visit order, so the
// iterations x colours primal sweeps STREAM it (three fully coalesced 16-byte records per visit) instead of gathering it by contact
// id — gathered, a sweep moved a third more DRAM bytes (ncu, 1M-box grid: 2.06 GB against 1.56 GB per iteration: a pile has about two
// contacts per manifold, so a gather uses half of every 64-byte DRAM granule, and a body's "I am B" manifolds are scattered).
// Already in the VISITING body's frame: a = {r_self, C0n}, b = {r_other, C0t.x}, n = {n, C0t.y}.  nVisits lives on the device.
__global__ void visit_geometry(const int4* __restrict__ visits, const int* __restrict__ nVisits, ManifoldSet ms, VisitGeom vg) {
    cudaGridDependencySynchronize();
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= *nVisits) return;
    int4 e = visits[v];
    float4 a = ms.cA[e.x], b = ms.cB[e.x];
    bool isA = (e.z & 1) != 0;
    vg.a[v] = isA ? a : make_float4(b.x, b.y, b.z, a.w);
    vg.b[v] = isA ? b : make_float4(a.x, a.y, a.z, b.w);
    vg.n[v] = ms.cN[e.x];
}

// ------------------------------------------------------------------ colouring
__device__ __forceinline__ unsigned mix32(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}
// Priority order of the greedy colouring: rank first (higher goes first), ties by a hash of the WORLD-LOCAL index.  By default every
// body has rank 0 and the order is the hashed one: the colouring is a pure function of the graph.  With AVBD_ITERATED_COLOUR=1 the
// rank is the body's colour in the PREVIOUS colouring of the same body set + 1 (0: it had none) — walking the old colour classes
// one after another is Culberson's iterated greedy: on an unchanged graph the new colouring never needs more colours than the old
// one and often fewer, and because the members of an old class are (nearly all) mutually non-adjacent the Jones-Plassmann rounds
// finish in about as many rounds as there are classes.  Rank and hash are world-local quantities, so a world colours the same way
// wherever it sits in a batch (ensemble runs are partition-invariant).  See run_colour (avbd_engine.cu) for why it is opt-in.
__device__ __forceinline__ bool outranks(int rankA, int localA, int rankB, int localB) {
    if (rankA != rankB) return rankA > rankB;
    unsigned ha = mix32((unsigned)localA + 0x9e3779b9u), hb = mix32((unsigned)localB + 0x9e3779b9u);
    return ha != hb ? ha > hb : localA > localB;
}
// Work word of the rounds: low byte = state (0 static, 1 uncoloured, 2 + c coloured c), the rest = priority rank (old colour + 1, 0 = none).
__device__ __forceinline__ int colour_word(int rank, int state) { return (rank << 8) | state; }

// Everything the graph stage wants zeroed or initialised before its first real kernel, in ONE launch instead of six memset nodes and
// colour_init (a step of a small world is a chain of dependent launches of a few microseconds each): adjacency ranges, entry counts, colour
// ranges, the visit counts' terminator, the free-body counters, the colouring's list cursors, and — initMode 1: fresh, 2: ranked by the
// previous colouring, 0: leave the work words alone — the colouring's work words.
__global__ void graph_prologue(const int* flags, int n, int nDyn, int4* adjRange, int* deg, int2* colRange, int* visitCount, int* colCursor,
                               Counters* cnt, int initMode, int* word, int* colour) {
    cudaGridDependencySynchronize();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) deg[i] = 0;
    if (i < n) {
        adjRange[i] = make_int4(0, 0, 0, 0);
        if (initMode) {
            const bool dyn = (flags[i] & kDynamic) != 0;
            int old = initMode == 2 ? colour[i] : -1;
            word[i] = dyn ? colour_word(old >= 0 ? old + 1 : 0, 1) : 0;
            colour[i] = dyn ? -1 : -2;
        }
    }
    if (i < 64) colRange[i] = make_int2(0, 0);
    if (i < 4) colCursor[i] = 0;
    if (i == 0) { visitCount[nDyn] = 0; cnt->nFree = 0; cnt->nLinkedFree = 0; }
}

// oldColour: the previous colouring to rank by (nullptr / negative entries: none).  May alias `colour`.
__global__ void colour_init(const int* flags, int n, const int* oldColour, int* word, int* colour) {
    cudaGridDependencySynchronize();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool dyn = (flags[i] & kDynamic) != 0;
    int old = oldColour ? oldColour[i] : -1;
    word[i] = dyn ? colour_word(old >= 0 ? old + 1 : 0, 1) : 0;
    colour[i] = dyn ? -1 : -2;
}

// Kept colouring (the default once a world has one): the contact graph of a pile changes by a percent or two per step, so last step's
// colouring is still valid almost everywhere.  A dynamic body keeps its colour unless a neighbour of the NEW graph that outranks it (hashed
// world-local priority) holds the same one; the bodies that lose theirs — and only those — go through the Jones-Plassmann rounds below,
// which see the kept colours as already taken.  Decided from the OLD colours alone (nothing is written that another thread reads here),
// so the result does not depend on timing, and from world-local quantities only, so a world colours the same way wherever it sits in a
// batch.  The colouring is then a function of the world's history, not of the current graph alone: snapshots carry it.
__device__ __forceinline__ int kept_word(int i, const int* flags, const int* estart, const int* entries, const ForceView& fv,
                                         const int* localIdx, const int* oldColour) {
    if (!(flags[i] & kDynamic)) return 0;
    const int c = oldColour[i];
    if (c < 0 || c > 61) return colour_word(0, 1);
    const int li = localIdx[i];
    bool keep = true;
    auto visit = [&](int other) {
        if (other < 0) return;
        if (oldColour[other] == c && (flags[other] & kDynamic) && outranks(0, localIdx[other], 0, li)) keep = false;
    };
    for (int e = estart[i], e1 = estart[i + 1]; e < e1 && keep; ++e) visit(entries[e]);
    if (fv.adjStart) {
        for (int k = fv.adjStart[i]; k < fv.adjStart[i + 1] && keep; ++k) {
            int e = fv.adj[k]; int idx = e >> 2; bool isA = e & 1;
            int other = (e & 2) ? (isA ? fv.springs[idx].b : fv.springs[idx].a) : (isA ? fv.joints[idx].b : fv.joints[idx].a);
            visit(other);
        }
    }
    return keep ? colour_word(0, 2 + c) : colour_word(0, 1);
}

// Stand-alone form (fallback path with one launch per round; the one-launch round kernels do this in their prologue).
__global__ void colour_keep(const int* flags, int n, const int* estart, const int* entries, ForceView fv, const int* localIdx, int* word, const int* colour) {
    cudaGridDependencySynchronize();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) word[i] = kept_word(i, flags, estart, entries, fv, localIdx, colour);
}

// One Jones-Plassmann attempt for body i: it takes the smallest colour unused by its neighbours once every
// higher-priority neighbour is coloured.  A lower-priority neighbour cannot be coloured before i is, so what i sees
// coloured is exactly its higher-priority neighbourhood whenever it succeeds: the result is the sequential greedy
// colouring in priority order, independent of timing (timing only changes how many attempts it takes).  Reading a
// neighbour's word while that neighbour writes it is part of the scheme: either value leads to the same colouring.
// Returns whether the body is coloured after the attempt.
__device__ __forceinline__ bool try_colour(int i, const int* estart, const int* entries, const ForceView& fv,
                                           const int* localIdx, volatile int* word, int* colour, Counters* cnt) {
    const int wi = word[i];
    if ((wi & 255) != 1) return true;
    const int li = localIdx[i], ri = wi >> 8;
    unsigned long long used = 0ull;
    bool ready = true;
    auto visit = [&](int other) {
        if (other < 0) return;
        const int wo = word[other], so = wo & 255;
        if (so >= 2) used |= 1ull << (so - 2);
        else if (so == 1 && outranks(wo >> 8, localIdx[other], ri, li)) ready = false;
    };
    // four neighbours at a time: their ids, then their words and priorities, are fetched side by side — a body's walk is a chain of
    // dependent gathers (id -> word -> priority) and the colouring kernels are made of nothing else
    for (int e = estart[i], e1 = estart[i + 1]; e < e1 && ready; e += 4) {
        int o[4], wv[4], lo[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) o[u] = e + u < e1 ? entries[e + u] : -1;
#pragma unroll
        for (int u = 0; u < 4; ++u) { wv[u] = o[u] >= 0 ? word[o[u]] : 0; lo[u] = o[u] >= 0 ? localIdx[o[u]] : 0; }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int so = wv[u] & 255;
            if (so >= 2) used |= 1ull << (so - 2);
            else if (so == 1 && outranks(wv[u] >> 8, lo[u], ri, li)) ready = false;
        }
    }
    if (fv.adjStart) {
        for (int k = fv.adjStart[i]; k < fv.adjStart[i + 1] && ready; ++k) {
            int e = fv.adj[k]; int idx = e >> 2; bool isA = e & 1;
            int other = (e & 2) ? (isA ? fv.springs[idx].b : fv.springs[idx].a) : (isA ? fv.joints[idx].b : fv.joints[idx].a);
            visit(other);
        }
    }
    if (!ready) return false;
    int c = __ffsll((long long)~used) - 1;
    if (c < 0) { c = 63; atomicOr(&cnt->overflow, 4); }
    word[i] = colour_word(ri, 2 + c);
    colour[i] = c;
    return true;
}

__global__ void colour_round(const int* dynList, int nDyn, const int* estart, const int* entries,
                             ForceView fv, const int* localIdx, volatile int* word, int* colour, Counters* cnt, bool countLeft) {
    cudaGridDependencySynchronize();
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nDyn) return;
    if (!try_colour(dynList[t], estart, entries, fv, localIdx, word, colour, cnt) && countLeft) {
        cg::coalesced_group grp = cg::coalesced_threads();
        if (grp.thread_rank() == 0) atomicAdd(&cnt->nUncoloured, (int)grp.size());
    }
}

// Small worlds: ALL rounds in one block (block barrier between rounds instead of a launch, no host check of the
// uncoloured count).  Same attempts, hence the same colouring as colour_round.  When the world's work words fit (nBodies <=
// kColourSmemBodies) the rounds run on a shared-memory copy: a round is then a handful of shared-memory reads per body instead of a
// chain of L2 round trips (Stress1000: 30 -> ~8 us).
#ifndef AVBD_COLOUR_ATTEMPTS
#define AVBD_COLOUR_ATTEMPTS 4
#endif
constexpr int kColourAttempts = AVBD_COLOUR_ATTEMPTS;        // attempts per body per round: the neighbours a body waits for are being coloured in the same round
constexpr int kColourBlockThreads = 1024;
constexpr int kColourSmemBodies = 10240;          // 40 KB of static shared memory
__global__ void __launch_bounds__(kColourBlockThreads) colour_rounds_block(const int* dynList, int nDyn, const int* estart, const int* entries,
                                                                           ForceView fv, const int* localIdx, int* word, int* colour, Counters* cnt, int nBodies,
                                                                           const int* keepFlags, unsigned* key, int* val) {
    cudaGridDependencySynchronize();
    __shared__ int sWord[kColourSmemBodies];
    volatile int* wd = word;
    if (keepFlags) {                               // kept colouring: the work words come from last step's colours (colour_init was not launched)
        for (int i = threadIdx.x; i < nBodies; i += blockDim.x) word[i] = kept_word(i, keepFlags, estart, entries, fv, localIdx, colour);
        __syncthreads();
    }
    if (nBodies <= kColourSmemBodies) {
        for (int i = threadIdx.x; i < nBodies; i += blockDim.x) sWord[i] = word[i];
        __syncthreads();
        wd = sWord;
    }
    int left = 1;
    for (int round = 0; round < 4096 && left; ++round) {
        int mine = 0;
        for (int t = threadIdx.x; t < nDyn; t += blockDim.x) {
            bool done = false;
#pragma unroll 1
            for (int attempt = 0; attempt < kColourAttempts && !done; ++attempt) done = try_colour(dynList[t], estart, entries, fv, localIdx, wd, colour, cnt);
            if (!done) mine = 1;
        }
        left = __syncthreads_or(mine);           // also makes this round's colours visible to the whole block
    }
    if (threadIdx.x == 0) cnt->nUncoloured = left;
    // the colour sort's keys (colour_keys), while the colours are at hand
    for (int t = threadIdx.x; t < nDyn; t += blockDim.x) { const int i = dynList[t]; key[t] = (unsigned)colour[i]; val[t] = i; }
}

// Large worlds: ALL rounds in one cooperative launch (grid barrier between rounds, no host round trip to learn how many bodies are
// left).  Each round walks a work list of the still uncoloured bodies and appends the ones it could not colour to the next list
// (warp-aggregated; the list's order is irrelevant to the result).  Same attempts, same colouring as colour_round.  Cursors rotate
// over three slots: the slot a round fills was last READ two barriers ago.
constexpr int kColourGridThreads = 256;
__global__ void __launch_bounds__(kColourGridThreads) colour_rounds_grid(const int* dynList, int nDyn, const int* estart, const int* entries,
                                                                         ForceView fv, const int* localIdx, volatile int* word, int* colour, Counters* cnt,
                                                                         int* listA, int* listB, int* cursors, const int* keepFlags, int nBodies,
                                                                         unsigned* key, int* val) {
    cg::grid_group grid = cg::this_grid();
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsize = gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    const int* list = dynList; int count = nDyn;
    if (keepFlags) {
        // kept colouring: work words from last step's colours; the bodies that lost theirs are the first work list (listB, count in cursors[3])
        const int rounded = (nBodies + 31) & ~31;
        for (int i = gtid; i < rounded; i += gsize) {
            bool left = false;
            if (i < nBodies) { const int wv = kept_word(i, keepFlags, estart, entries, fv, localIdx, colour); word[i] = wv; left = (wv & 255) == 1; }
            unsigned vote = __ballot_sync(0xffffffffu, left);
            if (vote) {
                int base = 0;
                if (lane == 0) base = atomicAdd(cursors + 3, __popc(vote));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (left) listB[base + __popc(vote & ((1u << lane) - 1u))] = i;
            }
        }
        grid.sync();
        count = *(volatile int*)(cursors + 3);
        list = listB;
        if (gtid == 0) cnt->colourKept = nDyn - count;
    }
    int round = 0;
    for (; round < 4096 && count > 0; ++round) {
        int* out = (round & 1) ? listB : listA;
        int* cur = cursors + round % 3;
        if (gtid == 0) cursors[(round + 1) % 3] = 0;            // next round's cursor (last read after the barrier of round - 2)
        const int rounded = (count + 31) & ~31;
        for (int t = gtid; t < rounded; t += gsize) {
            bool left = false; int i = 0;
            // a few attempts per round: the neighbours a body waits for are being coloured by other threads of this very round, and an
            // attempt is cheaper than a trip through the work list and a grid barrier
            if (t < count) {
                i = list[t]; left = true;
#pragma unroll 1
                for (int attempt = 0; attempt < kColourAttempts && left; ++attempt) left = !try_colour(i, estart, entries, fv, localIdx, word, colour, cnt);
            }
            unsigned vote = __ballot_sync(0xffffffffu, left);
            if (vote) {
                int base = 0;
                if (lane == 0) base = atomicAdd(cur, __popc(vote));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (left) out[base + __popc(vote & ((1u << lane) - 1u))] = i;
            }
        }
        grid.sync();
        count = *(volatile int*)cur;
        list = out;
    }
    if (gtid == 0) { cnt->nUncoloured = count; cnt->colourRounds = round; }
    // the colour sort's keys (colour_keys), while the grid is here: every round ended with a grid barrier, all colours are visible
    for (int t = gtid; t < nDyn; t += gsize) { const int i = dynList[t]; key[t] = (unsigned)__ldcg(colour + i); val[t] = i; }
}

// Work list of the bodies still uncoloured (order is irrelevant: the colouring does not depend on it).
__global__ void colour_compact(const int* list, int n, const int* colour, int* out, int* outCount) {
    cudaGridDependencySynchronize();
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    bool keep = false; int i = 0;
    if (t < n) { i = list[t]; keep = colour[i] < 0; }
    unsigned vote = __ballot_sync(0xffffffffu, keep);
    if (!vote) return;
    int lane = threadIdx.x & 31, base = 0;
    if (lane == 0) base = atomicAdd(outCount, __popc(vote));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (keep) out[base + __popc(vote & ((1u << lane) - 1u))] = i;
}

// Sort keys of the colour order.
__global__ void colour_keys(const int* dynList, int nDyn, const int* colour, unsigned* key, int* val) {
    cudaGridDependencySynchronize();
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nDyn) return;
    int i = dynList[t];
    key[t] = (unsigned)colour[i];
    val[t] = i;
}
// colourRange[c] = {first, last+1} in the colour-sorted body order; must be zeroed first.
__global__ void colour_bounds(const unsigned* keySorted, int nDyn, int2* colourRange, Counters* cnt) {
    cudaGridDependencySynchronize();
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nDyn) return;
    unsigned c = keySorted[t];
    if (t == 0 || keySorted[t - 1] != c) colourRange[c].x = t;
    if (t == nDyn - 1 || keySorted[t + 1] != c) colourRange[c].y = t + 1;
    if (t == nDyn - 1) cnt->nColours = (int)c + 1;
}

// colour_bounds and visit_count in one launch (both walk the colour-sorted bodies, neither reads what the other writes)
__global__ void colour_bounds_visit_count(const unsigned* keySorted, int nDyn, int2* colourRange, const int* colOrder, const int4* adjRange, const int* bList,
                                          const int4* hdr, int* visitCount, ForceView fv, int* freeList, int* linkedList, Counters* cnt, const int* bodyContacts) {
    cudaGridDependencySynchronize();
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nDyn) return;
    unsigned c = keySorted[t];
    if (t == 0 || keySorted[t - 1] != c) colourRange[c].x = t;
    if (t == nDyn - 1 || keySorted[t + 1] != c) colourRange[c].y = t + 1;
    if (t == nDyn - 1) cnt->nColours = (int)c + 1;
    visit_count_one(t, colOrder, adjRange, bList, hdr, visitCount, fv, freeList, linkedList, cnt, bodyContacts);
}

// out[c] = {first visit, one past the last visit} of colour c (a colour's bodies, hence its visits, are contiguous)
__global__ void colour_visit_bounds(const int2* colourRange, const Counters* cnt, const int* visitStart, int2* out) {
    cudaGridDependencySynchronize();
    int c = threadIdx.x;
    if (c >= 64) return;
    int2 r = c < cnt->nColours ? colourRange[c] : make_int2(0, 0);
    out[c] = r.y > r.x ? make_int2(visitStart[r.x], visitStart[r.y]) : make_int2(0, 0);
}

// ------------------------------------------------------------------ the whole graph stage of a SMALL world in one block
// A small world's step is a chain of dependent launches of a few microseconds each, and the graph stage alone was sixteen of them
// (Stress1000: 0.107 of 0.67 ms).  Up to kSmallGraphMax manifolds and the one-block colouring's body limit one block of 1024 threads runs every phase itself
// — the very per-element routines of the kernels above, a block barrier where they have a kernel boundary, cub::BlockRadixSort (stable,
// like the device-wide sorts it replaces) and a block prefix sum — so adjacency order, colours, colour order and visit lists come out
// identical to the multi-launch path's.  Fresh colouring only (the default).
constexpr int kSmallGraphThreads = 1024;
constexpr int kSmallGraphItems = 4;
constexpr int kSmallGraphMax = kSmallGraphThreads * kSmallGraphItems;
// exclusive prefix sum of in[0, n) by the whole block (1024 threads); sWarp: 32 ints of shared memory.  in != out.
__device__ __forceinline__ void block_scan_exclusive(const int* __restrict__ in, int* __restrict__ out, int n, int* sWarp) {
    const int per = (n + kSmallGraphThreads - 1) / kSmallGraphThreads;      // consecutive items per thread
    const int b = threadIdx.x * per, e = b + per < n ? b + per : n;
    int sum = 0;
    for (int i = b; i < e; ++i) sum += in[i];
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    int incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int up = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += up; }
    __syncthreads();                                       // sWarp may still be read from a previous call
    if (lane == 31) sWarp[wp] = incl;
    __syncthreads();
    if (wp == 0) {
        int v = sWarp[lane], sc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { int up = __shfl_up_sync(0xffffffffu, sc, d); if (lane >= d) sc += up; }
        sWarp[lane] = sc - v;
    }
    __syncthreads();
    int run = sWarp[wp] + incl - sum;
    for (int i = b; i < e; ++i) { int x = in[i]; out[i] = run; run += x; }
}
struct SmallGraph {
    const int* flags; int n; const int* dynList; int nDyn; const int* localIdx;
    const int4* hdr; const int* cstart; int nM; int sortBits;
    int4* adjRange; int* bList; unsigned* bKeySorted; int* deg; int* estart; int* entries;
    int* colour; unsigned* colKeySorted; int* colOrder; int2* colRange;
    int* visitCount; int* visitStart; int4* visits; int2* colVisit;
    int* freeList; int* linkedList; const BodyAux* aux; Counters* cnt;
};
__global__ void __launch_bounds__(kSmallGraphThreads) graph_small(SmallGraph a, ForceView fv) {
    cudaGridDependencySynchronize();
    using Sort = cub::BlockRadixSort<unsigned, kSmallGraphThreads, kSmallGraphItems, int>;
    __shared__ union Scratch { typename Sort::TempStorage sort; int word[kSmallGraphMax + 64]; Scratch() {} } sm;   // the sorts and the colouring never overlap
    __shared__ int sWarp[32];
    const int tid = threadIdx.x;
    constexpr int T = kSmallGraphThreads;
    // ---- prologue
    for (int i = tid; i <= a.n; i += T) {
        a.deg[i] = 0;
        if (i < a.n) { a.adjRange[i] = make_int4(0, 0, 0, 0); a.colour[i] = (a.flags[i] & kDynamic) ? -1 : -2; }
    }
    if (tid < 64) a.colRange[tid] = make_int2(0, 0);
    if (tid == 0) { a.visitCount[a.nDyn] = 0; a.cnt->nFree = 0; a.cnt->nLinkedFree = 0; }
    __syncthreads();
    // ---- adjacency: A runs, manifolds stably sorted by B, B runs
    unsigned keys[kSmallGraphItems]; int vals[kSmallGraphItems];
#pragma unroll
    for (int k = 0; k < kSmallGraphItems; ++k) {
        const int m = tid * kSmallGraphItems + k;
        keys[k] = m < a.nM ? adj_a_one(m, a.hdr, a.nM, a.flags, a.n, a.adjRange) : 0xffffffffu;
        vals[k] = m;
    }
    Sort(sm.sort).Sort(keys, vals, 0, a.sortBits);
#pragma unroll
    for (int k = 0; k < kSmallGraphItems; ++k) {
        const int idx = tid * kSmallGraphItems + k;
        if (idx < a.nM) { a.bKeySorted[idx] = keys[k]; a.bList[idx] = vals[k]; }
    }
    __syncthreads();
    for (int t = tid; t < a.nM; t += T) adj_b_one(t, a.bKeySorted, a.nM, a.n, a.adjRange);
    __syncthreads();
    // ---- body -> neighbour entries
    for (int t = tid; t < a.nDyn; t += T) entry_count_one(t, a.dynList, a.adjRange, a.bList, a.hdr, a.deg, nullptr);
    __syncthreads();
    block_scan_exclusive(a.deg, a.estart, a.n + 1, sWarp);
    __syncthreads();
    for (int t = tid; t < a.nDyn; t += T) entry_fill_one(t, a.dynList, a.adjRange, a.bList, a.hdr, a.estart, a.entries);
    // ---- colouring: Jones-Plassmann rounds on work words in shared memory (colour_rounds_block)
    for (int i = tid; i < a.n; i += T) sm.word[i] = (a.flags[i] & kDynamic) ? colour_word(0, 1) : 0;
    __syncthreads();
    {
        int left = 1;
        for (int round = 0; round < 4096 && left; ++round) {
            int mine = 0;
            for (int t = tid; t < a.nDyn; t += T) {
                bool done = false;
#pragma unroll 1
                for (int attempt = 0; attempt < kColourAttempts && !done; ++attempt) done = try_colour(a.dynList[t], a.estart, a.entries, fv, a.localIdx, sm.word, a.colour, a.cnt);
                if (!done) mine = 1;
            }
            left = __syncthreads_or(mine);
        }
        if (tid == 0) a.cnt->nUncoloured = left;
    }
    __syncthreads();
    // ---- colour order (stable by colour), colour ranges
#pragma unroll
    for (int k = 0; k < kSmallGraphItems; ++k) {
        const int t = tid * kSmallGraphItems + k;
        const int i = t < a.nDyn ? a.dynList[t] : -1;
        keys[k] = i >= 0 ? (unsigned)a.colour[i] : 0xffffffffu;
        vals[k] = i;
    }
    Sort(sm.sort).Sort(keys, vals, 0, 7);
#pragma unroll
    for (int k = 0; k < kSmallGraphItems; ++k) {
        const int t = tid * kSmallGraphItems + k;
        if (t < a.nDyn) { a.colKeySorted[t] = keys[k]; a.colOrder[t] = vals[k]; }
    }
    __syncthreads();
    for (int t = tid; t < a.nDyn; t += T) {
        unsigned c = a.colKeySorted[t];
        if (t == 0 || a.colKeySorted[t - 1] != c) a.colRange[c].x = t;
        if (t == a.nDyn - 1 || a.colKeySorted[t + 1] != c) a.colRange[c].y = t + 1;
        if (t == a.nDyn - 1) a.cnt->nColours = (int)c + 1;
    }
    __syncthreads();
    // ---- contact visits in colour order
    for (int t = tid; t < a.nDyn; t += T) visit_count_one(t, a.colOrder, a.adjRange, a.bList, a.hdr, a.visitCount, fv, a.freeList, a.linkedList, a.cnt);
    __syncthreads();
    block_scan_exclusive(a.visitCount, a.visitStart, a.nDyn + 1, sWarp);
    __syncthreads();
    for (int t = tid; t < a.nDyn; t += T) visit_fill_one(t, a.colOrder, a.adjRange, a.bList, a.hdr, a.cstart, a.visitStart, a.aux, a.colour, a.visits);
    if (tid < 64) {
        int2 r = tid < a.cnt->nColours ? a.colRange[tid] : make_int2(0, 0);
        a.colVisit[tid] = r.y > r.x ? make_int2(a.visitStart[r.x], a.visitStart[r.y]) : make_int2(0, 0);
    }
}

// ------------------------------------------------------------------ predict / warm-start decay of user forces
__global__ void predict_bodies(BodyView b, SolveParams prm, Diag* diag) {
    cudaGridDependencySynchronize();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int dyn = 0, ev = 0, world = -1;
    if (i < b.n) {
        world = b.worldId[i];
        BodyPose pose = b.pose[i]; BodyVel vel = b.vel[i]; BodyAux aux = b.aux[i]; BodyInit init;
        float4 pl = b.prevLin[i];
        ev = predict_body(pose, vel, pl, aux, init, prm);
        b.pose[i] = pose; b.vel[i] = vel; b.init[i] = init;
        b.aux[i].posI = aux.posI; b.aux[i].rotI = aux.rotI;
        dyn = aux.mass.y > 0.0f ? 1 : 0;
    }
    WorldGroup wg(world);
    dyn = wg.sum(dyn); ev = wg.sum(ev);
    if (wg.leader && world >= 0) {
        if (dyn) atomicAdd(&diag[world].dynamicBodies, dyn);
        if (ev) atomicAdd(&diag[world].nanEvents, ev);
    }
}

__global__ void decay_user_forces(ForceView fv, SolveParams prm) {
    cudaGridDependencySynchronize();
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < fv.nJoints) {
        JointRec& j = fv.joints[t];
        for (int r = 0; r < 6; ++r) decay_row(j.lambda[r], j.penalty[r], j.stiffness[r], prm);
    } else if (t - fv.nJoints < fv.nSprings) {
        SpringRec& s = fv.springs[t - fv.nJoints];
        decay_row(s.lambda, s.penalty, s.k, prm);
    }
}

// ------------------------------------------------------------------ velocity + diagnostics
__global__ void velocity_bodies(BodyView b, SolveParams prm, Diag* diag) {
    cudaGridDependencySynchronize();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    float ls = 0.0f, as = 0.0f; int ev = 0; int world = -1;
    if (i < b.n) {
        world = b.worldId[i];
        if (b.pose[i].pos.w > 0.0f) {
            BodyVel vel = b.vel[i]; float4 pl;
            ev = velocity_body(b.pose[i], b.init[i], vel, pl, prm.dt, ls, as);
            b.vel[i] = vel; b.prevLin[i] = pl;
        }
    }
    WorldGroup wg(world);
    ls = wg.max_nonneg(ls); as = wg.max_nonneg(as); ev = wg.sum(ev);
    if (wg.leader && world >= 0) {
        Diag* d = diag + world;
        if (ls > 0.0f) atomic_max_nonneg(&d->maxLinearSpeed, ls);
        if (as > 0.0f) atomic_max_nonneg(&d->maxAngularSpeed, as);
        if (ev) atomicAdd(&d->nanEvents, ev);
    }
}

// solver.cpp:472-497, one thread per live contact (only when the step's last dual pass did not already reduce them)
__global__ void diagnostics_contacts(BodyView b, ManifoldSet ms, int nContacts, Diag* diag) {
    cudaGridDependencySynchronize();
    int ci = blockIdx.x * blockDim.x + threadIdx.x;
    float sepn = 0.0f, lam = 0.0f; int nm = 0, nv = 0; int world = -1;
    if (ci < nContacts) {
        int m = ms.cM[ci];
        int4 h = ms.hdr[m];
        world = b.worldId[h.x];
        nm = (ci == 0 || ms.cM[ci - 1] != m) ? 1 : 0;
        BodyPose pa = b.pose[h.x], pb = b.pose[h.y];
        nv = (pa.pos.w > 0.0f ? 1 : 0) + (pb.pos.w > 0.0f ? 1 : 0);
        float4 a4 = ms.cA[ci], b4 = ms.cB[ci], n4 = ms.cN[ci];
        V3 pA = xyz(pa.pos) + qrot(quat(pa.rot), xyz(a4));
        V3 pB = xyz(pb.pos) + qrot(quat(pb.rot), xyz(b4));
        sepn = dot(pA - pB, xyz(n4));
        lam = ms.lp[ci].l.x;
    }
    reduce_contact_diag_block(world, sepn, lam, world >= 0 ? 1 : 0, nm, nv, diag);
}

// Rigid public state <-> the 13-float-per-body host layout (pos3 quat4 lin3 ang3), on the device so the
// host side of Solver::step() is one DMA each way.
__global__ void pack_state(BodyView b, float* out13) {
    cudaGridDependencySynchronize();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.n) return;
    BodyPose p = b.pose[i]; BodyVel v = b.vel[i];
    float* o = out13 + 13 * (size_t)i;
    o[0] = p.pos.x; o[1] = p.pos.y; o[2] = p.pos.z; o[3] = p.rot.x; o[4] = p.rot.y; o[5] = p.rot.z; o[6] = p.rot.w;
    o[7] = v.lin.x; o[8] = v.lin.y; o[9] = v.lin.z; o[10] = v.ang.x; o[11] = v.ang.y; o[12] = v.ang.z;
}
__global__ void unpack_state(BodyView b, const float* in13) {
    cudaGridDependencySynchronize();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.n) return;
    const float* o = in13 + 13 * (size_t)i;
    float invMass = b.aux[i].mass.y;
    BodyPose p; p.pos = make_float4(o[0], o[1], o[2], invMass); p.rot = make_float4(o[3], o[4], o[5], o[6]);
    BodyVel v; v.lin = make_float4(o[7], o[8], o[9], 0.f); v.ang = make_float4(o[10], o[11], o[12], 0.f);
    b.pose[i] = p; b.vel[i] = v;
}

__global__ void unpack_state_range(BodyView b, const float* in13, int first, int count) {
    cudaGridDependencySynchronize();
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    int i = first + k;
    const float* o = in13 + 13 * (size_t)i;
    float invMass = b.aux[i].mass.y;
    BodyPose p; p.pos = make_float4(o[0], o[1], o[2], invMass); p.rot = make_float4(o[3], o[4], o[5], o[6]);
    BodyVel v; v.lin = make_float4(o[7], o[8], o[9], 0.f); v.ang = make_float4(o[10], o[11], o[12], 0.f);
    b.pose[i] = p; b.vel[i] = v;
}

// Solver::pick (solver.cpp:145-228): ray against every dynamic OBB (slab test in body space); the closest hit wins,
// ties going to the newest body like the reference's list walk.  best = (tHit bits << 32) | (0xFFFFFFFF - index).
AVBD_HD bool ray_obb(V3 origin, V3 rayDir, V3 pos, Q4 rot, V3 size, float& tHit, V3& local) {
    const float eps = 1.0e-6f;
    Q4 inv = qconj(rot);
    V3 lo = qrot(inv, origin - pos), ld = qrot(inv, rayDir), half = size * 0.5f;
    float tEnter = 0.0f, tExit = FLT_MAX;
    for (int axis = 0; axis < 3; ++axis) {
        float o = comp(lo, axis), d = comp(ld, axis), mn = -comp(half, axis), mx = comp(half, axis);
        if (fabsf(d) < eps) { if (o < mn || o > mx) return false; continue; }
        float invD = 1.0f / d, t0 = (mn - o) * invD, t1 = (mx - o) * invD;
        if (t0 > t1) { float t = t0; t0 = t1; t1 = t; }
        tEnter = fmax2(tEnter, t0); tExit = fmin2(tExit, t1);
        if (tEnter > tExit) return false;
    }
    tHit = (tEnter >= 0.0f) ? tEnter : tExit;
    if (tHit < 0.0f) return false;
    local = lo + ld * tHit;
    return true;
}
__global__ void pick_bodies(BodyView b, V3 origin, V3 rayDir, unsigned long long* best) {
    cudaGridDependencySynchronize();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.n) return;
    BodyPose p = b.pose[i];
    if (p.pos.w <= 0.0f) return;
    float t; V3 local;
    if (!ray_obb(origin, rayDir, xyz(p.pos), quat(p.rot), xyz(b.size[i]), t, local)) return;
    unsigned long long key = ((unsigned long long)__float_as_uint(t) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)i);
    atomicMin(best, key);
}

} // namespace avbd
