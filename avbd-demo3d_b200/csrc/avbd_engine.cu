// avbd_engine.cu — world state, stage orchestration and the C ABI of
// include/avbd_b200.h.  One avbd_world = one CUDA stream on one GPU holding one
// world or a batch of independent worlds (ensemble).  No CPU fallback: every
// entry point fails when no CUDA device is usable.
#include "../../include/avbd_b200.h"
#include "avbd_kernels_graph.cuh"

#include <cub/cub.cuh>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <string>
#include <vector>

using namespace avbd;

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }

#define CK(expr)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (expr);                                                                     \
        if (e_ != cudaSuccess) return fail(AVBD_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
    } while (0)
#define TRY(expr)                \
    do {                         \
        int r_ = (expr);         \
        if (r_ < 0) return r_;   \
    } while (0)

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    int ensure(size_t n, bool keep, cudaStream_t s) {
        if (n <= cap) return 0;
        size_t ncap = std::max<size_t>(n + n / 2, 256);
        T* q = nullptr;
        CK(cudaMalloc(&q, ncap * sizeof(T)));
        if (keep && p && cap) {
            CK(cudaMemcpyAsync(q, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, s));
            CK(cudaStreamSynchronize(s));
        }
        if (p) cudaFree(p);
        p = q; cap = ncap;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

constexpr int kColourBlockMaxBodies = 2048;      // at most this many dynamic bodies: colouring rounds run in one block
static int colour_block_max() { const char* e = getenv("AVBD_COLOUR_BLOCK_MAX"); return e ? atoi(e) : kColourBlockMaxBodies; }      // read every time (tests switch it)
inline int blocks_for(long long n, int threads = kThreads) { return (int)std::max<long long>(1, (n + threads - 1) / threads); }

struct HostBody {   // what the host must remember to (re)classify bodies; dynamic state lives on the device
    float radius, invMass;
    int world, local;
};
struct HostForce { int type, a, b; };   // 0 joint 1 spring 2 ignore

} // namespace

struct avbd_world {
    int device = 0;
    cudaStream_t stream = nullptr;
    SolveParams prm{};
    long long launches = 0;      // this library's own kernels
    long long libLaunches = 0;   // CUB sort / scan passes (library code, counted apart)

    // bodies
    int n = 0, nDyn = 0, nWorlds = 1;
    bool anyUnvisited = false;   // some world holds two static bodies (a contact between them is visited by no body)
    std::vector<HostBody> hb;
    DevBuf<BodyPose> pose; DevBuf<BodyAux> aux; DevBuf<BodyVel> vel; DevBuf<BodyInit> init;
    DevBuf<float4> prevLin, size;
    int colouredBodies = -1;     // body count the `colour` array holds a valid colouring for (-1: none)
    bool freshColour = true;          // AVBD_ITERATED_COLOUR=1 clears it: rank the bodies by the previous colouring (see run_colour)
    bool keepColour = false;          // AVBD_KEEP_COLOUR=1: keep last step's colours where the new graph allows, colour only the bodies that lost theirs (see run_colour)
    bool coloursRestored = false;     // `colour` came from a snapshot and no graph has been built since
    bool topoSameAsLast = false;      // this step's manifolds have last step's slots and contact counts (np_build)
    DevBuf<int> colourWord;           // work words of the colouring rounds
    std::vector<int> savedColours;    // colouring read from a snapshot, uploaded by prepare()
    std::vector<cudaEvent_t> chunkEvents;   // avbd_download_state_chunked
    DevBuf<int> flags, worldId, localIdx, dynList, colWorkA, colWorkB;      // colWork*: uncoloured-body work lists of the colouring rounds
    bool topoDirty = true;
    bool contactDiagDone = false;   // the step's last dual pass reduced the contact diagnostics already
    int keyShift = 1;

    // broadphase
    float cell = 1.0f; unsigned tableSize = 256;
    DevBuf<unsigned> cellKey, cellKeySorted; DevBuf<int> cellVal, cellValSorted; DevBuf<int2> cellRange;
    DevBuf<int2> sortedCell; DevBuf<float4> sortedPos, sortedFrame, bodyFrame;
    DevBuf<int> largeList, worldLargeStart; int nLarge = 0;
    DevBuf<unsigned long long> pairs, cand, candSorted; int nCand = 0, nPairs = 0; long long lastPairs = 0, satLaunched = 0;
    DevBuf<unsigned long long> buildTiles;      // np_build's chained scan over its blocks
    DevBuf<int> mcount, visitCount, visitStart; DevBuf<int4> visits; int nContacts = 0;      // per-contact visit lists: small worlds (cluster loop) only
    DevBuf<int> bodyContacts; DevBuf<int> deg, estart, colCursor, sweepRange, freeList, linkedList; DevBuf<int> entries; DevBuf<int2> colVisit; DevBuf<float4> vgA, vgB, vgN;
    int2 hColVisit[64]; int sweepWarps[64] = {0}, sweepOff[64] = {0};      // per colour: its visit range, warps of the sweep, offset of its warp ranges
    int nFree = 0, nLinkedFree = 0; bool visitGeomStale = true, sweepRangesValid = false;            // contact geometry in visit order (VisitGeom), refreshed once per step
         // body -> manifold entries CSR (graph stage): colouring adjacency + the large-world sweep's work list

    // manifolds (ping-pong)
    struct MBuf { DevBuf<unsigned long long> key; DevBuf<int4> hdr; DevBuf<int> cstart, cM; DevBuf<float4> cA, cB, cN; DevBuf<ContactLP> lp; } mb[2];
    int cur = 0, nM = 0, nMPrev = 0;

    // graph
    DevBuf<int4> adjRange; DevBuf<unsigned> bKey, bKeySorted; DevBuf<int> bVal, bList;
    DevBuf<int> colour; DevBuf<unsigned> colKey, colKeySorted; DevBuf<int> colVal, colOrder; DevBuf<int2> colRange;
    int2 hColRange[64]; int nColours = 0; bool graphValid = false; bool forceRegraph = false; int maxColourCount = 0;
    long long graphReuses = 0; int persistentMaxBodies = 64;     // up to this many dynamic bodies the iteration loop is ONE cluster launch (solve_loop_cluster; AVBD_PERSISTENT_MAX_BODIES overrides): measured (tools/loop_modes.py, profiles/r02_stress1000_and_ensemble.md) it wins on the tiniest scenes only — TwoBlockDrop 8.1k vs 7.7k steps/s, Stack 5.5k vs 5.1k, Pyramid (55 bodies) 3.88k vs 3.78k, Wall (64) 3.44k vs 3.37k — and loses above: Stress1000 1.36k vs 1.52k, 1728 bodies 1.76k vs 2.08k, 8000 bodies 0.80k vs 1.47k.  Same additions in the same order: results are bit-identical either way
    size_t tilesCleared = 0;  // scan tiles bp_cells cleared at the start of this step's collision stage
    bool cellRaw = true;    // default: the per-cell sweep emits the sphere pairs, np_sat culls them.  AVBD_BROADPHASE=body: the per-body sweep instead
                            // (1M-box grid: bp_sweep 472 us against bp_sweep_cells<false> 328 us; small worlds: no difference)
    bool bodySweep = true;  // AVBD_BROADPHASE=cell: fused per-cell sweep + SAT cull instead of the per-body sweep + separate cull
    double hostLoopSec = 0.0; long long hostLoopSteps = 0;   // AVBD_DEBUG: host time spent issuing the iteration loop's launches
    int loopMode = 0;      // AVBD_LOOP: 0 auto, 1 per-colour launches, 2 cooperative grid loop, 3 tile cluster loop where eligible, 4 warp-pipeline cluster loop

    // user forces
    std::vector<JointRec> hJoints; std::vector<SpringRec> hSprings; std::vector<HostForce> hForces;
    DevBuf<JointRec> joints; DevBuf<SpringRec> springs; DevBuf<int> fadjStart, fadj; DevBuf<unsigned long long> excl;
    int nExcl = 0; bool forcesDirty = false; int uploadedJoints = 0, uploadedSprings = 0;

    // counters / diagnostics
    Counters* dCnt = nullptr; Counters* hCnt = nullptr;
    DevBuf<Diag> dDiag; Diag* hDiag = nullptr; size_t hDiagCap = 0;
    DevBuf<float> dx;
    DevBuf<char> temp;

    // per-kernel profiling (avbd_set_profiling): events around every primal sweep and dual pass
    bool profiling = false;
    // one record per profiled step, resolved at avbd_get_profile: no host wait inside or between the profiled steps
    struct ProfStep { cudaEvent_t ev[7]; std::vector<cudaEvent_t> dual; int timedDuals, total, colours, duals, deferred, rebuilt; long long n, nDyn, pairs, cand, nM, nMPrev, contacts, visits; };
    std::vector<ProfStep> profSteps; int profUsed = 0; ProfStep* profCur = nullptr;
    avbd_profile prof{};
    DevBuf<float> stateDev;

    // timing
    cudaEvent_t ev[9] = {};
    bool timed = false;
    avbd_step_stats stats{};

    ManifoldSet mset(int which) {
        MBuf& b = mb[which];
        ManifoldSet s; s.key = b.key.p; s.hdr = b.hdr.p; s.cstart = b.cstart.p; s.cM = b.cM.p; s.cA = b.cA.p; s.cB = b.cB.p; s.cN = b.cN.p; s.lp = b.lp.p;
        return s;
    }
    int ensure_manifolds(int which, size_t m) {
        MBuf& b = mb[which];
        TRY(b.key.ensure(m, false, stream)); TRY(b.hdr.ensure(m, false, stream)); TRY(b.cstart.ensure(m + 1, false, stream)); TRY(b.cM.ensure(4 * m, false, stream));
        TRY(b.cA.ensure(4 * m, false, stream)); TRY(b.cB.ensure(4 * m, false, stream)); TRY(b.cN.ensure(4 * m, false, stream));
        TRY(b.lp.ensure(4 * m, false, stream));
        return 0;
    }
    VisitGeom vgeom() { VisitGeom g; g.a = vgA.p; g.b = vgB.p; g.n = vgN.p; return g; }
    BodyView bview() {
        BodyView v; v.pose = pose.p; v.aux = aux.p; v.vel = vel.p; v.init = init.p; v.prevLin = prevLin.p; v.size = size.p;
        v.flags = flags.p; v.worldId = worldId.p; v.localIdx = localIdx.p; v.n = n;
        return v;
    }
    GridView gview() {
        GridView g; g.cell = cell; g.tableMask = tableSize - 1; g.key = cellKey.p; g.keySorted = cellKeySorted.p;
        g.val = cellVal.p; g.valSorted = cellValSorted.p; g.cellRange = cellRange.p;
        g.sortedCell = sortedCell.p; g.sortedPos = sortedPos.p; g.sortedFrame = bodySweep ? nullptr : sortedFrame.p; g.bodyFrame = bodyFrame.p; g.largeList = largeList.p; g.worldLargeStart = worldLargeStart.p;
        return g;
    }
    ForceView fview() {
        ForceView f; f.joints = joints.p; f.nJoints = (int)hJoints.size(); f.springs = springs.p; f.nSprings = (int)hSprings.size();
        bool any = f.nJoints + f.nSprings > 0;
        f.adjStart = any ? fadjStart.p : nullptr; f.adj = any ? fadj.p : nullptr;
        return f;
    }
};

namespace {

constexpr size_t kPickSlotOffset = (sizeof(Counters) + 7) / 8 * 8;     // 8-byte aligned scratch word after the counters (avbd_pick)

// Stage boundary k of the current step: the last-step events avbd_get_step_stats reads, and the profiled step's own record.
inline void stage_event(avbd_world* w, int k) {
    if (w->timed) cudaEventRecord(w->ev[k], w->stream);
    if (w->profCur) cudaEventRecord(w->profCur->ev[k], w->stream);
}

int read_counters(avbd_world* w) {
    CK(cudaMemcpyAsync(w->hCnt, w->dCnt, sizeof(Counters), cudaMemcpyDeviceToHost, w->stream));
    CK(cudaStreamSynchronize(w->stream));
    return 0;
}

template <class K, class V>
int sort_pairs(avbd_world* w, const K* kin, K* kout, const V* vin, V* vout, int n, int bits) {
    if (n <= 0) return 0;
    size_t bytes = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, kin, kout, vin, vout, n, 0, bits, w->stream));
    TRY(w->temp.ensure(bytes, false, w->stream));
    CK(cub::DeviceRadixSort::SortPairs(w->temp.p, bytes, kin, kout, vin, vout, n, 0, bits, w->stream));
    w->libLaunches += 1 + (bits + 7) / 8 * 2;
    return 0;
}
template <class K>
int sort_keys(avbd_world* w, const K* kin, K* kout, int n, int bits) {
    if (n <= 0) return 0;
    size_t bytes = 0;
    CK(cub::DeviceRadixSort::SortKeys(nullptr, bytes, kin, kout, n, 0, bits, w->stream));
    TRY(w->temp.ensure(bytes, false, w->stream));
    CK(cub::DeviceRadixSort::SortKeys(w->temp.p, bytes, kin, kout, n, 0, bits, w->stream));
    w->libLaunches += 1 + (bits + 7) / 8 * 2;
    return 0;
}
// Exclusive prefix sum of a short array in ONE block (CUB's device scan is two launches; a small world's step is a chain of launches).
constexpr int kScanSmallMax = 16384;
__global__ void __launch_bounds__(1024) scan_small(const int* __restrict__ in, int* __restrict__ out, int n) {
    cudaGridDependencySynchronize();
    __shared__ int sWarp[32];
    const int per = (n + 1023) / 1024;                     // consecutive items per thread
    const int b = threadIdx.x * per, e = b + per < n ? b + per : n;
    int sum = 0;
    for (int i = b; i < e; ++i) sum += in[i];
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    int incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int up = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += up; }
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
int exclusive_scan(avbd_world* w, const int* in, int* out, int n) {
    if (n <= 0) return 0;
    if (n <= kScanSmallMax && in != out) { launch_dep(scan_small, dim3(1), dim3(1024), 0, w->stream, in, out, n); w->launches++; return 0; }
    size_t bytes = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, w->stream));
    TRY(w->temp.ensure(bytes, false, w->stream));
    CK(cub::DeviceScan::ExclusiveSum(w->temp.p, bytes, in, out, n, w->stream));
    w->libLaunches += 2;
    return 0;
}

int bits_for(unsigned long long maxValue) { int b = 1; while ((1ull << b) <= maxValue) ++b; return b; }

__global__ void rekey_manifolds(ManifoldSet ms, int nM, int keyShift) {
    cudaGridDependencySynchronize();
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m < nM) ms.key[m] = ((unsigned long long)(unsigned)ms.hdr[m].x << keyShift) | (unsigned)ms.hdr[m].y;
}

int np_sat_launch(avbd_world* w, const BodyView& bv, const PairSink& raw, int expect, const PairSink& out) {
    // persistent warps striding over the candidates (the count is read on the device): enough blocks to fill the machine, no more
    static int residentDev[64] = {0};
    int& resident = residentDev[w->device & 63];
    if (!resident) {
        int per = 0, sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, w->device);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, np_sat, kThreads, 0) != cudaSuccess || per < 1) { cudaGetLastError(); per = 4; }
        resident = sms * per;
    }
    int grid = std::max(1, std::min(resident, blocks_for(expect, 32 * (kThreads / 32))));
    launch_dep(np_sat, dim3(grid), dim3(kThreads), 0, w->stream, bv, (const float4*)w->bodyFrame.p, raw.keys, raw.count, raw.cap, w->keyShift, w->excl.p, w->nExcl, out);
    w->satLaunched = (long long)raw.cap;        // every candidate the list holds is tested, whatever `expect` said
    w->launches++;
    return 0;
}

// Rebuilds everything that depends on the body / user-force SET (not on poses).
int prepare(avbd_world* w) {
    if (!w->topoDirty && !w->forcesDirty) return 0;
    cudaStream_t s = w->stream;
    int n = w->n;
    if (w->topoDirty && n > 0) {
        std::vector<float> radii(n);
        for (int i = 0; i < n; ++i) radii[i] = w->hb[i].radius;
        std::vector<float> sorted = radii;
        std::nth_element(sorted.begin(), sorted.begin() + n / 2, sorted.end());
        float median = sorted[n / 2];
        float thresh = 4.0f * median;
        float maxSmall = 0.0f;
        std::vector<int> flags(n), wid(n), lidx(n), dyn, large;
        int nWorlds = 1;
        for (int i = 0; i < n; ++i) {
            bool isLarge = radii[i] > thresh;
            if (!isLarge) maxSmall = std::max(maxSmall, radii[i]);
            flags[i] = (w->hb[i].invMass > 0.0f ? kDynamic : 0) | (isLarge ? kLarge : 0);
            wid[i] = w->hb[i].world; lidx[i] = w->hb[i].local;
            nWorlds = std::max(nWorlds, wid[i] + 1);
            if (flags[i] & kDynamic) dyn.push_back(i);
            if (isLarge) large.push_back(i);
        }
        w->nWorlds = nWorlds; w->nDyn = (int)dyn.size(); w->nLarge = (int)large.size();
        {   // a contact no dynamic body visits needs two static bodies in one world
            std::vector<int> statics(nWorlds, 0);
            w->anyUnvisited = false;
            for (int i = 0; i < n; ++i) if (!(flags[i] & kDynamic) && ++statics[wid[i]] > 1) w->anyUnvisited = true;
        }
        w->cell = std::max(2.02f * maxSmall, 1e-3f);
        unsigned table = 256; while (table < 2u * (unsigned)n) table <<= 1;
        w->tableSize = table;
        std::vector<int> wls(nWorlds + 1, 0);
        for (int l : large) wls[wid[l] + 1]++;
        for (int k = 0; k < nWorlds; ++k) wls[k + 1] += wls[k];   // large is ascending by index => grouped by world
        TRY(w->flags.ensure(n, false, s)); TRY(w->worldId.ensure(n, false, s)); TRY(w->localIdx.ensure(n, false, s));
        TRY(w->dynList.ensure(std::max(1, w->nDyn), false, s)); TRY(w->largeList.ensure(std::max(1, w->nLarge), false, s));
        TRY(w->worldLargeStart.ensure(nWorlds + 1, false, s));
        CK(cudaMemcpyAsync(w->flags.p, flags.data(), n * sizeof(int), cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(w->worldId.p, wid.data(), n * sizeof(int), cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(w->localIdx.p, lidx.data(), n * sizeof(int), cudaMemcpyHostToDevice, s));
        if (w->nDyn) CK(cudaMemcpyAsync(w->dynList.p, dyn.data(), dyn.size() * sizeof(int), cudaMemcpyHostToDevice, s));
        if (w->nLarge) CK(cudaMemcpyAsync(w->largeList.p, large.data(), large.size() * sizeof(int), cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(w->worldLargeStart.p, wls.data(), wls.size() * sizeof(int), cudaMemcpyHostToDevice, s));
        CK(cudaStreamSynchronize(s));
        // per-body scratch
        TRY(w->cellKey.ensure(n, false, s)); TRY(w->cellKeySorted.ensure(n, false, s)); TRY(w->cellVal.ensure(n, false, s));
        TRY(w->cellValSorted.ensure(n, false, s)); TRY(w->sortedCell.ensure(n, false, s)); TRY(w->sortedPos.ensure(n, false, s));
        if (!w->bodySweep) TRY(w->sortedFrame.ensure(6 * (size_t)n, false, s));
        TRY(w->bodyFrame.ensure(6 * (size_t)n, false, s));
        TRY(w->cellRange.ensure(table, false, s));
        TRY(w->adjRange.ensure(n, false, s)); TRY(w->colour.ensure(n, false, s));
        TRY(w->colKey.ensure(std::max(1, w->nDyn), false, s)); TRY(w->colKeySorted.ensure(std::max(1, w->nDyn), false, s));
        TRY(w->colVal.ensure(std::max(1, w->nDyn), false, s)); TRY(w->colOrder.ensure(std::max(1, w->nDyn), false, s));
        TRY(w->colRange.ensure(64, false, s));
        TRY(w->dDiag.ensure(nWorlds, false, s));
        if ((size_t)nWorlds > w->hDiagCap) {
            if (w->hDiag) cudaFreeHost(w->hDiag);
            CK(cudaMallocHost(&w->hDiag, sizeof(Diag) * nWorlds));
            w->hDiagCap = nWorlds;
            std::memset(w->hDiag, 0, sizeof(Diag) * nWorlds);
        }
        int shift = bits_for((unsigned long long)std::max(1, n - 1));
        if (shift != w->keyShift) {
            w->keyShift = shift;
            if (w->nM > 0) { launch_dep(rekey_manifolds, dim3(blocks_for(w->nM)), dim3(kThreads), 0, s, w->mset(w->cur), w->nM, shift); w->launches++; }
            w->forcesDirty = true;   // exclusion keys are packed with keyShift too
        }
        w->graphValid = false;
        w->colouredBodies = -1;          // the colour array was reallocated / the body set changed
        if ((int)w->savedColours.size() == n && n > 0) {                     // ... unless a snapshot brought the colouring along
            CK(cudaMemcpyAsync(w->colour.p, w->savedColours.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, s));
            CK(cudaStreamSynchronize(s));
            w->colouredBodies = n; w->coloursRestored = true;
        }
        w->savedColours.clear();
    }
    if (w->forcesDirty || w->topoDirty) {
        int nj = (int)w->hJoints.size(), ns = (int)w->hSprings.size();
        // device records: only append new ones so lambda/penalty of existing rows survive
        TRY(w->joints.ensure(std::max(1, nj), true, s)); TRY(w->springs.ensure(std::max(1, ns), true, s));
        if (nj > w->uploadedJoints)
            CK(cudaMemcpyAsync(w->joints.p + w->uploadedJoints, w->hJoints.data() + w->uploadedJoints, (nj - w->uploadedJoints) * sizeof(JointRec), cudaMemcpyHostToDevice, s));
        if (ns > w->uploadedSprings)
            CK(cudaMemcpyAsync(w->springs.p + w->uploadedSprings, w->hSprings.data() + w->uploadedSprings, (ns - w->uploadedSprings) * sizeof(SpringRec), cudaMemcpyHostToDevice, s));
        w->uploadedJoints = nj; w->uploadedSprings = ns;
        // CSR body -> user force rows, and the sorted pair-exclusion list (rigid.cpp:61-69 for non-manifold forces)
        std::vector<int> start(n + 1, 0), adj;
        std::vector<unsigned long long> ex;
        std::vector<std::vector<int>> per(n);
        int ji = 0, si = 0;
        for (const HostForce& f : w->hForces) {
            int idx = f.type == 0 ? ji++ : (f.type == 1 ? si++ : 0);
            if (f.type != 2) {
                if (f.a >= 0) per[f.a].push_back(idx * 4 + f.type * 2 + 1);
                if (f.b >= 0) per[f.b].push_back(idx * 4 + f.type * 2 + 0);
            }
            if (f.a >= 0 && f.b >= 0) {
                unsigned hi = (unsigned)std::max(f.a, f.b), lo = (unsigned)std::min(f.a, f.b);
                ex.push_back(((unsigned long long)hi << w->keyShift) | lo);
            }
        }
        for (int i = 0; i < n; ++i) { start[i + 1] = start[i] + (int)per[i].size(); adj.insert(adj.end(), per[i].begin(), per[i].end()); }
        std::sort(ex.begin(), ex.end());
        ex.erase(std::unique(ex.begin(), ex.end()), ex.end());
        w->nExcl = (int)ex.size();
        TRY(w->fadjStart.ensure(n + 1, false, s)); TRY(w->fadj.ensure(std::max<size_t>(1, adj.size()), false, s));
        TRY(w->excl.ensure(std::max<size_t>(1, ex.size()), false, s));
        CK(cudaMemcpyAsync(w->fadjStart.p, start.data(), start.size() * sizeof(int), cudaMemcpyHostToDevice, s));
        if (!adj.empty()) CK(cudaMemcpyAsync(w->fadj.p, adj.data(), adj.size() * sizeof(int), cudaMemcpyHostToDevice, s));
        if (!ex.empty()) CK(cudaMemcpyAsync(w->excl.p, ex.data(), ex.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, s));
        CK(cudaStreamSynchronize(s));
        w->graphValid = false;
    }
    w->topoDirty = false; w->forcesDirty = false;
    return 0;
}

// Broadphase (+ SAT cull): leaves the key-sorted pairs in candSorted (and, with `sat`, their winning SAT axis in
// the upper bits of each key).  sat = false stops at the sphere-overlap pairs (stage API: the reference's solver.cpp:262-266
// candidate set).  sat = true merges last step's manifolds whose spheres no longer overlap, applies the exclusion list
// and the 15-axis test, so what comes out is exactly the set of manifolds to build.  One host sync (sizes + overflow).
int run_broadphase(avbd_world* w, bool sat, bool clearStepScratch = false) {
    cudaStream_t s = w->stream;
    int n = w->n;
    w->nCand = 0; w->nPairs = 0;
    if (n == 0) return 0;
    BodyView bv = w->bview(); GridView gv = w->gview();
    // clearStepScratch (a whole step): the per-world diagnostics and the manifold build's scan tiles are cleared on the way
    launch_dep(bp_cells, dim3(blocks_for(n)), dim3(kThreads), 0, s, bv, gv, w->dCnt,
               clearStepScratch ? (int*)w->dDiag.p : (int*)nullptr, clearStepScratch ? (int)(sizeof(Diag) / sizeof(int)) * w->nWorlds : 0,
               clearStepScratch ? (int*)w->buildTiles.p : (int*)nullptr, clearStepScratch ? 2 * (int)w->buildTiles.cap : 0);
    w->tilesCleared = clearStepScratch ? w->buildTiles.cap : 0;
    int tbits = bits_for(w->tableSize);   // sentinel bucket == tableSize needs one more bit
    TRY(sort_pairs(w, w->cellKey.p, w->cellKeySorted.p, w->cellVal.p, w->cellValSorted.p, n, tbits));
    launch_dep(bp_cell_bounds, dim3(blocks_for(n)), dim3(kThreads), 0, s, bv, gv);
    w->launches += 2;
    if (w->pairs.cap == 0) TRY(w->pairs.ensure((size_t)std::max(1024, 12 * n), false, s));
    if (w->cand.cap == 0) TRY(w->cand.ensure((size_t)std::max(1024, 3 * n), false, s));
    for (int attempt = 0; attempt < 8; ++attempt) {
        if (attempt > 0) CK(cudaMemsetAsync(w->dCnt, 0, sizeof(Counters), s));        // the first attempt's counters were cleared by bp_cells
        PairSink raw; raw.keys = w->pairs.p; raw.codes = nullptr; raw.cap = (int)w->pairs.cap; raw.keyShift = w->keyShift; raw.codeShift = 0;
        raw.count = &w->dCnt->nPairs; raw.cnt = w->dCnt; raw.overflowBit = 1;
        PairSink out; out.keys = w->cand.p; out.codes = nullptr; out.cap = (int)w->cand.cap; out.keyShift = w->keyShift; out.codeShift = 2 * w->keyShift;   // 2 * keyShift + 5 <= 64
        out.count = &w->dCnt->nCand; out.cnt = w->dCnt; out.overflowBit = 2;
        // small-vs-small pairs: one warp per 32 cell-sorted bodies; with `sat` the cull is fused in and only survivors are written.
        // AVBD_BROADPHASE=cell selects the fused per-cell kernel (A/B measurements; default is the per-body sweep + separate cull).
        // bodies per warp: 32 on a world that fills the machine; a small one is cut finer (a warp walks its cells one after another)
        const int bpw = n >= 37888 ? 32 : (n >= 18944 ? 16 : (n >= 9472 ? 8 : 4));
        const int cellBlocks = blocks_for(n, bpw * (kThreads / 32));
        if (sat && !w->bodySweep) launch_dep(bp_sweep_cells<true>, dim3(cellBlocks), dim3(kThreads), 0, s, bv, gv, (const unsigned long long*)w->excl.p, w->nExcl, out, bpw);
        else if (!w->bodySweep || w->cellRaw) launch_dep(bp_sweep_cells<false>, dim3(cellBlocks), dim3(kThreads), 0, s, bv, gv, (const unsigned long long*)nullptr, 0, raw, bpw);
        else launch_dep(bp_sweep, dim3(blocks_for(16ll * n)), dim3(kThreads), 0, s, bv, gv, raw);
        const int nOld = sat ? w->nM : 0;
        if (w->nLarge || nOld > 0) { launch_dep(bp_side_pairs, dim3(blocks_for(std::max(w->nLarge ? n : 0, nOld))), dim3(kThreads), 0, s, bv, gv, w->mset(w->cur), nOld, raw, w->nLarge ? 1 : 0); w->launches++; }
        w->launches += 1;
        if (sat) {
            // sized by the pair count of the previous step (+ slack); the kernel reads the real count, a shortfall shows as overflow bit 16
            long long expect = std::min<long long>((long long)raw.cap, std::max<long long>(w->lastPairs + w->lastPairs / 8 + 4096, 1024));
            np_sat_launch(w, bv, raw, (int)expect, out);
        }
        TRY(read_counters(w));
        bool rawOver = w->hCnt->nPairs > raw.cap, satShort = sat && w->hCnt->nPairs > w->satLaunched, outOver = sat && w->hCnt->nCand > (int)w->cand.cap;
        w->lastPairs = w->hCnt->nPairs;
        if (!rawOver && !satShort && !outOver) { w->nPairs = w->hCnt->nPairs + w->hCnt->nSphere; w->nCand = sat ? w->hCnt->nCand : w->hCnt->nPairs; break; }
        if (rawOver) TRY(w->pairs.ensure((size_t)w->hCnt->nPairs + w->hCnt->nPairs / 4 + 1024, false, s));
        if (outOver) TRY(w->cand.ensure((size_t)w->hCnt->nCand + w->hCnt->nCand / 4 + 1024, false, s));
        if (attempt == 7) return fail(AVBD_ERR_CAPACITY, "pair buffer kept overflowing");
    }
    if (sat) {
        TRY(w->candSorted.ensure(w->cand.cap, false, s));
        TRY(sort_keys(w, w->cand.p, w->candSorted.p, w->nCand, 2 * w->keyShift));       // the SAT codes ride in the keys' upper bits
    } else {
        TRY(w->candSorted.ensure(w->pairs.cap, false, s));
        TRY(sort_keys(w, w->pairs.p, w->candSorted.p, w->nCand, 2 * w->keyShift));
    }
    CK(cudaGetLastError());
    return 0;
}

int run_collide(avbd_world* w) {
    cudaStream_t s = w->stream;
    TRY(prepare(w));
    stage_event(w, 0);
    w->nMPrev = w->nM;
    TRY(run_broadphase(w, true, true));
    stage_event(w, 1);
    int nSurv = w->nCand;
    int nxt = w->cur ^ 1;
    if (nSurv > 0) {
        TRY(w->ensure_manifolds(nxt, nSurv));
        static bool polySmem[64] = {false};          // a function attribute is per device
        if (!polySmem[w->device & 63]) {
            cudaFuncSetAttribute(np_build, cudaFuncAttributeMaxDynamicSharedMemorySize, kBuildThreads * kPolyFloatsPerThread * (int)sizeof(float));
            polySmem[w->device & 63] = true;
        }
        // contacts go straight to their dense place, in manifold order (chained scan over the build's blocks)
        int buildBlocks = blocks_for(nSurv, kBuildThreads);
        TRY(w->buildTiles.ensure((size_t)buildBlocks, false, s));
        if (w->tilesCleared < (size_t)buildBlocks)      // grown since bp_cells cleared it (or not cleared at all)
            CK(cudaMemsetAsync(w->buildTiles.p, 0, (size_t)buildBlocks * sizeof(unsigned long long), s));
        launch_dep(np_build, dim3(buildBlocks), dim3(kBuildThreads), kBuildThreads * kPolyFloatsPerThread * sizeof(float), s,
            w->bview(), w->candSorted.p, (const int*)nullptr, nSurv, w->keyShift, w->mset(w->cur), w->nM, w->mset(nxt), w->prm, w->dCnt, w->buildTiles.p);
        w->launches++;
        TRY(read_counters(w));
        w->nContacts = w->hCnt->nContacts;
        if (w->hCnt->overflow & 8) return fail(AVBD_ERR_CUDA, "broadphase emitted a pair twice");
    } else {
        w->nContacts = 0;
    }
    // adjacency, colouring and visit lists only depend on (pair, contact count) per slot: keep them when nothing moved
    w->topoSameAsLast = nSurv == w->nM && nSurv > 0 && !w->hCnt->topoChanged;
    bool sameTopology = w->graphValid && w->topoSameAsLast && !w->forceRegraph;
    w->cur = nxt; w->nM = nSurv;
    ForceView fv = w->fview();
    if (fv.nJoints + fv.nSprings > 0) {
        launch_dep(decay_user_forces, dim3(blocks_for(fv.nJoints + fv.nSprings)), dim3(kThreads), 0, s, fv, w->prm);
        w->launches++;
    }
    w->graphValid = sameTopology;
    w->visitGeomStale = true;         // every contact was rebuilt
    stage_event(w, 2);
    CK(cudaGetLastError());
    return 0;
}

int run_predict(avbd_world* w) {
    TRY(prepare(w));
    if (w->n == 0) return 0;
    launch_dep(predict_bodies, dim3(blocks_for(w->n)), dim3(kThreads), 0, w->stream, w->bview(), w->prm, w->dDiag.p);
    w->launches++;
    CK(cudaGetLastError());
    return 0;
}

// Body-aligned warp ranges of every colour's visits (avbd_solve.cu: warp_ranges), one launch, no host wait.
int build_sweep_ranges(avbd_world* w) {
    cudaStream_t s = w->stream;
    int total = 0;
    for (int c = 0; c < w->nColours; ++c) {
        int nv = w->hColVisit[c].y - w->hColVisit[c].x;
        w->sweepWarps[c] = nv > 0 ? primal_sweep_warps(nv) : 0;
        w->sweepOff[c] = total; total += w->sweepWarps[c] + 1;
    }
    TRY(w->sweepRange.ensure((size_t)std::max(1, total), false, s));
    int nw[64], off[64];
    for (int c = 0; c < 64; ++c) { nw[c] = c < w->nColours ? std::max(1, w->sweepWarps[c]) : 1; off[c] = c < w->nColours ? w->sweepOff[c] : 0; }
    if (w->nColours > 0 && w->nContacts > 0) { launch_warp_ranges(s, w->visitStart.p, w->colRange.p, w->nColours, nw, off, w->sweepRange.p); w->launches++; }
    w->sweepRangesValid = true;
    return 0;
}

int run_colour(avbd_world* w) {
    cudaStream_t s = w->stream;
    TRY(prepare(w));
    w->nColours = 0;
    if (w->n == 0 || w->nDyn == 0) { w->graphValid = true; return 0; }
    int nM = w->nM, n = w->n;
    ManifoldSet ms = w->mset(w->cur);
    // (how the colouring starts is decided here because the stage's one prologue launch also initialises its work words; see below)
    const bool keep = w->keepColour && w->freshColour && w->colouredBodies == n;
    const bool havePrev = !w->freshColour && w->colouredBodies == n;
    const bool keepSaved = havePrev && w->coloursRestored && w->topoSameAsLast;
    w->coloursRestored = false;
    TRY(w->deg.ensure((size_t)n + 1, false, s)); TRY(w->estart.ensure((size_t)n + 1, false, s));
    TRY(w->colourWord.ensure(n, false, s)); TRY(w->colCursor.ensure(4, false, s));
    TRY(w->visitCount.ensure((size_t)w->nDyn + 1, false, s)); TRY(w->visitStart.ensure((size_t)w->nDyn + 1, false, s));
    // a small world: the whole stage in one block (avbd_kernels_graph.cuh: graph_small)
    const bool smallGraph = nM <= kSmallGraphMax && w->nDyn <= colour_block_max() && w->nDyn <= kSmallGraphMax && n <= kSmallGraphMax + 64     /* its colouring is the one-block one: same size limit (2744 bodies: 1729 against 1808 steps/s) */ && !keep && !havePrev && !getenv("AVBD_NO_SMALL_GRAPH");
    if (smallGraph) {
        TRY(w->bKeySorted.ensure(std::max(1, nM), false, s)); TRY(w->bList.ensure(std::max(1, nM), false, s));
        TRY(w->entries.ensure((size_t)std::max(1, 2 * nM), false, s));
        TRY(w->visits.ensure((size_t)std::max(1, 2 * w->nContacts), false, s));
        TRY(w->freeList.ensure((size_t)w->nDyn, false, s)); TRY(w->linkedList.ensure((size_t)w->nDyn, false, s));
        TRY(w->colVisit.ensure(64, false, s));
        SmallGraph g;
        g.flags = w->flags.p; g.n = n; g.dynList = w->dynList.p; g.nDyn = w->nDyn; g.localIdx = w->localIdx.p;
        g.hdr = ms.hdr; g.cstart = ms.cstart; g.nM = nM; g.sortBits = bits_for((unsigned long long)n);
        g.adjRange = w->adjRange.p; g.bList = w->bList.p; g.bKeySorted = w->bKeySorted.p; g.deg = w->deg.p; g.estart = w->estart.p; g.entries = w->entries.p;
        g.colour = w->colour.p; g.colKeySorted = w->colKeySorted.p; g.colOrder = w->colOrder.p; g.colRange = w->colRange.p;
        g.visitCount = w->visitCount.p; g.visitStart = w->visitStart.p; g.visits = w->visits.p; g.colVisit = w->colVisit.p;
        g.freeList = w->freeList.p; g.linkedList = w->linkedList.p; g.aux = w->aux.p; g.cnt = w->dCnt;
        launch_dep(graph_small, dim3(1), dim3(kSmallGraphThreads), 0, s, g, w->fview());
        w->launches++;
        w->colouredBodies = n;
    } else {
    launch_dep(graph_prologue, dim3(blocks_for((long long)n + 1)), dim3(kThreads), 0, s, w->flags.p, n, w->nDyn, w->adjRange.p, w->deg.p, w->colRange.p, w->visitCount.p,
               w->colCursor.p, w->dCnt, (keepSaved || keep) ? 0 : (havePrev ? 2 : 1), w->colourWord.p, w->colour.p);
    w->launches++;
    if (nM > 0) {
        TRY(w->bKey.ensure(nM, false, s)); TRY(w->bKeySorted.ensure(nM, false, s)); TRY(w->bVal.ensure(nM, false, s)); TRY(w->bList.ensure(nM, false, s));
        launch_dep(adj_a_ranges, dim3(blocks_for(nM)), dim3(kThreads), 0, s, ms.hdr, nM, w->flags.p, n, w->adjRange.p, w->bKey.p, w->bVal.p);
        TRY(sort_pairs(w, w->bKey.p, w->bKeySorted.p, w->bVal.p, w->bList.p, nM, bits_for((unsigned long long)n)));
        launch_dep(adj_b_ranges, dim3(blocks_for(nM)), dim3(kThreads), 0, s, w->bKeySorted.p, nM, n, w->adjRange.p);
        w->launches += 2;
    } else {
        TRY(w->bList.ensure(1, false, s));
    }
    // body -> manifold entries (CSR by body): the colouring's adjacency and the large-world sweep's work list
    TRY(w->entries.ensure((size_t)std::max(1, 2 * nM), false, s));
    TRY(w->bodyContacts.ensure((size_t)n, false, s));
    launch_dep(entry_count, dim3(blocks_for(w->nDyn)), dim3(kThreads), 0, s, w->dynList.p, w->nDyn, w->adjRange.p, w->bList.p, ms.hdr, w->deg.p, w->bodyContacts.p);
    TRY(exclusive_scan(w, w->deg.p, w->estart.p, n + 1));
    launch_dep(entry_fill, dim3(blocks_for(w->nDyn)), dim3(kThreads), 0, s, w->dynList.p, w->nDyn, w->adjRange.p, w->bList.p, ms.hdr, w->estart.p, w->entries.p);
    w->launches += 2;
    ForceView fv = w->fview();
    // Default: hashed priority order only — the colouring is a pure function of the current graph.
    // Opt-in (AVBD_ITERATED_COLOUR=1): the previous colouring of the same body set ranks the bodies (avbd_kernels_graph.cuh: outranks),
    // Culberson's iterated greedy.  Fewer colours and fewer rounds (1M-box grid 9 -> 7 colours, step 7.73 -> 7.59 ms; 8192-world
    // ensemble 4 -> 3, 3.55 -> 3.39 ms; 8000-box grid 6 -> 5, 0.81 -> 0.73 ms), but a chain of bodies ends up two-coloured and a
    // red/black sweep carries a load change two links per iteration: the 10-box Stack then rests 1.8e-3 m from the reference's heights
    // instead of < 1e-3 (tests/test_gpu_scenes.py), so it is not the default.  The colouring then also depends on the world's history:
    // a snapshot carries the colours, and a restored world whose first step finds the topology it was saved with keeps them as they are.
    // Opt-in (AVBD_KEEP_COLOUR=1; kept colouring, avbd_kernels_graph.cuh: kept_word): a body keeps last step's colour unless a
    // higher-priority neighbour of the new graph holds the same one; only the bodies that lose theirs are coloured again.  Measured: the
    // graph stage gets cheaper (1M-box grid 0.95 -> 0.65 ms, 8192-world ensemble 0.42 -> 0.34, Stress1000 0.135 -> 0.118) but the colour
    // count creeps up while a pile settles (a re-coloured body takes the smallest colour its kept neighbours leave, nobody ever moves
    // down): 1M grid 9 -> 11-12 colours, Stress1000 5 -> 7, and every colour is a launch per sweep — the sweeps lose what the graph stage
    // gains (1M grid solve 4.92 -> 5.22 ms, Stress1000 0.71 -> 0.88 ms per step), and the settled Pyramid misses its rest-height gate.
    // Hence not the default.
    const int* keepFlags = keep ? (const int*)w->flags.p : (const int*)nullptr;
    w->colouredBodies = -1;
    // Jones-Plassmann rounds (= the sequential greedy colouring in hashed-priority order, whatever the timing), all in ONE launch:
    // one block for small worlds, a cooperative grid with a grid barrier per round otherwise — no host check of the uncoloured count
    // between rounds.  If the cooperative launch is refused, rounds are launched in batches with a host check per batch.
    bool coloured = keepSaved;
    bool keysDone = false;          // the one-launch round kernels also write the colour sort's keys
    if (coloured) {
    } else if (w->nDyn <= colour_block_max()) {
        launch_dep(colour_rounds_block, dim3(1), dim3(kColourBlockThreads), 0, s, w->dynList.p, w->nDyn, w->estart.p, w->entries.p, fv, w->localIdx.p, w->colourWord.p, w->colour.p, w->dCnt, n, keepFlags, w->colKey.p, w->colVal.p);
        w->launches++;
        coloured = true; keysDone = true;
    } else {
        static int residentDev[64] = {0};
        int& resident = residentDev[w->device & 63];
        if (!resident) {
            int per = 0, sms = 148;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, w->device);
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, colour_rounds_grid, kColourGridThreads, 0) != cudaSuccess || per < 1) { cudaGetLastError(); per = 0; }
            int coop = 0; cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, w->device);
            resident = (coop && per > 0) ? sms * per : -1;
        }
        if (resident > 0 && !getenv("AVBD_NO_COOP_COLOUR")) {
            TRY(w->colWorkA.ensure((size_t)w->nDyn, false, s)); TRY(w->colWorkB.ensure((size_t)w->nDyn, false, s));
            const int* dynList = w->dynList.p; int nDyn = w->nDyn; const int* estart = w->estart.p; const int* entries = w->entries.p;
            const int* localIdx = w->localIdx.p; volatile int* word = w->colourWord.p; int* colour = w->colour.p; Counters* cnt = w->dCnt;
            int* listA = w->colWorkA.p; int* listB = w->colWorkB.p; int* cursors = w->colCursor.p;
            int nBodies = n;
            unsigned* ckey = w->colKey.p; int* cval = w->colVal.p;
            void* args[] = {&dynList, &nDyn, &estart, &entries, &fv, &localIdx, &word, &colour, &cnt, &listA, &listB, &cursors, &keepFlags, &nBodies, &ckey, &cval};
            int grid = std::min(resident, blocks_for(w->nDyn, kColourGridThreads));
            cudaError_t e = cudaLaunchCooperativeKernel((void*)colour_rounds_grid, dim3(grid), dim3(kColourGridThreads), args, 0, s);
            if (e == cudaSuccess) { w->launches++; coloured = true; keysDone = true; } else { cudaGetLastError(); resident = -1; }
        }
    }
    if (!coloured) {
        if (keep) { launch_dep(colour_keep, dim3(blocks_for(n)), dim3(kThreads), 0, s, w->flags.p, n, w->estart.p, w->entries.p, fv, w->localIdx.p, w->colourWord.p, w->colour.p); w->launches++; }
        const int* list = w->dynList.p; int listCount = w->nDyn; int which = 0;
        for (int round = 0, batch = 6;;) {
            if (round > 4096) return fail(AVBD_ERR_CUDA, "graph colouring did not converge");
            CK(cudaMemsetAsync(&w->dCnt->nUncoloured, 0, sizeof(int), s));
            for (int k = 0; k < batch; ++k)
                launch_dep(colour_round, dim3(blocks_for(listCount)), dim3(kThreads), 0, s, list, listCount, w->estart.p, w->entries.p, fv, w->localIdx.p, w->colourWord.p, w->colour.p,
                                                                        w->dCnt, k == batch - 1);
            w->launches += batch; round += batch;
            TRY(read_counters(w));
            int left = w->hCnt->nUncoloured;
            if (left == 0) break;
            if (left * 2 < listCount && listCount > 4096) {
                DevBuf<int>& dst = which ? w->colWorkB : w->colWorkA;
                TRY(dst.ensure((size_t)left, false, s));
                CK(cudaMemsetAsync(&w->dCnt->nUncoloured, 0, sizeof(int), s));
                launch_dep(colour_compact, dim3(blocks_for(listCount)), dim3(kThreads), 0, s, list, listCount, w->colour.p, dst.p, &w->dCnt->nUncoloured);
                w->launches++;
                list = dst.p; listCount = left; which ^= 1;
            }
            batch = 8;          // the work list is short by now: spare rounds are cheaper than another host check
        }
    }
    w->colouredBodies = n;
    if (!keysDone) { launch_dep(colour_keys, dim3(blocks_for(w->nDyn)), dim3(kThreads), 0, s, w->dynList.p, w->nDyn, w->colour.p, w->colKey.p, w->colVal.p); w->launches++; }
    TRY(sort_pairs(w, w->colKey.p, w->colKeySorted.p, w->colVal.p, w->colOrder.p, w->nDyn, 7));
    // contact visits in colour order (the sweeps' work list): visitStart[k] belongs to colOrder[k]
    TRY(w->visits.ensure((size_t)std::max(1, 2 * w->nContacts), false, s));
    TRY(w->freeList.ensure((size_t)w->nDyn, false, s)); TRY(w->linkedList.ensure((size_t)w->nDyn, false, s));
    TRY(w->colVisit.ensure(64, false, s));
    launch_dep(colour_bounds_visit_count, dim3(blocks_for(w->nDyn)), dim3(kThreads), 0, s, w->colKeySorted.p, w->nDyn, w->colRange.p, w->colOrder.p, w->adjRange.p, w->bList.p,
               ms.hdr, w->visitCount.p, fv, w->freeList.p, w->linkedList.p, w->dCnt, (const int*)w->bodyContacts.p);
    TRY(exclusive_scan(w, w->visitCount.p, w->visitStart.p, w->nDyn + 1));
    launch_dep(visit_fill, dim3(blocks_for(w->nDyn)), dim3(kThreads), 0, s, w->colOrder.p, w->nDyn, w->adjRange.p, w->bList.p, ms.hdr, ms.cstart, w->visitStart.p, w->aux.p, w->colour.p, w->visits.p,
               (const int2*)w->colRange.p, (const Counters*)w->dCnt, w->colVisit.p);
    w->launches += 2;
    }
    w->visitGeomStale = true;
    // ONE host round trip for everything the launches of the sweeps need: colour ranges, their visit ranges, counters
    CK(cudaMemcpyAsync(w->hColRange, w->colRange.p, sizeof(int2) * 64, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(w->hColVisit, w->colVisit.p, sizeof(int2) * 64, cudaMemcpyDeviceToHost, s));
    TRY(read_counters(w));
    if (w->hCnt->nUncoloured != 0) return fail(AVBD_ERR_CUDA, "graph colouring did not converge");
    if (w->hCnt->overflow & 4) return fail(AVBD_ERR_CAPACITY, "more than 64 colours needed");
    w->nColours = w->hCnt->nColours;
    if (getenv("AVBD_DEBUG_COLOUR")) fprintf(stderr, "[avbd] colouring: %d colours, %d rounds, %d bodies kept\n", w->hCnt->nColours, w->hCnt->colourRounds, w->hCnt->colourKept);
    w->nFree = w->hCnt->nFree; w->nLinkedFree = w->hCnt->nLinkedFree;
    w->maxColourCount = 0;
    for (int c = 0; c < w->nColours; ++c) w->maxColourCount = std::max(w->maxColourCount, w->hColRange[c].y - w->hColRange[c].x);
    // the sweeps' body-aligned warp ranges: built at once for a large world, on first use for a world the cluster loop takes
    w->sweepRangesValid = false;
    if (w->nDyn > w->persistentMaxBodies || w->fview().nJoints + w->fview().nSprings > 0 || w->profiling) TRY(build_sweep_ranges(w));
    w->graphValid = true;
    CK(cudaGetLastError());
    return 0;
}

// biasDual >= 0: the sweep also applies the previous iteration's pending dual pass (deferred dual, avbd_solve.cu); it is that
// pass's clamp(1 - alpha, 0, 1) (manifold.cpp:179), which lies in [0, 1] for EVERY alpha a caller may set, so a negative value
// can only mean "nothing pending".
inline float dual_bias(float alpha) { float b = 1.0f - alpha; b = b > 0.0f ? b : 0.0f; return b < 1.0f ? b : 1.0f; }
int run_primal(avbd_world* w, float alpha, float* dxDev, float biasDual = -1.0f) {
    if (!w->graphValid) TRY(run_colour(w));
    cudaStream_t s = w->stream;
    ManifoldSet ms = w->mset(w->cur);
    ForceView fv = w->fview();
    bool staticsJustWritten = false;         // a static input of the sweeps was written by a launch with no host wait since
    if (!w->sweepRangesValid) { TRY(build_sweep_ranges(w)); staticsJustWritten = true; }
    if (w->visitGeomStale && w->nContacts > 0 && w->nDyn > 0) {
        size_t cap = w->visits.cap;
        TRY(w->vgA.ensure(cap, false, s)); TRY(w->vgB.ensure(cap, false, s)); TRY(w->vgN.ensure(cap, false, s));
        launch_dep(visit_geometry, dim3(blocks_for(2ll * w->nContacts)), dim3(kThreads), 0, s, w->visits.p, w->visitStart.p + w->nDyn, ms, w->vgeom());
        w->launches++;
        staticsJustWritten = true;
    }
    w->visitGeomStale = false;
    int freeLeft = w->nFree;                 // ride on the first sweep launch of the pass
    for (int c = 0; c < w->nColours; ++c) {
        int count = w->hColRange[c].y - w->hColRange[c].x;
        if (count <= 0) continue;
        if (w->nLinkedFree > 0) { launch_primal_free(s, w->bview(), fv, w->linkedList.p, w->nLinkedFree, w->colour.p, c, w->prm, dxDev, w->dDiag.p); w->launches++; }
        if (w->sweepWarps[c] > 0 || freeLeft > 0) {
            launch_primal_sweep(s, w->bview(), w->visits.p, w->vgeom(), ms, fv, w->sweepRange.p + w->sweepOff[c], w->sweepWarps[c], w->prm, alpha, biasDual, dxDev, w->dDiag.p,
                                w->freeList.p, freeLeft, staticsJustWritten);
            w->launches++;
            freeLeft = 0; staticsJustWritten = false;
        }
    }
    if (freeLeft > 0) {                      // no colour at all (nothing but free bodies cannot happen: every dynamic body has a colour), kept for safety
        launch_primal_sweep(s, w->bview(), w->visits.p, w->vgeom(), ms, fv, w->sweepRange.p, 0, w->prm, alpha, biasDual, dxDev, w->dDiag.p, w->freeList.p, freeLeft, staticsJustWritten);
        w->launches++;
    }
    CK(cudaGetLastError());
    return 0;
}

// contacts = false: only the user forces' dual (the manifold rows' pass was deferred into the next sweep).
// unvisitedReps / onlyUnvisited: see launch_dual.
int run_dual(avbd_world* w, float alpha, bool lastOfStep = false, bool contacts = true, int unvisitedReps = 0, bool onlyUnvisited = false) {
    cudaStream_t s = w->stream;
    if (w->nContacts > 0 && contacts) {
        launch_dual(s, w->bview(), w->mset(w->cur), w->nContacts, w->prm, alpha, unvisitedReps, onlyUnvisited, lastOfStep ? w->dDiag.p : nullptr);
        w->launches++;
        if (lastOfStep) w->contactDiagDone = true;
    }
    ForceView fv = w->fview();
    if (fv.nJoints + fv.nSprings > 0) {
        launch_dual_user_forces(s, w->bview(), fv, w->prm);
        w->launches++;
    }
    CK(cudaGetLastError());
    return 0;
}

int run_velocity(avbd_world* w) {
    cudaStream_t s = w->stream;
    if (w->n == 0) return 0;
    launch_dep(velocity_bodies, dim3(blocks_for(w->n)), dim3(kThreads), 0, s, w->bview(), w->prm, w->dDiag.p);
    w->launches++;
    if (w->nContacts > 0 && !w->contactDiagDone) {     // not already reduced by the step's last dual pass
        launch_dep(diagnostics_contacts, dim3(blocks_for(w->nContacts)), dim3(kThreads), 0, s, w->bview(), w->mset(w->cur), w->nContacts, w->dDiag.p);
        w->launches++;
    }
    w->contactDiagDone = false;
    CK(cudaMemcpyAsync(w->hDiag, w->dDiag.p, sizeof(Diag) * w->nWorlds, cudaMemcpyDeviceToHost, s));
    CK(cudaGetLastError());
    return 0;
}

int step_once(avbd_world* w) {
    cudaStream_t s = w->stream;
    w->profCur = nullptr;
    if (w->profiling) {
        if (w->profUsed == (int)w->profSteps.size()) {
            avbd_world::ProfStep ps{};
            for (auto& e : ps.ev) CK(cudaEventCreate(&e));
            w->profSteps.push_back(ps);
        }
        w->profCur = &w->profSteps[w->profUsed++];
    }
    TRY(run_collide(w));
    TRY(run_predict(w));
    stage_event(w, 3);
    bool rebuiltGraph = !w->graphValid;
    if (!w->graphValid) TRY(run_colour(w)); else w->graphReuses++;
    stage_event(w, 4);
    int total = w->prm.iterations + (w->prm.postStabilize ? 1 : 0);
    bool prof = w->profiling;
    ForceView fvAll = w->fview();
    const bool noUserForces = fvAll.nJoints + fvAll.nSprings == 0;
    bool persistent = !prof && w->nColours > 0 && w->nDyn <= w->persistentMaxBodies && noUserForces && (w->loopMode == 0 || w->loopMode == 3);
    bool gridLoop = !persistent && !prof && w->nColours > 0 && noUserForces && (w->loopMode == 2 || w->loopMode == 4) && w->prm.iterations > 0;
    if (gridLoop) {
        // the iteration loop in one cooperative launch of the sweep kernel's warp pipelines (grid barrier between colour phases)
        if (!w->sweepRangesValid) TRY(build_sweep_ranges(w));
        if (w->visitGeomStale && w->nContacts > 0 && w->nDyn > 0) {
            size_t cap = w->visits.cap;
            TRY(w->vgA.ensure(cap, false, s)); TRY(w->vgB.ensure(cap, false, s)); TRY(w->vgN.ensure(cap, false, s));
            launch_dep(visit_geometry, dim3(blocks_for(2ll * w->nContacts)), dim3(kThreads), 0, s, w->visits.p, w->visitStart.p + w->nDyn, w->mset(w->cur), w->vgeom());
            w->launches++;
        }
        w->visitGeomStale = false;
        gridLoop = w->loopMode == 4
            ? launch_solve_loop_warps(s, w->bview(), w->visits.p, w->vgeom(), w->mset(w->cur), fvAll, w->sweepRange.p, w->nColours, w->sweepWarps, w->sweepOff,
                                      w->prm, w->dDiag.p, w->freeList.p, w->nFree)
            : launch_solve_loop_grid(s, w->bview(), w->visits.p, w->vgeom(), w->mset(w->cur), fvAll, w->sweepRange.p, w->nColours, w->sweepWarps, w->sweepOff,
                                     w->prm, w->dDiag.p, w->freeList.p, w->nFree);
        if (gridLoop) {
            w->launches++;
            // what the loop could not apply: the last iteration's dual pass (+ contact diagnostics), or with postStabilize only the
            // contacts no dynamic body visits
            if (!w->prm.postStabilize) { TRY(run_dual(w, w->prm.alpha, true, true, w->prm.iterations, false)); }
            else if (w->anyUnvisited) { TRY(run_dual(w, 1.0f, false, true, w->prm.iterations, true)); }
        }
    }
    if (persistent) {
        // small world: the whole iteration loop in one cluster launch (cluster barriers instead of kernel boundaries)
        bool fuseDiag = !w->prm.postStabilize && w->prm.iterations > 0;
        persistent = launch_solve_loop(s, w->bview(), w->visitStart.p, w->visits.p, w->mset(w->cur), fvAll, w->colOrder.p, w->colRange.p,
                                       w->nColours, w->maxColourCount, w->nContacts, w->prm, w->dDiag.p, fuseDiag, w->anyUnvisited);
        if (persistent) { w->launches++; w->contactDiagDone = fuseDiag; } else cudaGetLastError();
    }
    // Profiling: the iteration loop is timed as a whole (stage events 4 -> 5) and only the stand-alone dual launches are bracketed by
    // their own events — an event between every pair of sweeps would serialise launches that otherwise overlap (programmatic
    // dependent launch) and inflate the very thing it measures.  ms_primal = loop - stand-alone duals.
    if (prof) while ((int)w->profCur->dual.size() < 2 * total + 2) { cudaEvent_t e; CK(cudaEventCreate(&e)); w->profCur->dual.push_back(e); }
    // Deferred dual: the manifold rows' dual pass of iteration k rides on sweep k+1 (each contact's first visit applies it);
    // only the pass after the LAST sweep runs as a kernel.  AVBD_SEPARATE_DUAL=1 keeps one dual launch per iteration.
    const char* sepEnv = getenv("AVBD_SEPARATE_DUAL");
    const bool separateDual = sepEnv && atoi(sepEnv) != 0;
    int duals = 0, deferred = 0, timedDuals = 0;
    float pendingBias = -1.0f;
    const auto hostT0 = std::chrono::steady_clock::now();
    for (int it = 0; it < total && !persistent && !gridLoop; ++it) {
        float a = w->prm.postStabilize ? (it < w->prm.iterations ? 1.0f : 0.0f) : w->prm.alpha;   // solver.cpp:340-342
        TRY(run_primal(w, a, nullptr, pendingBias));
        if (pendingBias >= 0.0f) ++deferred;
        pendingBias = -1.0f;
        if (it < w->prm.iterations) {
            bool lastSweep = it == total - 1;                 // nothing moves after this pass: it also reduces the contact diagnostics
            bool standalone = separateDual || lastSweep;
            if (prof && standalone) cudaEventRecord(w->profCur->dual[2 * timedDuals], s);
            if (separateDual) { TRY(run_dual(w, a, lastSweep)); ++duals; }
            else if (lastSweep) { TRY(run_dual(w, a, true, true, w->prm.iterations, false)); ++duals; }
            else { TRY(run_dual(w, a, false, false)); pendingBias = dual_bias(a); }
            if (prof && standalone) { cudaEventRecord(w->profCur->dual[2 * timedDuals + 1], s); ++timedDuals; }
        } else if (!separateDual && w->anyUnvisited && w->prm.iterations > 0) {
            TRY(run_dual(w, 1.0f, false, true, w->prm.iterations, true));    // postStabilize: contacts between static bodies only
        }
    }
    w->hostLoopSec += std::chrono::duration<double>(std::chrono::steady_clock::now() - hostT0).count(); w->hostLoopSteps++;
    stage_event(w, 5);
    TRY(run_velocity(w));
    stage_event(w, 6);
    if (prof) {          // sizes of this step (all known on the host); times are resolved at avbd_get_profile
        avbd_world::ProfStep& ps = *w->profCur;
        ps.timedDuals = timedDuals; ps.total = total; ps.colours = w->nColours; ps.duals = duals; ps.deferred = deferred; ps.rebuilt = rebuiltGraph ? 1 : 0;
        ps.n = w->n; ps.nDyn = w->nDyn; ps.pairs = w->nPairs; ps.cand = w->nCand; ps.nM = w->nM; ps.nMPrev = w->nMPrev; ps.contacts = w->nContacts;
        ps.visits = w->nColours > 0 ? w->hColVisit[w->nColours - 1].y : 0;
        w->profCur = nullptr;
    }
    return 0;
}

void fill_diag(const Diag& d, avbd_diagnostics* o) {
    o->maxPenetration = d.maxPenetration; o->maxConstraintViolation = d.maxViolation; o->maxLinearSpeed = d.maxLinearSpeed;
    o->maxAngularSpeed = d.maxAngularSpeed; o->maxNormalImpulse = d.maxNormalImpulse; o->activeContacts = d.activeContacts;
    o->activeManifolds = d.activeManifolds; o->dynamicBodies = d.dynamicBodies; o->nanEvents = d.nanEvents;
}

int use_device(int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) return fail(AVBD_ERR_NO_DEVICE, "no CUDA device available (this library has no CPU path)");
    if (device < 0 || device >= count) return fail(AVBD_ERR_ARG, "bad device index");
    CK(cudaSetDevice(device));
    return 0;
}

} // namespace

// ============================================================================= C ABI
extern "C" {

const char* avbd_last_error(void) { return g_err.c_str(); }

void* avbd_host_alloc(long long bytes) {
    void* p = nullptr;
    if (bytes <= 0 || cudaMallocHost(&p, (size_t)bytes) != cudaSuccess) { cudaGetLastError(); fail(AVBD_ERR_CUDA, "pinned host allocation failed"); return nullptr; }
    return p;
}
void avbd_host_free(void* p) { if (p) cudaFreeHost(p); }
int avbd_host_register(void* p, long long bytes) {
    if (!p || bytes <= 0) return fail(AVBD_ERR_ARG, "bad host range");
    if (cudaHostRegister(p, (size_t)bytes, cudaHostRegisterDefault) != cudaSuccess) { cudaGetLastError(); return fail(AVBD_ERR_CUDA, "cudaHostRegister failed"); }
    return 0;
}
int avbd_host_unregister(void* p) {
    if (!p) return 0;
    if (cudaHostUnregister(p) != cudaSuccess) { cudaGetLastError(); return fail(AVBD_ERR_CUDA, "cudaHostUnregister failed"); }
    return 0;
}

int avbd_device_count(void) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) { cudaGetLastError(); return 0; }
    return count;
}

avbd_world* avbd_world_create(int device) {
    if (use_device(device) < 0) return nullptr;
    avbd_world* w = new avbd_world();
    w->device = device;
    if (cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc(&w->dCnt, kPickSlotOffset + 8) != cudaSuccess || cudaMallocHost(&w->hCnt, sizeof(Counters)) != cudaSuccess) {
        fail(AVBD_ERR_CUDA, "world allocation failed");
        delete w;
        return nullptr;
    }
    for (auto& e : w->ev) cudaEventCreate(&e);
    if (const char* e = std::getenv("AVBD_PERSISTENT_MAX_BODIES")) w->persistentMaxBodies = std::atoi(e);
    if (const char* e = std::getenv("AVBD_FORCE_REGRAPH")) w->forceRegraph = std::atoi(e) != 0;
    if (const char* e = std::getenv("AVBD_BROADPHASE")) { w->bodySweep = std::strcmp(e, "cell") != 0; w->cellRaw = std::strcmp(e, "body") != 0 && w->bodySweep; }
    if (const char* e = std::getenv("AVBD_LOOP")) w->loopMode = !std::strcmp(e, "launch") ? 1 : (!std::strcmp(e, "grid") ? 2 : (!std::strcmp(e, "cluster") ? 3 : (!std::strcmp(e, "warps") ? 4 : 0)));
    if (const char* e = std::getenv("AVBD_ITERATED_COLOUR")) w->freshColour = std::atoi(e) == 0;
    if (const char* e = std::getenv("AVBD_KEEP_COLOUR")) w->keepColour = std::atoi(e) != 0;
    std::memset(w->hCnt, 0, sizeof(Counters));
    avbd_default_params(w);
    return w;
}

void avbd_world_destroy(avbd_world* w) {
    if (!w) return;
    cudaSetDevice(w->device);
    cudaStreamSynchronize(w->stream);
    if (std::getenv("AVBD_DEBUG") && w->hostLoopSteps > 0)
        std::fprintf(stderr, "avbd-demo3d_b200: host time issuing the iteration loop: %.1f us per step over %lld steps\n", 1e6 * w->hostLoopSec / w->hostLoopSteps, w->hostLoopSteps);
    w->pose.release(); w->aux.release(); w->vel.release(); w->init.release(); w->prevLin.release(); w->size.release();
    w->flags.release(); w->worldId.release(); w->localIdx.release(); w->dynList.release(); w->colWorkA.release(); w->colWorkB.release(); w->colourWord.release();
    w->cellKey.release(); w->cellKeySorted.release(); w->cellVal.release(); w->cellValSorted.release(); w->cellRange.release();
    w->sortedCell.release(); w->sortedPos.release(); w->sortedFrame.release(); w->bodyFrame.release(); w->largeList.release(); w->worldLargeStart.release();
    w->pairs.release(); w->cand.release(); w->candSorted.release();
    for (auto& b : w->mb) { b.key.release(); b.hdr.release(); b.cstart.release(); b.cM.release(); b.cA.release(); b.cB.release(); b.cN.release(); b.lp.release(); }
    w->adjRange.release(); w->bKey.release(); w->bKeySorted.release(); w->bVal.release(); w->bList.release();
    w->colour.release(); w->colKey.release(); w->colKeySorted.release(); w->colVal.release(); w->colOrder.release(); w->colRange.release();
    w->joints.release(); w->springs.release(); w->fadjStart.release(); w->fadj.release(); w->excl.release();
    w->dDiag.release(); w->dx.release(); w->temp.release(); w->stateDev.release(); w->bodyContacts.release(); w->deg.release(); w->estart.release(); w->colCursor.release(); w->entries.release(); w->sweepRange.release(); w->colVisit.release(); w->freeList.release(); w->linkedList.release(); w->vgA.release(); w->vgB.release(); w->vgN.release();
    w->mcount.release(); w->buildTiles.release(); w->visitCount.release(); w->visitStart.release(); w->visits.release();
    for (auto& ps : w->profSteps) { for (auto& e : ps.ev) cudaEventDestroy(e); for (auto& e : ps.dual) cudaEventDestroy(e); }
    for (cudaEvent_t e : w->chunkEvents) cudaEventDestroy(e);
    if (w->dCnt) cudaFree(w->dCnt);
    if (w->hCnt) cudaFreeHost(w->hCnt);
    if (w->hDiag) cudaFreeHost(w->hDiag);
    for (auto& e : w->ev) if (e) cudaEventDestroy(e);
    cudaStreamDestroy(w->stream);
    delete w;
}

int avbd_clear(avbd_world* w) {
    if (!w) return fail(AVBD_ERR_ARG, "null world");
    CK(cudaSetDevice(w->device));
    CK(cudaStreamSynchronize(w->stream));
    w->n = 0; w->nDyn = 0; w->nWorlds = 1; w->hb.clear(); w->nM = 0; w->nCand = 0; w->nPairs = 0; w->nContacts = 0;
    w->hJoints.clear(); w->hSprings.clear(); w->hForces.clear(); w->uploadedJoints = 0; w->uploadedSprings = 0; w->nExcl = 0;
    w->topoDirty = true; w->forcesDirty = true; w->graphValid = false; w->nColours = 0;
    w->colouredBodies = -1; w->coloursRestored = false; w->savedColours.clear();
    if (w->hDiag) std::memset(w->hDiag, 0, sizeof(Diag) * w->hDiagCap);
    return 0;
}

int avbd_set_params(avbd_world* w, float dt, const float* g, int iterations, float alpha, float beta, float gamma, int ps) {
    if (!w || !g) return fail(AVBD_ERR_ARG, "null argument");
    w->prm.dt = dt; w->prm.gx = g[0]; w->prm.gy = g[1]; w->prm.gz = g[2]; w->prm.iterations = iterations;
    w->prm.alpha = alpha; w->prm.beta = beta; w->prm.gamma = gamma; w->prm.postStabilize = ps ? 1 : 0;
    return 0;
}

int avbd_default_params(avbd_world* w) {       // solver.cpp:240-253
    const float g[3] = {0.0f, -10.0f, 0.0f};
    return avbd_set_params(w, 1.0f / 60.0f, g, 10, 0.95f, 100000.0f, 0.99f, 0);
}

int avbd_add_bodies(avbd_world* w, int count, const float* size3, const float* density, const float* friction, const float* pos3,
                    const float* quat4, const float* lin3, const float* ang3, const int* world_ids) {
    if (!w || count < 0 || (count > 0 && (!size3 || !density || !friction || !pos3 || !quat4 || !lin3 || !ang3)))
        return fail(AVBD_ERR_ARG, "bad argument to avbd_add_bodies");
    CK(cudaSetDevice(w->device));
    if (count == 0) return w->n;
    cudaStream_t s = w->stream;
    int first = w->n, total = first + count;
    std::vector<BodyPose> pose(count); std::vector<BodyAux> aux(count); std::vector<BodyVel> vel(count); std::vector<BodyInit> init(count);
    std::vector<float4> prev(count), size(count);
    if (world_ids) {       // validate everything before touching any state: a mid-loop failure would leave hb[] misaligned with the device
        int prevWorld = w->hb.empty() ? -1 : w->hb.back().world;
        for (int i = 0; i < count; ++i) {
            if (world_ids[i] < 0 || world_ids[i] < prevWorld || world_ids[i] > prevWorld + 1)
                return fail(AVBD_ERR_ARG, "world ids must start at 0 and be non-decreasing without gaps");
            prevWorld = world_ids[i];
        }
    }
    w->hb.reserve(total);
    for (int i = 0; i < count; ++i) {
        // Rigid::Rigid, rigid.cpp:12-41
        float sx = size3[3 * i], sy = size3[3 * i + 1], sz = size3[3 * i + 2];
        float mass = sx * sy * sz * density[i];
        float invMass = (mass > 0.0f) ? 1.0f / mass : 0.0f;
        float radius = sqrtf(sx * sx + sy * sy + sz * sz) * 0.5f;
        float ixx = 0.0f, iyy = 0.0f, izz = 0.0f;
        if (invMass > 0.0f) {
            ixx = (1.0f / 12.0f) * mass * (sy * sy + sz * sz);
            iyy = (1.0f / 12.0f) * mass * (sx * sx + sz * sz);
            izz = (1.0f / 12.0f) * mass * (sx * sx + sy * sy);
        }
        pose[i].pos = make_float4(pos3[3 * i], pos3[3 * i + 1], pos3[3 * i + 2], invMass);
        pose[i].rot = make_float4(quat4[4 * i], quat4[4 * i + 1], quat4[4 * i + 2], quat4[4 * i + 3]);
        vel[i].lin = make_float4(lin3[3 * i], lin3[3 * i + 1], lin3[3 * i + 2], 0.0f);
        vel[i].ang = make_float4(ang3[3 * i], ang3[3 * i + 1], ang3[3 * i + 2], 0.0f);
        prev[i] = vel[i].lin;
        init[i].pos0 = make_float4(0.f, 0.f, 0.f, 0.f); init[i].rot0 = make_float4(0.f, 0.f, 0.f, 1.f);
        aux[i].posI = make_float4(0.f, 0.f, 0.f, 0.f); aux[i].rotI = make_float4(0.f, 0.f, 0.f, 1.f);
        aux[i].mass = make_float4(mass, invMass, friction[i], radius);
        aux[i].inert = make_float4(ixx, iyy, izz, 0.0f);
        size[i] = make_float4(sx, sy, sz, friction[i]);
        HostBody hb; hb.radius = radius; hb.invMass = invMass; hb.world = world_ids ? world_ids[i] : 0;
        int prevWorld = w->hb.empty() ? -1 : w->hb.back().world;
        hb.local = (hb.world == prevWorld) ? w->hb.back().local + 1 : 0;
        w->hb.push_back(hb);
    }
    TRY(w->pose.ensure(total, true, s)); TRY(w->aux.ensure(total, true, s)); TRY(w->vel.ensure(total, true, s)); TRY(w->init.ensure(total, true, s));
    TRY(w->prevLin.ensure(total, true, s)); TRY(w->size.ensure(total, true, s));
    CK(cudaMemcpyAsync(w->pose.p + first, pose.data(), count * sizeof(BodyPose), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(w->aux.p + first, aux.data(), count * sizeof(BodyAux), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(w->vel.p + first, vel.data(), count * sizeof(BodyVel), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(w->init.p + first, init.data(), count * sizeof(BodyInit), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(w->prevLin.p + first, prev.data(), count * sizeof(float4), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(w->size.p + first, size.data(), count * sizeof(float4), cudaMemcpyHostToDevice, s));
    CK(cudaStreamSynchronize(s));
    w->n = total;
    w->topoDirty = true;
    return first;
}

int avbd_num_bodies(const avbd_world* w) { return w ? w->n : 0; }

namespace {
// current pose of one body (joint / spring constructors capture it, joint.cpp:17-19, spring.cpp:19-23)
int fetch_pose(avbd_world* w, int i, BodyPose& out) {
    CK(cudaMemcpyAsync(&out, w->pose.p + i, sizeof(BodyPose), cudaMemcpyDeviceToHost, w->stream));
    CK(cudaStreamSynchronize(w->stream));
    return 0;
}
}

namespace {
int push_joint(avbd_world* w, int a, int b, float4 rA, float4 rB, float4 rel0, float linK, float angK) {
    JointRec j{};
    j.a = a; j.b = b; j.rA = rA; j.rB = rB; j.rel0 = rel0;
    for (int r = 0; r < 6; ++r) { j.lambda[r] = 0.0f; j.penalty[r] = kPenaltyMin; j.stiffness[r] = r < 3 ? linK : angK; j.motor[r] = 0.0f; }
    w->hJoints.push_back(j);
    w->hForces.push_back(HostForce{0, a, b});
    w->forcesDirty = true;
    return (int)w->hJoints.size() - 1;
}
}

int avbd_add_joint(avbd_world* w, int a, int b, const float* anchorA, const float* anchorB, float linK, float angK) {
    if (!w || b < 0 || b >= w->n || a >= w->n || a == b || !anchorA) return fail(AVBD_ERR_ARG, "bad joint");
    CK(cudaSetDevice(w->device));
    BodyPose pb; TRY(fetch_pose(w, b, pb));
    Q4 qB = quat(pb.rot);
    if (a >= 0) {
        if (!anchorB) return fail(AVBD_ERR_ARG, "bad joint");
        BodyPose pa; TRY(fetch_pose(w, a, pa));
        return push_joint(w, a, b, make_float4(anchorA[0], anchorA[1], anchorA[2], 0.f), make_float4(anchorB[0], anchorB[1], anchorB[2], 0.f),
                          f4(qmul(qconj(quat(pa.rot)), qB)), linK, angK);                    // joint.cpp:19
    }
    V3 wa = mk3(anchorA[0], anchorA[1], anchorA[2]);
    M3 R = qmat(qB);                                                            // joint.cpp:47: transpose(R) * (anchor - pos)
    V3 d = wa - xyz(pb.pos);
    M3 Rt = m3(mk3(R.c[0].x, R.c[1].x, R.c[2].x), mk3(R.c[0].y, R.c[1].y, R.c[2].y), mk3(R.c[0].z, R.c[1].z, R.c[2].z));
    return push_joint(w, a, b, f4(wa, 0.f), f4(mv(Rt, d), 0.f), f4(qB), linK, angK);
}

int avbd_add_joint_raw(avbd_world* w, int a, int b, const float* rA3, const float* rB3, const float* rel0, float linK, float angK) {
    if (!w || b < 0 || b >= w->n || a >= w->n || a < -1 || a == b || !rA3 || !rB3 || !rel0) return fail(AVBD_ERR_ARG, "bad joint");
    return push_joint(w, a, b, make_float4(rA3[0], rA3[1], rA3[2], 0.f), make_float4(rB3[0], rB3[1], rB3[2], 0.f),
                      make_float4(rel0[0], rel0[1], rel0[2], rel0[3]), linK, angK);
}

int avbd_add_spring(avbd_world* w, int a, int b, const float* anchorA, const float* anchorB, float k, float rest) {
    if (!w || a < 0 || b < 0 || a >= w->n || b >= w->n || a == b || !anchorA || !anchorB) return fail(AVBD_ERR_ARG, "bad spring");
    CK(cudaSetDevice(w->device));
    SpringRec sp{};
    sp.a = a; sp.b = b; sp.k = k; sp.rest = rest;
    sp.rA = make_float4(anchorA[0], anchorA[1], anchorA[2], 0.f); sp.rB = make_float4(anchorB[0], anchorB[1], anchorB[2], 0.f);
    if (rest < 0) {                                                                  // spring.cpp:19-23
        BodyPose pa, pb; TRY(fetch_pose(w, a, pa)); TRY(fetch_pose(w, b, pb));
        V3 pA = xyz(pa.pos) + qrot(quat(pa.rot), xyz(sp.rA)), pB = xyz(pb.pos) + qrot(quat(pb.rot), xyz(sp.rB));
        sp.rest = len(pA - pB);
    }
    sp.lambda = 0.0f; sp.penalty = kPenaltyMin; sp.motor = 0.0f;
    w->hSprings.push_back(sp);
    w->hForces.push_back(HostForce{1, a, b});
    w->forcesDirty = true;
    return (int)w->hSprings.size() - 1;
}

int avbd_add_ignore(avbd_world* w, int a, int b) {
    if (!w || a < 0 || b < 0 || a >= w->n || b >= w->n || a == b) return fail(AVBD_ERR_ARG, "bad ignore pair");
    w->hForces.push_back(HostForce{2, a, b});
    w->forcesDirty = true;
    return 0;
}

int avbd_set_force_rows(avbd_world* w, int kind, int index, const float* lambda, const float* penalty, const float* motor, const float* stiffness) {
    if (!w || kind < 0 || kind > 1 || index < 0 || index >= (int)(kind == 0 ? w->hJoints.size() : w->hSprings.size()))
        return fail(AVBD_ERR_ARG, "bad force row edit");
    CK(cudaSetDevice(w->device));
    TRY(prepare(w));                         // the record is on the device from here on (lambda / penalty live there)
    cudaStream_t s = w->stream;
    if (kind == 0) {
        JointRec j;
        CK(cudaMemcpyAsync(&j, w->joints.p + index, sizeof(j), cudaMemcpyDeviceToHost, s)); CK(cudaStreamSynchronize(s));
        for (int r = 0; r < 6; ++r) {
            if (lambda) j.lambda[r] = lambda[r];
            if (penalty) j.penalty[r] = penalty[r];
            if (motor) j.motor[r] = motor[r];
            if (stiffness) j.stiffness[r] = stiffness[r];
        }
        w->hJoints[index] = j;
        CK(cudaMemcpyAsync(w->joints.p + index, &w->hJoints[index], sizeof(j), cudaMemcpyHostToDevice, s)); CK(cudaStreamSynchronize(s));
    } else {
        SpringRec sp;
        CK(cudaMemcpyAsync(&sp, w->springs.p + index, sizeof(sp), cudaMemcpyDeviceToHost, s)); CK(cudaStreamSynchronize(s));
        if (lambda) sp.lambda = lambda[0];
        if (penalty) sp.penalty = penalty[0];
        if (motor) sp.motor = motor[0];
        if (stiffness) sp.k = stiffness[0];
        w->hSprings[index] = sp;
        CK(cudaMemcpyAsync(w->springs.p + index, &w->hSprings[index], sizeof(sp), cudaMemcpyHostToDevice, s)); CK(cudaStreamSynchronize(s));
    }
    return 0;
}

int avbd_get_force_rows(avbd_world* w, int kind, int index, float* lambda, float* penalty, float* motor, float* stiffness) {
    if (!w || kind < 0 || kind > 1 || index < 0 || index >= (int)(kind == 0 ? w->hJoints.size() : w->hSprings.size()))
        return fail(AVBD_ERR_ARG, "bad force row query");
    CK(cudaSetDevice(w->device));
    TRY(prepare(w));
    cudaStream_t s = w->stream;
    if (kind == 0) {
        JointRec j;
        CK(cudaMemcpyAsync(&j, w->joints.p + index, sizeof(j), cudaMemcpyDeviceToHost, s)); CK(cudaStreamSynchronize(s));
        for (int r = 0; r < 6; ++r) {
            if (lambda) lambda[r] = j.lambda[r];
            if (penalty) penalty[r] = j.penalty[r];
            if (motor) motor[r] = j.motor[r];
            if (stiffness) stiffness[r] = j.stiffness[r];
        }
    } else {
        SpringRec sp;
        CK(cudaMemcpyAsync(&sp, w->springs.p + index, sizeof(sp), cudaMemcpyDeviceToHost, s)); CK(cudaStreamSynchronize(s));
        if (lambda) lambda[0] = sp.lambda;
        if (penalty) penalty[0] = sp.penalty;
        if (motor) motor[0] = sp.motor;
        if (stiffness) stiffness[0] = sp.k;
    }
    return 0;
}

int avbd_num_joints(const avbd_world* w) { return w ? (int)w->hJoints.size() : 0; }
int avbd_num_springs(const avbd_world* w) { return w ? (int)w->hSprings.size() : 0; }

int avbd_download_user_rows(avbd_world* w, float* joints12, float* springs2) {
    if (!w) return fail(AVBD_ERR_ARG, "null world");
    CK(cudaSetDevice(w->device));
    TRY(prepare(w));
    cudaStream_t s = w->stream;
    size_t nj = w->hJoints.size(), ns = w->hSprings.size();
    if (nj && joints12) CK(cudaMemcpyAsync(w->hJoints.data(), w->joints.p, nj * sizeof(JointRec), cudaMemcpyDeviceToHost, s));
    if (ns && springs2) CK(cudaMemcpyAsync(w->hSprings.data(), w->springs.p, ns * sizeof(SpringRec), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (joints12) for (size_t k = 0; k < nj; ++k) for (int r = 0; r < 6; ++r) { joints12[12 * k + r] = w->hJoints[k].lambda[r]; joints12[12 * k + 6 + r] = w->hJoints[k].penalty[r]; }
    if (springs2) for (size_t k = 0; k < ns; ++k) { springs2[2 * k] = w->hSprings[k].lambda; springs2[2 * k + 1] = w->hSprings[k].penalty; }
    return 0;
}

int avbd_upload_manifolds(avbd_world* w, int count, const int* ints, const int* feats, const int* stick, const float* flts) {
    if (!w || count < 0 || (count > 0 && (!ints || !feats || !stick || !flts))) return fail(AVBD_ERR_ARG, "bad argument to avbd_upload_manifolds");
    CK(cudaSetDevice(w->device));
    TRY(prepare(w));
    cudaStream_t s = w->stream;
    std::vector<int> idx(count);
    size_t nC = 0;
    for (int m = 0; m < count; ++m) {
        int a = ints[3 * m], b = ints[3 * m + 1], nc = ints[3 * m + 2];
        if (a <= b || b < 0 || a >= w->n || nc < 1 || nc > 4) return fail(AVBD_ERR_ARG, "bad manifold (need n > idxA > idxB >= 0, 1..4 contacts)");
        idx[m] = m; nC += (size_t)nc;
    }
    auto key_of = [&](int m) { return ((unsigned long long)(unsigned)ints[3 * m] << w->keyShift) | (unsigned)ints[3 * m + 1]; };
    std::sort(idx.begin(), idx.end(), [&](int x, int y) { return key_of(x) < key_of(y); });
    for (int m = 1; m < count; ++m) if (key_of(idx[m]) == key_of(idx[m - 1])) return fail(AVBD_ERR_ARG, "duplicate manifold pair");
    std::vector<unsigned long long> key(count); std::vector<int4> hdr(count); std::vector<int> cstart(count + 1), cM(nC);
    std::vector<float4> cA(nC), cB(nC), cN(nC); std::vector<ContactLP> lp(nC);
    size_t ci = 0;
    for (int q = 0; q < count; ++q) {
        int m = idx[q];
        const float* f = flts + 81 * (size_t)m;
        int nc = ints[3 * m + 2];
        int muBits; std::memcpy(&muBits, &f[0], 4);
        key[q] = key_of(m); hdr[q] = make_int4(ints[3 * m], ints[3 * m + 1], nc, muBits); cstart[q] = (int)ci;
        for (int c = 0; c < nc; ++c, ++ci) {
            const float* g = f + 1 + 14 * c;
            cA[ci] = make_float4(g[0], g[1], g[2], g[10]); cB[ci] = make_float4(g[3], g[4], g[5], g[11]); cN[ci] = make_float4(g[6], g[7], g[8], g[12]);
            float featBits; std::memcpy(&featBits, &feats[4 * m + c], 4);
            lp[ci].l = make_float4(f[57 + 3 * c], f[58 + 3 * c], f[59 + 3 * c], stick[4 * m + c] ? 1.0f : 0.0f);
            lp[ci].p = make_float4(f[69 + 3 * c], f[70 + 3 * c], f[71 + 3 * c], featBits);
            cM[ci] = q;
        }
    }
    cstart[count] = (int)ci;
    w->nM = count; w->nContacts = (int)nC;
    if (count > 0) {
        TRY(w->ensure_manifolds(w->cur, count));
        ManifoldSet ms = w->mset(w->cur);
        CK(cudaMemcpyAsync(ms.key, key.data(), count * sizeof(unsigned long long), cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(ms.hdr, hdr.data(), count * sizeof(int4), cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(ms.cstart, cstart.data(), (count + 1) * sizeof(int), cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(ms.cM, cM.data(), nC * sizeof(int), cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(ms.cA, cA.data(), nC * sizeof(float4), cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(ms.cB, cB.data(), nC * sizeof(float4), cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(ms.cN, cN.data(), nC * sizeof(float4), cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(ms.lp, lp.data(), nC * sizeof(ContactLP), cudaMemcpyHostToDevice, s));
        CK(cudaStreamSynchronize(s));
    }
    w->graphValid = false; w->visitGeomStale = true; w->contactDiagDone = false;
    return 0;
}

// ---- snapshot / restore -------------------------------------------------------------------------------------------------------
namespace {
struct SnapHeader {
    char magic[8];
    int version, n, nM, nContacts, nJoints, nSprings, nForces, keyShift;
    SolveParams prm;
    long long bytes;
};
constexpr char kSnapMagic[8] = {'A', 'V', 'B', 'D', 'S', 'N', 'P', '3'};
size_t snap_bytes(const avbd_world* w) {
    size_t n = w->n, m = w->nM, c = w->nContacts;
    return sizeof(SnapHeader) + n * (sizeof(HostBody) + sizeof(BodyPose) + sizeof(BodyAux) + sizeof(BodyVel) + sizeof(BodyInit) + 2 * sizeof(float4))
         + w->hJoints.size() * sizeof(JointRec) + w->hSprings.size() * sizeof(SpringRec) + w->hForces.size() * sizeof(HostForce)
         + m * (sizeof(unsigned long long) + sizeof(int4)) + (m + 1) * sizeof(int) + c * (sizeof(int) + 3 * sizeof(float4) + sizeof(ContactLP))
         + sizeof(int) + (w->colouredBodies == w->n ? n * sizeof(int) : 0);        // the colouring the next one is ranked by
}
}

long long avbd_snapshot_bytes(avbd_world* w) { return w ? (long long)snap_bytes(w) : 0; }

int avbd_snapshot(avbd_world* w, void* buf, long long cap) {
    if (!w || !buf) return fail(AVBD_ERR_ARG, "null argument");
    CK(cudaSetDevice(w->device));
    TRY(prepare(w));                                   // user-force records are on the device
    size_t need = snap_bytes(w);
    if ((long long)need > cap) return fail(AVBD_ERR_CAPACITY, "snapshot buffer too small (see avbd_snapshot_bytes)");
    cudaStream_t s = w->stream;
    char* o = static_cast<char*>(buf);
    SnapHeader h{};
    std::memcpy(h.magic, kSnapMagic, 8);
    h.version = 3; h.n = w->n; h.nM = w->nM; h.nContacts = w->nContacts; h.nJoints = (int)w->hJoints.size(); h.nSprings = (int)w->hSprings.size();
    h.nForces = (int)w->hForces.size(); h.keyShift = w->keyShift; h.prm = w->prm; h.bytes = (long long)need;
    std::memcpy(o, &h, sizeof(h)); o += sizeof(h);
    auto put_host = [&](const void* src, size_t bytes) { if (bytes) std::memcpy(o, src, bytes); o += bytes; };
    auto put_dev = [&](const void* src, size_t bytes) -> int {
        if (bytes) CK(cudaMemcpyAsync(o, src, bytes, cudaMemcpyDeviceToHost, s));
        o += bytes; return 0;
    };
    size_t n = w->n, m = w->nM, c = w->nContacts;
    put_host(w->hb.data(), n * sizeof(HostBody));
    TRY(put_dev(w->pose.p, n * sizeof(BodyPose))); TRY(put_dev(w->aux.p, n * sizeof(BodyAux))); TRY(put_dev(w->vel.p, n * sizeof(BodyVel)));
    TRY(put_dev(w->init.p, n * sizeof(BodyInit))); TRY(put_dev(w->prevLin.p, n * sizeof(float4))); TRY(put_dev(w->size.p, n * sizeof(float4)));
    TRY(put_dev(w->joints.p, w->hJoints.size() * sizeof(JointRec))); TRY(put_dev(w->springs.p, w->hSprings.size() * sizeof(SpringRec)));
    put_host(w->hForces.data(), w->hForces.size() * sizeof(HostForce));
    ManifoldSet ms = w->mset(w->cur);
    if (m) {
        TRY(put_dev(ms.key, m * sizeof(unsigned long long))); TRY(put_dev(ms.hdr, m * sizeof(int4))); TRY(put_dev(ms.cstart, (m + 1) * sizeof(int)));
        TRY(put_dev(ms.cM, c * sizeof(int))); TRY(put_dev(ms.cA, c * sizeof(float4))); TRY(put_dev(ms.cB, c * sizeof(float4)));
        TRY(put_dev(ms.cN, c * sizeof(float4))); TRY(put_dev(ms.lp, c * sizeof(ContactLP)));
    } else {
        int zero = 0; put_host(&zero, sizeof(int));
    }
    const int hasColours = (w->colouredBodies == w->n && n > 0) ? 1 : 0;
    put_host(&hasColours, sizeof(int));
    if (hasColours) TRY(put_dev(w->colour.p, n * sizeof(int)));
    CK(cudaStreamSynchronize(s));
    return 0;
}

int avbd_restore(avbd_world* w, const void* buf, long long bytes) {
    if (!w || !buf || bytes < (long long)sizeof(SnapHeader)) return fail(AVBD_ERR_ARG, "bad snapshot");
    CK(cudaSetDevice(w->device));
    SnapHeader h;
    std::memcpy(&h, buf, sizeof(h));
    if (std::memcmp(h.magic, kSnapMagic, 8) != 0 || h.version != 3 || h.bytes > bytes || h.n < 0 || h.nM < 0 || h.nContacts < 0)
        return fail(AVBD_ERR_ARG, "not a snapshot of this library build");
    TRY(avbd_clear(w));
    cudaStream_t s = w->stream;
    const char* o = static_cast<const char*>(buf) + sizeof(h);
    size_t n = h.n, m = h.nM, c = h.nContacts;
    w->prm = h.prm;
    w->hb.assign(reinterpret_cast<const HostBody*>(o), reinterpret_cast<const HostBody*>(o) + n); o += n * sizeof(HostBody);
    auto get_dev = [&](void* dst, size_t nbytes) -> int {
        if (nbytes) CK(cudaMemcpyAsync(dst, o, nbytes, cudaMemcpyHostToDevice, s));
        o += nbytes; return 0;
    };
    TRY(w->pose.ensure(n, false, s)); TRY(w->aux.ensure(n, false, s)); TRY(w->vel.ensure(n, false, s)); TRY(w->init.ensure(n, false, s));
    TRY(w->prevLin.ensure(n, false, s)); TRY(w->size.ensure(n, false, s));
    TRY(get_dev(w->pose.p, n * sizeof(BodyPose))); TRY(get_dev(w->aux.p, n * sizeof(BodyAux))); TRY(get_dev(w->vel.p, n * sizeof(BodyVel)));
    TRY(get_dev(w->init.p, n * sizeof(BodyInit))); TRY(get_dev(w->prevLin.p, n * sizeof(float4))); TRY(get_dev(w->size.p, n * sizeof(float4)));
    w->n = (int)n;
    w->hJoints.assign(reinterpret_cast<const JointRec*>(o), reinterpret_cast<const JointRec*>(o) + h.nJoints); o += (size_t)h.nJoints * sizeof(JointRec);
    w->hSprings.assign(reinterpret_cast<const SpringRec*>(o), reinterpret_cast<const SpringRec*>(o) + h.nSprings); o += (size_t)h.nSprings * sizeof(SpringRec);
    w->hForces.assign(reinterpret_cast<const HostForce*>(o), reinterpret_cast<const HostForce*>(o) + h.nForces); o += (size_t)h.nForces * sizeof(HostForce);
    w->uploadedJoints = 0; w->uploadedSprings = 0;            // prepare() uploads the records (with their saved lambda / penalty)
    w->topoDirty = true; w->forcesDirty = true;
    w->keyShift = h.keyShift;                                 // the saved keys are packed with it; prepare() re-derives the same value from n
    w->cur = 0; w->nM = (int)m; w->nContacts = (int)c;
    if (m) {
        TRY(w->ensure_manifolds(0, m));
        ManifoldSet ms = w->mset(0);
        TRY(get_dev(ms.key, m * sizeof(unsigned long long))); TRY(get_dev(ms.hdr, m * sizeof(int4))); TRY(get_dev(ms.cstart, (m + 1) * sizeof(int)));
        TRY(get_dev(ms.cM, c * sizeof(int))); TRY(get_dev(ms.cA, c * sizeof(float4))); TRY(get_dev(ms.cB, c * sizeof(float4)));
        TRY(get_dev(ms.cN, c * sizeof(float4))); TRY(get_dev(ms.lp, c * sizeof(ContactLP)));
    } else {
        o += sizeof(int);
    }
    int hasColours = 0;
    std::memcpy(&hasColours, o, sizeof(int)); o += sizeof(int);
    w->savedColours.clear(); w->coloursRestored = false;
    if (hasColours) { w->savedColours.assign(reinterpret_cast<const int*>(o), reinterpret_cast<const int*>(o) + n); o += n * sizeof(int); }
    CK(cudaStreamSynchronize(s));
    w->graphValid = false; w->visitGeomStale = true; w->contactDiagDone = false;      // lastPairs (a sizing hint) is kept: buffers are still there
    return 0;
}

int avbd_step(avbd_world* w, int nSteps) {
    if (!w) return fail(AVBD_ERR_ARG, "null world");
    CK(cudaSetDevice(w->device));
    for (int i = 0; i < nSteps; ++i) TRY(step_once(w));
    return 0;
}

int avbd_step_timed(avbd_world* w, int nSteps, float* ms) {
    if (!w || !ms) return fail(AVBD_ERR_ARG, "null argument");
    CK(cudaSetDevice(w->device));
    CK(cudaStreamSynchronize(w->stream));
    CK(cudaEventRecord(w->ev[7], w->stream));
    for (int i = 0; i < nSteps; ++i) TRY(step_once(w));
    CK(cudaEventRecord(w->ev[8], w->stream));
    CK(cudaStreamSynchronize(w->stream));
    CK(cudaEventElapsedTime(ms, w->ev[7], w->ev[8]));
    return 0;
}

int avbd_set_profiling(avbd_world* w, int on) {
    if (!w) return fail(AVBD_ERR_ARG, "null world");
    CK(cudaSetDevice(w->device));
    CK(cudaStreamSynchronize(w->stream));
    w->profiling = on != 0;
    w->prof = avbd_profile{};
    w->profUsed = 0; w->profCur = nullptr;
    return 0;
}

int avbd_get_profile(avbd_world* w, avbd_profile* out) {
    if (!w || !out) return fail(AVBD_ERR_ARG, "null argument");
    CK(cudaSetDevice(w->device));
    CK(cudaStreamSynchronize(w->stream));
    for (int k = 0; k < w->profUsed; ++k) {              // the profiled steps since the last call
        avbd_world::ProfStep& ps = w->profSteps[k];
        float t[6] = {0, 0, 0, 0, 0, 0}, tot = 0, dualMs = 0;
        for (int q = 0; q < 6; ++q) cudaEventElapsedTime(&t[q], ps.ev[q], ps.ev[q + 1]);
        cudaEventElapsedTime(&tot, ps.ev[0], ps.ev[6]);
        for (int q = 0; q < ps.timedDuals; ++q) { float b = 0; cudaEventElapsedTime(&b, ps.dual[2 * q], ps.dual[2 * q + 1]); dualMs += b; }
        cudaGetLastError();
        avbd_profile& p = w->prof;
        p.ms_broadphase += t[0]; p.ms_narrowphase += t[1]; p.ms_predict += t[2]; p.ms_graph += t[3]; p.ms_solve += t[4]; p.ms_velocity += t[5]; p.ms_step += tot;
        p.ms_primal += t[4] - dualMs; p.ms_dual += dualMs;
        p.steps += 1;
        p.primal_sweeps += ps.total; p.primal_launches += (long long)ps.total * ps.colours;
        p.primal_bodies += (long long)ps.total * ps.nDyn; p.primal_visits += (long long)ps.total * ps.visits;
        p.dual_launches += ps.duals; p.dual_contacts += (long long)ps.duals * ps.contacts;
        p.deferred_dual_contacts += (long long)ps.deferred * ps.contacts;
        p.bodies += ps.n; p.pairs += ps.pairs; p.candidates += ps.cand; p.manifolds += ps.nM; p.manifolds_prev += ps.nMPrev;
        p.contacts += ps.contacts; p.visits += ps.visits; p.graph_builds += ps.rebuilt;
    }
    w->profUsed = 0;
    *out = w->prof;
    out->kernel_launches = w->launches; out->library_launches = w->libLaunches;
    return 0;
}

int avbd_sync(avbd_world* w) {
    if (!w) return fail(AVBD_ERR_ARG, "null world");
    CK(cudaSetDevice(w->device));
    CK(cudaStreamSynchronize(w->stream));
    return 0;
}

int avbd_download_state(avbd_world* w, float* out) {
    if (!w || (!out && w->n)) return fail(AVBD_ERR_ARG, "null argument");
    CK(cudaSetDevice(w->device));
    int n = w->n; if (!n) return 0;
    TRY(w->stateDev.ensure((size_t)n * 13, false, w->stream));
    launch_dep(pack_state, dim3(blocks_for(n)), dim3(kThreads), 0, w->stream, w->bview(), w->stateDev.p);
    w->launches++;
    CK(cudaMemcpyAsync(out, w->stateDev.p, (size_t)n * 13 * sizeof(float), cudaMemcpyDeviceToHost, w->stream));
    CK(cudaStreamSynchronize(w->stream));
    return 0;
}

int avbd_download_state_chunked(avbd_world* w, float* out, int chunkBodies, avbd_chunk_fn landed, void* user) {
    if (!w || (!out && w->n) || chunkBodies <= 0) return fail(AVBD_ERR_ARG, "bad argument");
    CK(cudaSetDevice(w->device));
    int n = w->n; if (!n) return 0;
    TRY(w->stateDev.ensure((size_t)n * 13, false, w->stream));
    launch_dep(pack_state, dim3(blocks_for(n)), dim3(kThreads), 0, w->stream, w->bview(), w->stateDev.p);
    w->launches++;
    int chunks = (n + chunkBodies - 1) / chunkBodies;
    if (chunks > 64) { chunkBodies = (n + 63) / 64; chunks = (n + chunkBodies - 1) / chunkBodies; }
    while ((int)w->chunkEvents.size() < chunks) { cudaEvent_t e; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); w->chunkEvents.push_back(e); }
    for (int k = 0; k < chunks; ++k) {
        int first = k * chunkBodies, count = std::min(chunkBodies, n - first);
        CK(cudaMemcpyAsync(out + (size_t)first * 13, w->stateDev.p + (size_t)first * 13, (size_t)count * 13 * sizeof(float), cudaMemcpyDeviceToHost, w->stream));
        CK(cudaEventRecord(w->chunkEvents[k], w->stream));
    }
    for (int k = 0; k < chunks; ++k) {
        int first = k * chunkBodies, count = std::min(chunkBodies, n - first);
        CK(cudaEventSynchronize(w->chunkEvents[k]));
        if (landed) landed(first, count, user);
    }
    return 0;
}

int avbd_upload_state(avbd_world* w, const float* in) {
    if (!w || (!in && w->n)) return fail(AVBD_ERR_ARG, "null argument");
    CK(cudaSetDevice(w->device));
    int n = w->n; if (!n) return 0;
    TRY(w->stateDev.ensure((size_t)n * 13, false, w->stream));
    CK(cudaMemcpyAsync(w->stateDev.p, in, (size_t)n * 13 * sizeof(float), cudaMemcpyHostToDevice, w->stream));
    launch_dep(unpack_state, dim3(blocks_for(n)), dim3(kThreads), 0, w->stream, w->bview(), w->stateDev.p);
    w->launches++;
    CK(cudaStreamSynchronize(w->stream));
    return 0;
}

int avbd_upload_state_range(avbd_world* w, int first, int count, const float* in) {
    if (!w || first < 0 || count < 0 || first + count > w->n || (!in && count)) return fail(AVBD_ERR_ARG, "bad body range");
    CK(cudaSetDevice(w->device));
    if (!count) return 0;
    TRY(w->stateDev.ensure((size_t)w->n * 13, false, w->stream));
    CK(cudaMemcpyAsync(w->stateDev.p + (size_t)first * 13, in, (size_t)count * 13 * sizeof(float), cudaMemcpyHostToDevice, w->stream));
    launch_dep(unpack_state_range, dim3(blocks_for(count)), dim3(kThreads), 0, w->stream, w->bview(), w->stateDev.p, first, count);
    w->launches++;
    CK(cudaStreamSynchronize(w->stream));
    return 0;
}

int avbd_download_prev_linvel(avbd_world* w, float* out) {
    if (!w || (!out && w->n)) return fail(AVBD_ERR_ARG, "null argument");
    CK(cudaSetDevice(w->device));
    int n = w->n; if (!n) return 0;
    std::vector<float4> p(n);
    CK(cudaMemcpyAsync(p.data(), w->prevLin.p, n * sizeof(float4), cudaMemcpyDeviceToHost, w->stream));
    CK(cudaStreamSynchronize(w->stream));
    for (int i = 0; i < n; ++i) { out[3 * i] = p[i].x; out[3 * i + 1] = p[i].y; out[3 * i + 2] = p[i].z; }
    return 0;
}

int avbd_upload_prev_linvel(avbd_world* w, const float* in) {
    if (!w || (!in && w->n)) return fail(AVBD_ERR_ARG, "null argument");
    CK(cudaSetDevice(w->device));
    int n = w->n; if (!n) return 0;
    std::vector<float4> p(n);
    for (int i = 0; i < n; ++i) p[i] = make_float4(in[3 * i], in[3 * i + 1], in[3 * i + 2], 0.f);
    CK(cudaMemcpyAsync(w->prevLin.p, p.data(), n * sizeof(float4), cudaMemcpyHostToDevice, w->stream));
    CK(cudaStreamSynchronize(w->stream));
    return 0;
}

int avbd_download_body_props(avbd_world* w, float* out) {
    if (!w || (!out && w->n)) return fail(AVBD_ERR_ARG, "null argument");
    CK(cudaSetDevice(w->device));
    int n = w->n; if (!n) return 0;
    std::vector<BodyAux> aux(n); std::vector<float4> size(n);
    CK(cudaMemcpyAsync(aux.data(), w->aux.p, n * sizeof(BodyAux), cudaMemcpyDeviceToHost, w->stream));
    CK(cudaMemcpyAsync(size.data(), w->size.p, n * sizeof(float4), cudaMemcpyDeviceToHost, w->stream));
    CK(cudaStreamSynchronize(w->stream));
    for (int i = 0; i < n; ++i) {
        float* o = out + 10 * (size_t)i;
        o[0] = size[i].x; o[1] = size[i].y; o[2] = size[i].z; o[3] = aux[i].mass.x; o[4] = aux[i].mass.y;
        o[5] = aux[i].inert.x; o[6] = aux[i].inert.y; o[7] = aux[i].inert.z; o[8] = aux[i].mass.z; o[9] = aux[i].mass.w;
    }
    return 0;
}

int avbd_num_worlds(const avbd_world* w) { return w ? w->nWorlds : 0; }

int avbd_get_world_diagnostics(avbd_world* w, avbd_diagnostics* out, int count) {
    if (!w || !out) return fail(AVBD_ERR_ARG, "null argument");
    CK(cudaSetDevice(w->device));
    CK(cudaStreamSynchronize(w->stream));
    for (int k = 0; k < count; ++k) {
        if (k < w->nWorlds && w->hDiag) fill_diag(w->hDiag[k], out + k);
        else std::memset(out + k, 0, sizeof(avbd_diagnostics));
    }
    return 0;
}

int avbd_get_diagnostics(avbd_world* w, avbd_diagnostics* out) {
    if (!w || !out) return fail(AVBD_ERR_ARG, "null argument");
    CK(cudaSetDevice(w->device));
    CK(cudaStreamSynchronize(w->stream));
    Diag acc{};
    for (int k = 0; k < w->nWorlds && w->hDiag; ++k) {
        const Diag& d = w->hDiag[k];
        acc.maxPenetration = std::max(acc.maxPenetration, d.maxPenetration); acc.maxViolation = std::max(acc.maxViolation, d.maxViolation);
        acc.maxLinearSpeed = std::max(acc.maxLinearSpeed, d.maxLinearSpeed); acc.maxAngularSpeed = std::max(acc.maxAngularSpeed, d.maxAngularSpeed);
        acc.maxNormalImpulse = std::max(acc.maxNormalImpulse, d.maxNormalImpulse);
        acc.activeContacts += d.activeContacts; acc.activeManifolds += d.activeManifolds; acc.dynamicBodies += d.dynamicBodies; acc.nanEvents += d.nanEvents;
    }
    fill_diag(acc, out);
    return 0;
}

int avbd_world_diagnostics_device_ptr(avbd_world* w, void** ptr, int* count) {
    if (!w || !ptr || !count) return fail(AVBD_ERR_ARG, "null argument");
    *ptr = w->dDiag.p; *count = w->nWorlds;
    return 0;
}

int avbd_get_step_stats(avbd_world* w, avbd_step_stats* out) {
    if (!w || !out) return fail(AVBD_ERR_ARG, "null argument");
    CK(cudaSetDevice(w->device));
    CK(cudaStreamSynchronize(w->stream));
    avbd_step_stats st{};
    if (w->timed) {
        float t[6] = {0, 0, 0, 0, 0, 0};
        for (int k = 0; k < 6; ++k) cudaEventElapsedTime(&t[k], w->ev[k], w->ev[k + 1]);
        st.ms_broadphase = t[0]; st.ms_narrowphase = t[1]; st.ms_predict = t[2]; st.ms_graph = t[3]; st.ms_primal = t[4]; st.ms_velocity = t[5];
        cudaEventElapsedTime(&st.ms_total, w->ev[0], w->ev[6]);
        cudaGetLastError();
    }
    w->timed = true;   // stage events are recorded from the next step on
    st.bodies = w->n; st.dynamicBodies = w->nDyn; st.pairs = w->nPairs; st.candidates = w->nCand; st.manifolds = w->nM;
    avbd_diagnostics d; avbd_get_diagnostics(w, &d); st.contacts = d.activeContacts;
    st.colours = w->nColours; st.iterations = w->prm.iterations; st.kernelLaunches = w->launches;
    { long long v = 0; for (int k = 0; k < w->nWorlds && w->hDiag; ++k) v += w->hDiag[k].contactVisits; st.contactVisits = (int)v; }
    *out = st;
    return 0;
}

int avbd_num_manifolds(avbd_world* w) { return w ? w->nM : 0; }

int avbd_download_manifolds(avbd_world* w, int* ints, int* feats, int* stick, float* flts) {
    if (!w) return fail(AVBD_ERR_ARG, "null world");
    CK(cudaSetDevice(w->device));
    int nM = w->nM; if (!nM) return 0;
    size_t nC = (size_t)std::max(1, w->nContacts);
    std::vector<int4> hdr(nM); std::vector<int> cstart(nM + 1); std::vector<float4> cA(nC), cB(nC), cN(nC); std::vector<ContactLP> lp(nC);
    ManifoldSet ms = w->mset(w->cur);
    cudaStream_t s = w->stream;
    CK(cudaMemcpyAsync(hdr.data(), ms.hdr, nM * sizeof(int4), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(cstart.data(), ms.cstart, (nM + 1) * sizeof(int), cudaMemcpyDeviceToHost, s));
    if (w->nContacts > 0) {
        CK(cudaMemcpyAsync(cA.data(), ms.cA, nC * sizeof(float4), cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(cB.data(), ms.cB, nC * sizeof(float4), cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(cN.data(), ms.cN, nC * sizeof(float4), cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(lp.data(), ms.lp, nC * sizeof(ContactLP), cudaMemcpyDeviceToHost, s));
    }
    CK(cudaStreamSynchronize(s));
    int live = 0;
    for (int m = 0; m < nM; ++m) {
        int nc = hdr[m].z;
        if (nc <= 0) continue;              // SAT passed but no contact survived: the reference deletes these (solver.cpp:274-279)
        ints[3 * live] = hdr[m].x; ints[3 * live + 1] = hdr[m].y; ints[3 * live + 2] = nc;
        float* f = flts + 81 * (size_t)live;
        float mu; std::memcpy(&mu, &hdr[m].w, 4);
        f[0] = mu;
        for (int c = 0; c < 4; ++c) {
            bool on = c < nc; int ci = on ? cstart[m] + c : 0;
            int feat; std::memcpy(&feat, &lp[ci].p.w, 4);
            feats[4 * live + c] = on ? feat : 0;
            stick[4 * live + c] = on ? (lp[ci].l.w != 0.0f ? 1 : 0) : 0;
            float* o = f + 1 + 14 * c;
            const float v[14] = {cA[ci].x, cA[ci].y, cA[ci].z, cB[ci].x, cB[ci].y, cB[ci].z, cN[ci].x, cN[ci].y, cN[ci].z,
                                 0.0f /* penetration is draw-only state, not kept on the device */, cA[ci].w, cB[ci].w, cN[ci].w, 0.0f};
            for (int k = 0; k < 14; ++k) o[k] = on ? v[k] : 0.0f;
            for (int k = 0; k < 3; ++k) {
                f[57 + 3 * c + k] = on ? (&lp[ci].l.x)[k] : 0.0f;
                f[69 + 3 * c + k] = on ? (&lp[ci].p.x)[k] : 0.0f;
            }
        }
        ++live;
    }
    return live;
}

int avbd_stage_broadphase(avbd_world* w) {
    if (!w) return fail(AVBD_ERR_ARG, "null world");
    CK(cudaSetDevice(w->device));
    TRY(prepare(w));
    TRY(run_broadphase(w, false));
    return w->nCand;
}

int avbd_download_pairs(avbd_world* w, int* pairs, int cap) {
    if (!w) return fail(AVBD_ERR_ARG, "null world");
    CK(cudaSetDevice(w->device));
    int nc = w->nCand;
    if (nc == 0) return 0;
    std::vector<unsigned long long> k(nc);
    CK(cudaMemcpyAsync(k.data(), w->candSorted.p, nc * sizeof(unsigned long long), cudaMemcpyDeviceToHost, w->stream));
    CK(cudaStreamSynchronize(w->stream));
    int out = 0;
    const unsigned long long pairMask = (1ull << (2 * w->keyShift)) - 1ull;      // after a step the SAT codes ride above the pair bits
    for (int i = 0; i < nc; ++i) {
        k[i] &= pairMask;
        if (i > 0 && k[i] == k[i - 1]) continue;
        if (out < cap) { pairs[2 * out] = (int)(k[i] >> w->keyShift); pairs[2 * out + 1] = (int)(k[i] & ((1ull << w->keyShift) - 1ull)); }
        ++out;
    }
    return out;
}

int avbd_stage_collide(avbd_world* w) { if (!w) return fail(AVBD_ERR_ARG, "null world"); CK(cudaSetDevice(w->device)); return run_collide(w); }
int avbd_stage_predict(avbd_world* w) { if (!w) return fail(AVBD_ERR_ARG, "null world"); CK(cudaSetDevice(w->device)); return run_predict(w); }
int avbd_stage_colour(avbd_world* w) { if (!w) return fail(AVBD_ERR_ARG, "null world"); CK(cudaSetDevice(w->device)); return w->graphValid ? 0 : run_colour(w); }

int avbd_download_colours(avbd_world* w, int* colour_of, int* num_colours) {
    if (!w) return fail(AVBD_ERR_ARG, "null world");
    CK(cudaSetDevice(w->device));
    if (num_colours) *num_colours = w->nColours;
    if (colour_of && w->n) {
        CK(cudaMemcpyAsync(colour_of, w->colour.p, w->n * sizeof(int), cudaMemcpyDeviceToHost, w->stream));
        CK(cudaStreamSynchronize(w->stream));
    }
    return 0;
}

int avbd_stage_primal(avbd_world* w, float alpha, float* dx_out) {
    if (!w) return fail(AVBD_ERR_ARG, "null world");
    CK(cudaSetDevice(w->device));
    float* dxDev = nullptr;
    if (dx_out && w->n) {
        TRY(w->dx.ensure((size_t)w->n * 6, false, w->stream));
        CK(cudaMemsetAsync(w->dx.p, 0, (size_t)w->n * 6 * sizeof(float), w->stream));
        dxDev = w->dx.p;
    }
    TRY(run_primal(w, alpha, dxDev));
    if (dxDev) {
        CK(cudaMemcpyAsync(dx_out, dxDev, (size_t)w->n * 6 * sizeof(float), cudaMemcpyDeviceToHost, w->stream));
        CK(cudaStreamSynchronize(w->stream));
    }
    return 0;
}

int avbd_stage_dual(avbd_world* w, float alpha) { if (!w) return fail(AVBD_ERR_ARG, "null world"); CK(cudaSetDevice(w->device)); return run_dual(w, alpha); }
int avbd_stage_velocity(avbd_world* w) { if (!w) return fail(AVBD_ERR_ARG, "null world"); CK(cudaSetDevice(w->device)); return run_velocity(w); }

int avbd_collide_pairs(int device, int n, const float* a10, const float* b10, int* counts, int* feats4, float* geom36) {
    TRY(use_device(device));
    if (n <= 0) return 0;
    float *da = nullptr, *db = nullptr, *dg = nullptr; int *dc = nullptr, *df = nullptr;
    CK(cudaMalloc(&da, n * 10 * sizeof(float))); CK(cudaMalloc(&db, n * 10 * sizeof(float))); CK(cudaMalloc(&dg, (size_t)n * 36 * sizeof(float)));
    CK(cudaMalloc(&dc, n * sizeof(int))); CK(cudaMalloc(&df, n * 4 * sizeof(int)));
    CK(cudaMemcpy(da, a10, n * 10 * sizeof(float), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, b10, n * 10 * sizeof(float), cudaMemcpyHostToDevice));
    cudaFuncSetAttribute(np_collide_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * kPolyFloatsPerThread * (int)sizeof(float));
    np_collide_batch<<<blocks_for(n, 128), 128, 128 * kPolyFloatsPerThread * sizeof(float)>>>(da, db, n, dc, df, dg);
    CK(cudaGetLastError());
    CK(cudaMemcpy(counts, dc, n * sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(feats4, df, n * 4 * sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(geom36, dg, (size_t)n * 36 * sizeof(float), cudaMemcpyDeviceToHost));
    cudaFree(da); cudaFree(db); cudaFree(dg); cudaFree(dc); cudaFree(df);
    return 0;
}

int avbd_solve6x6(int device, int n, const float* lhs36, const float* rhs6, float* out6) {
    TRY(use_device(device));
    if (n <= 0) return 0;
    float *dl = nullptr, *dr = nullptr, *dout = nullptr;
    CK(cudaMalloc(&dl, (size_t)n * 36 * sizeof(float))); CK(cudaMalloc(&dr, n * 6 * sizeof(float))); CK(cudaMalloc(&dout, n * 6 * sizeof(float)));
    CK(cudaMemcpy(dl, lhs36, (size_t)n * 36 * sizeof(float), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dr, rhs6, n * 6 * sizeof(float), cudaMemcpyHostToDevice));
    launch_solve6_batch(nullptr, dl, dr, n, dout);
    CK(cudaGetLastError());
    CK(cudaMemcpy(out6, dout, n * 6 * sizeof(float), cudaMemcpyDeviceToHost));
    cudaFree(dl); cudaFree(dr); cudaFree(dout);
    return 0;
}

int avbd_pick(avbd_world* w, const float* origin3, const float* dir3, float* local3) {
    if (!w || !origin3 || !dir3 || !local3) return fail(AVBD_ERR_ARG, "null argument");
    CK(cudaSetDevice(w->device));
    if (w->n == 0) return -1;
    V3 origin = mk3(origin3[0], origin3[1], origin3[2]), dir = mk3(dir3[0], dir3[1], dir3[2]);
    float l2 = len2(dir);
    if (l2 < 1.0e-6f) return -1;                                    // solver.cpp:153-157
    V3 rayDir = dir / sqrtf(l2);
    unsigned long long* dBest = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(w->dCnt) + kPickSlotOffset);   // scratch slot after the counters
    unsigned long long init = ~0ull, best = ~0ull;
    CK(cudaMemcpyAsync(dBest, &init, sizeof(init), cudaMemcpyHostToDevice, w->stream));
    launch_dep(pick_bodies, dim3(blocks_for(w->n)), dim3(kThreads), 0, w->stream, w->bview(), origin, rayDir, dBest);
    w->launches++;
    CK(cudaMemcpyAsync(&best, dBest, sizeof(best), cudaMemcpyDeviceToHost, w->stream));
    CK(cudaStreamSynchronize(w->stream));
    if (best == ~0ull) return -1;
    int i = (int)(0xFFFFFFFFu - (unsigned)(best & 0xFFFFFFFFull));
    BodyPose p; float4 sz;
    CK(cudaMemcpyAsync(&p, w->pose.p + i, sizeof(p), cudaMemcpyDeviceToHost, w->stream));
    CK(cudaMemcpyAsync(&sz, w->size.p + i, sizeof(sz), cudaMemcpyDeviceToHost, w->stream));
    CK(cudaStreamSynchronize(w->stream));
    float t; V3 local;
    ray_obb(origin, rayDir, xyz(p.pos), quat(p.rot), xyz(sz), t, local);   // same function the kernel ran, for the local hit point
    local3[0] = local.x; local3[1] = local.y; local3[2] = local.z;
    return i;
}

} // extern "C"
