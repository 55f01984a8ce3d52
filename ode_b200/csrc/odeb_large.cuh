// odeb_large.cuh -- the large-world path (ODEB_MODE_CANONICAL, one world of 10^3..10^5+ bodies).
//
// The batched path gives each world to one warp lane pair; a single big world needs parallelism INSIDE the
// world instead.  Every serial list walk of the reference is replaced by a sort / scan / union-find that yields
// the same sets, and the order-sensitive SOR sweep keeps its sequential semantics through a colouring of the row groups:
//
//   broadphase   k_bp_keys + radix sort on (float)aabb.min[0] + k_bp_sweep / k_bp_big + radix sort of the pair keys
//                -> the same pair set as k_pair_pass (collision_sapspace.cpp:521-582 BoxPruning is itself a sort + sweep;
//                   the hash space reports the same set as an exhaustive AABB test), in (geomA<geomB) order
//   contacts     scan of the per-pair contact counts -> creation-order numbering (ode.cpp:1192-1200)
//   auto-disable k_lw_autodisable, thread per body (util.cpp:427-561)
//   islands      lock-free union-find over all joints (util.cpp:724-860 finds the same components with a DFS);
//                island number = rank of the component's highest enabled body, descending (world->firstbody order)
//   order        bodies (island, descending index), joints (island, ascending id) by radix sort; row offsets by scan
//   solve        once per step the row GROUPS (rows of the contacts of one geom pair / of one joint: same two bodies) are coloured so
//                that groups of one colour touch disjoint bodies (Jones-Plassmann rounds with first fit, priorities =
//                odeb_canon_key(seed, island, 0, group): deterministic, the oracle runs the same rounds); the canonical sweep order
//                is colour-major (the colours in a per-phase order), and a launch relaxes all groups of a colour side by side, one
//                lane per group with the two bodies' accumulators in registers: bit-identical to the sequential sweep in that order
//                (quickstep.cpp:2917-3033), no tickets, no fences, no polling -- ~9 colours on a brick wall, i.e. ~9 short launches
//                per sweep.  The records are re-laid once per step into TILES (32 groups of a colour, lane-interleaved, in sweep
//                order), so a sweep is one coalesced stream over HBM (k_lwt_sweep: register double buffer; k_lwt_sweep_tma:
//                cp.async.bulk + mbarrier ring).
//                (round 1 walked a hash-sorted order with per-body tickets: 0.58 ms per sweep on the 100k-box wall, bound by the
//                L2 round trips of the release / acquire hand-over along the dependency paths; a thread per group reading its
//                records where k_rows left them: 0.5 ms per sweep, bound by the number of outstanding scattered requests per SM)
//   control      k_lw_body_check + k_lw_island_ctl after every sweep (quickstep.cpp:1823-1856, :3253-3285), per island
#ifndef ODEB_LARGE_CUH
#define ODEB_LARGE_CUH
#include <cub/cub.cuh>

typedef unsigned long long u64;
#define LW_NOKEY 0xFFFFFFFFFFFFFFFFull

enum { LWC_NBIG = 0, LWC_NPAIRS = 1, LWC_NCONTACTS = 2, LWC_NORDERED = 3, LWC_NJORD = 4, LWC_MROWS = 5, LWC_NISLANDS = 6,
       LWC_NACTIVE = 7, LWC_NGROUPS = 8, LWC_UNCOLORED = 9, LWC_NCOLORS = 10, LWC_NTILES = 11, LWC_SEED = 12, LWC_GBAR = 13, LWC_ITER = 14, LWC_EXTRA = 15, LWC_TERM = 16, LWC_PBAR = 20 /* one barrier counter per phase launch, 12 of them */, LWC_COUNT = 32 };

struct LargePtrs {
    int *counters;                               // [LWC_COUNT]
    u64 *draws;                                  // [4]: dRand draws the reference's reorders would have consumed, sweeps, row-sweeps
    // broadphase
    unsigned *bp_key, *bp_key_s; int *bp_idx, *bp_idx_s, *bp_big;   // [NG]
    Real4 *bp_yz;                                // [NG] (min, max) on axes 1 and 2 in sorted order
    u64 *pair_key, *pair_key_s;                  // [MP]
    int *pc_base;                                // [MP]
    // islands
    int *parent, *maxen, *head_scan, *deg;       // [NB]
    u64 *bkey, *bkey_s; int *bval;               // [NB]
    int *isl_nb, *isl_m, *isl_bstart, *isl_rstart, *isl_done, *isl_viol;   // [NB + 1]
    u64 *jkey, *jkey_s; int *jmv, *jmv_s, *jrow; // [NJT]
    // solver
    int *row_island;                             // [MR]
    int *row_group;                              // [MR] first row of the row's group
    int *gsize, *heads;                          // [MR] rows of the group (at its first row); first rows of all groups, compacted
    int *ginc_ofs, *ginc_cur; int2 *ginc;        // [NB + 2], [NB + 2], [2 MR]: body (order position) -> (group, its priority) of the groups acting on it
    unsigned *gkey; int *gcolor, *gwin;          // [MR] at the group's first row: priority of the phase, colour (-1 none yet, -2 island finished), winner flag
    unsigned *skey, *skey_s;                     // [MR] sort keys of the groups: (colour, rows descending)
    int *clist, *ccount, *cofs, *tstart;         // [MR] groups by (colour, rows descending); [ODEB_CANON_COLOURS] groups per colour; [ODEB_CANON_COLOURS + 1] first list position / first tile of every colour
    int *theight, *tbase, *tgroup;               // [MR/32 + colours + 4] rows of a tile, its first tile row; [MR + 32 colours + 64] first row of the group of every tile lane (-1: none)
    int4 *tginfo;                                // beside tgroup: (rows, accumulator slot of body 1, of body 2, island)
    unsigned char *trec;                         // [trcap * LWT_ROW_BYTES]: per tile row the lane-interleaved compact records (LWT_QUADS x 32 quads), then the 32 lambdas
    Real *pinvm;                                 // [NB + 1] inverse mass by body order position
    int trcap;                                   // tile rows the two buffers hold
    void *tmp; size_t tmp_bytes;                 // cub scratch
};

__device__ __forceinline__ unsigned lw_float_key(Real v)
{   // order-preserving map float -> uint; +0.0f folds -0 onto +0 so that key order == float order incl. ties
    float f = (float)v + 0.0f;
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ bool lw_is_big(const Real *a) { return a[0] == -R_INF || a[1] == R_INF; }

// ------------------------------------------------------------------------------------------------ broadphase
__global__ void k_bp_keys(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D, const __grid_constant__ LargePtrs L)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= P.NG) return;
    const Real *a = D.aabb + 6 * (size_t)g;
    unsigned key;
    if (lw_is_big(a)) { L.bp_big[atomicAdd(&L.counters[LWC_NBIG], 1)] = g; key = 0xFFFFFFFFu; }
    else key = lw_float_key(a[0]);
    L.bp_key[g] = key; L.bp_idx[g] = g;
}

__device__ __forceinline__ void lw_emit_pair(const DevParams &P, const LargePtrs &L, int lo, int hi)
{
    int k = atomicAdd(&L.counters[LWC_NPAIRS], 1);
    if (k < P.MP) L.pair_key[k] = ((u64)(unsigned)lo << 32) | (unsigned)hi;
}

// sweep along axis 0 in float-sorted order: every pair whose axis-0 intervals overlap is visited from its earlier member.  A thread keeps
// the pairs it finds (a box of a wall meets ~4 later neighbours) and the warp reserves room for all of them with ONE atomic at the end:
// one atomicAdd per pair on the single pair counter serialised the kernel (0.68 ms for 393 k pairs).
#define LW_PAIRBUF 8
// the y / z extents in sorted order: the sweep reads its candidates' boxes as one sequential, L1-friendly stream and rejects most of them
// (a wall: ~600 candidates with overlapping axis-0 intervals per box, ~4 hits) before touching anything that is indexed by geom
__global__ void k_bp_gather(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D, const __grid_constant__ LargePtrs L)
{
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= P.NG) return;
    const Real *a = D.aabb + 6 * (size_t)L.bp_idx_s[q];
    const Real4 v = { a[2], a[3], a[4], a[5] };
    L.bp_yz[q] = v;
}
__global__ void k_bp_sweep(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D, const __grid_constant__ LargePtrs L)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
    u64 buf[LW_PAIRBUF]; int nbuf = 0;
    if (p < P.NG && L.bp_key_s[p] != 0xFFFFFFFFu) {
        const int i = L.bp_idx_s[p];
        Real ai[6];
        for (int k = 0; k < 6; k++) ai[k] = D.aabb[6 * (size_t)i + k];
        const unsigned kmax = lw_float_key(ai[1]);
        for (int q = p + 1; q < P.NG; q++) {
            if (L.bp_key_s[q] > kmax) break;
            const Real4 yz = L.bp_yz[q];
            if (ai[3] < yz.x || yz.y < ai[2] || ai[5] < yz.z || yz.w < ai[4]) continue;    // no overlap on axis 1 or 2: no pair in any space flavour
            const int j = L.bp_idx_s[q];
            Real aj[6];
            for (int k = 0; k < 6; k++) aj[k] = D.aabb[6 * (size_t)j + k];
            const bool hit = i < j ? pair_hit(P, D, ai, aj, i, j) : pair_hit(P, D, aj, ai, j, i);
            if (hit) {
                if (nbuf == LW_PAIRBUF) {                            // a crowded neighbourhood: this thread's buffer goes out on its own
                    const int base = atomicAdd(&L.counters[LWC_NPAIRS], nbuf);
                    for (int k = 0; k < nbuf; k++) if (base + k < P.MP) L.pair_key[base + k] = buf[k];
                    nbuf = 0;
                }
                buf[nbuf++] = i < j ? (((u64)(unsigned)i << 32) | (unsigned)j) : (((u64)(unsigned)j << 32) | (unsigned)i);
            }
        }
    }
    int ofs = nbuf;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int u = __shfl_up_sync(0xffffffffu, ofs, d); if (lane >= d) ofs += u; }
    const int total = __shfl_sync(0xffffffffu, ofs, 31);
    int base = 0;
    if (lane == 31 && total > 0) base = atomicAdd(&L.counters[LWC_NPAIRS], total);
    base = __shfl_sync(0xffffffffu, base, 31) + ofs - nbuf;
    for (int k = 0; k < nbuf; k++) if (base + k < P.MP) L.pair_key[base + k] = buf[k];
}

// geoms of unbounded axis-0 extent (planes) against everything
__global__ void k_bp_big(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D, const __grid_constant__ LargePtrs L)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P.NG) return;
    const int nbig = L.counters[LWC_NBIG];
    if (nbig == 0) return;
    Real aj[6];
    for (int k = 0; k < 6; k++) aj[k] = D.aabb[6 * (size_t)j + k];
    const bool jbig = lw_is_big(aj);
    for (int t = 0; t < nbig; t++) {
        const int b = L.bp_big[t];
        if (b == j || (jbig && j < b)) continue;          // big x big pairs once, from the lower index
        Real ab[6];
        for (int k = 0; k < 6; k++) ab[k] = D.aabb[6 * (size_t)b + k];
        if (b < j) { if (pair_hit(P, D, ab, aj, b, j)) lw_emit_pair(P, L, b, j); }
        else if (pair_hit(P, D, aj, ab, j, b)) lw_emit_pair(P, L, j, b);
    }
}

__global__ void k_bp_unpack(int np, const u64 *keys, int2 *pairs, int *npairs)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p == 0) *npairs = np;
    if (p >= np) return;
    u64 k = keys[p];
    pairs[p] = make_int2((int)(k >> 32), (int)(k & 0xFFFFFFFFu));
}

// ------------------------------------------------------------------------------------------------ contacts
__global__ void k_lw_contact_fill(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D, const __grid_constant__ LargePtrs L, int np)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= np) return;
    const int cnt = D.pc_count[p], base = L.pc_base[p];
    if (p == np - 1) {
        int nc = base + cnt;
        if (nc > P.MC) { atomicExch(D.overflow, 2); nc = P.MC; }
        D.ncontacts[0] = nc; L.counters[LWC_NCONTACTS] = nc;
    }
    if (!cnt) return;
    int2 pr = D.pairs[p];
    int b1 = D.gbody[pr.x], b2 = D.gbody[pr.y], rev = 0;
    if (b1 < 0) { b1 = b2; b2 = -1; rev = 1; }                 // dJointAttach ode.cpp:1404-1411
    for (int k = 0; k < cnt; k++) if (base + k < P.MC) D.cinfo[base + k] = make_int4(p * P.maxc + k, b1, b2, rev);
}

__global__ void k_lw_degree(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D, const __grid_constant__ LargePtrs L)
{
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= L.counters[LWC_NCONTACTS]) return;
    int4 v = D.cinfo[c];
    atomicAdd(&L.deg[v.y], 1);
    if (v.z >= 0) atomicAdd(&L.deg[v.z], 1);
}

// dInternalHandleAutoDisabling util.cpp:427-561, one thread per body (bodies are independent of each other here)
__global__ void k_lw_autodisable(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D, const __grid_constant__ LargePtrs L)
{
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= P.NB) return;
    int fl = D.bflags[b];
    if ((fl & (BF_AUTO_DISABLE | BF_DISABLED)) != BF_AUTO_DISABLE) return;
    if (L.deg[b] == 0 && D.sadj_ofs[b + 1] == D.sadj_ofs[b]) return;
    Real4 lv = D.lvel[b], av = D.avel[b];
    Real *buf = D.avg_buf + (size_t)b * 6 * P.adis_samples;
    int cnt = D.avg_counter[b], ready = D.avg_ready[b];
    buf[6 * cnt + 0] = lv.x; buf[6 * cnt + 1] = lv.y; buf[6 * cnt + 2] = lv.z;
    buf[6 * cnt + 3] = av.x; buf[6 * cnt + 4] = av.y; buf[6 * cnt + 5] = av.z;
    cnt++;
    if (cnt >= P.adis_samples) { cnt = 0; ready = 1; }
    D.avg_counter[b] = cnt; D.avg_ready[b] = ready;
    int idle = 0;
    if (ready) {
        idle = 1;
        Real al[3] = { buf[0], buf[1], buf[2] }, aa[3] = { buf[3], buf[4], buf[5] };
        if (P.adis_samples > 1) {
            for (int i = 1; i < P.adis_samples; i++)
                for (int k = 0; k < 3; k++) { al[k] += buf[6 * i + k]; aa[k] += buf[6 * i + 3 + k]; }
            Real r1 = R_(1.0) / (Real)P.adis_samples;
            for (int k = 0; k < 3; k++) { al[k] *= r1; aa[k] *= r1; }
        }
        Real ls = dot3(al, al);
        if (ls > P.adis_lin) idle = 0;
        else { Real as = dot3(aa, aa); if (as > P.adis_ang) idle = 0; }
    }
    int sl = D.adis_steps[b]; Real tl = D.adis_time[b];
    if (idle) { sl--; tl -= P.h; } else { sl = P.adis_steps; tl = P.adis_time; }
    D.adis_steps[b] = sl; D.adis_time[b] = tl;
    if (sl <= 0 && tl <= 0) {
        D.bflags[b] = fl | BF_DISABLED;
        Real4 z = { 0, 0, 0, 0 };
        D.lvel[b] = z; D.avel[b] = z;
    }
}

// ------------------------------------------------------------------------------------------------ islands
__device__ __forceinline__ int uf_find(int *par, int x)
{
    int p = ((volatile int *)par)[x];
    while (p != x) {
        int gp = ((volatile int *)par)[p];
        if (gp != p) par[x] = gp;       // path halving; parents only ever move towards the root
        x = p; p = gp;
    }
    return x;
}
__device__ __forceinline__ void uf_union(int *par, int u, int v)
{
    while (true) {
        u = uf_find(par, u); v = uf_find(par, v);
        if (u == v) return;
        if (u < v) { int t = u; u = v; v = t; }           // larger root goes under the smaller: no cycles
        if (atomicCAS(&par[u], u, v) == u) return;
    }
}
__global__ void k_lw_init_bodies(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D, const __grid_constant__ LargePtrs L)
{
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > P.NB) return;
    L.isl_nb[b] = 0; L.isl_m[b] = 0; L.isl_done[b] = 0; L.isl_viol[b] = 0; L.ginc_cur[b] = 0;
    if (b == P.NB) return;
    L.parent[b] = b; L.maxen[b] = -1;
}
__global__ void k_lw_union(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D, const __grid_constant__ LargePtrs L)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P.NJ + L.counters[LWC_NCONTACTS]) return;
    int b0, b1;
    if (j < P.NJ) { b0 = D.joints[j].b0; b1 = D.joints[j].b1; }
    else {
        int4 v = D.cinfo[j - P.NJ]; b0 = v.y; b1 = v.z;
        if (v.x % P.maxc != 0) return;                            // the first contact of the geom pair has made this union already
    }
    if (b1 >= 0) uf_union(L.parent, b0, b1);
}
__global__ void k_lw_roots(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D, const __grid_constant__ LargePtrs L)
{
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= P.NB) return;
    int r = uf_find(L.parent, b);
    L.bval[b] = r;                                             // root of b (bval doubles as scratch until the body sort)
    if (!(D.bflags[b] & BF_DISABLED)) atomicMax(&L.maxen[r], b);
}
__global__ void k_lw_heads(const __grid_constant__ DevParams P, const __grid_constant__ LargePtrs L)
{
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= P.NB) return;
    L.deg[b] = (L.maxen[L.bval[b]] == b) ? 1 : 0;              // deg is free again after auto-disable: head flags
}
// island number of a component = how many island heads have a higher body index (BuildIslands walks world->firstbody,
// i.e. descending creation index, and opens an island at every enabled body not yet reached, util.cpp:746-769)
__global__ void k_lw_label(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D, const __grid_constant__ LargePtrs L)
{
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= P.NB) return;
    const int T = L.head_scan[P.NB - 1];
    if (b == 0) { D.nislands[0] = T; L.counters[LWC_NISLANDS] = T; }
    const int h = L.maxen[L.bval[b]];
    int is = -1;
    u64 key = LW_NOKEY;
    if (h >= 0) {
        is = T - L.head_scan[h];
        D.bflags[b] &= ~BF_DISABLED;                          // bodies reached by the traversal are re-enabled (util.cpp:786-790)
        // one atomic per island and warp (a wall is ONE island: 10^5 atomics on one address otherwise)
        const unsigned peers = __match_any_sync(__activemask(), is);
        if ((threadIdx.x & 31) == __ffs(peers) - 1) { atomicAdd(&L.isl_nb[is], __popc(peers)); atomicAdd(&L.counters[LWC_NORDERED], __popc(peers)); }
        key = ((u64)(unsigned)is << 32) | (0xFFFFFFFFu - (unsigned)b);
    }
    D.body_island[b] = is;
    D.body_pos[b] = -1;
    L.bkey[b] = key;
}
__global__ void k_lw_iota(int n, int *v) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) v[i] = i; }
__global__ void k_lw_body_pos(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D, const __grid_constant__ LargePtrs L)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k == 0) D.nordered[0] = L.counters[LWC_NORDERED];
    if (k >= P.NB || L.bkey_s[k] == LW_NOKEY) return;
    D.body_pos[D.body_order[k]] = k;
}
__global__ void k_lw_joint_keys(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D, const __grid_constant__ LargePtrs L)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int nj = P.NJ + L.counters[LWC_NCONTACTS];
    if (j >= P.NJT) return;
    u64 key = LW_NOKEY; int m = 0;
    if (j < nj) {
        int b0;
        if (j < P.NJ) { b0 = D.joints[j].b0; m = D.jm[j]; }
        else { b0 = D.cinfo[j - P.NJ].y; m = P.m_contact; }
        const int is = D.body_island[b0];
        if (is >= 0 && m > 0) {
            key = ((u64)(unsigned)is << 32) | (unsigned)j;
            // one atomic per (island, rows per joint) and warp: a wall is one island, a million contact joints would queue on one address
            const unsigned peers = __match_any_sync(__activemask(), ((u64)(unsigned)is << 8) | (unsigned)m);
            if ((threadIdx.x & 31) == __ffs(peers) - 1) {
                const int n = __popc(peers);
                atomicAdd(&L.isl_m[is], m * n);
                atomicAdd(&L.counters[LWC_NJORD], n);
                atomicAdd(&L.counters[LWC_MROWS], m * n);
            }
        } else m = 0;
    }
    L.jkey[j] = key; L.jmv[j] = m;
}
__global__ void k_lw_joint_final(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D, const __grid_constant__ LargePtrs L, int nj)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k == 0) {
        D.njord[0] = L.counters[LWC_NJORD];
        int mr = L.counters[LWC_MROWS];
        if (mr > P.MR) { atomicExch(D.overflow, 3); }
        D.mrows[0] = mr;
    }
    if (k >= nj || L.jkey_s[k] == LW_NOKEY) return;
    const u64 key = L.jkey_s[k];
    const int is = (int)(key >> 32);
    D.joint_order[k] = (int)(key & 0xFFFFFFFFu);
    D.joint_island[k] = is;
    D.joint_row[k] = L.jrow[k] - L.isl_rstart[is];
}
__global__ void k_lw_island_info(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D, const __grid_constant__ LargePtrs L)
{
    int is = blockIdx.x * blockDim.x + threadIdx.x;
    if (is >= L.counters[LWC_NISLANDS]) return;
    const int m = L.isl_m[is];
    D.island_info[is] = make_int4(L.isl_bstart[is], L.isl_nb[is], L.isl_rstart[is], m);
    if (m == 0) L.isl_done[is] = 1; else atomicAdd(&L.counters[LWC_NACTIVE], 1);
}

// ------------------------------------------------------------------------------------------------ solve: groups, colours, sweeps
__global__ void k_lw_zero_cur(int n, int *v) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) v[i] = 0; }

// once per step: size of every group, the compact list of groups (first rows), how many groups act on every body, lambda = 0
__global__ void k_lwc_groups(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D, const __grid_constant__ LargePtrs L, int mrows)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= mrows) return;
    D.lambda[r] = 0;
    const int2 rb = D.rbody[r];                                   // accumulator slots of the row's two bodies (k_rows_t<false>)
    const int g = L.row_group[r];
    atomicAdd(&L.gsize[g], 1);
    if (g != r) return;
    L.heads[atomicAdd(&L.counters[LWC_NGROUPS], 1)] = r;
    atomicAdd(&L.ginc_cur[rb.x], 1);
    if (rb.y != P.NB) atomicAdd(&L.ginc_cur[rb.y], 1);
}
__global__ void k_lwc_ginc_fill(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D, const __grid_constant__ LargePtrs L)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= L.counters[LWC_NGROUPS]) return;
    const int g = L.heads[t];
    const int2 rb = D.rbody[g];
    // (group, its priority): the colouring rounds compare priorities without a second dependent load
    const unsigned is = (unsigned)L.row_island[g];
    const int2 e = make_int2(g, (int)odebi_canon_key(D.seed[0], is, 0u, (unsigned)(g - L.isl_rstart[is])));
    L.ginc[L.ginc_ofs[rb.x] + atomicAdd(&L.ginc_cur[rb.x], 1)] = e;
    if (rb.y != P.NB) L.ginc[L.ginc_ofs[rb.y] + atomicAdd(&L.ginc_cur[rb.y], 1)] = e;
}

// Colouring of one phase (see oracle/orc_world.cpp canonical_order: the same rounds).  A group is "above" another one when its
// (key, first row) is larger.  Round = mark (every uncoloured group that has no uncoloured neighbour above it wins; reads only
// colours of earlier rounds) + assign (winners, never neighbours of each other, take the smallest colour none of their coloured
// neighbours holds): the result does not depend on thread timing (k_lwc_color_rounds).
__global__ void k_lwc_color_init(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D, const __grid_constant__ LargePtrs L)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < ODEB_CANON_COLOURS) L.ccount[t] = 0;
    if (t >= L.counters[LWC_NGROUPS]) return;
    const int g = L.heads[t];
    const unsigned is = (unsigned)L.row_island[g];
    L.gkey[g] = odebi_canon_key(D.seed[0], is, 0u, (unsigned)(g - L.isl_rstart[is]));
    L.gcolor[g] = -1;
    atomicAdd(&L.counters[LWC_UNCOLORED], 1);
}
// all rounds of the colouring in ONE cooperative launch (grid barrier between mark and assign): a round is two passes over the groups that
// still matter, and as separate launches the ~30 rounds of a wall cost 64 launches of ~14 us each plus a host round trip every 8 rounds
__device__ __forceinline__ void lwc_grid_sync(unsigned *bar, unsigned &target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();
        atomicAdd(bar, 1u);
        unsigned v;                                               // relaxed polls (an acquire load invalidates the SM's L1 on every poll), one fence at the end
        do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory"); } while ((int)(v - target) < 0);
        __threadfence();
    }
    __syncthreads();
}
__global__ void __launch_bounds__(1024) k_lwc_color_rounds(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D, const __grid_constant__ LargePtrs L)
{
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, GT = gridDim.x * blockDim.x;
    const int ng = L.counters[LWC_NGROUPS];
    unsigned *gbar = (unsigned *)&L.counters[LWC_GBAR];
    unsigned target = 0;
    for (int round = 0; round < (1 << 20); round++) {
        // mark: every uncoloured group without an uncoloured neighbour above it wins (reads only colours of earlier rounds).  The neighbours
        // of a side are examined four at a time: their colours are independent loads (one exposed L2 latency per four neighbours)
        for (int t = gtid; t < ng; t += GT) {
            const int g = L.heads[t];
            if (__ldcg(&L.gcolor[g]) != -1) continue;
            const unsigned kg = L.gkey[g];
            const int2 rb = D.rbody[g];
            bool top = true;
            for (int side = 0; side < 2 && top; side++) {
                const int b = side ? rb.y : rb.x;
                if (b == P.NB) continue;
                const int e1 = L.ginc_ofs[b + 1];
                for (int e = L.ginc_ofs[b]; e < e1 && top; e += 4) {
                    int2 h[4]; int c[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) h[k] = e + k < e1 ? L.ginc[e + k] : make_int2(g, 0);
#pragma unroll
                    for (int k = 0; k < 4; k++) c[k] = h[k].x != g ? __ldcg(&L.gcolor[h[k].x]) : 0;
#pragma unroll
                    for (int k = 0; k < 4; k++) if (h[k].x != g && c[k] == -1 && ((unsigned)h[k].y > kg || ((unsigned)h[k].y == kg && h[k].x > g))) top = false;
                }
            }
            __stcg(&L.gwin[g], top ? 1 : 0);
        }
        lwc_grid_sync(gbar, target);
        // assign: the winners (never neighbours of each other) take the smallest colour none of their coloured neighbours holds
        for (int t = gtid; t < ng; t += GT) {
            const int g = L.heads[t];
            if (__ldcg(&L.gcolor[g]) != -1 || !__ldcg(&L.gwin[g])) continue;
            const int2 rb = D.rbody[g];
            u64 used[ODEB_CANON_COLOURS / 64];
#pragma unroll
            for (int k = 0; k < ODEB_CANON_COLOURS / 64; k++) used[k] = 0;
            for (int side = 0; side < 2; side++) {
                const int b = side ? rb.y : rb.x;
                if (b == P.NB) continue;
                const int e1 = L.ginc_ofs[b + 1];
                for (int e = L.ginc_ofs[b]; e < e1; e += 4) {
                    int c[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) c[k] = e + k < e1 ? __ldcg(&L.gcolor[L.ginc[e + k].x]) : -1;
#pragma unroll
                    for (int k = 0; k < 4; k++) if (c[k] >= 0 && c[k] < ODEB_CANON_COLOURS) {
#pragma unroll
                        for (int q = 0; q < ODEB_CANON_COLOURS / 64; q++) if ((c[k] >> 6) == q) used[q] |= 1ull << (c[k] & 63);
                    }
                }
            }
            int c = ODEB_CANON_COLOURS - 1;                        // the smallest colour no coloured neighbour holds
#pragma unroll
            for (int q = ODEB_CANON_COLOURS / 64 - 1; q >= 0; q--) if (~used[q] != 0ull) { const int f = 64 * q + __ffsll((long long)~used[q]) - 1; if (f < c) c = f; }
            if (c >= ODEB_CANON_COLOURS - 1) atomicExch(D.overflow, 4);   // a body with more than ODEB_CANON_COLOURS - 2 groups
            __stcg(&L.gcolor[g], c);
            atomicSub(&L.counters[LWC_UNCOLORED], 1);
            atomicMax(&L.counters[LWC_NCOLORS], c + 1);
            atomicAdd(&L.ccount[c], 1);
        }
        lwc_grid_sync(gbar, target);
        if (__ldcg(&L.counters[LWC_UNCOLORED]) <= 0) break;
    }
}

__device__ __forceinline__ Real4 ldcg4(const Real4 *p)
{
#if defined(ODEB_DOUBLE)
    double2 a = __ldcg((const double2 *)p), b = __ldcg((const double2 *)p + 1);
    Real4 r = { a.x, a.y, b.x, b.y };
#else
    float4 a = __ldcg((const float4 *)p);
    Real4 r = { a.x, a.y, a.z, a.w };
#endif
    return r;
}
__device__ __forceinline__ void stcg4(Real4 *p, const Real4 &v)
{
#if defined(ODEB_DOUBLE)
    __stcg((double2 *)p, make_double2(v.x, v.y)); __stcg((double2 *)p + 1, make_double2(v.z, v.w));
#else
    __stcg((float4 *)p, make_float4(v.x, v.y, v.z, v.w));
#endif
}

// ---- tiles.  The groups are sorted by (colour, rows descending) and cut into TILES of 32 groups of one colour; a tile is as high as its
// largest group.  The records of a tile are stored lane-interleaved in a second buffer: quad q (16 bytes) of row k of the tile's group
// `lane` sits at quad (q * 32 + lane) of tile row tbase + k, its lambda behind the row's LWT_QUADS x 32 quads.  The whole tile is ONE
// contiguous block of HBM in exactly the order the sweep walks it, every warp access is a full line set, and the colouring is fixed for
// the step, so the layout is built once per step (k_lwt_finish) and the 40 sweeps stream it.
// The tile record is COMPACT, 20 reals instead of the 32 of the batched kernels' record: the sweeps of a big world are bound by HBM
// bandwidth, and 12 of the 32 reals (iMJ = invM * J^T) are cheap to recompute from the unscaled J and the two bodies' inverse mass /
// inertia, which a lane loads once per group (L2 hits).  Every recomputed value comes out of the same expression on the same inputs as
// in finish_row (compute_invM_JT quickstep.cpp:859-897, Ad scaling :2251-2316), so the bits are the same.
//   c0 = J1l.xyz J1a.x | c1 = J1a.yz, rhs * Ad, cfm * Ad | c2 = J2l.xyz J2a.x | c3 = J2a.yz, lo, hi | c4 = Ad, row - findex, modmax 1, modmax 2
#define LWT_QUADS 5
#define LWT_ROW_BYTES (32 * LWT_QUADS * (int)sizeof(Real4) + 32 * (int)sizeof(Real))
#define LWT_LAM_OFS (32 * LWT_QUADS * (int)sizeof(Real4))
__global__ void k_lwt_sort_keys(const __grid_constant__ LargePtrs L)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= L.counters[LWC_NGROUPS]) return;
    const int g = L.heads[t];
    const int sz = L.gsize[g] < 4095 ? L.gsize[g] : 4095;
    L.skey[t] = ((unsigned)L.gcolor[g] << 12) | (unsigned)(4095 - sz);      // 8 + 12 bits
}
// list offsets and first tile of every colour (one thread)
__global__ void k_lwt_color_scan(const __grid_constant__ DevPtrs D, const __grid_constant__ LargePtrs L)
{
    int o = 0, t = 0;
    for (int c = 0; c < ODEB_CANON_COLOURS; c++) { L.cofs[c] = o; L.tstart[c] = t; o += L.ccount[c]; t += (L.ccount[c] + 31) >> 5; }
    L.cofs[ODEB_CANON_COLOURS] = o; L.tstart[ODEB_CANON_COLOURS] = t;
    L.counters[LWC_NTILES] = t;
    L.counters[LWC_SEED] = (int)D.seed[0];
}
// one warp per tile: the groups of its lanes (rows, accumulator slots of the two bodies, island), the tile's height
__global__ void __launch_bounds__(128) k_lwt_tiles(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D, const __grid_constant__ LargePtrs L)
{
    const int tile = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (tile >= L.counters[LWC_NTILES]) return;
    int c = 0;
    {   // the colour whose tile range holds this tile (tstart is ascending)
        int lo = 0, hi = ODEB_CANON_COLOURS - 1;
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (tile >= L.tstart[mid]) lo = mid; else hi = mid - 1; }
        c = lo;
        while (c < ODEB_CANON_COLOURS - 1 && L.tstart[c + 1] <= tile) c++;   // skip empty colours that start at the same tile
    }
    const int pos = L.cofs[c] + ((tile - L.tstart[c]) << 5) + lane;
    int4 gi = make_int4(0, 0, P.NB, 0);                            // (rows, body 1, body 2, island); rows == 0: empty lane
    int g = -1;
    if (pos < L.cofs[c + 1]) {
        g = L.clist[pos];
        const int2 rb = D.rbody[g];
        gi = make_int4(L.gsize[g], rb.x, rb.y, L.row_island[g]);
    }
    L.tginfo[(size_t)tile * 32 + lane] = gi;
    L.tgroup[(size_t)tile * 32 + lane] = g;
    const int h = __reduce_max_sync(0xffffffffu, gi.x);
    if (lane == 0) { L.theight[tile] = h; if (tile == 0) L.theight[L.counters[LWC_NTILES]] = 0; }
}
// one block per tile: Stage2c + compute_invM_JT + Ad (finish_row's arithmetic) of the half-built records k_rows_t<false> left, written as
// compact tile records; lambda = 0; the bodies' inverse masses by order position for the sweeps
__global__ void __launch_bounds__(128) k_lwt_finish(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D, const __grid_constant__ LargePtrs L)
{
    const int tile = blockIdx.x, lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int base = L.tbase[tile];
    if (base + L.theight[tile] > L.trcap) { if (threadIdx.x == 0) atomicExch(D.overflow, 3); return; }
    const int g = L.tgroup[(size_t)tile * 32 + lane];
    if (g < 0) return;
    const int4 gi = L.tginfo[(size_t)tile * 32 + lane];
    const int sz = gi.x;
    const bool two = gi.z != P.NB;
    Real in0[6], invI0[12], im0, in1[6], invI1[12], im1 = 0;
    body_rhs_tmp(P, D, 0, gi.y, in0, invI0, &im0);
    if (two) body_rhs_tmp(P, D, 0, gi.z, in1, invI1, &im1);
    if (wib == 0) { L.pinvm[gi.y] = im0; if (two) L.pinvm[gi.z] = im1; }
    for (int k = wib; k < sz; k += 4) {
        const Real4 *src = D.rows + (size_t)(g + k) * 8;
        const Real4 v0 = ldcg4(src), v1 = ldcg4(src + 1), v2 = ldcg4(src + 2), v3 = ldcg4(src + 3);
        const Real q[16] = { v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w, v3.x, v3.y, v3.z, v3.w };
        Real imj[14];
        Real sum = R_(0.0);
        for (int c = 0; c < 6; c++) sum += q[C_J1L + c] * in0[c];
        for (int c = 0; c < 3; c++) imj[c] = im0 * q[C_J1L + c];
        mul0_331(imj + 3, invI0, q + C_J1A);
        imj[6] = P.dyn_enabled ? modmax6(imj) : R_(0.0);
        for (int c = 7; c < 14; c++) imj[c] = 0;
        if (two) {
            for (int c = 0; c < 6; c++) sum += q[C_J2L + c] * in1[c];
            for (int c = 0; c < 3; c++) imj[7 + c] = im1 * q[C_J2L + c];
            mul0_331(imj + 10, invI1, q + C_J2A);
            imj[13] = P.dyn_enabled ? modmax6(imj + 7) : R_(0.0);
        }
        const Real rhs = q[C_RHS] + sum;
        Real s2 = R_(0.0);
        for (int c = 0; c < 6; c++) s2 += imj[c] * q[C_J1L + c];
        if (two) for (int c = 0; c < 6; c++) s2 += imj[7 + c] * q[C_J2L + c];
        const Real cfm_i = q[C_CFM];
        const Real Ad = P.sor_w / (s2 + cfm_i);
        const int fi = D.findex[g + k];
        Real4 c1 = v1, c4 = { Ad, 0, imj[6], imj[13] };
        c1.z = rhs * Ad; c1.w = cfm_i * Ad;
        *(int *)&c4.y = fi == -1 ? 0 : g + k - fi;
        unsigned char *trow = L.trec + (size_t)(base + k) * LWT_ROW_BYTES;
        Real4 *dst = (Real4 *)trow + lane;
        dst[0] = v0; dst[32] = c1; dst[64] = v2; dst[96] = v3; dst[128] = c4;
        ((Real *)(trow + LWT_LAM_OFS))[lane] = 0;
    }
}
// feedback only: the final lambdas back in row order
__global__ void __launch_bounds__(128) k_lwt_lambda_out(const __grid_constant__ DevPtrs D, const __grid_constant__ LargePtrs L)
{
    const int tile = blockIdx.x, lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int base = L.tbase[tile];
    const int g = L.tgroup[(size_t)tile * 32 + lane];
    const int sz = g >= 0 ? L.gsize[g] : 0;
    for (int k = wib; k < sz; k += 4) D.lambda[g + k] = ((const Real *)(L.trec + (size_t)(base + k) * LWT_ROW_BYTES + LWT_LAM_OFS))[lane];
}

// The row update (Stage4LCP_IterationStep quickstep.cpp:2917-3033) of one row of a group whose two bodies' accumulators, inverse masses
// and inverse inertias are in registers, from the compact tile record C0..C4.  `fd` = row - findex (0: none); the row it points at is an
// earlier row of the same group (findex is joint-local), normally the latest row without a friction index, whose new lambda is kept in a
// register.
#define LWT_ROW(C0, C1, C2, C3, C4, OLD, K)                                                                              \
    {                                                                                                                    \
        const Real Ad = (C4).x;                                                                                          \
        const int fd = *(const int *)&(C4).y;                                                                            \
        const Real ja1[3] = { (C0).w, (C1).x, (C1).y };                                                                  \
        Real ma1[3];                                                                                                     \
        mul0_331(ma1, invI0, ja1);                                                                                       \
        Real delta = (C1).z - (OLD) * (C1).w;                                                                            \
        delta -= f1a.x * ((C0).x * Ad) + f1a.y * ((C0).y * Ad) + f1a.z * ((C0).z * Ad) + f1a.w * ((C0).w * Ad) + f1b.x * ((C1).x * Ad) + f1b.y * ((C1).y * Ad); \
        Real ma2[3] = { 0, 0, 0 };                                                                                       \
        if (two) {                                                                                                       \
            const Real ja2[3] = { (C2).w, (C3).x, (C3).y };                                                              \
            mul0_331(ma2, invI1, ja2);                                                                                   \
            delta -= f2a.x * ((C2).x * Ad) + f2a.y * ((C2).y * Ad) + f2a.z * ((C2).z * Ad) + f2a.w * ((C2).w * Ad) + f2b.x * ((C3).x * Ad) + f2b.y * ((C3).y * Ad); \
        }                                                                                                                \
        Real hi_act, lo_act;                                                                                             \
        if (fd != 0) { hi_act = RFABS((C3).w * ((K) - fd == free_k ? free_lambda : __ldcg((const Real *)(lamp + (size_t)((K) - fd) * LWT_ROW_BYTES)))); lo_act = -hi_act; } \
        else { hi_act = (C3).w; lo_act = (C3).z; }                                                                       \
        Real new_lambda = (OLD) + delta;                                                                                 \
        if (new_lambda < lo_act) { delta = lo_act - (OLD); new_lambda = lo_act; }                                        \
        else if (new_lambda > hi_act) { delta = hi_act - (OLD); new_lambda = hi_act; }                                   \
        *(Real *)(lamp + (size_t)(K) * LWT_ROW_BYTES) = new_lambda;                                                      \
        if (fd == 0) { free_k = (K); free_lambda = new_lambda; }                                                         \
        if (delta != 0) {                                                                                                \
            f1a.x += delta * (im0 * (C0).x); f1a.y += delta * (im0 * (C0).y); f1a.z += delta * (im0 * (C0).z); f1a.w += delta * ma1[0]; \
            f1b.x += delta * ma1[1]; f1b.y += delta * ma1[2];                                                            \
            if (delta > 0) f1b.w += delta * (C4).z; else f1b.z += delta * (C4).z;                                        \
            if (two) {                                                                                                   \
                if (delta > 0) f2b.w += delta * (C4).w; else f2b.z += delta * (C4).w;                                    \
                f2a.x += delta * (im1 * (C2).x); f2a.y += delta * (im1 * (C2).y); f2a.z += delta * (im1 * (C2).z); f2a.w += delta * ma2[0]; \
                f2b.x += delta * ma2[1]; f2b.y += delta * ma2[2];                                                        \
            }                                                                                                            \
        }                                                                                                                \
    }

// cp.async.bulk + mbarrier plumbing of the sweep kernel's per-warp ring
__device__ __forceinline__ void lwt_mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void lwt_mbar_expect_tx(unsigned bar, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void lwt_bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void lwt_mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tLWT_WAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra LWT_DONE_%=;\n\tbra LWT_WAIT_%=;\n\tLWT_DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
#define LWT_STAGES 6
#ifndef LWT_WARPS
#if defined(ODEB_DOUBLE)
#define LWT_WARPS 2
#else
#define LWT_WARPS 4
#endif
#endif
#define LWT_STAGE_BYTES LWT_ROW_BYTES
// small worlds: ONE block of LWT_SMALL_WARPS warps holds every tile of a colour, the colours are separated by __syncthreads (~0.1 us) instead of
// the grid barrier (~2 us of L2 round trips): a 1000-body pile spends most of its step in those barriers otherwise
#if defined(ODEB_DOUBLE)
#define LWT_SMALL_WARPS 8
#else
#define LWT_SMALL_WARPS 16
#endif
#define LWT_SMALL_STAGES 4
// A whole PHASE (up to 8 sweeps: every colour in the phase's order, the per-body convergence test and the per-island iteration control
// after each sweep) as one persistent cooperative launch.  ncu on the launch-per-colour version: a colour costs ~9 us however few tiles
// it has (launch, tile lookup, first DRAM round trip, 12 dependent rows) and only ~8 us more for 46 MB of records -- latency, not
// bandwidth.  Here a warp owns the same tiles in every sweep, the colours are separated by a grid barrier (one atomic + one spin per
// block), and the records of a warp's NEXT tile are requested (tile lookup + the first bulk copies of its ring) BEFORE the warp enters the
// barrier: records and lambdas of a tile only change when that tile itself runs, so the only loads left behind a barrier are the two
// bodies' accumulators (L2 hits).  Everything another block may have written is read with ld.cg (L1 is not coherent across the barrier).
struct LwPhase {
    int norder;                                  // colours that have tiles, in the phase's visiting order
    int corder[ODEB_CANON_COLOURS];
    int tstart[ODEB_CANON_COLOURS + 1];          // first tile of every colour
    int nordered, nislands;
    int phase;                                   // 0, 1, ...: this launch runs the sweeps 8 phase .. 8 phase + 7 (if the solve gets that far)
};
// a single-block launch (small worlds: every colour fits the block's warps) synchronises with the block barrier alone
__device__ __forceinline__ void lwt_sync(unsigned *bar, unsigned &target)
{
    if (gridDim.x == 1) __syncthreads(); else lwc_grid_sync(bar, target);
}
template <int WARPS, int STAGES>
__global__ void __launch_bounds__(32 * WARPS) k_lwt_phase_t(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D, const __grid_constant__ LargePtrs L, const __grid_constant__ LwPhase ph)
{
    extern __shared__ __align__(128) unsigned char lwt_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int gw = blockIdx.x * WARPS + wib, TW = gridDim.x * WARPS;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, GT = gridDim.x * blockDim.x;
    unsigned char *ring = lwt_smem + (size_t)wib * STAGES * LWT_STAGE_BYTES;
    unsigned long long *bars = (unsigned long long *)(lwt_smem + (size_t)WARPS * STAGES * LWT_STAGE_BYTES) + wib * STAGES;
    const unsigned bar0 = (unsigned)__cvta_generic_to_shared(bars);
    const unsigned ring0 = (unsigned)__cvta_generic_to_shared(ring);
    constexpr unsigned RB = 32 * LWT_QUADS * (unsigned)sizeof(Real4), LB = 32 * (unsigned)sizeof(Real);
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) lwt_mbar_init(bar0 + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    // All phases of a step are queued back to back (no host round trip in between); a phase whose predecessors ended the solve returns at
    // once.  Iteration state travels in the counters (written by the previous launch: kernel boundary).
    if (L.counters[LWC_TERM] || L.counters[LWC_NACTIVE] == 0 || L.counters[LWC_ITER] != 8 * ph.phase) return;
    unsigned *gbar = (unsigned *)&L.counters[LWC_PBAR + ph.phase];
    unsigned gtarget = 0;
    unsigned ri = 0, rc = 0;                                       // tile rows requested / consumed by this warp so far
    unsigned is = 0, cs = 0, cpar = 0;                             // ring stage of the next request / of the next row consumed, and that stage's mbarrier parity
    // the warp's pending item (sweep, position in the colour order, tile) and the state of the tile that was begun for it
    int it_sw = 0, it_ci = ph.norder, it_tile = 0;
    for (int ci = 0; ci < ph.norder; ci++) { const int c = ph.corder[ci]; if (ph.tstart[c] + gw < ph.tstart[c + 1]) { it_ci = ci; it_tile = ph.tstart[c] + gw; break; } }
    const int first_ci = it_ci, first_tile = it_tile;
    const bool any_items = first_ci < ph.norder;
    int height = 0, issued = 0; int4 gi = make_int4(0, 0, P.NB, 0);
    unsigned char *rec = 0;
    Real invI0[12], invI1[12], im0 = 0, im1 = 0;
#pragma unroll
    for (int c = 0; c < 12; c++) { invI0[c] = 0; invI1[c] = 0; }
    auto tile_begin = [&](int tile) {
        height = L.theight[tile];
        const int base = L.tbase[tile];
        rec = L.trec + (size_t)base * LWT_ROW_BYTES;
        gi = L.tginfo[(size_t)tile * 32 + lane];
        // the two bodies' inverse inertia (world frame, k_body_pre) and inverse mass: constant during the solve
        if (gi.x > 0) {
            const Real4 *ii = (const Real4 *)(D.invIw + 12 * (size_t)gi.y);
            const Real4 r0 = ii[0], r1 = ii[1], r2 = ii[2];
            invI0[0] = r0.x; invI0[1] = r0.y; invI0[2] = r0.z; invI0[4] = r1.x; invI0[5] = r1.y; invI0[6] = r1.z; invI0[8] = r2.x; invI0[9] = r2.y; invI0[10] = r2.z;
            im0 = L.pinvm[gi.y];
            if (gi.z != P.NB) {
                const Real4 *jj = (const Real4 *)(D.invIw + 12 * (size_t)gi.z);
                const Real4 t0 = jj[0], t1 = jj[1], t2 = jj[2];
                invI1[0] = t0.x; invI1[1] = t0.y; invI1[2] = t0.z; invI1[4] = t1.x; invI1[5] = t1.y; invI1[6] = t1.z; invI1[8] = t2.x; invI1[9] = t2.y; invI1[10] = t2.z;
                im1 = L.pinvm[gi.z];
            }
        }
        issued = height < STAGES - 1 ? height : STAGES - 1;
        if (lane == 0) {
            unsigned s = is;
            for (int k = 0; k < issued; k++) {
                lwt_mbar_expect_tx(bar0 + 8 * s, RB + LB);
                lwt_bulk_g2s(ring0 + s * LWT_STAGE_BYTES, rec + (size_t)k * LWT_ROW_BYTES, RB + LB, bar0 + 8 * s);
                s = s + 1 == STAGES ? 0 : s + 1;
            }
        }
        ri += issued;
        is = (is + issued) % STAGES;
    };
    if (any_items) tile_begin(it_tile);
    unsigned iteration = (unsigned)L.counters[LWC_ITER], extra = (unsigned)L.counters[LWC_EXTRA];
    Real exit_delta = extra ? P.extra_delta : P.premature_delta;
    int terminated = 0;
    Real4 *cf = D.cforce;
    for (int sw = 0; sw < 8; sw++) {
        for (int ci = 0; ci < ph.norder; ci++) {
            while (any_items && it_sw == sw && it_ci == ci) {
                // ---- run the begun tile
                const int sz = gi.x;
                const bool two = gi.z != P.NB;
                Real4 f1a = { 0, 0, 0, 0 }, f1b = f1a, f2a = f1a, f2b = f1a;
                bool run = false;
                if (sz > 0) {
                    f1a = ldcg4(&cf[2 * gi.y]); f1b = ldcg4(&cf[2 * gi.y + 1]);
                    if (two) { f2a = ldcg4(&cf[2 * gi.z]); f2b = ldcg4(&cf[2 * gi.z + 1]); }
                    run = __ldcg(&L.isl_done[gi.w]) == 0;
                }
                unsigned char *lamp = rec + LWT_LAM_OFS + lane * sizeof(Real);
                int free_k = -1; Real free_lambda = 0;
                for (int k = 0; k < height; k++) {
                    const unsigned s = cs;
                    lwt_mbar_wait(bar0 + 8 * s, cpar);
                    if (run && k < sz) {
                        const Real4 *st = (const Real4 *)(ring + (size_t)s * LWT_STAGE_BYTES) + lane;
                        const Real4 c0 = st[0], c1 = st[32], c2 = st[64], c3 = st[96], c4 = st[128];
                        const Real old_lambda = ((const Real *)(ring + (size_t)s * LWT_STAGE_BYTES + RB))[lane];
                        LWT_ROW(c0, c1, c2, c3, c4, old_lambda, k)
                    }
                    rc++;
                    if (++cs == STAGES) { cs = 0; cpar ^= 1u; }
                    __syncwarp();
                    if (issued < height) {
                        if (lane == 0) {
                            const unsigned sn = is;
                            lwt_mbar_expect_tx(bar0 + 8 * sn, RB + LB);
                            lwt_bulk_g2s(ring0 + sn * LWT_STAGE_BYTES, rec + (size_t)issued * LWT_ROW_BYTES, RB + LB, bar0 + 8 * sn);
                        }
                        ri++; issued++;
                        is = is + 1 == STAGES ? 0 : is + 1;
                    }
                }
                if (run) {
                    stcg4(&cf[2 * gi.y], f1a); stcg4(&cf[2 * gi.y + 1], f1b);
                    if (two) { stcg4(&cf[2 * gi.z], f2a); stcg4(&cf[2 * gi.z + 1], f2b); }
                }
                // ---- the warp's next item; its tile is begun right away (before the barrier).  The lambdas this tile just wrote are read
                // again by the bulk copies of its next visit (async proxy), possibly at once (a warp with a single tile)
                asm volatile("fence.proxy.async.global;" ::: "memory");
                __syncwarp();
                {
                    const int c = ph.corder[it_ci];
                    it_tile += TW;
                    if (it_tile >= ph.tstart[c + 1]) {
                        int cj = it_ci + 1;
                        for (; cj < ph.norder; cj++) { const int c2 = ph.corder[cj]; if (ph.tstart[c2] + gw < ph.tstart[c2 + 1]) { it_tile = ph.tstart[c2] + gw; break; } }
                        if (cj < ph.norder) it_ci = cj;
                        else { it_sw++; it_ci = first_ci; it_tile = first_tile; }
                    }
                    if (it_sw < 8) tile_begin(it_tile);
                }
            }
            lwt_sync(gbar, gtarget);
        }
        // ---- end of the sweep: iteration control quickstep.cpp:1832-1855, :3253-3285 (k_lw_body_check + k_lw_island_ctl)
        ++iteration;
        int terminate_all = 0, in_extra = 0;
        if (iteration - extra == P.num_iter) {
            if (extra != 0 || P.max_extra == 0) { terminate_all = 1; in_extra = extra != 0; }
            else { extra = P.max_extra; exit_delta = P.extra_delta; }
        }
        if (P.dyn_enabled && !terminate_all) {
            for (int k = gtid; k < ph.nordered; k += GT) {
                const int is = D.body_island[D.body_order[k]];
                if (__ldcg(&L.isl_done[is])) continue;
                Real4 v = ldcg4(&cf[2 * k + 1]);
                if (!(v.w < exit_delta) || !(-v.z < exit_delta)) __stcg(&L.isl_viol[is], 1);
                v.z = 0; v.w = 0;
                stcg4(&cf[2 * k + 1], v);
            }
            lwt_sync(gbar, gtarget);
        }
        for (int is = gtid; is < ph.nislands; is += GT) {
            if (__ldcg(&L.isl_done[is])) continue;
            const int m = L.isl_m[is];
            atomicAdd(&L.draws[1], 1ull); atomicAdd(&L.draws[2], (u64)m);
            unsigned *st = D.stats;
            bool done = false;
            if (terminate_all) { if (in_extra) atomicAdd(&st[3], 1u); done = true; }
            else if (P.dyn_enabled) {
                const bool hit = (exit_delta == 0) || __ldcg(&L.isl_viol[is]);
                __stcg(&L.isl_viol[is], 0);
                if (!hit) {
                    if (iteration < P.num_iter) atomicAdd(&st[1], 1u);
                    else if (iteration > P.num_iter) atomicAdd(&st[2], 1u);
                    done = true;
                }
            }
            if (done) { __stcg(&L.isl_done[is], 1); atomicSub(&L.counters[LWC_NACTIVE], 1); }
            else if (iteration >= 8 && (iteration & 7) == 0) atomicAdd(&L.draws[0], (u64)(m - 1));
        }
        lwt_sync(gbar, gtarget);
        if (terminate_all) { terminated = 1; break; }
        if (__ldcg(&L.counters[LWC_NACTIVE]) == 0) break;
    }
    // bulk copies of a tile that was begun for a sweep that does not take place must land before the block retires
    while (rc < ri) { lwt_mbar_wait(bar0 + 8 * cs, cpar); rc++; if (++cs == STAGES) { cs = 0; cpar ^= 1u; } }
    if (gtid == 0) { L.counters[LWC_ITER] = (int)iteration; L.counters[LWC_EXTRA] = (int)extra; L.counters[LWC_TERM] = terminated; }
}

// after a sweep: per-body convergence test + reset (CheckForMaximumToBeLessThanLimitAndResetMaxAdjustments quickstep.cpp:3253-3285)
__global__ void k_lw_body_check(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D, const __grid_constant__ LargePtrs L, Real exit_delta, int check)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= L.counters[LWC_NORDERED]) return;
    const int is = D.body_island[D.body_order[k]];
    if (L.isl_done[is] || !check) return;
    Real4 v = D.cforce[2 * k + 1];
    if (!(v.w < exit_delta) || !(-v.z < exit_delta)) L.isl_viol[is] = 1;
    v.z = 0; v.w = 0;
    D.cforce[2 * k + 1] = v;
}
// per-island iteration control (quickstep.cpp:1832-1855). All unfinished islands share the sweep counter, so the host
// passes the scalar part of the decision (`iteration` after the increment, whether the sweep limit was reached).
__global__ void k_lw_island_ctl(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D, const __grid_constant__ LargePtrs L,
                                unsigned iteration, int terminate_all, int in_extra, Real exit_delta)
{
    int is = blockIdx.x * blockDim.x + threadIdx.x;
    if (is >= L.counters[LWC_NISLANDS]) return;
    if (L.isl_done[is]) return;
    const int m = L.isl_m[is];
    atomicAdd(&L.draws[1], 1ull); atomicAdd(&L.draws[2], (u64)m);
    unsigned *st = D.stats;
    bool done = false;
    if (terminate_all) { if (in_extra) atomicAdd(&st[3], 1u); done = true; }
    else if (P.dyn_enabled) {
        const bool hit = (exit_delta == 0) || L.isl_viol[is];
        L.isl_viol[is] = 0;
        if (!hit) {
            if (iteration < P.num_iter) atomicAdd(&st[1], 1u);
            else if (iteration > P.num_iter) atomicAdd(&st[2], 1u);
            done = true;
        }
    }
    if (done) { L.isl_done[is] = 1; atomicSub(&L.counters[LWC_NACTIVE], 1); }
    else if (iteration >= 8 && (iteration & 7) == 0) atomicAdd(&L.draws[0], (u64)(m - 1));   // the reorder this island is about to do
}
// end of step: Stage5 statistics, and the dRand stream advanced by the draws the reference's Fisher-Yates reorders consume
__global__ void k_lw_finish(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D, const __grid_constant__ LargePtrs L)
{
    D.stats[0] += (unsigned)L.counters[LWC_NISLANDS];
    u64 n = L.draws[0];
    unsigned acc_mult = 1u, acc_plus = 0u, cur_mult = 1664525u, cur_plus = 1013904223u;
    while (n) {
        if (n & 1ull) { acc_mult *= cur_mult; acc_plus = acc_plus * cur_mult + cur_plus; }
        cur_plus = (cur_mult + 1u) * cur_plus; cur_mult *= cur_mult;
        n >>= 1;
    }
    D.seed[0] = acc_mult * D.seed[0] + acc_plus;
    D.sweeps[0] = L.draws[1]; D.sweeps[1] = L.draws[2];
}
#endif
