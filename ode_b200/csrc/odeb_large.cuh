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
//   solve        per phase of 8 sweeps the row GROUPS (rows of the contacts of one geom pair / of one joint: same two bodies) are
//                coloured so that groups of one colour touch disjoint bodies (Jones-Plassmann rounds with first fit, priorities =
//                odeb_canon_key(seed, island, phase, group): deterministic, the oracle runs the same rounds); the canonical sweep
//                order is colour-major, and k_lwc_sweep relaxes all groups of a colour side by side, one thread per group with the
//                two bodies' accumulators in registers: bit-identical to the sequential sweep in that order (quickstep.cpp:2917-3033),
//                no tickets, no fences, no polling -- ~9 colours on a brick wall, i.e. ~9 short launches per sweep
//                (round 1 walked a hash-sorted order with per-body tickets: 0.58 ms per sweep on the 100k-box wall, bound by the
//                L2 round trips of the release / acquire hand-over along the dependency paths)
//   control      k_lw_body_check + k_lw_island_ctl after every sweep (quickstep.cpp:1823-1856, :3253-3285), per island
#ifndef ODEB_LARGE_CUH
#define ODEB_LARGE_CUH
#include <cub/cub.cuh>

typedef unsigned long long u64;
#define LW_NOKEY 0xFFFFFFFFFFFFFFFFull

enum { LWC_NBIG = 0, LWC_NPAIRS = 1, LWC_NCONTACTS = 2, LWC_NORDERED = 3, LWC_NJORD = 4, LWC_MROWS = 5, LWC_NISLANDS = 6,
       LWC_NACTIVE = 7, LWC_NGROUPS = 8, LWC_UNCOLORED = 9, LWC_NCOLORS = 10, LWC_BIGGROUPS = 11, LWC_CHUNKS = 12, LWC_COUNT = 16 };

struct LargePtrs {
    int *counters;                               // [LWC_COUNT]
    u64 *draws;                                  // [4]: dRand draws the reference's reorders would have consumed, sweeps, row-sweeps
    // broadphase
    unsigned *bp_key, *bp_key_s; int *bp_idx, *bp_idx_s, *bp_big;   // [NG]
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
    int *ginc_ofs, *ginc_cur, *ginc;             // [NB + 2], [NB + 2], [2 MR]: body (order position) -> groups acting on it
    unsigned *gkey; int *gcolor, *gwin;          // [MR] at the group's first row: priority of the phase, colour (-1 none yet, -2 island finished), winner flag
    int *clist, *ccount, *cofs;                  // [MR] groups by colour; [64] groups per colour; [65] first list position of every colour
    int4 *cinfo;                                 // [MR] beside clist: (first row, rows, accumulator slot of body 1, of body 2) of the group
    int *slot_row, *cstart;                      // [3 MR + 4096] row of every lane of every 32-row chunk (-1: empty); [65] first chunk of every colour
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

// sweep along axis 0 in float-sorted order: every pair whose axis-0 intervals overlap is visited from its earlier member
__global__ void k_bp_sweep(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D, const __grid_constant__ LargePtrs L)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.NG) return;
    if (L.bp_key_s[p] == 0xFFFFFFFFu) return;
    const int i = L.bp_idx_s[p];
    Real ai[6];
    for (int k = 0; k < 6; k++) ai[k] = D.aabb[6 * (size_t)i + k];
    const unsigned kmax = lw_float_key(ai[1]);
    for (int q = p + 1; q < P.NG; q++) {
        if (L.bp_key_s[q] > kmax) break;
        const int j = L.bp_idx_s[q];
        Real aj[6];
        for (int k = 0; k < 6; k++) aj[k] = D.aabb[6 * (size_t)j + k];
        if (i < j) { if (pair_hit(P, D, ai, aj, i, j)) lw_emit_pair(P, L, i, j); }
        else if (pair_hit(P, D, aj, ai, j, i)) lw_emit_pair(P, L, j, i);
    }
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
    else { int4 v = D.cinfo[j - P.NJ]; b0 = v.y; b1 = v.z; }
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
        atomicAdd(&L.isl_nb[is], 1);
        atomicAdd(&L.counters[LWC_NORDERED], 1);
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
            atomicAdd(&L.isl_m[is], m);
            atomicAdd(&L.counters[LWC_NJORD], 1);
            atomicAdd(&L.counters[LWC_MROWS], m);
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
    const int g = L.row_group[r];
    atomicAdd(&L.gsize[g], 1);
    if (g != r) return;
    L.heads[atomicAdd(&L.counters[LWC_NGROUPS], 1)] = r;
    const int2 rb = D.rbody[r];
    atomicAdd(&L.ginc_cur[rb.x], 1);
    if (rb.y != P.NB) atomicAdd(&L.ginc_cur[rb.y], 1);
}
__global__ void k_lwc_ginc_fill(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D, const __grid_constant__ LargePtrs L)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= L.counters[LWC_NGROUPS]) return;
    const int g = L.heads[t];
    const int2 rb = D.rbody[g];
    L.ginc[L.ginc_ofs[rb.x] + atomicAdd(&L.ginc_cur[rb.x], 1)] = g;
    if (rb.y != P.NB) L.ginc[L.ginc_ofs[rb.y] + atomicAdd(&L.ginc_cur[rb.y], 1)] = g;
}

// Colouring of one phase (see oracle/orc_world.cpp canonical_order: the same rounds).  A group is "above" another one when its
// (key, first row) is larger.  Round = k_lwc_mark (every uncoloured group that has no uncoloured neighbour above it wins; reads only
// colours of earlier rounds) + k_lwc_assign (winners, never neighbours of each other, take the smallest colour none of their coloured
// neighbours holds): the result does not depend on thread timing.
__global__ void k_lwc_color_init(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D, const __grid_constant__ LargePtrs L, int phase)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < 64) L.ccount[t] = 0;
    if (t >= L.counters[LWC_NGROUPS]) return;
    const int g = L.heads[t];
    const unsigned is = (unsigned)L.row_island[g];
    if (L.isl_done[is]) { L.gcolor[g] = -2; return; }
    L.gkey[g] = odebi_canon_key(D.seed[0], is, (unsigned)phase, (unsigned)(g - L.isl_rstart[is]));
    L.gcolor[g] = -1;
    atomicAdd(&L.counters[LWC_UNCOLORED], 1);
}
__global__ void k_lwc_mark(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D, const __grid_constant__ LargePtrs L)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= L.counters[LWC_NGROUPS]) return;
    const int g = L.heads[t];
    if (L.gcolor[g] != -1) { L.gwin[g] = 0; return; }
    const unsigned kg = L.gkey[g];
    const int2 rb = D.rbody[g];
    bool top = true;
    for (int side = 0; side < 2 && top; side++) {
        const int b = side ? rb.y : rb.x;
        if (b == P.NB) continue;
        for (int e = L.ginc_ofs[b], e1 = L.ginc_ofs[b + 1]; e < e1; e++) {
            const int h = L.ginc[e];
            if (h == g || L.gcolor[h] != -1) continue;
            const unsigned kh = L.gkey[h];
            if (kh > kg || (kh == kg && h > g)) { top = false; break; }
        }
    }
    L.gwin[g] = top ? 1 : 0;
}
__global__ void k_lwc_assign(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D, const __grid_constant__ LargePtrs L)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= L.counters[LWC_NGROUPS]) return;
    const int g = L.heads[t];
    if (!L.gwin[g]) return;
    const int2 rb = D.rbody[g];
    unsigned long long used = 0;
    for (int side = 0; side < 2; side++) {
        const int b = side ? rb.y : rb.x;
        if (b == P.NB) continue;
        for (int e = L.ginc_ofs[b], e1 = L.ginc_ofs[b + 1]; e < e1; e++) {
            const int c = L.gcolor[L.ginc[e]];                 // winners of this round are never neighbours: every neighbour's colour is final or -1
            if (c >= 0 && c < 64) used |= 1ull << c;
        }
    }
    int c = 0;
    while (c < 63 && ((used >> c) & 1ull)) c++;
    if (c >= 63) atomicExch(D.overflow, 4);                    // a body with more than 62 groups: beyond the colour mask
    L.gcolor[g] = c;
    atomicSub(&L.counters[LWC_UNCOLORED], 1);
    atomicMax(&L.counters[LWC_NCOLORS], c + 1);
    atomicAdd(&L.ccount[c], 1);
}
// groups by colour: list offsets (one thread), then the lists (order inside a colour is irrelevant: disjoint bodies)
__global__ void k_lwc_color_scan(const __grid_constant__ LargePtrs L)
{
    int o = 0;
    for (int c = 0; c < 64; c++) { L.cofs[c] = o; o += L.ccount[c]; L.ccount[c] = 0; }
    L.cofs[64] = o;
}
__global__ void k_lwc_color_fill(const __grid_constant__ DevPtrs D, const __grid_constant__ LargePtrs L)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= L.counters[LWC_NGROUPS]) return;
    const int g = L.heads[t];
    const int c = L.gcolor[g];
    if (c < 0) return;
    const int pos = L.cofs[c] + atomicAdd(&L.ccount[c], 1);
    L.clist[pos] = g;
    const int2 rb = D.rbody[g];
    L.cinfo[pos] = make_int4(g, L.gsize[g], rb.x, rb.y);
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
__device__ __forceinline__ Real4 shfl_up4(const Real4 &v)
{
    Real4 r;
    r.x = __shfl_up_sync(0xffffffffu, v.x, 1); r.y = __shfl_up_sync(0xffffffffu, v.y, 1);
    r.z = __shfl_up_sync(0xffffffffu, v.z, 1); r.w = __shfl_up_sync(0xffffffffu, v.w, 1);
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

// Chunks: the groups of a colour are packed into chunks of 32 row slots (a group never straddles two chunks), so that a warp can give
// every row of a chunk its own lane.  One warp packs 32 consecutive groups of the colour's list greedily and reserves its chunks with one
// atomic on the global chunk cursor; the colours are packed one launch after the other, so the chunks of a colour are consecutive from
// cstart[colour] (k_lwc_chunk_start).  Next-fit leaves every chunk but a warp's last more than half full: 3 MR slots always suffice.
// Which group lands in which chunk depends on the list order (atomics), which is irrelevant: groups of a colour touch disjoint bodies.
__global__ void k_lwc_chunk_start(const __grid_constant__ LargePtrs L, int colour) { L.cstart[colour] = L.counters[LWC_CHUNKS]; }
__global__ void __launch_bounds__(256) k_lwc_chunks(const __grid_constant__ LargePtrs L, int colour)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
    const int lo = L.cofs[colour], n = L.cofs[colour + 1] - lo;
    if ((t & ~31) >= n) return;
    const int g = t < n ? L.clist[lo + t] : -1;
    const int sz = g >= 0 ? L.gsize[g] : 0;
    if (sz > 32) atomicAdd(&L.counters[LWC_BIGGROUPS], 1);
    int chunk = 0, fill = 0, my_chunk = 0, my_start = 0;
    for (int k = 0; k < 32; k++) {                                 // every lane replays the same greedy packing
        const int sk = __shfl_sync(0xffffffffu, sz, k);
        if (sk == 0 || sk > 32) continue;
        if (fill + sk > 32) { chunk++; fill = 0; }
        if (k == lane) { my_chunk = chunk; my_start = fill; }
        fill += sk;
    }
    const int used = chunk + (fill > 0 ? 1 : 0);
    int base = 0;
    if (lane == 0 && used > 0) base = atomicAdd(&L.counters[LWC_CHUNKS], used);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (g >= 0 && sz <= 32) {
        int *slot = L.slot_row + ((size_t)(base + my_chunk) * 32 + my_start);
        for (int i = 0; i < sz; i++) slot[i] = (g + i) | (i == 0 ? 0x40000000 : 0) | (i == sz - 1 ? (int)0x80000000u : 0);   // bit 30: first, bit 31: last row of its group
    }
}

// One colour of one sweep, lane per row: a warp takes one chunk, every lane loads its row's record (all rows of a colour are in flight
// at once: the sweep streams the rows from HBM at bandwidth instead of one dependent row after the other per thread); the rows of a group
// then execute in order, one per step, the two bodies' accumulators travelling from lane to lane by shuffle and a contact's normal-row
// lambda forwarded to its friction rows; the last row of the group writes the accumulators back.
__global__ void __launch_bounds__(128) k_lwc_sweep_rows(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D, const __grid_constant__ LargePtrs L, int colour)
{
    const int chunk = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    const int first = L.cstart[colour];
    if (chunk >= L.cstart[colour + 1] - first) return;
    const int slot = L.slot_row[(size_t)(first + chunk) * 32 + lane];
    const bool any = slot != -1;
    const int r = any ? (slot & 0x3fffffff) : 0;
    const bool head = any && (slot & 0x40000000), tail = any && (slot < 0);
    Real4 *cf = D.cforce;
    Real *lam = D.lambda;
    // The 32 row records of the chunk (128 B each, scattered group by group) are fetched COALESCED: in each of 8 passes the warp reads
    // 4 whole records (8 lanes x 16 B per record) and parks them in shared memory with a padded row stride, then every lane reads its own
    // record from there (both sides free of bank conflicts: stride = 9 x 16 B).  A lane loading its own record directly touches 32
    // different lines per instruction: measured 42-50 us per colour instead of ~10.
    constexpr int RS = (int)sizeof(Real4) * 9;                                  // padded record stride in shared memory (bytes)
    extern __shared__ __align__(16) unsigned char lwc_smem[];
    unsigned char *wbuf = lwc_smem + (size_t)(threadIdx.x >> 5) * 32 * RS;
    {
        Real4 v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int rs = 4 * k + (lane >> 3);
            const int rr = __shfl_sync(0xffffffffu, r, rs);
            v[k] = ldcg4(D.rows + (size_t)rr * 8 + (lane & 7));
        }
#pragma unroll
        for (int k = 0; k < 8; k++) *(Real4 *)(wbuf + (size_t)(4 * k + (lane >> 3)) * RS + (lane & 7) * sizeof(Real4)) = v[k];
    }
    __syncwarp();
    const Real4 *rec = (const Real4 *)(wbuf + (size_t)lane * RS);
    const Real4 a0 = rec[0], a1 = rec[1], a2 = rec[2], a3 = rec[3], b0 = rec[4], b1q = rec[5], b2q = rec[6], b3 = rec[7];
    const int2 rb = any ? D.rbody[r] : make_int2(0, P.NB);
    const int fi = any ? D.findex[r] : -1;
    const Real old_lambda = any ? lam[r] : R_(0.0);
    const int isl = any ? L.row_island[r] : 0;
    const bool two = rb.y != P.NB;
    Real4 f1a = { 0, 0, 0, 0 }, f1b = f1a, f2a = f1a, f2b = f1a;
    if (head) {
        f1a = ldcg4(&cf[2 * rb.x]); f1b = ldcg4(&cf[2 * rb.x + 1]);
        if (two) { f2a = ldcg4(&cf[2 * rb.y]); f2b = ldcg4(&cf[2 * rb.y + 1]); }
    }
    const bool live = any && !L.isl_done[isl];
    const unsigned heads = __ballot_sync(0xffffffffu, head);
    const int headlane = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));       // first lane of this lane's run
    // the friction-index row is an earlier row of the same group, i.e. an earlier lane of the same run: its new lambda is forwarded
    const int fwd_src = (live && fi != -1) ? lane - (r - fi) : lane;
    const bool fwd = live && fi != -1 && fwd_src >= headlane && fwd_src < lane;
    Real my_lambda = 0;
    bool have = live && head;
    for (;;) {
        const bool exec = have;
        const Real lam_fw = __shfl_sync(0xffffffffu, my_lambda, fwd ? fwd_src : lane);
        if (exec) {
            Real delta = a1.z - old_lambda * a1.w;
            delta -= f1a.x * a0.x + f1a.y * a0.y + f1a.z * a0.z + f1a.w * a0.w + f1b.x * a1.x + f1b.y * a1.y;
            if (two) delta -= f2a.x * b0.x + f2a.y * b0.y + f2a.z * b0.z + f2a.w * b0.w + f2b.x * b1q.x + f2b.y * b1q.y;
            Real hi_act, lo_act;
            if (fi != -1) { hi_act = RFABS(b1q.w * (fwd ? lam_fw : lam[fi])); lo_act = -hi_act; }
            else { hi_act = b1q.w; lo_act = b1q.z; }
            Real new_lambda = old_lambda + delta;
            if (new_lambda < lo_act) { delta = lo_act - old_lambda; new_lambda = lo_act; }
            else if (new_lambda > hi_act) { delta = hi_act - old_lambda; new_lambda = hi_act; }
            lam[r] = new_lambda;
            my_lambda = new_lambda;
            if (delta != 0) {
                f1a.x += delta * a2.x; f1a.y += delta * a2.y; f1a.z += delta * a2.z; f1a.w += delta * a2.w;
                f1b.x += delta * a3.x; f1b.y += delta * a3.y;
                if (delta > 0) f1b.w += delta * a3.z; else f1b.z += delta * a3.z;
                if (two) {
                    if (delta > 0) f2b.w += delta * b3.z; else f2b.z += delta * b3.z;
                    f2a.x += delta * b2q.x; f2a.y += delta * b2q.y; f2a.z += delta * b2q.z; f2a.w += delta * b2q.w;
                    f2b.x += delta * b3.x; f2b.y += delta * b3.y;
                }
            }
            if (tail) {
                stcg4(&cf[2 * rb.x], f1a); stcg4(&cf[2 * rb.x + 1], f1b);
                if (two) { stcg4(&cf[2 * rb.y], f2a); stcg4(&cf[2 * rb.y + 1], f2b); }
            }
        }
        const unsigned pass = __ballot_sync(0xffffffffu, exec && !tail);
        if (pass == 0) break;
        const Real4 n1a = shfl_up4(f1a), n1b = shfl_up4(f1b), n2a = shfl_up4(f2a), n2b = shfl_up4(f2b);
        have = lane > 0 && ((pass >> (lane - 1)) & 1u);
        if (have) { f1a = n1a; f1b = n1b; f2a = n2a; f2b = n2b; }
    }
}

// One colour of one sweep, thread per group with the records staged through shared memory.  ncu on the two simpler kernels says why:
// a thread that walks its group's rows straight from HBM pays one memory latency per row (34 us per colour, 25 k threads in flight);
// a lane per row fetches everything at once but then passes the accumulators from lane to lane, 12 steps of ~130 instructions with a
// third of the lanes busy (issue-bound, 36 M warp instructions per colour, 36-50 us).  Here a warp takes 32 consecutive groups of the
// colour, fetches ALL their records coalesced (8 lanes x 16 B per record, 4 records per instruction, every request in flight at once)
// into shared memory with a padded record stride, and then every lane relaxes its own group from shared memory: full lanes, one
// exposed memory latency per warp.  Groups whose records do not fit the warp's buffer (LWC_ROWS_PER_WARP) are relaxed straight from HBM.
#if defined(ODEB_DOUBLE)
#define LWC_ROWS_PER_WARP 160
#else
#define LWC_ROWS_PER_WARP 288
#endif
#define LWC_WARPS 4
__global__ void __launch_bounds__(32 * LWC_WARPS) k_lwc_sweep_staged(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D, const __grid_constant__ LargePtrs L, int colour)
{
    constexpr int RS = (int)sizeof(Real4) * 9;                                  // padded record stride (bytes): conflict-free on both sides
    extern __shared__ __align__(16) unsigned char lwc_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    unsigned char *wbuf = lwc_smem + (size_t)wib * (LWC_ROWS_PER_WARP * RS + LWC_ROWS_PER_WARP * sizeof(int));
    int *srow = (int *)(wbuf + (size_t)LWC_ROWS_PER_WARP * RS);
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int lo = L.cofs[colour], n = L.cofs[colour + 1] - lo;
    if ((t & ~31) >= n) return;
    const int g = t < n ? L.clist[lo + t] : -1;
    const int sz = g >= 0 ? L.gsize[g] : 0;
    // slots of the warp's buffer: exclusive prefix of the group sizes; groups beyond the buffer stay in HBM
    int ofs = sz;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int u = __shfl_up_sync(0xffffffffu, ofs, d); if (lane >= d) ofs += u; }
    const int total = __shfl_sync(0xffffffffu, ofs, 31);
    ofs -= sz;
    const bool staged = g >= 0 && ofs + sz <= LWC_ROWS_PER_WARP;
    const int nst = total < LWC_ROWS_PER_WARP ? total : LWC_ROWS_PER_WARP;
    if (staged) for (int i = 0; i < sz; i++) srow[ofs + i] = g + i;
    else if (g >= 0) for (int i = 0; ofs + i < LWC_ROWS_PER_WARP && i < sz; i++) srow[ofs + i] = g + i;      // a straddling group: harmless filler
    __syncwarp();
    for (int s0 = 0; s0 < nst; s0 += 16) {                        // 4 passes of 4 records per trip: 4 requests in flight per lane
        Real4 v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) { const int s = s0 + 4 * k + (lane >> 3); if (s < nst) v[k] = ldcg4(D.rows + (size_t)srow[s] * 8 + (lane & 7)); }
#pragma unroll
        for (int k = 0; k < 4; k++) { const int s = s0 + 4 * k + (lane >> 3); if (s < nst) *(Real4 *)(wbuf + (size_t)s * RS + (lane & 7) * sizeof(Real4)) = v[k]; }
    }
    if (g < 0) return;                                            // (no warp-level primitive below)
    const bool skip = L.isl_done[L.row_island[g]] != 0;
    const int2 rb = D.rbody[g];
    const bool two = rb.y != P.NB;
    Real4 *cf = D.cforce;
    Real *lam = D.lambda;
    Real4 f1a = ldcg4(&cf[2 * rb.x]), f1b = ldcg4(&cf[2 * rb.x + 1]);
    Real4 f2a = { 0, 0, 0, 0 }, f2b = f2a;
    if (two) { f2a = ldcg4(&cf[2 * rb.y]); f2b = ldcg4(&cf[2 * rb.y + 1]); }
    __syncwarp();                                                 // the records are in shared memory
    if (skip) return;
    int free_row = -1; Real free_lambda = 0;                      // the latest row of the group without a friction index and its new lambda
    for (int k = 0; k < sz; k++) {
        const int r = g + k;
        const Real4 *rec = staged ? (const Real4 *)(wbuf + (size_t)(ofs + k) * RS) : D.rows + (size_t)r * 8;
        const Real4 a0 = rec[0], a1 = rec[1], a2 = rec[2], a3 = rec[3], b0 = rec[4], b1q = rec[5], b2q = rec[6], b3 = rec[7];
        const Real old_lambda = lam[r];
        const int fi = D.findex[r];
        Real delta = a1.z - old_lambda * a1.w;
        delta -= f1a.x * a0.x + f1a.y * a0.y + f1a.z * a0.z + f1a.w * a0.w + f1b.x * a1.x + f1b.y * a1.y;
        if (two) delta -= f2a.x * b0.x + f2a.y * b0.y + f2a.z * b0.z + f2a.w * b0.w + f2b.x * b1q.x + f2b.y * b1q.y;
        Real hi_act, lo_act;
        if (fi != -1) { hi_act = RFABS(b1q.w * (fi == free_row ? free_lambda : lam[fi])); lo_act = -hi_act; }
        else { hi_act = b1q.w; lo_act = b1q.z; }
        Real new_lambda = old_lambda + delta;
        if (new_lambda < lo_act) { delta = lo_act - old_lambda; new_lambda = lo_act; }
        else if (new_lambda > hi_act) { delta = hi_act - old_lambda; new_lambda = hi_act; }
        lam[r] = new_lambda;
        if (fi == -1) { free_row = r; free_lambda = new_lambda; }
        if (delta != 0) {
            f1a.x += delta * a2.x; f1a.y += delta * a2.y; f1a.z += delta * a2.z; f1a.w += delta * a2.w;
            f1b.x += delta * a3.x; f1b.y += delta * a3.y;
            if (delta > 0) f1b.w += delta * a3.z; else f1b.z += delta * a3.z;
            if (two) {
                if (delta > 0) f2b.w += delta * b3.z; else f2b.z += delta * b3.z;
                f2a.x += delta * b2q.x; f2a.y += delta * b2q.y; f2a.z += delta * b2q.z; f2a.w += delta * b2q.w;
                f2b.x += delta * b3.x; f2b.y += delta * b3.y;
            }
        }
    }
    stcg4(&cf[2 * rb.x], f1a); stcg4(&cf[2 * rb.x + 1], f1b);
    if (two) { stcg4(&cf[2 * rb.y], f2a); stcg4(&cf[2 * rb.y + 1], f2b); }
}

// One colour of one sweep, thread per group, the records brought in by TMA bulk copies.  The records of a group are one contiguous
// block of HBM (gsize x 32 reals), so every lane issues ONE cp.async.bulk for its whole group into the warp's shared-memory pool,
// all of them complete on the warp's mbarrier (lane 0 armed it with the byte count), and after one wait every lane relaxes its own
// group from shared memory.  All records of the colour are in flight at once and no register or LSU instruction is spent per 16 bytes:
// what made the thread-per-group kernel slow was one DRAM latency per row (ncu: stall_long_sb 86 %, issue active 7 %), what made the
// lane-per-row kernel slow was passing accumulators between lanes (issue-bound).  Group j of the warp starts at row offset ofs_j in
// the pool, shifted by j x 16 bytes: the eight lanes of a shared-memory phase then read eight different bank groups (records are 8 or
// 16 x 16 bytes).  Groups that do not fit the pool read their records from HBM.
__device__ __forceinline__ void lwc_mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void lwc_mbar_expect_tx(unsigned bar, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void lwc_bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ bool lwc_mbar_try_wait(unsigned bar, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
#define LWC_TMA_WARPS 2
#define LWC_AUX_ROWS 24                                     // rows per group whose lambda / findex ranges fit the per-lane slots
#define LWC_AUX_BYTES ((LWC_AUX_ROWS + 8) * (int)sizeof(Real))
__global__ void __launch_bounds__(32 * LWC_TMA_WARPS) k_lwc_sweep_tma(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D, const __grid_constant__ LargePtrs L, int colour)
{
    constexpr int REC = (int)sizeof(Real4) * 8;                                 // bytes of a record
    constexpr int POOL = LWC_ROWS_PER_WARP * REC + 32 * 16 + 32 * 2 * LWC_AUX_BYTES;   // pool of a warp: records + the per-group 16-byte shifts + lambda / findex ranges
    extern __shared__ __align__(128) unsigned char lwc_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    unsigned char *pool = lwc_smem + (size_t)wib * POOL;
    unsigned char *aux = pool + LWC_ROWS_PER_WARP * REC + 32 * 16;
    unsigned long long *bars = (unsigned long long *)(lwc_smem + (size_t)LWC_TMA_WARPS * POOL);
    const unsigned bar = (unsigned)__cvta_generic_to_shared(bars + wib);
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int lo = L.cofs[colour], n = L.cofs[colour + 1] - lo;
    if ((t & ~31) >= n) return;
    if (lane == 0) lwc_mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    const int4 gi = t < n ? L.cinfo[lo + t] : make_int4(-1, 0, 0, P.NB);
    const int g = gi.x, sz = gi.y;
    int ofs = sz;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int u = __shfl_up_sync(0xffffffffu, ofs, d); if (lane >= d) ofs += u; }
    ofs -= sz;
    const bool staged = g >= 0 && ofs + sz <= LWC_ROWS_PER_WARP && sz <= LWC_AUX_ROWS;
    // lambda and the friction indices of the group's rows come along as two more bulk copies (the 16-byte aligned ranges that hold them):
    // a load of lambda inside the row loop would put one memory latency on every row
    const int a0i = g & ~3, a1i = (g + sz + 3) & ~3;              // aligned element range [a0i, a1i) of both arrays
    const unsigned abytes = staged ? (unsigned)(a1i - a0i) * 4u : 0u;
    const unsigned bytes = staged ? (unsigned)sz * REC : 0u;
    const unsigned lbytes = staged ? (unsigned)(a1i - a0i) * (unsigned)sizeof(Real) : 0u;
    const unsigned total = __reduce_add_sync(0xffffffffu, bytes + lbytes + abytes);
    unsigned char *mine = pool + (size_t)ofs * REC + lane * 16;
    unsigned char *auxl = aux + (size_t)lane * (2 * LWC_AUX_BYTES);          // [lambda range | findex range] of this lane's group
    if (lane == 0) lwc_mbar_expect_tx(bar, total);
    __syncwarp();
    if (staged) {
        lwc_bulk_g2s((unsigned)__cvta_generic_to_shared(mine), D.rows + (size_t)g * 8, bytes, bar);
        lwc_bulk_g2s((unsigned)__cvta_generic_to_shared(auxl), D.lambda + a0i, lbytes, bar);
        lwc_bulk_g2s((unsigned)__cvta_generic_to_shared(auxl + LWC_AUX_BYTES), D.findex + a0i, abytes, bar);
    }
    // what does not come through the pool is requested meanwhile
    bool skip = true;
    const int2 rb = make_int2(gi.z, gi.w);
    Real4 f1a = { 0, 0, 0, 0 }, f1b = f1a, f2a = f1a, f2b = f1a;
    Real4 *cf = D.cforce;
    Real *lam = D.lambda;
    if (g >= 0) {
        f1a = ldcg4(&cf[2 * rb.x]); f1b = ldcg4(&cf[2 * rb.x + 1]);
        if (rb.y != P.NB) { f2a = ldcg4(&cf[2 * rb.y]); f2b = ldcg4(&cf[2 * rb.y + 1]); }
        skip = L.isl_done[L.row_island[g]] != 0;
    }
    const bool two = rb.y != P.NB;
    const Real *slam = (const Real *)auxl + (g - a0i);
    const int *sfi = (const int *)(auxl + LWC_AUX_BYTES) + (g - a0i);
    while (!lwc_mbar_try_wait(bar, 0)) { }
    if (skip) return;
    int free_row = -1; Real free_lambda = 0;                      // the latest row of the group without a friction index and its new lambda
    for (int k = 0; k < sz; k++) {
        const int r = g + k;
        const Real4 *rec = staged ? (const Real4 *)(mine + (size_t)k * REC) : D.rows + (size_t)r * 8;
        const Real4 a0 = rec[0], a1 = rec[1], a2 = rec[2], a3 = rec[3], b0 = rec[4], b1q = rec[5], b2q = rec[6], b3 = rec[7];
        const Real old_lambda = staged ? slam[k] : lam[r];
        const int fi = staged ? sfi[k] : D.findex[r];
        Real delta = a1.z - old_lambda * a1.w;
        delta -= f1a.x * a0.x + f1a.y * a0.y + f1a.z * a0.z + f1a.w * a0.w + f1b.x * a1.x + f1b.y * a1.y;
        if (two) delta -= f2a.x * b0.x + f2a.y * b0.y + f2a.z * b0.z + f2a.w * b0.w + f2b.x * b1q.x + f2b.y * b1q.y;
        Real hi_act, lo_act;
        if (fi != -1) { hi_act = RFABS(b1q.w * (fi == free_row ? free_lambda : lam[fi])); lo_act = -hi_act; }
        else { hi_act = b1q.w; lo_act = b1q.z; }
        Real new_lambda = old_lambda + delta;
        if (new_lambda < lo_act) { delta = lo_act - old_lambda; new_lambda = lo_act; }
        else if (new_lambda > hi_act) { delta = hi_act - old_lambda; new_lambda = hi_act; }
        lam[r] = new_lambda;
        if (fi == -1) { free_row = r; free_lambda = new_lambda; }
        if (delta != 0) {
            f1a.x += delta * a2.x; f1a.y += delta * a2.y; f1a.z += delta * a2.z; f1a.w += delta * a2.w;
            f1b.x += delta * a3.x; f1b.y += delta * a3.y;
            if (delta > 0) f1b.w += delta * a3.z; else f1b.z += delta * a3.z;
            if (two) {
                if (delta > 0) f2b.w += delta * b3.z; else f2b.z += delta * b3.z;
                f2a.x += delta * b2q.x; f2a.y += delta * b2q.y; f2a.z += delta * b2q.z; f2a.w += delta * b2q.w;
                f2b.x += delta * b3.x; f2b.y += delta * b3.y;
            }
        }
    }
    stcg4(&cf[2 * rb.x], f1a); stcg4(&cf[2 * rb.x + 1], f1b);
    if (two) { stcg4(&cf[2 * rb.y], f2a); stcg4(&cf[2 * rb.y + 1], f2b); }
}

// One colour of one sweep: a thread relaxes the rows of one group in row order, the two bodies' accumulators (and the most recent
// normal-row lambda, which the contact's friction rows clamp against) in registers; the next row's record is requested while the current
// one is computed.  No other group of the colour touches these bodies.  Arithmetic = Stage4LCP_IterationStep quickstep.cpp:2917-3033.
#ifndef LWC_BLK
#define LWC_BLK 2
#endif
__global__ void __launch_bounds__(128) k_lwc_sweep(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D, const __grid_constant__ LargePtrs L, int colour)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int lo = L.cofs[colour];
    if (t >= L.cofs[colour + 1] - lo) return;
    const int4 gi = L.cinfo[lo + t];                              // (first row, rows, accumulator slots of the two bodies): one load
    const int g = gi.x, n = gi.y;
    const int2 rb = make_int2(gi.z, gi.w);
    const bool two = rb.y != P.NB;
    Real4 *cf = D.cforce;
    Real *lam = D.lambda;
    // A launch holds only as many threads as the colour has groups (~45 k on the 100k-box wall: 10 warps per SM), so registers are free
    // and latency is everything: the records are fetched LWC_BLK rows at a time (8 independent 16-byte loads per row in flight, together
    // with their lambdas and friction indices) -- one exposed memory latency per LWC_BLK rows instead of one per row.
    const Real4 *rec = D.rows + (size_t)g * 8;
    Real4 q[LWC_BLK][8]; Real ol[LWC_BLK]; int fx[LWC_BLK];
#pragma unroll
    for (int j = 0; j < LWC_BLK; j++) if (j < n) {
#pragma unroll
        for (int c = 0; c < 8; c++) q[j][c] = ldcg4(rec + 8 * j + c);
        ol[j] = lam[g + j]; fx[j] = D.findex[g + j];
    }
    Real4 f1a = ldcg4(&cf[2 * rb.x]), f1b = ldcg4(&cf[2 * rb.x + 1]);
    Real4 f2a = { 0, 0, 0, 0 }, f2b = f2a;
    if (two) { f2a = ldcg4(&cf[2 * rb.y]); f2b = ldcg4(&cf[2 * rb.y + 1]); }
    if (L.isl_done[L.row_island[g]]) return;
    int free_row = -1; Real free_lambda = 0;                   // the latest row of the group without a friction index and its new lambda
    for (int k0 = 0; k0 < n; k0 += LWC_BLK) {
#pragma unroll
        for (int j = 0; j < LWC_BLK; j++) if (k0 + j < n) {
            const int r = g + k0 + j;
            const Real4 a0 = q[j][0], a1 = q[j][1], a2 = q[j][2], a3 = q[j][3], b0 = q[j][4], b1q = q[j][5], b2q = q[j][6], b3 = q[j][7];
            const Real old_lambda = ol[j];
            const int fi = fx[j];
            Real delta = a1.z - old_lambda * a1.w;
            delta -= f1a.x * a0.x + f1a.y * a0.y + f1a.z * a0.z + f1a.w * a0.w + f1b.x * a1.x + f1b.y * a1.y;
            if (two) delta -= f2a.x * b0.x + f2a.y * b0.y + f2a.z * b0.z + f2a.w * b0.w + f2b.x * b1q.x + f2b.y * b1q.y;
            Real hi_act, lo_act;
            if (fi != -1) { hi_act = RFABS(b1q.w * (fi == free_row ? free_lambda : lam[fi])); lo_act = -hi_act; }
            else { hi_act = b1q.w; lo_act = b1q.z; }
            Real new_lambda = old_lambda + delta;
            if (new_lambda < lo_act) { delta = lo_act - old_lambda; new_lambda = lo_act; }
            else if (new_lambda > hi_act) { delta = hi_act - old_lambda; new_lambda = hi_act; }
            lam[r] = new_lambda;
            if (fi == -1) { free_row = r; free_lambda = new_lambda; }
            if (delta != 0) {
                f1a.x += delta * a2.x; f1a.y += delta * a2.y; f1a.z += delta * a2.z; f1a.w += delta * a2.w;
                f1b.x += delta * a3.x; f1b.y += delta * a3.y;
                if (delta > 0) f1b.w += delta * a3.z; else f1b.z += delta * a3.z;
                if (two) {
                    if (delta > 0) f2b.w += delta * b3.z; else f2b.z += delta * b3.z;
                    f2a.x += delta * b2q.x; f2a.y += delta * b2q.y; f2a.z += delta * b2q.z; f2a.w += delta * b2q.w;
                    f2b.x += delta * b3.x; f2b.y += delta * b3.y;
                }
            }
        }
        if (k0 + LWC_BLK < n) {
            const Real4 *nr = rec + 8 * (size_t)(k0 + LWC_BLK);
#pragma unroll
            for (int j = 0; j < LWC_BLK; j++) if (k0 + LWC_BLK + j < n) {
#pragma unroll
                for (int c = 0; c < 8; c++) q[j][c] = ldcg4(nr + 8 * j + c);
                ol[j] = lam[g + k0 + LWC_BLK + j]; fx[j] = D.findex[g + k0 + LWC_BLK + j];
            }
        }
    }
    stcg4(&cf[2 * rb.x], f1a); stcg4(&cf[2 * rb.x + 1], f1b);
    if (two) { stcg4(&cf[2 * rb.y], f2a); stcg4(&cf[2 * rb.y + 1], f2b); }
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
