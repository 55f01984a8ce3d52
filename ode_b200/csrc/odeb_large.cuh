// odeb_large.cuh -- the large-world path (ODEB_MODE_CANONICAL, one world of 10^3..10^5+ bodies).
//
// The batched path gives each world to one warp lane pair; a single big world needs parallelism INSIDE the
// world instead.  Every serial list walk of the reference is replaced by a sort / scan / union-find that yields
// the same sets, and the order-sensitive SOR sweep keeps its sequential semantics through per-body tickets:
//
//   broadphase   k_bp_keys + radix sort on (float)aabb.min[0] + k_bp_sweep / k_bp_big + radix sort of the pair keys
//                -> the same pair set as k_pair_pass (collision_sapspace.cpp:521-582 BoxPruning is itself a sort + sweep;
//                   the hash space reports the same set as an exhaustive AABB test), in (geomA<geomB) order
//   contacts     scan of the per-pair contact counts -> creation-order numbering (ode.cpp:1192-1200)
//   auto-disable k_lw_autodisable, thread per body (util.cpp:427-561)
//   islands      lock-free union-find over all joints (util.cpp:724-860 finds the same components with a DFS);
//                island number = rank of the component's highest enabled body, descending (world->firstbody order)
//   order        bodies (island, descending index), joints (island, ascending id) by radix sort; row offsets by scan
//   solve        per phase of 8 sweeps: row order = radix sort of (island, class|hash key); per body the rows that
//                touch it are ranked by their position ("tickets"); k_lw_sweep walks the order with all SMs, a row
//                runs when both of its bodies' counters have reached its tickets, i.e. after exactly the rows that
//                precede it on those bodies in the sequential sweep (quickstep.cpp:2917-3033) -> bit-identical to the
//                sequential sweep in that order, no level barriers
//   control      k_lw_body_check + k_lw_island_ctl after every sweep (quickstep.cpp:1823-1856, :3253-3285), per island
#ifndef ODEB_LARGE_CUH
#define ODEB_LARGE_CUH
#include <cub/cub.cuh>

typedef unsigned long long u64;
#define LW_NOKEY 0xFFFFFFFFFFFFFFFFull

enum { LWC_NBIG = 0, LWC_NPAIRS = 1, LWC_NCONTACTS = 2, LWC_NORDERED = 3, LWC_NJORD = 4, LWC_MROWS = 5, LWC_NISLANDS = 6,
       LWC_NACTIVE = 7, LWC_CURSOR = 8, LWC_COUNT = 16 };

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
    u64 *okey, *okey_s; int *oval, *ord, *rpos;  // [MR]
    int *inc_ofs, *inc_cur, *inc, *inc_pos;      // [NB + 2], [NB + 1], [2 MR], [2 MR]
    int2 *ticket;                                // [MR] at the head position of a run: the run's ticket on body 1, on body 2
    int *row_group, *head_pos;                   // [MR] first row of the row's group; first position of the position's run
    unsigned *cnt;                               // [NB + 1] runs executed on each body in the running sweep
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
    L.isl_nb[b] = 0; L.isl_m[b] = 0; L.isl_done[b] = 0; L.isl_viol[b] = 0; L.cnt[b] = 0; L.inc_cur[b] = 0;
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

// ------------------------------------------------------------------------------------------------ solve: order + tickets
__global__ void k_lw_inc_count(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D, const __grid_constant__ LargePtrs L, int mrows)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= mrows) return;
    int2 rb = D.rbody[r];
    atomicAdd(&L.inc_cur[rb.x], 1);
    if (rb.y != P.NB) atomicAdd(&L.inc_cur[rb.y], 1);
    D.lambda[r] = 0;
}
__global__ void k_lw_inc_fill(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D, const __grid_constant__ LargePtrs L, int mrows)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= mrows) return;
    int2 rb = D.rbody[r];
    L.inc[L.inc_ofs[rb.x] + atomicAdd(&L.inc_cur[rb.x], 1)] = 2 * r;
    if (rb.y != P.NB) L.inc[L.inc_ofs[rb.y] + atomicAdd(&L.inc_cur[rb.y], 1)] = 2 * r + 1;
}
__global__ void k_lw_zero_cur(int n, int *v) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) v[i] = 0; }

// Row order of one phase.  Rows are grouped: the rows of the contacts of one geom pair form a group (so do the rows of one
// permanent joint); a group's key is odeb_canon_key(seed, island, phase, island-local index of the group's first row).
// Phase 0 keeps ReorderPrep's two classes (rows without a friction index first, quickstep.cpp:2329-2355) and sorts each
// class by group key; phase k >= 1 (the reorder at sweep 8k) sorts all rows by group key. Ties: ascending row index, so
// the rows of a group stay adjacent and in joint order.
__global__ void k_lw_order_keys(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D, const __grid_constant__ LargePtrs L, int mrows, int phase)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= mrows) return;
    const unsigned is = (unsigned)L.row_island[r];
    const unsigned glocal = (unsigned)(L.row_group[r] - L.isl_rstart[is]);
    u64 low = odebi_canon_key(D.seed[0], is, (unsigned)phase, glocal);
    if (phase == 0 && D.findex[r] != -1) low |= 1ull << 32;
    L.okey[r] = ((u64)is << 33) | low;
    L.oval[r] = r;
}
// Runs: maximal stretches of consecutive positions inside one 32-position chunk that belong to one group (and, in phase 0,
// one class). All rows of a run act on the same two bodies, so a run is the unit of scheduling: head_pos[p] = position
// of the first row of p's run. Also the inverse permutation rpos.
__global__ void k_lw_runs(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D, const __grid_constant__ LargePtrs L, int mrows, int phase)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;          // blockDim is a multiple of 32: warps are chunk-aligned
    const int lane = threadIdx.x & 31;
    int key = -1;
    if (p < mrows) {
        const int r = L.ord[p];
        L.rpos[r] = p;
        key = 2 * L.row_group[r] + ((phase == 0 && D.findex[r] != -1) ? 1 : 0);
    }
    const int prev = __shfl_up_sync(0xffffffffu, key, 1);
    const bool start = lane == 0 || key != prev;
    const unsigned heads = __ballot_sync(0xffffffffu, start);
    if (p < mrows) L.head_pos[p] = p - lane + (31 - __clz(heads & (0xffffffffu >> (31 - lane))));
}
// tickets: rank of every run among the runs acting on the same body, by position in the sweep order
__global__ void k_lw_tickets(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D, const __grid_constant__ LargePtrs L)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= L.counters[LWC_NORDERED]) return;
    const int lo = L.inc_ofs[k], hi = L.inc_ofs[k + 1];
    for (int e = lo; e < hi; e++) {
        const int pe = L.rpos[L.inc[e] >> 1], hp = L.head_pos[pe];
        L.inc_pos[e] = 2 * hp + (pe == hp ? 1 : 0);
    }
    for (int e = lo; e < hi; e++) {
        const int ve = L.inc_pos[e];
        if (!(ve & 1)) continue;
        int rank = 0;
        for (int f = lo; f < hi; f++) { const int vf = L.inc_pos[f]; rank += ((vf & 1) && vf < ve) ? 1 : 0; }
        int *t = (int *)&L.ticket[ve >> 1];
        t[L.inc[e] & 1] = rank;
    }
}

// ------------------------------------------------------------------------------------------------ solve: sweep
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned *p, unsigned v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_u32(unsigned *p, unsigned v)
{
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
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

__device__ __forceinline__ Real4 shfl_up4(const Real4 &v)
{
    Real4 r;
    r.x = __shfl_up_sync(0xffffffffu, v.x, 1); r.y = __shfl_up_sync(0xffffffffu, v.y, 1);
    r.z = __shfl_up_sync(0xffffffffu, v.z, 1); r.w = __shfl_up_sync(0xffffffffu, v.w, 1);
    return r;
}

#ifndef ODEB_LW_BLOCKS
#define ODEB_LW_BLOCKS 3             // resident blocks per SM (80 registers): 27.6 ms per step on the 100k-box wall, 2 -> 30.6, 4 (spills) -> 43.2
#endif
// One sweep over all rows of all unfinished islands, in the phase's order. Warps claim consecutive chunks of 32 positions
// through one cursor, so every claimed position only ever waits for positions that are already claimed by a running warp
// (or finished): the walk cannot deadlock whatever the grid size. Inside a chunk a lane owns one row (its record is in
// registers before any waiting starts). The head lane of a run polls the two bodies' counters; when both have reached
// the run's tickets it loads the bodies' accumulators, and the run then executes in lockstep, one row per step, the
// accumulators travelling from lane to lane by shuffle; the tail lane writes them back and bumps the counters.
// Arithmetic = Stage4LCP_IterationStep quickstep.cpp:2917-3033.
__global__ void __launch_bounds__(256, ODEB_LW_BLOCKS) k_lw_sweep(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D, const __grid_constant__ LargePtrs L, int mrows)
{
    if (L.counters[LWC_NACTIVE] == 0) return;
    const int lane = threadIdx.x & 31;
    const int nchunks = (mrows + 31) >> 5;
    Real4 *cf = D.cforce;
    Real *lam = D.lambda;
#ifndef ODEB_LW_CLAIM
#define ODEB_LW_CLAIM 1             // chunks claimed per cursor atomic. Must stay 1: a warp that holds chunks it is not yet working on makes
                                   // every later position wait for it (measured: 4 -> 9.8 s per step instead of 38 ms)
#endif
    for (int sub = ODEB_LW_CLAIM, base = 0;; sub++) {
        if (sub == ODEB_LW_CLAIM) {
            if (lane == 0) base = atomicAdd(&L.counters[LWC_CURSOR], ODEB_LW_CLAIM);
            base = __shfl_sync(0xffffffffu, base, 0);
            sub = 0;
        }
        const int chunk = base + sub;
        if (chunk >= nchunks) break;
        const int p = chunk * 32 + lane;
        bool pending = p < mrows;
        int r = 0, fi = -1; int2 rb = make_int2(0, 0), tk = make_int2(0, 0);
        bool head = false, tail = false;
        int hl = lane;                           // lane of the head of this lane's run
        Real old_lambda = 0;
        Real4 a0, a1, a2, a3, b0, b1q, b2q, b3;
        if (pending) {
            r = L.ord[p];
            if (L.isl_done[L.row_island[r]]) pending = false;
        }
        if (pending) {
            const Real4 *rec = D.rows + (size_t)r * 8;
            a0 = rec[0]; a1 = rec[1]; a2 = rec[2]; a3 = rec[3]; b0 = rec[4]; b1q = rec[5]; b2q = rec[6]; b3 = rec[7];
            rb = D.rbody[r]; fi = D.findex[r];
            hl = L.head_pos[p] - chunk * 32;                       // runs never cross a chunk (k_lw_runs)
            head = hl == lane;
            tail = (p + 1 >= mrows) || (L.head_pos[p + 1] == p + 1);
            if (head) tk = L.ticket[p];
            old_lambda = lam[r];            // each row is updated once per sweep: its own lambda cannot change under it
        }
        const bool two = rb.y != P.NB;
        Real4 f1a = { 0, 0, 0, 0 }, f1b = f1a, f2a = f1a, f2b = f1a;
        unsigned backoff = 0;
        // lambda of the friction-index row (the contact's normal row). When that row sits earlier in the same run its new value
        // is forwarded from the lane that computes it (fwd_src) instead of going through L2 once per friction row; otherwise
        // the row ran in an earlier run on the same two bodies (or runs later in the sweep), so its value is final for this
        // lane as soon as the run's tickets are up and is fetched together with the accumulators.
        int fwd_src = lane; bool fwd = false;
        {
            const int src = lane - (r - fi);
            const int r_src = __shfl_sync(0xffffffffu, r, (src >= 0 && src < 32) ? src : lane);
            if (pending && fi != -1 && src >= hl && src < lane && r_src == fi) { fwd = true; fwd_src = src; }
        }
        Real lam_fi_val = 0, my_lambda = 0;
        while (__any_sync(0xffffffffu, pending)) {
            bool have = false;
            if (pending && head) {
                bool ok = ld_acquire_u32(&L.cnt[rb.x]) == (unsigned)tk.x;
                if (ok && two) ok = ld_acquire_u32(&L.cnt[rb.y]) == (unsigned)tk.y;
                if (ok) {
                    f1a = ldcg4(&cf[2 * rb.x]); f1b = ldcg4(&cf[2 * rb.x + 1]);
                    if (two) { f2a = ldcg4(&cf[2 * rb.y]); f2b = ldcg4(&cf[2 * rb.y + 1]); }
                    have = true;
                }
            }
#ifndef ODEB_LW_BACKOFF_MAX
#define ODEB_LW_BACKOFF_MAX 128
#endif
            {
                const unsigned started = __ballot_sync(0xffffffffu, have);
                if (pending && fi != -1 && !fwd && ((started >> hl) & 1u)) lam_fi_val = __ldcg(&lam[fi]);
            }
            if (!__any_sync(0xffffffffu, have)) {
                if (ODEB_LW_BACKOFF_MAX > 0) { backoff = backoff < ODEB_LW_BACKOFF_MAX ? backoff + 16 : ODEB_LW_BACKOFF_MAX; __nanosleep(backoff); }
                continue;
            }
            backoff = 0;
            for (;;) {
                const bool exec = have;
                const Real lam_fw = __shfl_sync(0xffffffffu, my_lambda, fwd_src);
                if (exec) {
                    Real delta = a1.z - old_lambda * a1.w;
                    delta -= f1a.x * a0.x + f1a.y * a0.y + f1a.z * a0.z + f1a.w * a0.w + f1b.x * a1.x + f1b.y * a1.y;
                    if (two) delta -= f2a.x * b0.x + f2a.y * b0.y + f2a.z * b0.z + f2a.w * b0.w + f2b.x * b1q.x + f2b.y * b1q.y;
                    Real hi_act, lo_act;
                    if (fi != -1) { hi_act = RFABS(b1q.w * (fwd ? lam_fw : lam_fi_val)); lo_act = -hi_act; }
                    else { hi_act = b1q.w; lo_act = b1q.z; }
                    Real new_lambda = old_lambda + delta;
                    if (new_lambda < lo_act) { delta = lo_act - old_lambda; new_lambda = lo_act; }
                    else if (new_lambda > hi_act) { delta = hi_act - old_lambda; new_lambda = hi_act; }
                    __stcg(&lam[r], new_lambda);
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
                    pending = false;
                }
                __syncwarp();           // lambda written by an earlier row of the run is visible to the later rows (friction index)
                const unsigned pass = __ballot_sync(0xffffffffu, exec && !tail);
                // the run's tickets travel with the accumulators: the tail needs them for the release
                const Real4 n1a = shfl_up4(f1a), n1b = shfl_up4(f1b), n2a = shfl_up4(f2a), n2b = shfl_up4(f2b);
                const int ntx = __shfl_up_sync(0xffffffffu, tk.x, 1), nty = __shfl_up_sync(0xffffffffu, tk.y, 1);
                if (exec && tail) {             // one fence for the run's stores, then both counters
                    asm volatile("fence.acq_rel.gpu;" ::: "memory");        // (not __threadfence(): that is the sequentially consistent fence)
                    st_relaxed_u32(&L.cnt[rb.x], (unsigned)tk.x + 1u);
                    if (two) st_relaxed_u32(&L.cnt[rb.y], (unsigned)tk.y + 1u);
                }
                have = lane > 0 && ((pass >> (lane - 1)) & 1u);
                if (have) { f1a = n1a; f1b = n1b; f2a = n2a; f2b = n2b; tk.x = ntx; tk.y = nty; }
                if (pass == 0) break;
            }
        }
    }
}

// after a sweep: per-body convergence test + reset (CheckForMaximumToBeLessThanLimitAndResetMaxAdjustments quickstep.cpp:3253-3285)
__global__ void k_lw_body_check(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D, const __grid_constant__ LargePtrs L, Real exit_delta, int check)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k == 0) L.counters[LWC_CURSOR] = 0;
    if (k >= L.counters[LWC_NORDERED]) return;
    L.cnt[k] = 0;
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
