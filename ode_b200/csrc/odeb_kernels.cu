// odeb_kernels.cu -- the B200 (sm_100a) per-step world update behind include/ode_b200.h.
//
// One step of W independent worlds = the reference's
//     dSpaceCollide + near-callback (dCollide, dJointCreateContact, dJointAttach)
//     dWorldQuickStep (auto-disable, islands, dxQuickStepIsland stages 0..6)
//     dJointGroupEmpty
// executed as the kernel sequence below, all state resident in HBM as SoA buffers:
//
//   k_aabb          thread / geom     computeAABB                       (box.cpp:60, sphere.cpp:59, capsule.cpp:60, plane.cpp:80)
//   k_pair_count    thread / geom i   collideAABBs filter over j > i    (collision_space_internal.h:44-78)
//   k_pair_scan     thread / world    exclusive scan of the counts
//   k_pair_fill     thread / geom i   pair list in canonical (i<j lexicographic) order
//   k_narrow        thread / pair     dCollide                          (collision_kernel.cpp:292-338)
//   k_joint_info1   thread / joint    getInfo1 of hinge/universal       (hinge.cpp:54, universal.cpp:266)
//   k_islands       thread / world    contact numbering, dJointAttach list order, auto-disable, island DFS
//                                     (ode.cpp:1383-1439, util.cpp:427-561, util.cpp:724-860)
//   k_body_pre      thread / body     Stage0: gravity, invI_world, gyroscopic torque (quickstep.cpp:1176-1306)
//   k_rows          thread / joint    Stage2a: getInfo2 -> 16-wide rows (quickstep.cpp:1486-1610)
//   k_rows_finish   thread / row      Stage2b+2c, iMJ, Ad scaling       (quickstep.cpp:1644-1747, 859-897, 2251-2316)
//   k_solve         thread / world    SOR-LCP sweeps in the reference's row order, dRand replay,
//                                     dynamic iteration control        (quickstep.cpp:1823-1856, 2329-2355, 2578-2611, 2917-3033, 3253-3285)
//   k_integrate     thread / body     Stage4b, 6a, 6b + dxStepBody      (quickstep.cpp:3082-3108, 3299-3439, util.cpp:583-692)
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false (parity with the reference's
// FMA-free x86-64 build); -DODEB_DOUBLE selects double precision.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdarg>
#include <string>
#include <vector>
#include <algorithm>
#include <nvtx3/nvToolsExt.h>                 // header-only; a no-op unless a tool (nsys, ncu --nvtx) injects itself
#include "../../include/ode_b200.h"
#include "odeb_collide.cuh"
#include "odeb_joints.cuh"

// ------------------------------------------------------------------------------------------------
// device data model

struct alignas(sizeof(Real) * 4) Real4 { Real x, y, z, w; };
#define FINDEX_FLAG 0x40000000

// body flag values (meaning of ode/src/objects.h:49-57)
enum { BF_FINITE_ROT = 1, BF_DISABLED = 4, BF_NO_GRAVITY = 8, BF_AUTO_DISABLE = 16, BF_LIN_DAMP = 32,
       BF_ANG_DAMP = 64, BF_MAX_ANG_SPEED = 128, BF_GYRO = 256 };

struct DevParams {
    int W, NB, NG, NJ, MP, MC, MR, NJT;   // worlds, bodies, geoms, permanent joints, pair / contact / row capacity, NJ+MC
    int maxc, space_type, skip_connected;
    int hash_minlevel, hash_maxlevel;     // dxHashSpace levels (default -3..10, dHashSpaceSetLevels)
    int classic;                          // 1: contacts, their surfaces and the joint adjacency are supplied by the host every step (odeb_classic.inl)
    int m_contact;                        // rows per contact joint (contact.cpp:48-122, uniform under one policy)
    DSurface surf;
    Real gravity[3], erp, cfm, sor_w, premature_delta, extra_delta;
    unsigned num_iter, max_extra; int dyn_enabled;
    Real max_vel, min_depth;
    Real adis_lin, adis_ang, adis_time; int adis_steps, adis_samples;
    Real damp_lin_scale, damp_ang_scale, damp_lin_thr, damp_ang_thr, max_ang_speed;
    Real h, hrecip;
    int SR;                               // rows per island that fit the shared-memory solve path
};

struct DevPtrs {
    // body state [W*NB]
    Real4 *pos, *quat, *lvel, *avel, *facc, *tacc; Real4 *R;   // R: 3 Real4 per body
    int *bflags, *adis_steps; Real *adis_time; Real *avg_buf; int *avg_counter, *avg_ready;
    // template (per batch)
    Real *bmass, *binvmass; Real *bI, *binvI;    // [NB], [NB*12]
    int *gtype, *gbody; Real *gparam; unsigned *gcat, *gcol;   // [NG], gparam [NG*4]
    Real4 *gspose;                               // [NG*4] position + rotation rows of geoms that have no body; offset pose of body geoms with gofs
    int *gofs;                                   // [NG] 1: the geom sits at gspose relative to its body (dGeomSetOffset*)
    DJointT *joints;                             // [NJ]
    int *sadj_ofs, *sadj_joint, *sadj_other;     // static adjacency in attach order
    // collision scratch
    Real *aabb;                                  // [W*NG*6]
    int *pair_cnt, *pair_ofs, *npairs;           // [W*NG], [W*NG], [W]
    int2 *pairs;                                 // [W*MP]
    int *pc_count; Real4 *cgeom;                 // [W*MP], [W*MP*maxc*2] (pos,depth | normal,0)
    int *ray_count;                              // [W*MP] hits of pairs with a ray geom: reported (odeb_get_ray_hits), never turned into joints
    int *ray_geom; int nray;                     // [nray] geom index of every ray, in geom order
    Real *ray_range; int *ray_hit;               // [W*nray] nearest hit of every ray in the last collide pass (k_ray_ranges)
    int *ncontacts; int4 *cinfo;                 // [W], [W*MC] = (slot, b0, b1, reverse)
    DSurface *csurf;                             // [MC] per-contact surface parameters (classic mode)
    // joints dynamic
    int *jm; DLimitState *jlimit;                // [W*NJ]
    // islands
    int *c_ofs, *c_cur, *c_adj_c, *c_adj_o;      // [W*(NB+1)], [W*NB], [W*2*MC] x2
    signed char *btag, *jtag;                    // [W*NB], [W*NJT]
    int *stack;                                  // [W*NB]
    int *body_order, *body_pos, *body_island;    // [W*NB]
    int *joint_order, *joint_row, *joint_island; // [W*NJT]
    int4 *island_info;                           // [W*NB] = (bodyStart, nb, rowStart, m)
    int *nislands, *nordered, *njord, *mrows;    // [W]
    // rows
    Real4 *rows;                                 // [W*MR*8]: one 32-real record per row, split into a body-1 half and a body-2 half
                                                 //   (layout in odeb_solve.cuh)
    int2 *rbody;                                 // [W*MR] accumulator slots (order positions) of the row's two bodies; one-body rows: (p0, NB)
    int *findex, *order; Real *lambda;           // [W*MR]
    unsigned *wkey; int *wlist;                  // [W] work estimate of every world (k_reorder_prep) and the worlds sorted by it, heaviest first (k_solve6's block -> world map)
    unsigned *order0;                            // [W*MR] ReorderPrep order of every island as packed schedule entries (k_reorder_prep, read by k_solve6)
    Real4 *jcopy;                                // [W*MR*3] J1l J1a J2l J2a of every row before any scaling (joint feedback on), else null
    Real4 *jfb;                                  // [W*NJT*4] per joint id: {f1, state} {t1, 0} {f2, 0} {t2, 0} (quickstep.cpp:3108-3182)
    int *row_island, *row_group;                 // [MR] island of every row, first row of the row's group (large-world path only, else null)
    Real4 *cforce;                               // [W*(NB+1)*2]  (fc 6, fa 2) per order position + one dummy slot per world
    Real *invIw;                                 // [W*NB*12], indexed by order position
    unsigned *stats, *seed;                      // [W*4], [W]
    unsigned long long *sweeps;                  // [W*2] = (sweeps, row-sweeps) of the last step
    int *isl_done;                               // [W*NB] hybrid solve: 1 = the island was completed by k_solve, 0 = k_solve5 continues it
    int *overflow;                               // [4] capacity overflow flag; since the host last looked: largest island (rows), largest island
                                                 //     with rows (bodies), most islands with rows in one world
    int *maxpairs;                               // [1] most pairs of any world in this collide pass (k_pair_scan): k_narrow packs its threads by it
};

__device__ __forceinline__ Real4 ld4(const Real4 *p) { return *p; }
__device__ __forceinline__ void load_body(const DevPtrs &D, int gb, DBody &b)
{
    Real4 p = D.pos[gb], q = D.quat[gb], l = D.lvel[gb], a = D.avel[gb];
    b.pos[0] = p.x; b.pos[1] = p.y; b.pos[2] = p.z;
    b.q[0] = q.x; b.q[1] = q.y; b.q[2] = q.z; b.q[3] = q.w;
    b.lvel[0] = l.x; b.lvel[1] = l.y; b.lvel[2] = l.z;
    b.avel[0] = a.x; b.avel[1] = a.y; b.avel[2] = a.z;
    const Real4 *R = D.R + 3 * (size_t)gb;
    Real4 r0 = R[0], r1 = R[1], r2 = R[2];
    b.R[0] = r0.x; b.R[1] = r0.y; b.R[2] = r0.z; b.R[3] = 0;
    b.R[4] = r1.x; b.R[5] = r1.y; b.R[6] = r1.z; b.R[7] = 0;
    b.R[8] = r2.x; b.R[9] = r2.y; b.R[10] = r2.z; b.R[11] = 0;
}

__device__ __forceinline__ void load_geom(const DevParams &P, const DevPtrs &D, int w, int g, DGeom &G)
{
    G.type = D.gtype[g]; G.body = D.gbody[g];
    for (int k = 0; k < 4; k++) G.p[k] = D.gparam[4 * g + k];
    if (G.body >= 0) {
        size_t gb = (size_t)w * P.NB + G.body;
        Real4 p = D.pos[gb];
        G.pos[0] = p.x; G.pos[1] = p.y; G.pos[2] = p.z;
        const Real4 *R = D.R + 3 * gb;
        Real4 r0 = R[0], r1 = R[1], r2 = R[2];
        G.R[0] = r0.x; G.R[1] = r0.y; G.R[2] = r0.z; G.R[3] = 0;
        G.R[4] = r1.x; G.R[5] = r1.y; G.R[6] = r1.z; G.R[7] = 0;
        G.R[8] = r2.x; G.R[9] = r2.y; G.R[10] = r2.z; G.R[11] = 0;
        if (D.gofs[g]) {   // dxGeom::computePosr collision_kernel.cpp:455-466: final = body pose * offset pose
            const Real4 *sp = D.gspose + 4 * (size_t)g;
            Real4 op = sp[0], o0 = sp[1], o1 = sp[2], o2 = sp[3];
            Real ofs[3] = { op.x, op.y, op.z }, oR[12] = { o0.x, o0.y, o0.z, 0, o1.x, o1.y, o1.z, 0, o2.x, o2.y, o2.z, 0 }, fp[3], fR[12];
            mul0_331(fp, G.R, ofs);
            G.pos[0] = fp[0] + G.pos[0]; G.pos[1] = fp[1] + G.pos[1]; G.pos[2] = fp[2] + G.pos[2];
            mul0_333(fR, G.R, oR);
            for (int k = 0; k < 12; k++) G.R[k] = fR[k];
        }
    } else {
        const Real4 *sp = D.gspose + 4 * (size_t)g;
        Real4 p = sp[0], r0 = sp[1], r1 = sp[2], r2 = sp[3];
        G.pos[0] = p.x; G.pos[1] = p.y; G.pos[2] = p.z;
        G.R[0] = r0.x; G.R[1] = r0.y; G.R[2] = r0.z; G.R[3] = 0;
        G.R[4] = r1.x; G.R[5] = r1.y; G.R[6] = r1.z; G.R[7] = 0;
        G.R[8] = r2.x; G.R[9] = r2.y; G.R[10] = r2.z; G.R[11] = 0;
    }
}

// ------------------------------------------------------------------------------------------------
// collision

__global__ void k_aabb(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D)
{
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)P.W * P.NG) return;
    int w = (int)(t / P.NG), g = (int)(t % P.NG);
    if (t == 0 && D.maxpairs) *D.maxpairs = 0;
    DGeom G;
    load_geom(P, D, w, g, G);
    Real a[6];
    odeb_compute_aabb(G, a);
    Real *o = D.aabb + 6 * t;
    for (int k = 0; k < 6; k++) o[k] = a[k];
}

// pair predicate of the selected space, as a set (see oracle/orc_world.cpp find_pairs for the derivation)
__device__ __forceinline__ bool pair_hit(const DevParams &P, const DevPtrs &D, const Real *ai, const Real *aj, int i, int j)
{
    int bi = D.gbody[i], bj = D.gbody[j];
    if (bi == bj && bi >= 0) return false;
    if (((D.gcat[i] & D.gcol[j]) || (D.gcat[j] & D.gcol[i])) == 0) return false;
    if (P.space_type == ODEB_SPACE_SAP) {
        bool infi = (ai[1] == R_INF), infj = (aj[1] == R_INF);
        if (infi != infj) return true;
        if (!infi) {
            float min1 = (float)ai[0], max1 = (float)ai[1], min2 = (float)aj[0], max2 = (float)aj[1];
            bool ax0 = (min1 <= min2) ? (min2 <= max1) : (min1 <= max2);
            return ax0 && !(ai[3] < aj[2] || aj[3] < ai[2]) && !(ai[5] < aj[4] || aj[5] < ai[4]);
        }
    }
    if (ai[0] > aj[1] || ai[1] < aj[0] || ai[2] > aj[3] || ai[3] < aj[2] || ai[4] > aj[5] || ai[5] < aj[4]) return false;
    return P.space_type != ODEB_SPACE_HASH || odeb_hash_space_meets(ai, aj, P.hash_minlevel, P.hash_maxlevel);
}

template <bool FILL>
__global__ void k_pair_pass(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D)
{
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)P.W * P.NG) return;
    int w = (int)(t / P.NG), i = (int)(t % P.NG);
    const Real *A = D.aabb + 6 * (size_t)w * P.NG;
    Real ai[6];
    for (int k = 0; k < 6; k++) ai[k] = A[6 * i + k];
    int n = 0;
    int base = FILL ? D.pair_ofs[t] : 0;
    int2 *out = D.pairs + (size_t)w * P.MP;
    for (int j = i + 1; j < P.NG; j++) {
        Real aj[6];
        for (int k = 0; k < 6; k++) aj[k] = A[6 * j + k];
        if (pair_hit(P, D, ai, aj, i, j)) {
            if (FILL) { if (base + n < P.MP) out[base + n] = make_int2(i, j); }
            n++;
        }
    }
    if (!FILL) D.pair_cnt[t] = n;
}

__global__ void k_pair_scan(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D)
{
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= P.W) return;
    const int *c = D.pair_cnt + (size_t)w * P.NG;
    int *o = D.pair_ofs + (size_t)w * P.NG;
    int s = 0;
    for (int i = 0; i < P.NG; i++) { o[i] = s; s += c[i]; }
    if (s > P.MP) { atomicExch(D.overflow, 1); s = P.MP; }
    D.npairs[w] = s;
    if (D.maxpairs) atomicMax(D.maxpairs, s);
}

// packed: the pair slots of a world are numbered with the stride of the fullest world of this pass (rounded to a power of two) instead of
// the capacity MP, so that live threads sit next to each other (16 live pairs out of MP = 136 slots on a 16-box stack: full warps instead
// of half-empty ones followed by four idle warps per world)
__global__ void k_narrow(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D, const int packed)
{
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)P.W * P.MP) return;
    int w, p;
    if (packed) {
        const int mx = *D.maxpairs;
        int S = 1;
        while (S < mx) S <<= 1;
        if (S > P.MP) S = P.MP;
        if (t >= (size_t)P.W * S) return;
        w = (int)(t / S); p = (int)(t % S);
    } else { w = (int)(t / P.MP); p = (int)(t % P.MP); }
    if (p >= D.npairs[w]) return;
    t = (size_t)w * P.MP + p;
    int2 pr = D.pairs[t];
    int b1 = D.gbody[pr.x], b2 = D.gbody[pr.y];
    int n = 0;
    bool skip = (b1 < 0 && b2 < 0);
    if (!skip && P.skip_connected && P.NJ > 0 && b1 >= 0 && b2 >= 0)      // dAreConnectedExcluding(b1, b2, dJointTypeContact) ode.cpp:1569-1577
        for (int k = D.sadj_ofs[b1]; k < D.sadj_ofs[b1 + 1]; k++) if (D.sadj_other[k] == b2) { skip = true; break; }
    if (!skip) {
        DGeom g1, g2;
        load_geom(P, D, w, pr.x, g1);
        load_geom(P, D, w, pr.y, g2);
        DContactGeom c[8];
        n = odeb_collide(g1, g2, P.maxc, c);
        Real4 *out = D.cgeom + t * P.maxc * 2;
        for (int i = 0; i < n; i++) {
            Real4 a = { c[i].pos[0], c[i].pos[1], c[i].pos[2], c[i].depth };
            Real4 b = { c[i].normal[0], c[i].normal[1], c[i].normal[2], 0 };
            out[2 * i] = a; out[2 * i + 1] = b;
        }
    }
    if (D.ray_count) {      // sensor policy: a pair with a ray keeps its hits in the contact slots but contributes no contact joint
        const bool ray = !skip && (D.gtype[pr.x] == ODEB_RAY || D.gtype[pr.y] == ODEB_RAY);
        D.ray_count[t] = ray ? n : 0;
        if (ray) n = 0;
    }
    D.pc_count[t] = n;
}

// ------------------------------------------------------------------------------------------------
// joints: getInfo1 of permanent joints

__global__ void k_joint_info1(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D)
{
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)P.W * P.NJ) return;
    int w = (int)(t / P.NJ), j = (int)(t % P.NJ);
    const DJointT &jt = D.joints[j];
    int m; DLimitState ls;
    if (jt.type == ODEB_JOINT_BALL) { m = 3; ls.limit1 = ls.limit2 = 0; ls.err1 = ls.err2 = 0; }
    else {
        DBody b0, b1;
        load_body(D, w * P.NB + jt.b0, b0);
        if (jt.b1 >= 0) load_body(D, w * P.NB + jt.b1, b1);
        odeb_joint_info1(jt, b0, jt.b1 >= 0 ? &b1 : 0, &m, &ls);
    }
    D.jm[t] = m; D.jlimit[t] = ls;
}

// ------------------------------------------------------------------------------------------------
// contacts -> joint lists -> auto-disable -> islands (one thread replays one world's list mechanics)

template <bool SM> struct IslIdx { typedef int type; };
template <> struct IslIdx<true> { typedef unsigned short type; };
// shared memory of k_islands_t<true> per block of 32 worlds
__host__ __device__ inline size_t odeb_islands_smem(int NB, int MC, int NJT)
{
    return (size_t)32 * (sizeof(unsigned short) * ((size_t)(NB + 1) + NB + 4 * (size_t)MC + NB) + (size_t)NB + NJT);
}

// SM = true: the scratch of the list replay (contact adjacency CSR, tags, DFS stack) lives in shared memory, interleaved
// [element][thread] as 16-bit / 8-bit entries, instead of per-world global arrays: every step of the replay is a dependent
// read-modify-write, and a world is one thread, so the kernel is bound by the latency of those accesses.
// ONE = true (worlds of more than a few bodies): one WORLD PER WARP, lane 0 runs the replay and its scratch is that warp's own slice of
// shared memory (stride 1).  With a world per lane the 32 replays of a warp diverge at every branch (different contact lists, different
// DFS orders) and execute one after the other: 4096 x 64-body piles took 488 us (0.93 M cycles per warp = 32 x the ~29 k cycles of one
// world).  A warp per world has nothing to diverge from, and ~40 worlds are resident per SM (5 KB of scratch each) instead of 32.
template <bool SM, bool ONE>
__global__ void __launch_bounds__(ONE ? 128 : 32) k_islands_t(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D)
{
    extern __shared__ __align__(16) unsigned char isl_smem[];
    int w = ONE ? (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5) : (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (w >= P.W) return;
    if (ONE && (threadIdx.x & 31) != 0) return;
    const int NB = P.NB, NJ = P.NJ, MC = P.MC;
    constexpr int S = (SM && !ONE) ? 32 : 1;             // element stride of the scratch arrays
    typedef typename IslIdx<SM>::type idx_t;             // unsigned short in shared memory, int in global memory
    const int NONE = SM ? 0xffff : -1;                   // "no second body" in adj_o
    // 1. number the contact joints in creation order (pair order, then dCollide's contact order)
    int nc = 0;
    if (P.classic) nc = D.ncontacts[w];     // contact joints were created through dJointCreateContact: numbered by the host
    else {
        int np = D.npairs[w];
        const int2 *pairs = D.pairs + (size_t)w * P.MP;
        const int *pcc = D.pc_count + (size_t)w * P.MP;
        int4 *ci = D.cinfo + (size_t)w * MC;
        for (int p = 0; p < np; p++) {
            int cnt = pcc[p];
            if (!cnt) continue;
            int2 pr = pairs[p];
            int b1 = D.gbody[pr.x], b2 = D.gbody[pr.y], rev = 0;
            if (b1 < 0) { b1 = b2; b2 = -1; rev = 1; }       // dJointAttach ode.cpp:1404-1411
            for (int k = 0; k < cnt; k++) {
                if (nc < MC) ci[nc] = make_int4(p * P.maxc + k, b1, b2, rev);
                nc++;
            }
        }
        if (nc > MC) { atomicExch(D.overflow, 2); nc = MC; }
        D.ncontacts[w] = nc;
    }
    // 2. per-body contact adjacency (CSR, ascending contact index; walked descending = newest first)
    idx_t *cofs, *ccur, *adj_c, *adj_o, *stack;
    signed char *btag, *jtag;
    if (SM && ONE) {
        const size_t per_world = (odeb_islands_smem(NB, MC, P.NJT) / 32 + 15) / 16 * 16;
        idx_t *q = (idx_t *)(isl_smem + (threadIdx.x >> 5) * per_world);
        cofs = q; q += (size_t)(NB + 1); ccur = q; q += (size_t)NB; adj_c = q; q += (size_t)2 * MC;
        adj_o = q; q += (size_t)2 * MC; stack = q; q += (size_t)NB;
        btag = (signed char *)q; jtag = btag + (size_t)NB;
    } else if (SM) {
        idx_t *q = (idx_t *)isl_smem + threadIdx.x;
        cofs = q; q += (size_t)(NB + 1) * 32; ccur = q; q += (size_t)NB * 32; adj_c = q; q += (size_t)2 * MC * 32;
        adj_o = q; q += (size_t)2 * MC * 32; stack = q; q += (size_t)NB * 32;
        btag = (signed char *)q + threadIdx.x; jtag = btag + (size_t)NB * 32;
    } else {
        cofs = (idx_t *)(D.c_ofs + (size_t)w * (NB + 1)); ccur = (idx_t *)(D.c_cur + (size_t)w * NB);
        adj_c = (idx_t *)(D.c_adj_c + (size_t)w * 2 * MC); adj_o = (idx_t *)(D.c_adj_o + (size_t)w * 2 * MC);
        stack = (idx_t *)(D.stack + (size_t)w * NB);
        btag = D.btag + (size_t)w * NB; jtag = D.jtag + (size_t)w * P.NJT;
    }
    const int4 *ci = D.cinfo + (size_t)w * MC;
    for (int b = 0; b <= NB; b++) cofs[b * S] = 0;
    if (!P.classic) {       // classic mode: the host-built adjacency (sadj_*) already holds every joint in dJointAttach order
        for (int c = 0; c < nc; c++) { int4 v = ci[c]; cofs[(v.y + 1) * S]++; if (v.z >= 0) cofs[(v.z + 1) * S]++; }
        for (int b = 0; b < NB; b++) { cofs[(b + 1) * S] += cofs[b * S]; ccur[b * S] = cofs[b * S]; }
        for (int c = 0; c < nc; c++) {
            int4 v = ci[c];
            int k = ccur[v.y * S]++; adj_c[k * S] = (idx_t)c; adj_o[k * S] = (idx_t)(v.z >= 0 ? v.z : NONE);
            if (v.z >= 0) { k = ccur[v.z * S]++; adj_c[k * S] = (idx_t)c; adj_o[k * S] = (idx_t)v.y; }
        }
    }
    // 3. dInternalHandleAutoDisabling util.cpp:427-561 (world->firstbody order = reverse creation)
    int *bflags = D.bflags + (size_t)w * NB;
    if (P.adis_samples > 0) {
        for (int b = NB - 1; b >= 0; b--) {
            int fl = bflags[b];
            if ((fl & (BF_AUTO_DISABLE | BF_DISABLED)) != BF_AUTO_DISABLE) continue;
            if (cofs[(b + 1) * S] == cofs[b * S] && D.sadj_ofs[b + 1] == D.sadj_ofs[b]) continue;
            size_t gb = (size_t)w * NB + b;
            Real4 lv = D.lvel[gb], av = D.avel[gb];
            Real *buf = D.avg_buf + gb * 6 * P.adis_samples;
            int cnt = D.avg_counter[gb], ready = D.avg_ready[gb];
            buf[6 * cnt + 0] = lv.x; buf[6 * cnt + 1] = lv.y; buf[6 * cnt + 2] = lv.z;
            buf[6 * cnt + 3] = av.x; buf[6 * cnt + 4] = av.y; buf[6 * cnt + 5] = av.z;
            cnt++;
            if (cnt >= P.adis_samples) { cnt = 0; ready = 1; }
            D.avg_counter[gb] = cnt; D.avg_ready[gb] = ready;
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
            int sl = D.adis_steps[gb]; Real tl = D.adis_time[gb];
            if (idle) { sl--; tl -= P.h; } else { sl = P.adis_steps; tl = P.adis_time; }
            D.adis_steps[gb] = sl; D.adis_time[gb] = tl;
            if (sl <= 0 && tl <= 0) {
                bflags[b] = fl | BF_DISABLED;
                Real4 z = { 0, 0, 0, 0 };
                D.lvel[gb] = z; D.avel[gb] = z;
            }
        }
    }
    // 4. BuildIslands util.cpp:724-860: DFS from each untagged enabled body, LIFO body stack,
    //    each body's joints walked newest attachment first (contacts of this step, then permanent joints)
    int *border = D.body_order + (size_t)w * NB, *bpos = D.body_pos + (size_t)w * NB, *bisl = D.body_island + (size_t)w * NB;
    int *jorder = D.joint_order + (size_t)w * P.NJT, *jrow = D.joint_row + (size_t)w * P.NJT, *jisl = D.joint_island + (size_t)w * P.NJT;
    int4 *iinfo = D.island_info + (size_t)w * NB;
    const int *jm = D.jm + (size_t)w * NJ;
    for (int b = 0; b < NB; b++) { btag[b * S] = 0; bisl[b] = -1; bpos[b] = -1; }
    for (int j = 0; j < NJ + nc; j++) jtag[j * S] = 0;
    int nbo = 0, njo = 0, nis = 0, rows = 0, nis_rows = 0, max_mi = 0, max_nbi = 0;
    for (int bb = NB - 1; bb >= 0; bb--) {
        if (btag[bb * S]) continue;
        if (bflags[bb] & BF_DISABLED) { btag[bb * S] = -1; continue; }
        btag[bb * S] = 1;
        int bstart = nbo, rstart = rows, mi = 0;
        border[nbo] = bb; bpos[bb] = nbo; bisl[bb] = nis; nbo++;
        int sp = 0, b = bb;
        while (true) {
            for (int k = (int)cofs[(b + 1) * S] - 1, k0 = (int)cofs[b * S]; k >= k0; k--) {
                int jid = NJ + adj_c[k * S];
                if (!jtag[jid * S]) {
                    jtag[jid * S] = 1;
                    jorder[njo] = jid; jrow[njo] = mi; jisl[njo] = nis; njo++;
                    mi += P.m_contact;
                    int nb2 = adj_o[k * S];
                    if (nb2 != NONE && btag[nb2 * S] <= 0) { btag[nb2 * S] = 1; bflags[nb2] &= ~BF_DISABLED; stack[(sp++) * S] = (idx_t)nb2; }
                }
            }
            for (int k = D.sadj_ofs[b + 1] - 1; k >= D.sadj_ofs[b]; k--) {
                int jid = D.sadj_joint[k];
                if (!jtag[jid * S]) {
                    jtag[jid * S] = 1;
                    int m = jid < NJ ? jm[jid] : D.csurf[jid - NJ].the_m;
                    if (m != 0) { jorder[njo] = jid; jrow[njo] = mi; jisl[njo] = nis; njo++; mi += m; }
                    int nb2 = D.sadj_other[k];
                    if (nb2 >= 0 && btag[nb2 * S] <= 0) { btag[nb2 * S] = 1; bflags[nb2] &= ~BF_DISABLED; stack[(sp++) * S] = (idx_t)nb2; }
                }
            }
            if (sp == 0) break;
            b = stack[(--sp) * S];
            border[nbo] = b; bpos[b] = nbo; bisl[b] = nis; nbo++;
        }
        if (rstart + mi > P.MR) { atomicExch(D.overflow, 3); mi = 0; }
        iinfo[nis] = make_int4(bstart, nbo - bstart, rstart, mi);
        if (mi > 0) { nis_rows++; if (mi > max_mi) max_mi = mi; if (nbo - bstart > max_nbi) max_nbi = nbo - bstart; }
        rows += mi;
        nis++;
    }
    // largest island / most islands of the call: the host picks the solver kernel and its shared-memory budget from them
    if (max_mi > 0) { atomicMax(D.overflow + 1, max_mi); atomicMax(D.overflow + 2, max_nbi); atomicMax(D.overflow + 3, nis_rows); }
    D.nislands[w] = nis; D.nordered[w] = nbo; D.njord[w] = njo; D.mrows[w] = rows;
}

// ------------------------------------------------------------------------------------------------
// Stage 0 (bodies)

__global__ void k_body_pre(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D)
{
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)P.W * P.NB) return;
    // a capacity overflow anywhere in this step (pairs, contacts or rows truncated): no body moves, in this step or in the ones queued
    // behind it, until the host has seen the flag (odeb_sync / odeb_get_state) -- the state stays the one the last complete step left
    if (D.overflow[0] != 0) return;
    int w = (int)(t / P.NB), k = (int)(t % P.NB);
    if (k >= D.nordered[w]) return;
    int b = D.body_order[t];
    size_t gb = (size_t)w * P.NB + b;
    int fl = D.bflags[gb];
    Real4 fa = D.facc[gb];
    if (!(fl & BF_NO_GRAVITY)) {
        Real m = D.bmass[b];
        if (P.gravity[0]) fa.x += m * P.gravity[0];
        if (P.gravity[1]) fa.y += m * P.gravity[1];
        if (P.gravity[2]) fa.z += m * P.gravity[2];
        D.facc[gb] = fa;
    }
    Real R[12], tmp[12], ii[12];
    { const Real4 *Rp = D.R + 3 * gb; Real4 r0 = Rp[0], r1 = Rp[1], r2 = Rp[2];
      R[0] = r0.x; R[1] = r0.y; R[2] = r0.z; R[3] = 0; R[4] = r1.x; R[5] = r1.y; R[6] = r1.z; R[7] = 0; R[8] = r2.x; R[9] = r2.y; R[10] = r2.z; R[11] = 0; }
    const Real *invI = D.binvI + 12 * b;
    mul2_333(tmp, invI, R);
    mul0_333(ii, R, tmp);
    Real *out = D.invIw + 12 * t;
    for (int i = 0; i < 12; i++) out[i] = ((i & 3) == 3) ? R_(0.0) : ii[i];
    if ((fl & BF_GYRO) && D.binvmass[b] > 0) {
        Real I[12], L[3], Itild[12], itInv[12];
        Real4 av4 = D.avel[gb];
        Real av[3] = { av4.x, av4.y, av4.z };
        mul2_333(tmp, D.bI + 12 * b, R);
        mul0_333(I, R, tmp);
        I[3] = I[7] = I[11] = 0;
        mul0_331(L, I, av);
        for (int i = 0; i < 12; i++) Itild[i] = 0;
        Itild[1] = +L[2]; Itild[2] = -L[1]; Itild[4] = -L[2]; Itild[6] = +L[0]; Itild[8] = +L[1]; Itild[9] = -L[0];
        for (int i = 0; i < 12; i++) Itild[i] = Itild[i] * P.h + I[i];
        L[0] *= P.hrecip; L[1] *= P.hrecip; L[2] *= P.hrecip;
        if (invert3(itInv, Itild) != 0) {
            itInv[3] = itInv[7] = itInv[11] = 0;
            mul0_333(Itild, I, itInv);
            Itild[0] -= 1; Itild[5] -= 1; Itild[10] -= 1;
            Real tau0[3];
            mul0_331(tau0, Itild, L);
            Real4 ta = D.tacc[gb];
            ta.x += tau0[0]; ta.y += tau0[1]; ta.z += tau0[2];
            D.tacc[gb] = ta;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Stage 2a: rows

__device__ __forceinline__ void atomic_add_real(Real *p, Real v) { atomicAdd(p, v); }

__device__ __forceinline__ Real modmax6(const Real *v)
{   // dxCalculateModuloMaximum matrix.h:92-104
    Real r = RFABS(v[0]);
    for (int i = 1; i < 6; i++) { Real a = RFABS(v[i]); if (a > r) r = a; }
    return r;
}

// rhs_tmp of one body (Stage2b quickstep.cpp:1644-1690), recomputed per row instead of staged
__device__ __forceinline__ void body_rhs_tmp(const DevParams &P, const DevPtrs &D, int w, int pos, Real *out, Real *invI, Real *invMass)
{
    size_t go = (size_t)w * P.NB + pos;
    int b = D.body_order[go];
    size_t gb = (size_t)w * P.NB + b;
    Real im = D.binvmass[b];
    *invMass = im;
    Real4 fa = D.facc[gb], ta = D.tacc[gb], lv = D.lvel[gb], av = D.avel[gb];
    const Real *ii = D.invIw + 12 * go;
    for (int i = 0; i < 12; i++) invI[i] = ii[i];
    out[0] = -(fa.x * im + lv.x * P.hrecip);
    out[1] = -(fa.y * im + lv.y * P.hrecip);
    out[2] = -(fa.z * im + lv.z * P.hrecip);
    Real tv[3] = { ta.x, ta.y, ta.z }, r[3];
    mul0_331(r, invI, tv);
    out[3] = -(av.x * P.hrecip) - r[0];
    out[4] = -(av.y * P.hrecip) - r[1];
    out[5] = -(av.z * P.hrecip) - r[2];
}

// Stage2c (rhs += J * rhs_tmp, quickstep.cpp:1023-1055), compute_invM_JT (:859-897) and Stage4LCP_AdComputation (:2251-2316) of one row, and
// its final 8 x Real4 record (odeb_solve.cuh).  in/invI/im: body_rhs_tmp of the row's bodies.  Shared by k_rows_finish and the fused k_rows.
__device__ __forceinline__ void finish_row(const DevParams &P, Real *q, const Real *in0, const Real *invI0, Real im0,
                                           bool two, const Real *in1, const Real *invI1, Real im1, Real4 *Jp)
{
    Real imj[14];
    Real sum = R_(0.0);
    for (int k = 0; k < 6; k++) sum += q[C_J1L + k] * in0[k];
    for (int k = 0; k < 3; k++) imj[k] = im0 * q[C_J1L + k];
    mul0_331(imj + 3, invI0, q + C_J1A);
    imj[6] = P.dyn_enabled ? modmax6(imj) : R_(0.0);
    for (int k = 7; k < 14; k++) imj[k] = 0;
    if (two) {
        for (int k = 0; k < 6; k++) sum += q[C_J2L + k] * in1[k];
        for (int k = 0; k < 3; k++) imj[7 + k] = im1 * q[C_J2L + k];
        mul0_331(imj + 10, invI1, q + C_J2A);
        imj[13] = P.dyn_enabled ? modmax6(imj + 7) : R_(0.0);
    }
    q[C_RHS] += sum;
    Real s2 = R_(0.0);
    for (int k = 0; k < 6; k++) s2 += imj[k] * q[C_J1L + k];
    if (two) for (int k = 0; k < 6; k++) s2 += imj[7 + k] * q[C_J2L + k];
    Real cfm_i = q[C_CFM];
    Real Ad = P.sor_w / (s2 + cfm_i);
    q[C_CFM] = cfm_i * Ad;
    q[C_RHS] *= Ad;
    for (int k = 0; k < 6; k++) q[C_J1L + k] *= Ad;
    if (two) for (int k = 0; k < 6; k++) q[C_J2L + k] *= Ad;
    Real4 a0 = { q[0], q[1], q[2], q[3] }, a1 = { q[4], q[5], q[6], q[7] };
    Real4 a2 = { imj[0], imj[1], imj[2], imj[3] }, a3 = { imj[4], imj[5], imj[6], q[C_HI] };
    Real4 b0 = { q[8], q[9], q[10], q[11] }, b1 = { q[12], q[13], q[C_LO], q[C_HI] };
    Real4 b2 = { imj[7], imj[8], imj[9], imj[10] }, b3 = { imj[11], imj[12], imj[13], 0 };
    Jp[0] = a0; Jp[1] = a1; Jp[2] = a2; Jp[3] = a3; Jp[4] = b0; Jp[5] = b1; Jp[6] = b2; Jp[7] = b3;
}

// FUSED (worlds without permanent joints: no limit motor can add body forces after a row was built, so the grid-wide ordering
// Stage2a -> Stage2b of quickstep.cpp:1486-1690 is not needed): the thread finishes its rows itself (finish_row) and writes the final
// records once, instead of k_rows_finish reading the half-built records back and rewriting them.
// STD3 (with FUSED): every joint is a contact with exactly three rows (normal + two friction directions, DevParams::rows_std3): the row
// loops unroll and the rows live in registers instead of a 640-byte local array.
template <bool FUSED, bool STD3 = false>
__global__ void k_rows_t(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D)
{
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)P.W * P.NJT) return;
    int w = (int)(t / P.NJT), k = (int)(t % P.NJT);
    if (k >= D.njord[w]) return;
    int jid = D.joint_order[t];
    int4 isl = D.island_info[(size_t)w * P.NB + D.joint_island[t]];
    if (isl.w == 0) return;
    int rbase = D.joint_row[t];                 // island-local first row
    int row0 = isl.z + rbase;                   // world-local first row
    Real row[6 * ROWLEN];
    int findex[6];
    int m, b0i, b1i;
    DBody b0, b1;
    Real tq[3] = { 0, 0, 0 }, fq[3] = { 0, 0, 0 }, tboth[3] = { 0, 0, 0 };
    bool has_tq = false, has_f = false;
    if (STD3 || jid >= P.NJ) {
        int4 ci = D.cinfo[(size_t)w * P.MC + (jid - P.NJ)];
        const DSurface &surf = (!STD3 && P.classic) ? D.csurf[jid - P.NJ] : P.surf;
        m = STD3 ? 3 : surf.the_m; b0i = ci.y; b1i = ci.z;
#pragma unroll
        for (int r = 0; r < (STD3 ? 3 : m); r++) {
            Real *q = row + r * ROWLEN;
#pragma unroll
            for (int c = 0; c < ROWLEN; c++) q[c] = 0;
            q[C_CFM] = P.cfm; q[C_LO] = -R_INF; q[C_HI] = R_INF;
            findex[r] = -1;
        }
        load_body(D, w * P.NB + b0i, b0);
        if (b1i >= 0) load_body(D, w * P.NB + b1i, b1);
        const Real4 *cg = D.cgeom + ((size_t)w * (P.classic ? (size_t)P.MC : (size_t)P.MP * P.maxc) + ci.x) * 2;
        Real4 a = cg[0], n4 = cg[1];
        Real cpos[3] = { a.x, a.y, a.z }, cn[3] = { n4.x, n4.y, n4.z };
        odeb_contact_info2<STD3>(surf, cpos, cn, a.w, ci.w, b0, b1i >= 0 ? &b1 : 0, P.hrecip, P.erp, P.min_depth, P.max_vel, row, findex);
    } else {
        const DJointT &jt = D.joints[jid];
        m = D.jm[(size_t)w * P.NJ + jid]; b0i = jt.b0; b1i = jt.b1;
        for (int r = 0; r < m; r++) {
            Real *q = row + r * ROWLEN;
            for (int c = 0; c < ROWLEN; c++) q[c] = 0;
            q[C_CFM] = P.cfm; q[C_LO] = -R_INF; q[C_HI] = R_INF;
            findex[r] = -1;
        }
        load_body(D, w * P.NB + b0i, b0);
        if (b1i >= 0) load_body(D, w * P.NB + b1i, b1);
        odeb_joint_info2(jt, D.jlimit[(size_t)w * P.NJ + jid], b0, b1i >= 0 ? &b1 : 0, P.hrecip, P.erp, row, tq, &has_tq, fq, tboth, &has_f);
    }
    int p0 = D.body_pos[(size_t)w * P.NB + b0i], p1 = b1i >= 0 ? D.body_pos[(size_t)w * P.NB + b1i] : -1;
    Real4 *rec = D.rows + ((size_t)w * P.MR + row0) * 8;
    int *fi = D.findex + (size_t)w * P.MR + row0;
    Real in0[6], invI0[12], im0 = 0, in1[6], invI1[12], im1 = 0;
    if (FUSED) {
        body_rhs_tmp(P, D, w, p0, in0, invI0, &im0);
        if (p1 != -1) body_rhs_tmp(P, D, w, p1, in1, invI1, &im1);
    }
#pragma unroll
    for (int r = 0; r < (STD3 ? 3 : m); r++) {
        Real *q = row + r * ROWLEN;
        q[C_RHS] *= P.hrecip; q[C_CFM] *= P.hrecip;
        if (!FUSED) {
            Real4 v0 = { q[0], q[1], q[2], q[3] }, v1 = { q[4], q[5], q[6], q[7] }, v2 = { q[8], q[9], q[10], q[11] }, v3 = { q[12], q[13], q[14], q[15] };
            rec[8 * r] = v0; rec[8 * r + 1] = v1; rec[8 * r + 2] = v2; rec[8 * r + 3] = v3;
        }
        if (D.jcopy) {   // Jcopy quickstep.cpp:1562-1583
            Real4 *jc = D.jcopy + ((size_t)w * P.MR + row0 + r) * 3;
            Real4 c0 = { q[C_J1L], q[C_J1L + 1], q[C_J1L + 2], q[C_J1A] }, c1 = { q[C_J1A + 1], q[C_J1A + 2], q[C_J2L], q[C_J2L + 1] },
                  c2 = { q[C_J2L + 2], q[C_J2A], q[C_J2A + 1], q[C_J2A + 2] };
            jc[0] = c0; jc[1] = c1; jc[2] = c2;
        }
        fi[r] = findex[r] == -1 ? -1 : findex[r] + row0;
        if (D.row_island) {     // large-world path: the contacts of one geom pair (ids consecutive, same island) form one row group
            D.row_island[row0 + r] = D.joint_island[t];
            D.row_group[row0 + r] = (jid >= P.NJ) ? row0 - (D.cinfo[(size_t)w * P.MC + (jid - P.NJ)].x % P.maxc) * m : row0;
        }
        if (FUSED) {
            finish_row(P, q, in0, invI0, im0, p1 != -1, in1, invI1, im1, rec + 8 * r);
            D.rbody[(size_t)w * P.MR + row0 + r] = make_int2(p0, (p1 == -1) ? P.NB : p1);
        } else {
            // body order positions travel in the last two slots of the record (k_rows_finish) and in rbody (large-world path: k_lwc_groups)
            *(int *)&rec[8 * r + 7].z = p0; *(int *)&rec[8 * r + 7].w = p1;
            D.rbody[(size_t)w * P.MR + row0 + r] = make_int2(p0, (p1 == -1) ? P.NB : p1);
        }
    }
    if (has_f) {    // dBodyAddForce / dBodyAddTorque from a powered linear limit motor at its stop (joints/joint.cpp:688-704)
        Real *f0 = (Real *)&D.facc[(size_t)w * P.NB + b0i];
        atomic_add_real(f0, -fq[0]); atomic_add_real(f0 + 1, -fq[1]); atomic_add_real(f0 + 2, -fq[2]);
        if (b1i >= 0) {
            Real *f1 = (Real *)&D.facc[(size_t)w * P.NB + b1i], *t0 = (Real *)&D.tacc[(size_t)w * P.NB + b0i], *t1 = (Real *)&D.tacc[(size_t)w * P.NB + b1i];
            atomic_add_real(t0, tboth[0]); atomic_add_real(t0 + 1, tboth[1]); atomic_add_real(t0 + 2, tboth[2]);
            atomic_add_real(t1, tboth[0]); atomic_add_real(t1 + 1, tboth[1]); atomic_add_real(t1 + 2, tboth[2]);
            atomic_add_real(f1, fq[0]); atomic_add_real(f1 + 1, fq[1]); atomic_add_real(f1 + 2, fq[2]);
        }
    }
    if (has_tq) {   // dBodyAddTorque from a powered limit motor at its stop (joints/joint.cpp:677-705)
        Real *t0 = (Real *)&D.tacc[(size_t)w * P.NB + b0i];
        atomic_add_real(t0, -tq[0]); atomic_add_real(t0 + 1, -tq[1]); atomic_add_real(t0 + 2, -tq[2]);
        if (b1i >= 0) { Real *t1 = (Real *)&D.tacc[(size_t)w * P.NB + b1i]; atomic_add_real(t1, tq[0]); atomic_add_real(t1 + 1, tq[1]); atomic_add_real(t1 + 2, tq[2]); }
    }
}

__global__ void k_rows_finish(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D)
{
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)P.W * P.MR) return;
    int w = (int)(t / P.MR), i = (int)(t % P.MR);
    if (i >= D.mrows[w]) return;
    Real4 *Jp = D.rows + t * 8, *Mp = Jp + 4;
    Real q[16];
    { Real4 v0 = Jp[0], v1 = Jp[1], v2 = Jp[2], v3 = Jp[3];
      q[0] = v0.x; q[1] = v0.y; q[2] = v0.z; q[3] = v0.w; q[4] = v1.x; q[5] = v1.y; q[6] = v1.z; q[7] = v1.w;
      q[8] = v2.x; q[9] = v2.y; q[10] = v2.z; q[11] = v2.w; q[12] = v3.x; q[13] = v3.y; q[14] = v3.z; q[15] = v3.w; }
    int p0 = *(int *)&Mp[3].z, p1 = *(int *)&Mp[3].w;
    Real in0[6], invI0[12], im0, in1[6], invI1[12], im1 = 0;
    body_rhs_tmp(P, D, w, p0, in0, invI0, &im0);
    if (p1 != -1) body_rhs_tmp(P, D, w, p1, in1, invI1, &im1);
    finish_row(P, q, in0, invI0, im0, p1 != -1, in1, invI1, im1, Jp);
    D.rbody[t] = make_int2(p0, (p1 == -1) ? P.NB : p1);     // one-body rows address the dummy accumulator slot NB
}

// Stage4b joint feedback (quickstep.cpp:3108-3182, Multiply1_12q1 :153-184): f = sum over the joint's rows of Jcopy^T * lambda,
// one running sum per component in row order. Thread per ordered joint.
__global__ void k_feedback(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D)
{
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)P.W * P.NJT) return;
    int w = (int)(t / P.NJT), k = (int)(t % P.NJT);
    if (k >= D.njord[w]) return;
    int jid = D.joint_order[t];
    int4 isl = D.island_info[(size_t)w * P.NB + D.joint_island[t]];
    if (isl.w == 0) return;
    const int row0 = isl.z + D.joint_row[t];
    int m; bool two;
    if (jid >= P.NJ) {
        int4 ci = D.cinfo[(size_t)w * P.MC + (jid - P.NJ)];
        m = (P.classic ? D.csurf[jid - P.NJ] : P.surf).the_m; two = ci.z >= 0;
    } else { m = D.jm[(size_t)w * P.NJ + jid]; two = D.joints[jid].b1 >= 0; }
    const Real4 *jc = D.jcopy + ((size_t)w * P.MR + row0) * 3;
    const Real *lam = D.lambda + (size_t)w * P.MR + row0;
    Real a[6] = { 0, 0, 0, 0, 0, 0 }, b[6] = { 0, 0, 0, 0, 0, 0 };
    for (int r = 0; r < m; r++) {
        const Real4 c0 = jc[3 * r], c1 = jc[3 * r + 1], c2 = jc[3 * r + 2];
        const Real s = lam[r];
        a[0] += c0.x * s; a[1] += c0.y * s; a[2] += c0.z * s; a[3] += c0.w * s; a[4] += c1.x * s; a[5] += c1.y * s;
        if (two) { b[0] += c1.z * s; b[1] += c1.w * s; b[2] += c2.x * s; b[3] += c2.y * s; b[4] += c2.z * s; b[5] += c2.w * s; }
    }
    Real4 *o = D.jfb + ((size_t)w * P.NJT + jid) * 4;
    Real4 o0 = { a[0], a[1], a[2], two ? R_(2.0) : R_(1.0) }, o1 = { a[3], a[4], a[5], 0 }, o2 = { b[0], b[1], b[2], 0 }, o3 = { b[3], b[4], b[5], 0 };
    o[0] = o0; o[1] = o1; o[2] = o2; o[3] = o3;
}

#include "odeb_solve.cuh"
#include "odeb_solve_bl.cuh"
#include "odeb_solve5.cuh"
#include "odeb_solve6.cuh"
#include "odeb_large.cuh"

// ------------------------------------------------------------------------------------------------
// Stage 4b + 6a + 6b (dxStepBody)

__device__ __forceinline__ Real sinc_(Real x)
{
    if (RFABS(x) < 1.0e-4) return R_(1.0) - x * x * R_(0.166666666666666666667);
    return RSIN(x) / x;
}

__global__ void k_integrate(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D)
{
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)P.W * P.NB) return;
    // a capacity overflow anywhere in this step (pairs, contacts or rows truncated): no body moves, in this step or in the ones queued
    // behind it, until the host has seen the flag (odeb_sync / odeb_get_state) -- the state stays the one the last complete step left
    if (D.overflow[0] != 0) return;
    int w = (int)(t / P.NB), k = (int)(t % P.NB);
    if (k >= D.nordered[w]) return;
    int b = D.body_order[t];
    size_t gb = (size_t)w * P.NB + b;
    const Real h = P.h;
    Real4 lv = D.lvel[gb], av = D.avel[gb], fa = D.facc[gb], ta = D.tacc[gb];
    int4 info = D.island_info[(size_t)w * P.NB + D.body_island[gb]];
    if (info.w > 0) {   // Stage4b quickstep.cpp:3082-3108
        const Real4 *cfp = D.cforce + ((size_t)w * (P.NB + 1) + k) * 2;
        Real4 c0 = cfp[0], c1 = cfp[1];
        lv.x += h * c0.x; av.x += h * c0.w;
        lv.y += h * c0.y; av.y += h * c1.x;
        lv.z += h * c0.z; av.z += h * c1.y;
    }
    {   // Stage6a quickstep.cpp:3299-3337
        Real km = h * D.binvmass[b];
        lv.x += km * fa.x; lv.y += km * fa.y; lv.z += km * fa.z;
        Real tv[3] = { ta.x * h, ta.y * h, ta.z * h }, r[3];
        const Real *ii = D.invIw + 12 * t;
        Real invI[12];
        for (int i = 0; i < 12; i++) invI[i] = ii[i];
        mul0_331(r, invI, tv);
        av.x += r[0]; av.y += r[1]; av.z += r[2];
    }
    // dxStepBody util.cpp:583-692
    int fl = D.bflags[gb];
    if (fl & BF_MAX_ANG_SPEED) {
        const Real mas = P.max_ang_speed;
        const Real asp = av.x * av.x + av.y * av.y + av.z * av.z;
        if (asp > mas * mas) { const Real coef = mas / RSQRT(asp); av.x *= coef; av.y *= coef; av.z *= coef; }
    }
    Real4 ps = D.pos[gb], q4 = D.quat[gb];
    ps.x += h * lv.x; ps.y += h * lv.y; ps.z += h * lv.z;
    Real q[4] = { q4.x, q4.y, q4.z, q4.w };
    Real wv[3] = { av.x, av.y, av.z };
    if (fl & BF_FINITE_ROT) {
        Real qr[4], q2[4];
        Real wlen = RSQRT(wv[0] * wv[0] + wv[1] * wv[1] + wv[2] * wv[2]);
        Real hh = h * R_(0.5);
        Real theta = wlen * hh;
        qr[0] = RCOS(theta);
        Real s = sinc_(theta) * hh;
        qr[1] = wv[0] * s; qr[2] = wv[1] * s; qr[3] = wv[2] * s;
        qmul0(q2, qr, q);
        for (int j = 0; j < 4; j++) q[j] = q2[j];
    } else {
        Real dq[4];
        dq_from_w(dq, wv, q);
        for (int j = 0; j < 4; j++) q[j] += h * dq[j];
    }
    normalize4(q);
    Real R[12];
    r_from_q(R, q);
    if (fl & BF_LIN_DAMP) {
        const Real ls = lv.x * lv.x + lv.y * lv.y + lv.z * lv.z;
        if (ls > P.damp_lin_thr) { const Real kk = 1 - P.damp_lin_scale; lv.x *= kk; lv.y *= kk; lv.z *= kk; }
    }
    if (fl & BF_ANG_DAMP) {
        const Real as = av.x * av.x + av.y * av.y + av.z * av.z;
        if (as > P.damp_ang_thr) { const Real kk = 1 - P.damp_ang_scale; av.x *= kk; av.y *= kk; av.z *= kk; }
    }
    Real4 z = { 0, 0, 0, 0 };
    q4.x = q[0]; q4.y = q[1]; q4.z = q[2]; q4.w = q[3];
    D.pos[gb] = ps; D.quat[gb] = q4; D.lvel[gb] = lv; D.avel[gb] = av; D.facc[gb] = z; D.tacc[gb] = z;
    Real4 *Rp = D.R + 3 * gb;
    Real4 r0 = { R[0], R[1], R[2], 0 }, r1 = { R[4], R[5], R[6], 0 }, r2 = { R[8], R[9], R[10], 0 };
    Rp[0] = r0; Rp[1] = r1; Rp[2] = r2;
}

// state upload helper: normalise quaternion + rebuild R (dBodySetQuaternion ode.cpp:379-392)
__global__ void k_set_quat(int n, Real4 *quat, Real4 *Rout)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    Real4 q4 = quat[t];
    Real q[4] = { q4.x, q4.y, q4.z, q4.w }, R[12];
    normalize4(q);
    r_from_q(R, q);
    q4.x = q[0]; q4.y = q[1]; q4.z = q[2]; q4.w = q[3];
    quat[t] = q4;
    Real4 r0 = { R[0], R[1], R[2], 0 }, r1 = { R[4], R[5], R[6], 0 }, r2 = { R[8], R[9], R[10], 0 };
    Rout[3 * t] = r0; Rout[3 * t + 1] = r1; Rout[3 * t + 2] = r2;
}

// body state (Real4 SoA) -> tight arrays [pos 3n | quat 4n | lvel 3n | avel 3n] for one device->host transfer
__global__ void k_pack_state(size_t n, const Real4 *pos, const Real4 *quat, const Real4 *lvel, const Real4 *avel, Real *out)
{
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const Real4 p = pos[t], q = quat[t], l = lvel[t], a = avel[t];
    Real *o = out + 3 * t;            o[0] = p.x; o[1] = p.y; o[2] = p.z;
    o = out + 3 * n + 4 * t;          o[0] = q.x; o[1] = q.y; o[2] = q.z; o[3] = q.w;
    o = out + 7 * n + 3 * t;          o[0] = l.x; o[1] = l.y; o[2] = l.z;
    o = out + 10 * n + 3 * t;         o[0] = a.x; o[1] = a.y; o[2] = a.z;
}

// Range-sensor read-out: nearest hit (smallest dContactGeom.depth = distance along the ray, ray.cpp) of every ray geom of every world in
// the last collide pass, +inf / -1 when the ray saw nothing.  Thread per (world, ray); a world has a few dozen pairs at most.
__global__ void k_ray_ranges(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D)
{
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)P.W * D.nray) return;
    const int w = (int)(t / D.nray), g = D.ray_geom[t % D.nray];
    const int np = D.npairs[w];
    const int2 *pr = D.pairs + (size_t)w * P.MP;
    const int *rc = D.ray_count + (size_t)w * P.MP;
    Real best = R_INF; int other = -1;
    for (int p = 0; p < np; p++) {
        if (rc[p] <= 0 || (pr[p].x != g && pr[p].y != g)) continue;
        const Real d = D.cgeom[((size_t)w * P.MP + p) * P.maxc * 2].w;
        if (d < best) { best = d; other = pr[p].x == g ? pr[p].y : pr[p].x; }
    }
    D.ray_range[t] = best; D.ray_hit[t] = other;
}

// dBodyAddForce / dBodyAddTorque for every body from tight [force 3n | torque 3n] arrays
__global__ void k_add_ft(size_t n, Real4 *facc, Real4 *tacc, const Real *src)
{
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    if (facc) { Real4 a = facc[t]; const Real *f = src + 3 * t; a.x += f[0]; a.y += f[1]; a.z += f[2]; facc[t] = a; }
    if (tacc) { Real4 a = tacc[t]; const Real *f = src + 3 * n + 3 * t; a.x += f[0]; a.y += f[1]; a.z += f[2]; tacc[t] = a; }
}

#include "odeb_host.inl"
#include "odeb_large_host.inl"
#include "odeb_classic.inl"
