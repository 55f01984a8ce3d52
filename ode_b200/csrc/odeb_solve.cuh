// odeb_solve.cuh -- k_solve: the SOR-LCP sweeps of dxQuickStepIsland (quickstep.cpp:1823-1856 loop,
// :2329-2355 ReorderPrep, :2578-2611 random reorder via dRandInt, :2917-3033 IterationStep,
// :3253-3285 dynamic iteration control), islands of a world in the reference's order.
//
// The row update is a strict dependency chain (each row reads the constraint-force accumulators the
// previous row wrote), so the kernel is latency-bound per world and throughput comes from worlds in
// flight and from the length of the per-row instruction stream.  Mapping on B200:
//   * TWO lanes per world (16 worlds per warp): lane A owns the body-1 half of every row (J1, rhs, cfm,
//     iMJ1), lane B the body-2 half (J2, lo, hi, iMJ2).  Each lane forms its own 6-term dot product in the
//     reference's left-to-right order, the two partial results meet through one warp shuffle
//     (delta = ((rhs - lambda*cfm) - s1) - s2, exactly the reference's association), both lanes clamp
//     redundantly and each updates its own body's accumulators.  Per-lane instruction and shared-memory
//     traffic are half of a one-lane-per-world sweep.
//   * lambda, the per-position metadata (row index, friction index, the two body slots) and the per-body
//     accumulators (cforce + max-adjustment pair) live in shared memory, interleaved [element][world];
//   * the row half-records (16 reals: 64 B single / 128 B double) stream from L2/HBM through a per-lane ring
//     of cp.async (LDGSTS, L1-bypass) stages issued RING-1 rows ahead in solve order;
//   * the row update is branch-free; one-body rows address a dummy accumulator slot with zero J2/iMJ2.
// Islands too large for the shared-memory budget take the global-memory path (same arithmetic, lane A).
#ifndef ODEB_SOLVE_CUH
#define ODEB_SOLVE_CUH

#define ODEB_RING 8                                        // must stay 8: the row loop is unrolled by the ring depth
#define ODEB_HALF_CHUNKS ((int)(sizeof(Real) * 16 / 16))   // 16-byte chunks per half record
#define ODEB_WPW 16                                        // worlds per warp
// lane -> (world in warp, side). Side = upper half-warp: the 8 lanes of one LDS.128 / STS.128 phase then belong to 8
// different worlds (8 x 4 distinct banks), whereas interleaved sides (lane & 1) put both lanes of a world on the same
// 4 banks with different accumulator slots -> 2-way conflict on every accumulator access.
#if defined(ODEB_LANES_INTERLEAVED)
#define LANE_WL(lane) ((lane) >> 1)
#define LANE_SIDE(lane) ((lane) & 1)
#define LANE_A(lane) ((lane) & ~1)
#define ODEB_PARTNER 1
#else
#define LANE_WL(lane) ((lane) & 15)
#define LANE_SIDE(lane) ((lane) >> 4)
#define LANE_A(lane) ((lane) & 15)
#define ODEB_PARTNER 16
#endif

struct HalfRegs { Real4 q0, q1, q2, q3; };
// record of one row in HBM (8 Real4):
//   A half: q0 = {J1l.x, J1l.y, J1l.z, J1a.x}  q1 = {J1a.y, J1a.z, rhs, cfm}  q2 = {iMJ1l.xyz, iMJ1a.x}  q3 = {iMJ1a.y, iMJ1a.z, max|iMJ1|, hi}
//   B half: q0 = {J2l.x, J2l.y, J2l.z, J2a.x}  q1 = {J2a.y, J2a.z, lo, hi}    q2 = {iMJ2l.xyz, iMJ2a.x}  q3 = {iMJ2a.y, iMJ2a.z, max|iMJ2|, 0}
// (J row quickstep.cpp:267-322, iMJ row :374-431)

__device__ __forceinline__ void cp_async16(unsigned dst, const void *src)
{
    // .cg (L2 only): the L1-allocating .ca form measured 19 % slower here (profiles/r1_solver_variants.txt)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// iteration control (quickstep.cpp:1832-1855, :3253-3285); returns true when the island is finished.
// fa (max adjustments) live in the .z/.w of each body's second accumulator vector.
template <class CF>
__device__ __forceinline__ bool sweep_control(const DevParams &P, CF &cf, int bstart, int nb, unsigned &iteration, unsigned &extra,
                                              Real &exit_delta, unsigned &st1, unsigned &st2, unsigned &st3)
{
    const unsigned num_iterations = P.num_iter;
    if (iteration - extra == num_iterations) {
        if (extra != 0 || P.max_extra == 0) { if (extra != 0) st3++; return true; }
        extra = P.max_extra;
        exit_delta = P.extra_delta;
    }
    if (P.dyn_enabled) {
        bool hit = (exit_delta == 0);
        for (int k = 0; k < nb; k++) {
            Real4 v = cf.get(2 * (bstart + k) + 1);
            if (!(v.w < exit_delta) || !(-v.z < exit_delta)) hit = true;
            v.z = 0; v.w = 0;
            cf.set(2 * (bstart + k) + 1, v);
        }
        if (!hit) {
            if (iteration < num_iterations) st1++;
            else if (iteration > num_iterations) st2++;
            return true;
        }
    }
    return false;
}

struct CfGlobal { Real4 *p; __device__ Real4 get(int i) const { return p[i]; } __device__ void set(int i, const Real4 &v) { p[i] = v; } };
struct CfShared { Real4 *p; __device__ Real4 get(int i) const { return p[i * ODEB_WPW]; } __device__ void set(int i, const Real4 &v) { p[i * ODEB_WPW] = v; } };

__device__ __forceinline__ Real shfl_xor1(unsigned mask, Real v)
{
    return __shfl_xor_sync(mask, v, ODEB_PARTNER);
}

__device__ __forceinline__ void load_half(HalfRegs &h, const uint4 *slot)
{   // slot: chunk c at slot[c * 32]
#if defined(ODEB_DOUBLE)
    uint4 t[8];
#pragma unroll
    for (int c = 0; c < 8; c++) t[c] = slot[c * 32];
    const Real4 *r4 = (const Real4 *)t;
    h.q0 = r4[0]; h.q1 = r4[1]; h.q2 = r4[2]; h.q3 = r4[3];
#else
    const Real4 *s4 = (const Real4 *)slot;
    h.q0 = s4[0]; h.q1 = s4[32]; h.q2 = s4[64]; h.q3 = s4[96];
#endif
}

// per-position, per-lane metadata word: row index (bits 0..10) | friction-index row (bits 11..20) | this lane's
// accumulator slot (bits 21..31).  Rows without a friction index carry their own index in the second field.
#define META_IDX(mt) ((int)((mt) & 0x7ffu))
#define META_FI(mt) ((int)(((mt) >> 11) & 0x3ffu))
#define META_SLOT(mt) ((int)((mt) >> 21))
#define ODEB_META_PAD (2 * ODEB_RING)
#define ODEB_FULL 0xffffffffu

// One row of the sweep for one lane of the pair.  CUR/NXT are the two register sets of the software pipeline,
// K is the position of the row inside the unrolled group of ODEB_RING rows (ring stages become constants).
//   (1) loads that depend on row i-1's stores (lambda, this lane's body accumulators),
//   (2) bookkeeping that fills their latency (cp.async issue for row i+RING-1, register load of row i+1's
//       half record and metadata), (3) the arithmetic chain with one shuffle, (4) stores + warp sync.
// The whole warp runs the row loop in lockstep (trip count = longest island of its 16 worlds); positions past a
// world's own row count read the zero metadata word and have their stores predicated off.
#define ODEB_ROW(CUR, NXT, MT, MTN, K)                                                                                   \
    {                                                                                                                    \
        const int index = META_IDX(MT), fi = META_FI(MT);                                                                \
        Real4 *cfp = cf + (size_t)META_SLOT(MT) * (2 * ODEB_WPW);                                                       \
        const Real old_lambda = lam[index * ODEB_WPW];                                                                   \
        const Real lam_fi = lam[fi * ODEB_WPW];                                                                          \
        Real4 fa = cfp[0], fb = cfp[ODEB_WPW];                                                                           \
        {                                                                                                                \
            if (i + (K) + ODEB_RING - 1 < m) {                                                                           \
                const char *src = rec_base + (size_t)META_IDX(ma) * (sizeof(Real) * 32);                                 \
                const unsigned dst = ring_addr + (unsigned)((((K) + ODEB_RING - 1) & (ODEB_RING - 1)) * CH * 32 * 16);   \
                _Pragma("unroll") for (int c = 0; c < CH; c++) cp_async16(dst + c * 32 * 16, src + c * 16);             \
            }                                                                                                            \
            cp_async_commit();                                                                                           \
        }                                                                                                                \
        const Real lo_b = shfl_xor1(ODEB_FULL, CUR.q1.z);               /* lane A receives lo from lane B */             \
        const Real s = fa.x * CUR.q0.x + fa.y * CUR.q0.y + fa.z * CUR.q0.z + fa.w * CUR.q0.w + fb.x * CUR.q1.x + fb.y * CUR.q1.y; \
        const Real ta = (CUR.q1.z - old_lambda * CUR.q1.w) - s;         /* lane A: (rhs - lambda*cfm) - s1 */           \
        const Real mine = side ? s : ta;                                                                                 \
        const Real other = shfl_xor1(ODEB_FULL, mine);                                                                   \
        /* the next row's operands load while the shuffle is in flight */                                                \
        cp_async_wait<ODEB_RING - 2>();                                                                                  \
        load_half(NXT, ring + (size_t)(((K) + 1) & (ODEB_RING - 1)) * CH * 32);                                         \
        MTN = (i + (K) + 1 < m) ? mp[((K) + 1) * 32] : 0u;                                                               \
        ma = mp[((K) + ODEB_RING) * 32];                                /* metadata of the row prefetched next */        \
        Real delta = side ? (other - mine) : (mine - other);            /* ((rhs - lambda*cfm) - s1) - s2 */             \
        const Real hi = side ? CUR.q1.w : CUR.q3.w;                                                                      \
        const Real lo = side ? CUR.q1.z : lo_b;                                                                          \
        const bool hasfi = fi != index;                                                                                  \
        const Real hi_f = RFABS(hi * lam_fi);                                                                            \
        const Real hi_act = hasfi ? hi_f : hi;                                                                           \
        const Real lo_act = hasfi ? -hi_f : lo;                                                                          \
        Real new_lambda = old_lambda + delta;                                                                            \
        const bool c_lo = new_lambda < lo_act;                                                                           \
        const bool c_hi = !c_lo && (new_lambda > hi_act);                                                                \
        const Real lim = c_lo ? lo_act : hi_act;                                                                         \
        if (c_lo || c_hi) { delta = lim - old_lambda; new_lambda = lim; }                                                \
        const bool pos = delta > 0;                                                                                      \
        fa.x += delta * CUR.q2.x; fa.y += delta * CUR.q2.y; fa.z += delta * CUR.q2.z; fa.w += delta * CUR.q2.w;          \
        fb.x += delta * CUR.q3.x; fb.y += delta * CUR.q3.y;                                                              \
        {                                                                                                                \
            const Real t1 = delta * CUR.q3.z;                                                                            \
            const Real pv = fb.w + t1, nv = fb.z + t1;                                                                   \
            fb.w = pos ? pv : fb.w; fb.z = pos ? fb.z : nv;                                                              \
        }                                                                                                                \
        if (i + (K) < m) {                                                                                               \
            if (side == 0) lam[index * ODEB_WPW] = new_lambda;                                                           \
            cfp[0] = fa; cfp[ODEB_WPW] = fb;                                                                             \
        }                                                                                                                \
        __syncwarp();                                                                                                    \
    }

// Shared-memory solve of one island per world by the whole warp (lane pairs in lockstep). `m_own` is the island's
// row count for this lane's world, 0 when the world has no island for this round (the pair then idles through
// the row loop with stores predicated off).
template <int HY>     // HY = number of sweeps after which a hybrid solve hands over to k_solve5 (0: plain kernel, compiled without any of it)
__device__ __forceinline__ void solve_islands_warp(const DevParams &P, unsigned char *smem, int lane,
                                                   const Real4 *rows, const int *findex, const int2 *rbody, Real4 *cf_out, Real *lam_out,
                                                   int bstart, int nb, int rstart, int m_own, unsigned &seed,
                                                   unsigned &st1, unsigned &st2, unsigned &st3,
                                                   unsigned long long &sweeps, unsigned long long &rowsweeps,
                                                   const int pausable = 0, int *paused_out = 0)
{
    constexpr int CH = ODEB_HALF_CHUNKS;
    const int wl = LANE_WL(lane), side = LANE_SIDE(lane);
    uint4 *ring = (uint4 *)smem + lane;                                          // chunk (stage,c) at ring[(stage*CH+c)*32]
    unsigned char *p = smem + (size_t)ODEB_RING * CH * 32 * 16;
    Real4 *cf = (Real4 *)p + wl;  p += (size_t)2 * (P.NB + 1) * ODEB_WPW * sizeof(Real4);          // cf[k*16], k < 2*(NB+1)
    Real *lam = (Real *)p + wl;   p += (size_t)(P.SR + 1) * ODEB_WPW * sizeof(Real);               // lam[i*16]
    unsigned *meta = (unsigned *)p + lane;                                                            // meta[pos*32], one column per lane
    const unsigned ring_addr = (unsigned)__cvta_generic_to_shared(ring);
    const Real4 z4 = { 0, 0, 0, 0 };
    if (m_own > 0) {
        for (int k = side; k < 2 * nb; k += 2) cf[(2 * bstart + k) * ODEB_WPW] = z4;
        cf[(2 * P.NB + side) * ODEB_WPW] = z4;
        for (int i = side; i <= m_own; i += 2) lam[i * ODEB_WPW] = 0;
        // ReorderPrep quickstep.cpp:2329-2355: rows without a friction index first, then the others
        int nvalid = 0;
        for (int i = 0; i < m_own; i++) if (findex[rstart + i] != -1) nvalid++;
        int head = 0, tail = m_own - nvalid;
        for (int i = 0; i < m_own; i++) {
            const int fi = findex[rstart + i];
            const int2 rb = rbody[rstart + i];
            const unsigned slot = (unsigned)(side ? rb.y : rb.x);
            const unsigned mt = (unsigned)i | ((unsigned)(fi == -1 ? i : fi - rstart) << 11) | (slot << 21);
            if (fi == -1) meta[(head++) * 32] = mt; else meta[(tail++) * 32] = mt;
        }
    } else {
        lam[0] = 0;     // idle pairs address row 0 / slot 0 only
    }
    __syncwarp();
    const char *rec_base = (const char *)(rows + (size_t)rstart * 8) + side * (sizeof(Real) * 16);
    Real exit_delta = P.premature_delta;
    CfShared cfs = { cf };
    int done = m_own > 0 ? 0 : 1;
    int paused = 0;                 // hybrid solve (odeb_host.inl launch_dynamics): the island stops after `stop_after` sweeps and k_solve5 takes over
    unsigned iteration = 0, extra = 0;
    for (;;) {
        if (!done && iteration >= 8 && (iteration & 7) == 0) {
            // ConstraintsShuffling quickstep.cpp:2578-2611 with dRandInt misc.cpp:78-139; each lane permutes its own column
            for (int idx = 1; idx < m_own; idx++) {
                int sw = odeb_rand_int(&seed, idx + 1);
                unsigned a = meta[idx * 32], b = meta[sw * 32];
                meta[idx * 32] = b; meta[sw * 32] = a;
            }
        }
        __syncwarp();
        const int m = done ? 0 : m_own;
        const int m_max = __reduce_max_sync(ODEB_FULL, m);
#pragma unroll
        for (int k = 0; k < ODEB_RING - 1; k++) {
            if (k < m) {
                const char *src = rec_base + (size_t)META_IDX(meta[k * 32]) * (sizeof(Real) * 32);
                unsigned dst = ring_addr + (unsigned)(k * CH * 32 * 16);
#pragma unroll
                for (int c = 0; c < CH; c++) cp_async16(dst + c * 32 * 16, src + c * 16);
            }
            cp_async_commit();
        }
        cp_async_wait<ODEB_RING - 2>();
        HalfRegs r0, r1;
        load_half(r0, ring);
        unsigned mt0 = m > 0 ? meta[0] : 0u, mt1;
        unsigned ma = meta[(ODEB_RING - 1) * 32];
        const unsigned *mp = meta;
        for (int i = 0; i < m_max; i += ODEB_RING, mp += ODEB_RING * 32) {
            ODEB_ROW(r0, r1, mt0, mt1, 0)
            ODEB_ROW(r1, r0, mt1, mt0, 1)
            ODEB_ROW(r0, r1, mt0, mt1, 2)
            ODEB_ROW(r1, r0, mt1, mt0, 3)
            ODEB_ROW(r0, r1, mt0, mt1, 4)
            ODEB_ROW(r1, r0, mt1, mt0, 5)
            ODEB_ROW(r0, r1, mt0, mt1, 6)
            ODEB_ROW(r1, r0, mt1, mt0, 7)
        }
        cp_async_wait<0>();
        int d = 0;
        if (!done) {
            ++iteration; ++sweeps; rowsweeps += m_own;
            if (side == 0) d = sweep_control(P, cfs, bstart, nb, iteration, extra, exit_delta, st1, st2, st3) ? 1 : 0;
        }
        __syncwarp();
        d = __shfl_sync(ODEB_FULL, d, LANE_A(lane));
        done |= d;
        if (HY) { if (pausable && !done && iteration == (unsigned)HY) { paused = 1; done = 1; } }   // before the first reorder: the seed is untouched
        if (__all_sync(ODEB_FULL, done)) break;
    }
    if (HY) *paused_out = paused;
    if (m_own > 0) for (int k = side; k < 2 * nb; k += 2) cf_out[2 * bstart + k] = cf[(2 * bstart + k) * ODEB_WPW];
    if (lam_out && m_own > 0) for (int i = side; i < m_own; i += 2) lam_out[rstart + i] = lam[i * ODEB_WPW];   // joint feedback needs the final lambda
    __syncwarp();
}

// one SOR row update straight from HBM (islands beyond the shared-memory budget)
__device__ __forceinline__ void row_update_global(const Real4 *r, int index, int fi, int b1, int b2, Real4 *cf, Real *lam)
{
    const Real4 a0 = r[0], a1 = r[1], a2 = r[2], a3 = r[3], b0 = r[4], b1q = r[5], b2q = r[6], b3 = r[7];
    const Real old_lambda = lam[index];
    Real delta = a1.z - old_lambda * a1.w;
    Real4 f1a = cf[2 * b1], f1b = cf[2 * b1 + 1];
    delta -= f1a.x * a0.x + f1a.y * a0.y + f1a.z * a0.z + f1a.w * a0.w + f1b.x * a1.x + f1b.y * a1.y;
    Real4 f2a = cf[2 * b2], f2b = cf[2 * b2 + 1];      // one-body rows: b2 = dummy slot NB, J2 = iMJ2 = 0
    delta -= f2a.x * b0.x + f2a.y * b0.y + f2a.z * b0.z + f2a.w * b0.w + f2b.x * b1q.x + f2b.y * b1q.y;
    Real hi_act, lo_act;
    if (fi != -1) { hi_act = RFABS(b1q.w * lam[fi]); lo_act = -hi_act; }
    else { hi_act = b1q.w; lo_act = b1q.z; }
    Real new_lambda = old_lambda + delta;
    if (new_lambda < lo_act) { delta = lo_act - old_lambda; lam[index] = lo_act; }
    else if (new_lambda > hi_act) { delta = hi_act - old_lambda; lam[index] = hi_act; }
    else lam[index] = new_lambda;
    if (delta != 0) {
        f1a.x += delta * a2.x; f1a.y += delta * a2.y; f1a.z += delta * a2.z; f1a.w += delta * a2.w;
        f1b.x += delta * a3.x; f1b.y += delta * a3.y;
        if (delta > 0) f1b.w += delta * a3.z; else f1b.z += delta * a3.z;
        cf[2 * b1] = f1a; cf[2 * b1 + 1] = f1b;
        if (delta > 0) f2b.w += delta * b3.z; else f2b.z += delta * b3.z;
        f2a.x += delta * b2q.x; f2a.y += delta * b2q.y; f2a.z += delta * b2q.z; f2a.w += delta * b2q.w;
        f2b.x += delta * b3.x; f2b.y += delta * b3.y;
        cf[2 * b2] = f2a; cf[2 * b2 + 1] = f2b;
    }
}

// stop_after > 0: first half of the hybrid solve.  Islands of at most sr_b rows stop after `stop_after` sweeps (a multiple of 8 below
// num_iterations, i.e. before any dRand draw and before the extra phase), leave their accumulators in cforce, lambda in D.lambda and
// isl_done = 0; k_solve5 (resume) continues them.  Larger islands, and islands that finish early, are completed here (isl_done = 1).
template <int HY>
__global__ void __launch_bounds__(32) k_solve_t(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D, const int sr_b)
{
    constexpr int stop_after = HY;
    extern __shared__ __align__(32) unsigned char smem[];
    const int lane = threadIdx.x;
    const int side = LANE_SIDE(lane);
    const int wraw = blockIdx.x * ODEB_WPW + LANE_WL(lane);
    const bool valid = wraw < P.W;
    const int w = valid ? wraw : P.W - 1;          // lanes past the last world idle through the warp-wide loops
    unsigned seed = D.seed[w];
    unsigned st0 = 0, st1 = 0, st2 = 0, st3 = 0;
    unsigned long long sweeps = 0, rowsweeps = 0;
    const Real4 *rows = D.rows + (size_t)w * P.MR * 8;
    const int *findex = D.findex + (size_t)w * P.MR;
    const int2 *rbody = D.rbody + (size_t)w * P.MR;
    Real4 *cf_out = D.cforce + (size_t)w * (P.NB + 1) * 2;
    const int4 *iinfo = D.island_info + (size_t)w * P.NB;
    const int nis = valid ? D.nislands[w] : 0;
    const int nis_max = __reduce_max_sync(ODEB_FULL, nis);
    // Hybrid solve: islands are handed over to k_solve5 only if EVERY island of the world can be (or finishes before the hand-over).  An
    // island completed here draws its reorders from the world's seed at once; were an earlier island of the same world paused, it would
    // draw later (in k_solve5) what the reference gives it first, and both islands would see different random orders.
    // (decided from the world's total row count: sufficient, exact for single-island worlds, and one independent load instead of a scan
    //  of the island table at the start of the kernel)
    // (kept as a per-lane row threshold: the same test as a separate predicate costs 20 us per launch at 4096 worlds, measured)
    const int srb_w = (stop_after != 0 && valid && D.mrows[w] <= (sr_b < P.SR ? sr_b : P.SR)) ? sr_b : -1;
    for (int is = 0; is < nis_max; is++) {
        int4 info = make_int4(0, 0, 0, 0);
        if (is < nis) info = iinfo[is];
        const int bstart = info.x, nb = info.y, rstart = info.z, m = info.w;
        if (m > P.SR && side == 0) {
            // ---------------- global-memory path (islands beyond the shared-memory budget), lane A only
            CfGlobal cfg = { cf_out };
            Real *lamg = D.lambda + (size_t)w * P.MR + rstart;
            int *order = D.order + (size_t)w * P.MR + rstart;
            const Real4 z4 = { 0, 0, 0, 0 };
            for (int k = 0; k < 2 * nb; k++) cfg.set(2 * bstart + k, z4);
            cfg.set(2 * P.NB, z4); cfg.set(2 * P.NB + 1, z4);
            int nvalid = 0;
            for (int i = 0; i < m; i++) { lamg[i] = 0; if (findex[rstart + i] != -1) nvalid++; }
            {
                int head = 0, tail = m - nvalid;
                for (int i = 0; i < m; i++) { if (findex[rstart + i] == -1) order[head++] = i; else order[tail++] = i; }
            }
            const Real4 *rec = rows + (size_t)rstart * 8;
            Real exit_delta = P.premature_delta;
            for (unsigned iteration = 0, extra = 0;;) {
                if (iteration >= 8 && (iteration & 7) == 0) {
                    for (int idx = 1; idx < m; idx++) {
                        int sw = odeb_rand_int(&seed, idx + 1);
                        int a = order[idx], b = order[sw];
                        order[idx] = b; order[sw] = a;
                    }
                }
                for (int i = 0; i < m; i++) {
                    const int index = order[i];
                    const int fi = findex[rstart + index];
                    const int2 rb = rbody[rstart + index];
                    row_update_global(rec + (size_t)index * 8, index, fi == -1 ? -1 : fi - rstart, rb.x, rb.y, cf_out, lamg);
                }
                ++iteration; ++sweeps; rowsweeps += m;
                if (sweep_control(P, cfg, bstart, nb, iteration, extra, exit_delta, st1, st2, st3)) break;
            }
        }
        __syncwarp();
        seed = __shfl_sync(ODEB_FULL, seed, LANE_A(lane));     // lane B replays the same dRand stream in the shared-memory path
        const int m_smem = (m > 0 && m <= P.SR) ? m : 0;
        int paused = 0;
        if (__any_sync(ODEB_FULL, m_smem > 0))
            solve_islands_warp<HY>(P, smem, lane, rows, findex, rbody, cf_out, (D.jcopy || stop_after) ? D.lambda + (size_t)w * P.MR : 0, bstart, nb, rstart, m_smem, seed, st1, st2, st3, sweeps, rowsweeps,
                                   m_smem <= srb_w, &paused);
        if (stop_after && side == 0 && is < nis && valid) D.isl_done[(size_t)w * P.NB + is] = paused ? 0 : 1;
        if (is < nis) st0++;
    }
    if (side == 0 && valid) {
        D.seed[w] = seed;
        unsigned *st = D.stats + 4 * (size_t)w;
        st[0] += st0; st[1] += st1; st[2] += st2; st[3] += st3;
        D.sweeps[2 * (size_t)w] = sweeps; D.sweeps[2 * (size_t)w + 1] = rowsweeps;
    }
}
#define ODEB_HYBRID_SWEEPS 8
#define k_solve k_solve_t<0>
#define k_solve_hy k_solve_t<ODEB_HYBRID_SWEEPS>
#endif
