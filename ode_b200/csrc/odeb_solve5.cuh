// odeb_solve5.cuh -- k_solve5<P>: the SOR-LCP sweeps of dxQuickStepIsland with P row processors per world
// (quickstep.cpp:1823-1856 loop, :2329-2355 ReorderPrep, :2578-2611 random reorder via dRandInt,
//  :2917-3033 IterationStep, :3253-3285 dynamic iteration control).
//
// k_solve (odeb_solve.cuh) runs one row at a time per world: with a few thousand worlds the GPU holds every world
// at once and the step time is the length of the per-world dependency chain (m rows x N sweeps x ~420 cycles).
// The reference's sequential sweep only orders rows that share a body (a row reads and writes the accumulators
// of its two bodies and lambda of its own contact), so rows on disjoint bodies may run side by side with
// bit-identical results.  This kernel gives every world P lane pairs ("processors") and a STATIC schedule:
//   * after ReorderPrep and after every dRand reorder, the world's leader lane list-schedules the solve order
//     onto P in-order processors: slot(row) = first slot after the latest earlier row on either of its bodies in
//     which a processor is free.  The schedule is reused for the 8 sweeps until the next reorder;
//   * a sweep then is `nslots` lockstep slots (16-box stack: 136 for the initial order, ~52 after a shuffle for
//     P = 4, instead of 192), each slot the row update of k_solve: two lanes per row (body-1 / body-2 halves,
//     one shuffle), accumulators + lambda in shared memory, half records through a per-lane cp.async ring that runs
//     along the lane's own schedule column (4 stages: shared memory is what limits the number of resident warps; 8
//     stages and an additional prefetch.global.L2 ODEB5_FAR slots ahead were measured and bring nothing for P = 4,
//     profiles/r1_solver_variants.txt).  Idle slots carry the dummy body and have their copies and stores predicated off.
// Lane layout: lane = proc * 2 WPW + side * WPW + world (WPW = 16 / P worlds per warp), schedule column = proc * WPW + world.  The two
// lanes of a row (its body-1 / body-2 halves) are WPW lanes apart, i.e. in the same quarter-warp for P = 4: an LDS.128 / STS.128 phase then
// holds, per world, the two bodies of ONE row (different bodies, neighbours in the island order for stacks and chains, hence different
// CF5 colours) instead of the body-1 halves of two unrelated rows, which collided whenever their parities agreed (ODEB5_OLD_LANES: the
// previous layout, side = lane >> 4).
// One order / meta entry: row (bits 0..11) | row - friction-index row (12..14, 0 = none; the reference only ever points a
// friction row at the normal row of its own contact, 1 or 2 rows back) | body-1 slot (15..22) | body-2 slot (23..30).
#ifndef ODEB_SOLVE5_CUH
#define ODEB_SOLVE5_CUH

#ifndef ODEB5_RING
#define ODEB5_RING 4                                       // 4 or 8: the slot loop is unrolled by the ring depth
#endif
#ifndef ODEB5_FAR
#define ODEB5_FAR 12                                       // L2 prefetch distance in slots
#endif
#define ODEB5_PAD (ODEB5_RING + ODEB5_FAR + 4)
#define ODEB5_MAXBODIES 254                                // 8-bit body slots, the dummy slot NB included
#define ODEB5_MAXROWS 4095                                 // 12-bit row index
#define E5_ROW(e) ((int)((e) & 0xfffu))
#define E5_FI(e) (E5_ROW(e) - (int)(((e) >> 12) & 7u))     // friction-index row = row - delta, delta 0 = none (fi == row)
#define E5_B1(e) ((int)(((e) >> 15) & 0xffu))
#define E5_B2(e) ((int)(((e) >> 23) & 0xffu))

// shared memory per warp for a capacity of sr rows (= schedule slots) per island
__host__ __device__ inline size_t odeb5_smem(int P, int NB, int sr)
{
    const int wpw = 16 / P;
    size_t b = (size_t)ODEB5_RING * ODEB_HALF_CHUNKS * 32 * 16;              // ring
    b += (size_t)2 * (NB + 1) * wpw * sizeof(Real4);                          // accumulators
    b += (size_t)(sr + 1) * wpw * sizeof(Real);                               // lambda
    b = (b + 15) / 16 * 16;
    b += (size_t)(sr + ODEB5_PAD) * 16 * sizeof(unsigned);                    // schedule columns
    b += (size_t)sr * wpw * sizeof(unsigned);                                 // solve order
    b += (size_t)(NB + 1) * wpw * sizeof(unsigned short);                     // scheduler: last slot per body
    return (b + 31) / 32 * 32;
}

// accumulator vector j (0/1) of body slot b for world wl: the 16-byte unit inside the body's 8 (WPW = 4) units is rotated
// by a bit of the body index, so that two processors of one world meet in the same bank group half as often
#define CF5(b, j) cf[(size_t)(b) * (2 * WPW) + ((((j) * WPW) + (((b) & 1) * WPW)) & (2 * WPW - 1))]

template <int WPW> struct CfShared5 {     // element i = vector (i & 1) of body slot (i >> 1), see CF5
    Real4 *cf;
    __device__ Real4 get(int i) const { return CF5(i >> 1, i & 1); }
    __device__ void set(int i, const Real4 &v) { CF5(i >> 1, i & 1) = v; }
};

#if !defined(ODEB5_FAR_PREFETCH)     // measured: no gain (profiles/r1_solver_variants.txt), off by default
__device__ __forceinline__ void prefetch_l2(const void *) {}
#else
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p)); }
#endif

// One schedule slot.  CUR/NXT: register sets of the software pipeline; MT: this slot's entry; MTN: next slot's (loaded here);
// ma: entry RING-1 slots ahead (its cp.async is issued here), mf: entry FAR slots ahead (L2 prefetch).
#define ODEB5_ROW(CUR, NXT, MT, MTN, K)                                                                                  \
    {                                                                                                                    \
        const int index = E5_ROW(MT), fi = E5_FI(MT), b1 = E5_B1(MT);                                                    \
        const bool live = b1 != NBd;                                                                                     \
        const int bs = side ? E5_B2(MT) : b1;                                                                            \
        const Real old_lambda = lam[index * WPW];                                                                        \
        const Real lam_fi = lam[fi * WPW];                                                                               \
        Real4 fa = CF5(bs, 0), fb = CF5(bs, 1);                                                                          \
        MTN = mp[((K) + 1) * 16];                                                                                        \
        {                                                                                                                \
            if (E5_B1(ma) != NBd) {                                                                                      \
                const char *src = rec_base + (size_t)E5_ROW(ma) * (sizeof(Real) * 32);                                   \
                const unsigned dst = ring_addr + (unsigned)((((K) + ODEB5_RING - 1) & (ODEB5_RING - 1)) * CH * 32 * 16); \
                _Pragma("unroll") for (int c = 0; c < CH; c++) cp_async16(dst + c * 32 * 16, src + c * 16);             \
            }                                                                                                            \
            cp_async_commit();                                                                                           \
            if (E5_B1(mf) != NBd) prefetch_l2(rec_base + (size_t)E5_ROW(mf) * (sizeof(Real) * 32));                      \
        }                                                                                                                \
        ma = mp[((K) + ODEB5_RING) * 16];                               /* schedule entries for the next slot's copies */ \
        mf = mp[((K) + 1 + ODEB5_FAR) * 16];                                                                             \
        const Real lo_b = __shfl_xor_sync(ODEB_FULL, CUR.q1.z, PARTNER5); /* lane A receives lo from lane B */           \
        const Real s = fa.x * CUR.q0.x + fa.y * CUR.q0.y + fa.z * CUR.q0.z + fa.w * CUR.q0.w + fb.x * CUR.q1.x + fb.y * CUR.q1.y; \
        const Real ta = (CUR.q1.z - old_lambda * CUR.q1.w) - s;         /* lane A: (rhs - lambda*cfm) - s1 */           \
        const Real mine = side ? s : ta;                                                                                 \
        const Real other = __shfl_xor_sync(ODEB_FULL, mine, PARTNER5);                                                   \
        cp_async_wait<ODEB5_RING - 2>();                                                                                 \
        load_half(NXT, ring + (size_t)(((K) + 1) & (ODEB5_RING - 1)) * CH * 32);                                        \
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
        if (live) {                                                                                                      \
            if (side == 0) lam[index * WPW] = new_lambda;                                                                \
            CF5(bs, 0) = fa; CF5(bs, 1) = fb;                                                                            \
        }                                                                                                                \
        __syncwarp();                                                                                                    \
    }

// resume > 0: second half of the hybrid solve.  Islands k_solve paused after `resume` sweeps (D.isl_done == 0) are continued from their
// accumulators (cforce) and lambda (D.lambda): ReorderPrep is rebuilt (it is deterministic), the reorder that the reference performs before
// sweep resume + 1 is drawn here, and the sweeps go on under the P-processor schedule.  Islands with isl_done == 1 are left alone.
template <int P, int RESUME>
__global__ void __launch_bounds__(32) k_solve5_t(const __grid_constant__ DevParams Pm, const __grid_constant__ DevPtrs D, const int SR5)
{
    constexpr int resume = RESUME;
    extern __shared__ __align__(32) unsigned char smem[];
    constexpr int WPW = 16 / P;
    constexpr int CH = ODEB_HALF_CHUNKS;
    const int lane = threadIdx.x;
#if defined(ODEB5_OLD_LANES)
    constexpr int PARTNER5 = 16;
    const int side = lane >> 4, col = lane & 15;
    const int wl = col & (WPW - 1), proc = col / WPW;
#else
    constexpr int PARTNER5 = WPW;
    const int wl = lane & (WPW - 1), side = (lane / WPW) & 1, proc = lane / (2 * WPW);
    const int col = proc * WPW + wl;
#endif
    const int wid = lane / WPW;                                   // index of the lane among its world's 2P lanes
    const unsigned wmask = ((WPW == 1) ? 0xffffffffu : (WPW == 2) ? 0x55555555u : (WPW == 4) ? 0x11111111u : 0x01010101u) << wl;
    const unsigned below = (1u << lane) - 1u;
    const bool leader = (proc == 0) && (side == 0);               // lane index == wl
    const int NBd = Pm.NB;                                        // dummy accumulator slot
    const int wraw = blockIdx.x * WPW + wl;
    const bool valid = wraw < Pm.W;
    const int w = valid ? wraw : Pm.W - 1;

    // ---- shared memory
    uint4 *ring = (uint4 *)smem + lane;                           // chunk (stage, c) at ring[(stage * CH + c) * 32]
    const unsigned ring_addr = (unsigned)__cvta_generic_to_shared(ring);
    unsigned char *p = smem + (size_t)ODEB5_RING * CH * 32 * 16;
    Real4 *cf = (Real4 *)p + wl;    p += (size_t)2 * (Pm.NB + 1) * WPW * sizeof(Real4);        // cf[k * WPW]
    Real *lam = (Real *)p + wl;     p += (size_t)(SR5 + 1) * WPW * sizeof(Real);                // lam[i * WPW]
    p = smem + ((size_t)(p - smem) + 15) / 16 * 16;
    unsigned *meta = (unsigned *)p; p += (size_t)(SR5 + ODEB5_PAD) * 16 * sizeof(unsigned);     // meta[slot * 16 + col]
    unsigned *order = (unsigned *)p + wl; p += (size_t)SR5 * WPW * sizeof(unsigned);            // order[i * WPW]
    unsigned short *last = (unsigned short *)p + wl;                                            // last[body * WPW]
    const unsigned IDLE = ((unsigned)NBd << 15) | ((unsigned)NBd << 23);

    unsigned seed = D.seed[w];
    unsigned st0 = 0, st1 = 0, st2 = 0, st3 = 0;
    unsigned long long sweeps = 0, rowsweeps = 0;
    const Real4 *rows = D.rows + (size_t)w * Pm.MR * 8;
    const int *findex = D.findex + (size_t)w * Pm.MR;
    const int2 *rbody = D.rbody + (size_t)w * Pm.MR;
    Real4 *cf_out = D.cforce + (size_t)w * (Pm.NB + 1) * 2;
    const int4 *iinfo = D.island_info + (size_t)w * Pm.NB;
    const int nis = valid ? D.nislands[w] : 0;
    const int nis_max = __reduce_max_sync(ODEB_FULL, nis);
    // every schedule column starts idle (shared memory is not cleared between kernels; lanes of worlds that never solve an
    // island still walk their column)
    if (side == 0) for (int s = 0; s < SR5 + ODEB5_PAD; s++) meta[s * 16 + col] = IDLE;
    __syncwarp();
    int fill = 0;                                                 // slots of this world's columns that may hold entries

    for (int is = 0; is < nis_max; is++) {
        int4 info = make_int4(0, 0, 0, 0);
        if (is < nis) { info = iinfo[is]; if (leader && !resume) st0++; }
        if (resume && is < nis && D.isl_done[(size_t)w * Pm.NB + is]) info.w = 0;      // completed by k_solve
        const int bstart = info.x, nb = info.y, rstart = info.z, m = info.w;
        // islands beyond the row budget, or with a friction index the packed entry cannot hold, take the serial path
        const int m_try = (m > 0 && m <= SR5) ? m : 0;
        const int m_try_max = __reduce_max_sync(ODEB_FULL, m_try);
        int nfree = 0;
        unsigned badfi = 0;
        for (int i0 = 0; i0 < m_try_max; i0 += 2 * P) {
            const int i = i0 + wid;
            int fi = -1;
            if (i < m_try) fi = findex[rstart + i];
            const bool fr = (i < m_try) && fi == -1;
            const bool bad = (i < m_try) && fi != -1 && (unsigned)(i - (fi - rstart) - 1) > 6u;
            nfree += __popc(__ballot_sync(ODEB_FULL, fr) & wmask);
            badfi |= __ballot_sync(ODEB_FULL, bad) & wmask;
        }
        const bool big = m > SR5 || badfi != 0;
        if (big && leader) solve_island_serial(Pm, D, w, bstart, nb, rstart, m, seed, st1, st2, st3, sweeps, rowsweeps);
        __syncwarp();
        const int m_own = (m > 0 && !big) ? m : 0;
        const int m_max = __reduce_max_sync(ODEB_FULL, m_own);
        if (m_max == 0) continue;

        // ---- island set-up by the world's 2P lanes: accumulators and lambda = 0, ReorderPrep (stable partition:
        //      rows without a friction index first, quickstep.cpp:2329-2355)
        const Real4 z4 = { 0, 0, 0, 0 };
        if (m_own > 0) {
            if (resume) {
                const Real *lam_in = D.lambda + (size_t)w * Pm.MR + rstart;
                for (int k = wid; k < 2 * nb; k += 2 * P) CF5(bstart + (k >> 1), k & 1) = cf_out[2 * bstart + k];
                for (int i = wid; i <= m_own; i += 2 * P) lam[i * WPW] = i < m_own ? lam_in[i] : R_(0.0);
            } else {
                for (int k = wid; k < 2 * nb; k += 2 * P) CF5(bstart + (k >> 1), k & 1) = z4;
                for (int i = wid; i <= m_own; i += 2 * P) lam[i * WPW] = 0;
            }
            if (wid < 2) CF5(Pm.NB, wid) = z4;
        }
        {
            int head = 0, tail = nfree;
            for (int i0 = 0; i0 < m_max; i0 += 2 * P) {
                const int i = i0 + wid;
                const bool in = i < m_own;
                int fi = 0; int2 rb = make_int2(0, 0);
                if (in) { fi = findex[rstart + i]; rb = rbody[rstart + i]; }
                const bool fr = in && fi == -1;
                const unsigned bf = __ballot_sync(ODEB_FULL, fr) & wmask, bo = __ballot_sync(ODEB_FULL, in && !fr) & wmask;
                const unsigned e = (unsigned)i | ((unsigned)(fi == -1 ? 0 : i - (fi - rstart)) << 12) | ((unsigned)rb.x << 15) | ((unsigned)rb.y << 23);
                if (fr) order[(head + __popc(bf & below)) * WPW] = e;
                else if (in) order[(tail + __popc(bo & below)) * WPW] = e;
                head += __popc(bf); tail += __popc(bo);
            }
        }
        __syncwarp();

        const char *rec_base = (const char *)(rows + (size_t)rstart * 8) + side * (sizeof(Real) * 16);
        Real exit_delta = Pm.premature_delta;
        CfShared5<WPW> cfs = { cf };
        int done = m_own > 0 ? 0 : 1;
        unsigned iteration = (unsigned)resume, extra = 0;
        int nslots = 0;
        bool resched = !done;
        if (resume && !done && leader) {
            // the reorder the reference performs at the top of iteration `resume` (a multiple of 8), quickstep.cpp:2578-2611
            int sw = m_own > 1 ? odeb_rand_int(&seed, 2) : 0;        // the next draw is computed while this swap's loads are in flight
            for (int idx = 1; idx < m_own; idx++) {
                const unsigned a = order[idx * WPW], b = order[sw * WPW];
                const int swn = idx + 1 < m_own ? odeb_rand_int(&seed, idx + 2) : 0;
                order[idx * WPW] = b; order[sw * WPW] = a;
                sw = swn;
            }
        }
        __syncwarp();
        for (;;) {
            if (__any_sync(ODEB_FULL, resched)) {
                // ---- (re)build the schedule: clear the columns, then the leader list-schedules the order onto P processors
                const int clr = __reduce_max_sync(ODEB_FULL, resched ? fill : 0);
                if (side == 0) for (int s = 0; s < clr; s++) if (resched) meta[s * 16 + col] = IDLE;
                __syncwarp();
                int ns = 0;
                if (resched && leader) {
                    for (int k = 0; k < nb; k++) last[(bstart + k) * WPW] = 0;
                    last[NBd * WPW] = 0;
                    int pf[P];
#pragma unroll
                    for (int q = 0; q < P; q++) pf[q] = 0;
                    unsigned *mcol = meta + wl;
                    // software-pipelined: the next row's entry and its bodies' last slots are loaded before this row's stores, then
                    // patched if this row just moved them (the shared-memory round trip leaves the loop-carried chain)
                    unsigned e = m_own > 0 ? order[0] : IDLE;
                    int b1 = E5_B1(e), b2 = E5_B2(e);
                    int l1 = last[b1 * WPW], l2 = last[b2 * WPW];
                    for (int i = 0; i < m_own; i++) {
                        unsigned en = IDLE;
                        if (i + 1 < m_own) en = order[(i + 1) * WPW];
                        const int b1n = E5_B1(en), b2n = E5_B2(en);
                        int l1n = last[b1n * WPW], l2n = last[b2n * WPW];
                        const int r = l1 > l2 ? l1 : l2;
                        int best = -1, bestpf = -1, mn = 0, mnpf = pf[0];
#pragma unroll
                        for (int q = 0; q < P; q++) {
                            if (pf[q] <= r && pf[q] > bestpf) { best = q; bestpf = pf[q]; }
                            if (pf[q] < mnpf) { mn = q; mnpf = pf[q]; }
                        }
                        const int pq = best >= 0 ? best : mn;
                        const int slot = best >= 0 ? r : mnpf;
#pragma unroll
                        for (int q = 0; q < P; q++) if (q == pq) pf[q] = slot + 1;
                        mcol[slot * 16 + pq * WPW] = e;
                        last[b1 * WPW] = (unsigned short)(slot + 1);
                        if (b2 != NBd) last[b2 * WPW] = (unsigned short)(slot + 1);
                        ns = ns > slot + 1 ? ns : slot + 1;
                        if (b1n == b1 || b1n == b2) l1n = slot + 1;                     // b1n is a real body, so b1n == b2 implies b2 is one too
                        if (b2n != NBd && (b2n == b1 || b2n == b2)) l2n = slot + 1;     // the dummy slot's entry stays 0
                        e = en; b1 = b1n; b2 = b2n; l1 = l1n; l2 = l2n;
                    }
                }
                ns = __shfl_sync(ODEB_FULL, ns, wl);
                if (resched) { nslots = ns; fill = ns; }
                resched = false;
                __syncwarp();
            }
            const int ns_own = done ? 0 : nslots;
            const int ns_max = __reduce_max_sync(ODEB_FULL, ns_own);
            // a finished world must not execute its (still valid) schedule again: its lanes read the idle entry
            const unsigned *mp = done ? (meta + (size_t)SR5 * 16 + col) : (meta + col);   // [SR5, SR5 + PAD) stays idle for good
            // ---- prime the ring along the lane's schedule column
#pragma unroll
            for (int k = 0; k < ODEB5_RING - 1; k++) {
                const unsigned mk = mp[k * 16];
                if (E5_B1(mk) != NBd) {
                    const char *src = rec_base + (size_t)E5_ROW(mk) * (sizeof(Real) * 32);
                    const unsigned dst = ring_addr + (unsigned)(k * CH * 32 * 16);
#pragma unroll
                    for (int c = 0; c < CH; c++) cp_async16(dst + c * 32 * 16, src + c * 16);
                }
                cp_async_commit();
            }
            for (int k = ODEB5_RING - 1; k < ODEB5_FAR; k++) {
                const unsigned mk = mp[k * 16];
                if (E5_B1(mk) != NBd) prefetch_l2(rec_base + (size_t)E5_ROW(mk) * (sizeof(Real) * 32));
            }
            cp_async_wait<ODEB5_RING - 2>();
            HalfRegs r0, r1;
            load_half(r0, ring);
            unsigned mt0 = mp[0], mt1;
            unsigned ma = mp[(ODEB5_RING - 1) * 16], mf = mp[ODEB5_FAR * 16];
            for (int i = 0; i < ns_max; i += ODEB5_RING, mp = done ? mp : mp + ODEB5_RING * 16) {
                ODEB5_ROW(r0, r1, mt0, mt1, 0)
                ODEB5_ROW(r1, r0, mt1, mt0, 1)
                ODEB5_ROW(r0, r1, mt0, mt1, 2)
                ODEB5_ROW(r1, r0, mt1, mt0, 3)
#if ODEB5_RING == 8
                ODEB5_ROW(r0, r1, mt0, mt1, 4)
                ODEB5_ROW(r1, r0, mt1, mt0, 5)
                ODEB5_ROW(r0, r1, mt0, mt1, 6)
                ODEB5_ROW(r1, r0, mt1, mt0, 7)
#endif
            }
            cp_async_wait<0>();
            int d = 0;
            bool shuffle = false;
            if (!done) {
                ++iteration;
                if (leader) {
                    ++sweeps; rowsweeps += m_own;
                    d = sweep_control(Pm, cfs, bstart, nb, iteration, extra, exit_delta, st1, st2, st3) ? 1 : 0;
                    shuffle = !d && iteration >= 8 && (iteration & 7) == 0;
                    if (shuffle) {
                        // ConstraintsShuffling quickstep.cpp:2578-2611 with dRandInt misc.cpp:78-139
                        int sw = m_own > 1 ? odeb_rand_int(&seed, 2) : 0;
                        for (int idx = 1; idx < m_own; idx++) {
                            const unsigned a = order[idx * WPW], b = order[sw * WPW];
                            const int swn = idx + 1 < m_own ? odeb_rand_int(&seed, idx + 2) : 0;
                            order[idx * WPW] = b; order[sw * WPW] = a;
                            sw = swn;
                        }
                    }
                }
            }
            __syncwarp();
            d = __shfl_sync(ODEB_FULL, d, wl);
            resched = __shfl_sync(ODEB_FULL, shuffle ? 1 : 0, wl) != 0;
            done |= d;
            if (__all_sync(ODEB_FULL, done)) break;
        }
        if (m_own > 0) for (int k = wid; k < 2 * nb; k += 2 * P) cf_out[2 * bstart + k] = CF5(bstart + (k >> 1), k & 1);
        if (D.jcopy && m_own > 0) for (int i = wid; i < m_own; i += 2 * P) D.lambda[(size_t)w * Pm.MR + rstart + i] = lam[i * WPW];   // joint feedback
        __syncwarp();
    }
    if (leader && valid) {
        D.seed[w] = seed;
        unsigned *st = D.stats + 4 * (size_t)w;
        st[0] += st0; st[1] += st1; st[2] += st2; st[3] += st3;
        if (resume) { D.sweeps[2 * (size_t)w] += sweeps; D.sweeps[2 * (size_t)w + 1] += rowsweeps; }
        else { D.sweeps[2 * (size_t)w] = sweeps; D.sweeps[2 * (size_t)w + 1] = rowsweeps; }
    }
}
#endif
