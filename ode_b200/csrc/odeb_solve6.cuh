// odeb_solve6.cuh -- k_solve6<P>: the SOR-LCP sweeps of dxQuickStepIsland for worlds that hold SEVERAL islands
// (dxProcessIslands util.cpp:886-938 hands islands out one by one; quickstep.cpp:1823-1856 loop, :2329-2355 ReorderPrep,
//  :2578-2611 random reorder via dRandInt, :2917-3033 IterationStep, :3253-3285 dynamic iteration control).
//
// The reference threads ONE dRand stream (misc.cpp:33-49) through the islands of a world: island i+1 draws its reorders from the
// seed island i left, and how many numbers island i draws depends on how many sweeps it needed (1..40, decided by its own
// convergence; measured on 64-body piles the count of the previous step predicts it for only ~50-85 % of the islands, so
// speculating on it loses).  The islands of one world therefore form a chain, and what runs side by side is WORLDS.  k_solve5
// does that in lock step: the 16/P worlds of a warp walk island index 0, 1, 2, ... together, every island padded to the longest
// schedule and the most sweeps among them, and every serial section (reorder, list scheduling, iteration control) is paid by the
// whole warp.  Here every world of the warp is an independent WALKER with its own (island, sweep, slot) position:
//   * the warp's main loop executes ODEB6_RING schedule slots per trip for all walkers at once (the row update of k_solve5: two
//     lanes per row, accumulators + lambda in shared memory, half records through a per-lane cp.async ring);
//   * a walker whose sweep has ended takes the TRANSITION branch before the next trip; only the lanes of that world run it
//     (group masks in every warp primitive), the others wait at the reconvergence point.  What a transition costs was the
//     whole game (first version: 59 % of the kernel), so everything serial was moved out of it:
//       - ReorderPrep (the island's initial order) is computed for all islands at once by k_reorder_prep; set-up is a copy;
//       - the dRandInt Fisher-Yates pass of the NEXT reorder runs in the shadow of the current sweeps: the world's leader lane
//         performs ODEB6_FYSTEPS insertion steps per trip on a second order buffer (inside-out form of the reference's swap
//         loop: same draws, same result); a reorder commits that buffer (and the seed it reached), an island that finishes
//         earlier simply drops it -- the seed only ever advances by committed reorders, exactly like the reference;
//       - the P-processor schedule is built by all 2P lanes of the world together (window scheduler: the 2P oldest unscheduled
//         rows sit one per lane, a row is ready when no older window row shares a body, up to P ready rows fill a slot) instead
//         of a one-lane list-scheduling loop, and is stored compactly (rows in schedule order + one 16-bit word per slot);
//       - iteration control is a ballot over the island's bodies;
//       - consecutive sweeps over the same schedule need no set-up at all: the slot-info words wrap around, so the prefetch
//         pipeline runs across the sweep boundary and the next sweep starts with its first rows already in the ring.
//   * accumulator slots are relative to the island (dummy slot = NBI), so shared memory follows the largest island, not the world.
// P = 1 (16 worlds per warp, the solve order is the schedule) is the throughput shape for very large batches of small islands,
// P = 2 / 4 shorten the per-world chain.  Results are bit-identical to the sequential sweep (rows in one slot touch disjoint
// bodies and every row follows all earlier rows on its bodies).
#ifndef ODEB_SOLVE6_CUH
#define ODEB_SOLVE6_CUH

#ifndef E5_LIVE                             // idle schedule entries carry a sign bit (real entries end at bit 30): liveness is a sign test
#define E5_IDLE_BIT 0x80000000u
#define E5_LIVE(e) ((int)(e) >= 0)
#endif
#define ODEB6_RING 4
#ifndef ODEB6_FYSTEPS
#define ODEB6_FYSTEPS 1                    // shadow Fisher-Yates steps per trip: each costs ~2.5 % of the kernel on 64-body piles (1: 8.22, 2: 8.37, 4: 8.84 ms per step); the rest is caught up at the reorder
#endif
#define ODEB6_TAIL 16                      // entries behind a schedule: round-up to the ring, wrap-around copies, idle words
#define ODEB6_DUMMY 255u                   // body-2 slot of one-body rows in k_reorder_prep's entries

__host__ __device__ inline size_t odeb6_smem(int P, int nbi, int sr)
{
    const int wpw = 16 / P;
    size_t b = (size_t)ODEB6_RING * ODEB_HALF_CHUNKS * 32 * 16;              // ring
    b += (size_t)2 * (nbi + 1) * wpw * sizeof(Real4);                         // accumulators
    b += (size_t)(sr + 1) * wpw * sizeof(Real);                               // lambda
    b = (b + 15) / 16 * 16;
    if (P > 1) {
        b += (size_t)2 * sr * wpw * sizeof(unsigned);                         // solve order + rows in schedule order
        b += (size_t)(sr + ODEB6_TAIL) * wpw * sizeof(unsigned short);        // slot info
    } else {
        b += (size_t)2 * (sr + ODEB6_TAIL) * wpw * sizeof(unsigned);          // two order buffers (= schedules)
    }
    return (b + 31) / 32 * 32;
}

// ReorderPrep (quickstep.cpp:2329-2355: stable partition, rows without a friction index first) of every island of every world,
// as packed k_solve5 entries with island-relative body slots: one warp per world, ballot-ranked.  isl_done[w, island] = 1 marks
// islands whose friction-index distance does not fit the entry (serial path).  Also writes the world's sort key (see below).
__global__ void __launch_bounds__(128) k_reorder_prep(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D)
{
    const int w = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= P.W) return;
    const unsigned below = (1u << lane) - 1u;
    const int *findex = D.findex + (size_t)w * P.MR;
    const int2 *rbody = D.rbody + (size_t)w * P.MR;
    unsigned *out = D.order0 + (size_t)w * P.MR;
    const int4 *iinfo = D.island_info + (size_t)w * P.NB;
    const int nis = D.nislands[w];
    int est = 0, mmax = 0;
    for (int is = 0; is < nis; is++) {
        const int4 info = iinfo[is];
        const int bstart = info.x, rstart = info.z, m = info.w;
        if (m == 0) continue;
        est += m; mmax = m > mmax ? m : mmax;
        int nfree = 0;
        for (int i0 = 0; i0 < m; i0 += 32) {
            const int i = i0 + lane;
            nfree += __popc(__ballot_sync(ODEB_FULL, i < m && findex[rstart + i] == -1));
        }
        int head = 0, tail = nfree;
        bool bad = false;
        for (int i0 = 0; i0 < m; i0 += 32) {
            const int i = i0 + lane;
            const bool in = i < m;
            int fi = -1; int2 rb = make_int2(bstart, P.NB);
            if (in) { fi = findex[rstart + i]; rb = rbody[rstart + i]; }
            const bool fr = in && fi == -1;
            const unsigned bf = __ballot_sync(ODEB_FULL, fr), bo = __ballot_sync(ODEB_FULL, in && !fr);
            const unsigned dlt = fi == -1 ? 0u : (unsigned)(i - (fi - rstart));
            if (in && fi != -1 && dlt - 1u > 6u) bad = true;                 // the packed entry holds row - findex = 1..7
            const unsigned s2 = rb.y == P.NB ? ODEB6_DUMMY : (unsigned)(rb.y - bstart);
            const unsigned e = (unsigned)i | ((dlt & 7u) << 12) | (((unsigned)(rb.x - bstart) & 0xffu) << 15) | ((s2 & 0xffu) << 23);
            if (fr) out[rstart + head + __popc(bf & below)] = e;
            else if (in) out[rstart + tail + __popc(bo & below)] = e;
            head += __popc(bf); tail += __popc(bo);
        }
        bad = __any_sync(ODEB_FULL, bad);
        if (lane == 0) D.isl_done[(size_t)w * P.NB + is] = bad ? 1 : 0;
    }
    // work estimate: rows, the largest island counted twice (large islands run the full 40 sweeps)
    if (lane == 0) D.wkey[w] = (unsigned)(est + mmax);
}

// Worlds in descending order of estimated work (counting sort over 1024 buckets, one block): the heaviest worlds start first
// (shorter tail of the solver launch) and worlds of similar work share a warp (a warp lasts as long as its slowest world).
// The order inside a bucket is whatever the atomics produce: it only moves worlds between warps, never changes a result.
__global__ void __launch_bounds__(1024) k_world_sort(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D)
{
    __shared__ unsigned cnt[1024];
    __shared__ unsigned wsum[32];
    const int t = threadIdx.x;
    const unsigned scale = 2u * (unsigned)P.MR + 1u;
    cnt[t] = 0;
    __syncthreads();
    for (int w = t; w < P.W; w += 1024) {
        const unsigned e = D.wkey[w] < scale ? D.wkey[w] : scale - 1u;
        atomicAdd(&cnt[1023u - (unsigned)(((unsigned long long)e * 1023ull) / scale)], 1u);
    }
    __syncthreads();
    // exclusive scan of the 1024 counters
    const unsigned c = cnt[t];
    unsigned v = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const unsigned u = __shfl_up_sync(0xffffffffu, v, d); if ((t & 31) >= d) v += u; }
    if ((t & 31) == 31) wsum[t >> 5] = v;
    __syncthreads();
    if (t < 32) {
        unsigned x = wsum[t];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const unsigned u = __shfl_up_sync(0xffffffffu, x, d); if (t >= d) x += u; }
        wsum[t] = x;
    }
    __syncthreads();
    cnt[t] = v - c + ((t >> 5) ? wsum[(t >> 5) - 1] : 0u);
    __syncthreads();
    for (int w = t; w < P.W; w += 1024) {
        const unsigned e = D.wkey[w] < scale ? D.wkey[w] : scale - 1u;
        D.wlist[atomicAdd(&cnt[1023u - (unsigned)(((unsigned long long)e * 1023ull) / scale)], 1u)] = w;
    }
}

#if defined(ODEB6_PROF)   // cycle accounting (experiment builds, tools/build_variant.sh): [0] total, [1] transition wall (warp), [2] trips, [3] trips with a
// transition, [4] control + reorder commit, [5] Fisher-Yates catch-up steps, [6] island fetch, [7] schedule, [8] prime, [9] transitions, [10] serial fallbacks,
// [11] warps, [12] schedule slots built, [13] rows scheduled
__device__ unsigned long long odeb6_prof[16];
#define PROF_T(v) const long long v = clock64()
#define PROF_ADD(i, x) pf_[i] += (x)
#else
#define PROF_T(v)
#define PROF_ADD(i, x)
#endif

// One schedule slot.  EK: the entry executed now, refilled at the end with the entry RING positions further on; EC: the entry
// RING-1 positions ahead, whose half record is requested now.
#define ODEB6_SLOT(CUR, NXT, EK, EC, K)                                                                                  \
    {                                                                                                                    \
        const unsigned MT = EK;                                                                                          \
        const int index = E5_ROW(MT), fi = E5_FI(MT);                                                                    \
        const bool live = E5_LIVE(MT);                                                                                   \
        const int bs = (int)((MT >> bsh) & 0xffu);                      /* this lane's body: bits 15.. (side 0) or 23.. (side 1) */ \
        const Real old_lambda = lam[index * WPW];                                                                        \
        const Real lam_fi = lam[fi * WPW];                                                                               \
        /* CF5(bs, 0) / CF5(bs, 1) by byte offset: vector 1 is vector 0's slot with one bit flipped */                     \
        const unsigned cfo = ((unsigned)bs * (2u * WPW) + (((unsigned)bs & 1u) * WPW)) * (unsigned)sizeof(Real4);       \
        Real4 *const cfa = (Real4 *)((char *)cf + cfo), *const cfb = (Real4 *)((char *)cf + (cfo ^ (WPW * (unsigned)sizeof(Real4)))); \
        Real4 fa = *cfa, fb = *cfb;                                                                                      \
        {                                                                                                                \
            if (E5_LIVE(EC)) {                                                                                           \
                const char *src = rec_base + (size_t)E5_ROW(EC) * (sizeof(Real) * 32);                                   \
                const unsigned dst = ring_addr + (unsigned)((((K) + ODEB6_RING - 1) & (ODEB6_RING - 1)) * CH * 32 * 16); \
                _Pragma("unroll") for (int c = 0; c < CH; c++) cp_async16(dst + c * 32 * 16, src + c * 16);             \
            }                                                                                                            \
            cp_async_commit();                                                                                           \
        }                                                                                                                \
        if (P > 1) {                                                                                                     \
            EK = (proc < (int)(si & 15u)) ? sord[((si >> 4) + proc) * WPW] : IDLE;                                       \
            si = sip[(K) * WPW];                                                                                         \
        } else EK = ep[(K) * 16];                                                                                        \
        const Real lo_b = __shfl_xor_sync(ODEB_FULL, CUR.q1.z, PARTNER);  /* lane A receives lo from lane B */           \
        const Real s = fa.x * CUR.q0.x + fa.y * CUR.q0.y + fa.z * CUR.q0.z + fa.w * CUR.q0.w + fb.x * CUR.q1.x + fb.y * CUR.q1.y; \
        const Real ta = (CUR.q1.z - old_lambda * CUR.q1.w) - s;         /* lane A: (rhs - lambda*cfm) - s1 */           \
        const Real mine = side ? s : ta;                                                                                 \
        const Real other = __shfl_xor_sync(ODEB_FULL, mine, PARTNER);                                                    \
        cp_async_wait<ODEB6_RING - 2>();                                                                                 \
        load_half(NXT, ring + (size_t)(((K) + 1) & (ODEB6_RING - 1)) * CH * 32);                                        \
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
            *cfa = fa; *cfb = fb;                                                                                       \
        }                                                                                                                \
        __syncwarp();                                                                                                    \
    }

// one insertion step of the shadow Fisher-Yates pass (leader lane): the reference's `swap(order[idx], order[dRandInt(idx + 1)])`
// (quickstep.cpp:2596-2606) in inside-out form -- order[idx] is still the element the previous order holds there
// (P = 1, where the sweeps read the order itself, so the pass fills a second buffer); P > 1: the sweeps read the schedule, the
// order is idle between two schedule builds and the reference's swap runs on it in place
#define ODEB6_FY_STEP()                                                                                                  \
    {                                                                                                                    \
        const int sw = odeb_rand_int(&fy_seed, fy_idx + 1);                                                              \
        const unsigned v = oc[fy_idx * OST];                                                                             \
        if (P == 1) {                                                                                                    \
            const unsigned t = on[sw * OST];                                                                             \
            on[fy_idx * OST] = (sw == fy_idx) ? v : t;                                                                   \
            on[sw * OST] = v;                                                                                            \
        } else {                                                                                                         \
            const unsigned t = oc[sw * OST];                                                                             \
            oc[fy_idx * OST] = t;                                                                                        \
            oc[sw * OST] = v;                                                                                            \
        }                                                                                                                \
        fy_idx++;                                                                                                        \
    }

template <int P>
__global__ void __launch_bounds__(32, 1) k_solve6_t(const __grid_constant__ DevParams Pm, const __grid_constant__ DevPtrs D, const int SR6, const int NBI)
{
    extern __shared__ __align__(32) unsigned char smem[];
    constexpr int WPW = 16 / P;
    constexpr int CH = ODEB_HALF_CHUNKS;
    constexpr int PARTNER = WPW;
    constexpr int G = 2 * P;                                      // lanes per world
    constexpr int OST = WPW;                                      // stride of the per-world arrays
    const int lane = threadIdx.x;
    const int wl = lane & (WPW - 1), side = (lane / WPW) & 1, proc = lane / (2 * WPW);
    const int wid = lane / WPW;                                   // index of the lane among its world's 2P lanes
    const unsigned gmask = ((WPW == 1) ? 0xffffffffu : (WPW == 2) ? 0x55555555u : (WPW == 4) ? 0x11111111u : (WPW == 8) ? 0x01010101u : 0x00010001u) << wl;
    const bool leader = wid == 0;
    const int NBd = NBI;                                          // dummy accumulator slot
    const int wraw = blockIdx.x * WPW + wl;
    const bool valid = wraw < Pm.W;
    const int w = D.wlist[valid ? wraw : Pm.W - 1];               // worlds in descending order of estimated work

    // ---- shared memory
    uint4 *ring = (uint4 *)smem + lane;                           // chunk (stage, c) at ring[(stage * CH + c) * 32]
    const unsigned ring_addr = (unsigned)__cvta_generic_to_shared(ring);
    unsigned char *p = smem + (size_t)ODEB6_RING * CH * 32 * 16;
    Real4 *cf = (Real4 *)p + wl;    p += (size_t)2 * (NBI + 1) * WPW * sizeof(Real4);           // CF5(slot, j)
    Real *lam = (Real *)p + wl;     p += (size_t)(SR6 + 1) * WPW * sizeof(Real);                // lam[i * WPW]
    p = smem + ((size_t)(p - smem) + 15) / 16 * 16;
    const int OCAP = (P > 1) ? SR6 : SR6 + ODEB6_TAIL;            // entries per order buffer
    unsigned *oc = (unsigned *)p + wl;                            // current solve order, oc[i * OST]
    unsigned *on = oc + (size_t)OCAP * WPW;                       // P = 1: order of the next reorder (shadow Fisher-Yates)
    p += (size_t)(P > 1 ? 1 : 2) * OCAP * WPW * sizeof(unsigned);
    unsigned *sord = (unsigned *)p + wl;                          // P > 1: rows in schedule order
    unsigned short *sinfo = (unsigned short *)(p + (size_t)SR6 * WPW * sizeof(unsigned)) + wl;   // P > 1: slot s = (first entry << 4) | rows
    const unsigned below = (1u << lane) - 1u;
    const unsigned IDLE = ((unsigned)NBd << 15) | ((unsigned)NBd << 23) | E5_IDLE_BIT;
    const int bsh = side ? 23 : 15;
    const int IDLE_AT = SR6 + 8;                                  // [IDLE_AT, IDLE_AT + 8): idle for good (finished worlds park here)

    unsigned seed = D.seed[w];
    unsigned st0 = 0, st1 = 0, st2 = 0, st3 = 0;
    unsigned long long sweeps = 0, rowsweeps = 0;
    const Real4 *rows = D.rows + (size_t)w * Pm.MR * 8;
    Real4 *cf_out = D.cforce + (size_t)w * (Pm.NB + 1) * 2;
    const int4 *iinfo = D.island_info + (size_t)w * Pm.NB;
    const unsigned *order0 = D.order0 + (size_t)w * Pm.MR;
    const int nis = valid ? D.nislands[w] : 0;
    {
        const Real4 z4 = { 0, 0, 0, 0 };
        if (wid < 2) CF5(NBd, wid) = z4;                          // the dummy slot is read by idle lanes: keep it finite
        if (leader) {
            lam[0] = 0;
            for (int s = 0; s < 8; s++) { if (P > 1) sinfo[(IDLE_AT + s) * WPW] = 0; else { oc[(IDLE_AT + s) * OST] = IDLE; on[(IDLE_AT + s) * OST] = IDLE; } }
        }
    }
    __syncwarp();

    // ---- walker state (uniform over the world's lanes unless noted)
    int is = 0;                                                   // next island to fetch
    bool wdone = false, active = false;
    int bstart = 0, nb = 0, rstart = 0, m_own = 0;
    unsigned iteration = 0, extra = 0;
    Real exit_delta = Pm.premature_delta;
    int nslots = 0, spos = 0;
    int fy_idx = 0, fy_m = 0; unsigned fy_seed = 0;               // shadow Fisher-Yates (leader): next position, rows, working seed
    const char *rec_base = (const char *)rows + side * (sizeof(Real) * 16);
    const unsigned short *sip = sinfo + IDLE_AT * WPW;            // P > 1: slot info RING + 1 positions ahead of the trip
    unsigned si = 0;                                              //        slot info RING positions ahead
    const unsigned *ep = oc + IDLE_AT * OST;                      // P = 1: entry RING positions ahead of the trip
    HalfRegs r0, r1;
    unsigned e0 = IDLE, e1 = IDLE, e2 = IDLE, e3 = IDLE;          // entries of the trip's four slots
    {
        const Real4 z4 = { 0, 0, 0, 0 };
        r0.q0 = r0.q1 = r0.q2 = r0.q3 = z4; r1 = r0;
    }
#if defined(ODEB6_PROF)
    long long pf_[14] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 };
    const long long pf_start = clock64();
#endif

    for (;;) {
        PROF_T(t_a);
        const bool tr = !wdone && spos >= nslots;                 // this world's sweep has ended (or it has not started yet)
        const bool any_tr = __any_sync(ODEB_FULL, tr);
        if (any_tr) {
            // ================= transitions =================
            // Every warp primitive below is executed by the converged warp with the full mask (lanes of worlds that are not in a
            // transition carry false predicates): partial-mask shuffles and votes inside a divergent branch go through the
            // convergence-barrier slow path (WARPSYNC.COLLECTIVE per call), which made the first window scheduler cost ~2400
            // cycles per slot.  Divergent sections in between hold plain loads and stores only.
            bool fresh = false;                                   // a new order is in oc: schedule it and prime the pipeline
            PROF_T(t0);
            // ---- A: end of a sweep: iteration control quickstep.cpp:1832-1855, :3253-3285
            const bool ctl = tr && active;
            bool fin = false, dyn = false, hit = false, reord = false;
            if (ctl) {
                ++iteration;
                if (leader) { ++sweeps; rowsweeps += m_own; }
                if (iteration - extra == Pm.num_iter) {
                    if (extra != 0 || Pm.max_extra == 0) { if (extra != 0 && leader) st3++; fin = true; }
                    else { extra = Pm.max_extra; exit_delta = Pm.extra_delta; }
                }
                if (!fin && Pm.dyn_enabled) {
                    dyn = true;
                    hit = (exit_delta == 0);
                    // the max-adjustment pairs of up to 4 bodies per lane are loaded before any of them is reset (independent loads
                    // instead of a load -> store chain per body)
                    for (int k0 = wid; k0 < nb; k0 += 4 * G) {
                        Real4 v[4];
#pragma unroll
                        for (int j = 0; j < 4; j++) if (k0 + j * G < nb) v[j] = CF5(k0 + j * G, 1);
#pragma unroll
                        for (int j = 0; j < 4; j++) if (k0 + j * G < nb) {
                            if (!(v[j].w < exit_delta) || !(-v[j].z < exit_delta)) hit = true;
                            v[j].z = 0; v[j].w = 0;
                            CF5(k0 + j * G, 1) = v[j];
                        }
                    }
                }
            }
            const unsigned hb = __ballot_sync(ODEB_FULL, hit) & gmask;
            if (dyn && hb == 0) {
                if (leader) { if (iteration < Pm.num_iter) st1++; else if (iteration > Pm.num_iter) st2++; }
                fin = true;
            }
            if (ctl) {
                if (fin) {
                    // ---- island finished: results to HBM; the shadow reorder (and the draws it made) is dropped
                    for (int k = wid; k < 2 * nb; k += G) cf_out[2 * bstart + k] = CF5(k >> 1, k & 1);
                    if (D.jcopy) for (int i = wid; i < m_own; i += G) D.lambda[(size_t)w * Pm.MR + rstart + i] = lam[i * WPW];   // joint feedback
                    active = false; fy_m = 0;
                } else if (iteration >= 8 && (iteration & 7) == 0) {
                    // ---- ConstraintsShuffling quickstep.cpp:2578-2611 with dRandInt misc.cpp:78-139: commit the shadow pass
                    if (leader) {
                        while (fy_idx < fy_m) { ODEB6_FY_STEP() PROF_ADD(5, 1); }
                        seed = fy_seed;
                    }
                    reord = true;
                } else {
                    // ---- the same schedule again: the pipeline already holds its first slots (wrap-around copies)
                    spos = 0;
                    if (P > 1) sip = sinfo + (ODEB6_RING + 1) * WPW; else ep = oc + ODEB6_RING * OST;
                }
            }
            __syncwarp();
            if (reord) { if (P == 1) { unsigned *t = oc; oc = on; on = t; } fresh = true; }
            PROF_T(t1);
            // ---- B: fetch the next island that has rows
            while (__any_sync(ODEB_FULL, tr && !active && !wdone)) {
                if (tr && !active && !wdone) {
                    if (is >= nis) wdone = true;
                    else {
                        const int4 info = iinfo[is];
                        const int bad = D.isl_done[(size_t)w * Pm.NB + is];
                        is++;
                        if (leader) st0++;
                        bstart = info.x; nb = info.y; rstart = info.z; m_own = info.w;
                        if (m_own != 0) {
                            if (m_own > SR6 || nb > NBI || bad) {
                                // islands beyond the shared-memory budget, or with a friction index the packed entry cannot hold: serial path
                                PROF_ADD(10, 1);
                                if (leader) solve_island_serial(Pm, D, w, bstart, nb, rstart, m_own, seed, st1, st2, st3, sweeps, rowsweeps);
                            } else {
                                const Real4 z4 = { 0, 0, 0, 0 };
                                for (int k = wid; k < 2 * nb; k += G) CF5(k >> 1, k & 1) = z4;
                                for (int i = wid; i <= m_own; i += G) lam[i * WPW] = 0;
                                for (int i = wid; i < m_own; i += G) {             // ReorderPrep order from k_reorder_prep
                                    unsigned e = order0[rstart + i];
                                    if ((e >> 23) == ODEB6_DUMMY) e = (e & 0x007fffffu) | ((unsigned)NBd << 23);
                                    oc[i * OST] = e;
                                }
                                rec_base = (const char *)(rows + (size_t)rstart * 8) + side * (sizeof(Real) * 16);
                                iteration = 0; extra = 0; exit_delta = Pm.premature_delta;
                                active = true; fresh = true;
                            }
                        }
                    }
                }
                __syncwarp();
            }
            PROF_T(t2);
            // ---- C: schedule the new order and prime the pipeline
            const bool sch = tr && active && fresh;
            if (__any_sync(ODEB_FULL, sch)) {
                if (sch) {
                    cp_async_wait<0>();                           // prefetches of the old schedule's first slots may still be landing in the ring
                    if (leader) { fy_idx = 1; fy_m = m_own; fy_seed = seed; if (P == 1) on[0] = oc[0]; }   // the next reorder starts from this order and this seed
                }
                if (P > 1) {
                    // window scheduler: the 2P oldest unscheduled rows of the world sit one per lane, each with its order position q.
                    // A row is ready when no older window row shares a body; the (up to) P oldest ready rows fill the slot; freed
                    // lanes take the next rows of the order.  Branch-free: every lane broadcasts one word {body1, body2, q} and
                    // compares the four body combinations of each pair with one SIMD byte compare (one-body rows carry the
                    // sentinels 0xfe on the testing / 0xff on the tested side, so two dummies never match; real slots are <= 253).
                    int next = 0, slot = 0, placed = 0;
                    unsigned we = IDLE; bool have = false; int q = 0;
                    const bool chain = sch && nb == 1;              // one body: every row follows the previous one, the order is the schedule
                    if (chain) {
                        for (int i = wid; i < m_own; i += G) { sord[i * WPW] = oc[i * OST]; sinfo[i * WPW] = (unsigned short)((i << 4) | 1); }
                        slot = m_own; placed = m_own;
                    }
                    if (sch && !chain) { if (wid < m_own) { we = oc[wid * OST]; have = true; q = wid; } next = m_own < G ? m_own : G; }
                    for (;;) {
                        const unsigned hv = __ballot_sync(ODEB_FULL, have);
                        if (hv == 0) break;
                        const bool gl = (hv & gmask) != 0;          // this world is still scheduling
                        const unsigned b1 = (we >> 15) & 0xffu, b2 = (we >> 23) & 0xffu;
                        const unsigned b2m = b2 == (unsigned)NBd ? 0xfeu : b2;
                        const unsigned mine4 = (b1 | (b2m << 8)) * 0x00010001u;                            // [b1, b2, b1, b2]
                        const unsigned key = have ? (b1 | ((b2 == (unsigned)NBd ? 0xffu : b2) << 8) | ((unsigned)q << 16)) : 0xffffffffu;
                        unsigned oq[G - 1];
                        bool conflict = false;
#pragma unroll
                        for (int r = 1; r < G; r++) {
                            const unsigned ok = __shfl_sync(ODEB_FULL, key, ((wid + r) & (G - 1)) * WPW + wl);
                            oq[r - 1] = ok >> 16;
                            const unsigned theirs4 = __byte_perm(ok, 0, 0x1100);                           // [o1, o1, o2, o2]
                            conflict |= (__vcmpeq4(mine4, theirs4) != 0u) & (oq[r - 1] < (unsigned)q);
                        }
                        const bool ready = have && !conflict;
                        const unsigned rd = __ballot_sync(ODEB_FULL, ready);
                        int rank = 0;
#pragma unroll
                        for (int r = 1; r < G; r++) rank += (int)((rd >> (((wid + r) & (G - 1)) * WPW + wl)) & 1u) & (int)(oq[r - 1] < (unsigned)q);
                        const bool take = ready && rank < P;
                        const int ntake = min(__popc(rd & gmask), P);
                        if (take) sord[(placed + rank) * WPW] = we;
                        if (leader && gl) sinfo[slot * WPW] = (unsigned short)((placed << 4) | ntake);
                        if (gl) { placed += ntake; slot++; }
                        const bool freed = gl && (take || !have);
                        const unsigned fr = __ballot_sync(ODEB_FULL, freed) & gmask;
                        if (freed) {
                            q = next + __popc(fr & below);
                            have = q < m_own;
                            we = have ? oc[q * OST] : IDLE;
                        }
                        if (gl) next = min(m_own, next + __popc(fr));
                    }
                    __syncwarp();
                    if (sch) {
                        nslots = slot;
                        PROF_ADD(12, slot); PROF_ADD(13, placed);
                        if (leader) {
                            const int nr = (slot + ODEB6_RING - 1) & ~(ODEB6_RING - 1);
                            for (int s = slot; s < nr; s++) sinfo[s * WPW] = 0;
                            for (int j = 0; j <= ODEB6_RING; j++) sinfo[(nr + j) * WPW] = sinfo[j * WPW];     // wrap-around copies
                        }
                    }
                } else if (sch) {
                    nslots = m_own;
                    if (leader) {
                        const int nr = (m_own + ODEB6_RING - 1) & ~(ODEB6_RING - 1);
                        for (int s = m_own; s < nr; s++) oc[s * OST] = IDLE;
                        for (int j = 0; j < ODEB6_RING; j++) oc[(nr + j) * OST] = oc[j * OST];            // wrap-around copies
                    }
                }
                __syncwarp();
                PROF_T(t3);
                PROF_ADD(7, t3 - t2);
                if (sch) {
                    // ---- prime: entries of slots 0..3, half records of slots 0..2 on their way, slot 0's in registers
                    spos = 0;
                    if (P > 1) {
                        const unsigned s0 = sinfo[0], s1 = sinfo[WPW], s2 = sinfo[2 * WPW], s3 = sinfo[3 * WPW];
                        e0 = (proc < (int)(s0 & 15u)) ? sord[((s0 >> 4) + proc) * WPW] : IDLE;
                        e1 = (proc < (int)(s1 & 15u)) ? sord[((s1 >> 4) + proc) * WPW] : IDLE;
                        e2 = (proc < (int)(s2 & 15u)) ? sord[((s2 >> 4) + proc) * WPW] : IDLE;
                        e3 = (proc < (int)(s3 & 15u)) ? sord[((s3 >> 4) + proc) * WPW] : IDLE;
                        si = sinfo[ODEB6_RING * WPW];
                        sip = sinfo + (ODEB6_RING + 1) * WPW;
                    } else {
                        e0 = oc[0]; e1 = oc[OST]; e2 = oc[2 * OST]; e3 = oc[3 * OST];
                        ep = oc + ODEB6_RING * OST;
                    }
                    const unsigned pe[3] = { e0, e1, e2 };
#pragma unroll
                    for (int k = 0; k < ODEB6_RING - 1; k++) {
                        if (E5_LIVE(pe[k])) {
                            const char *src = rec_base + (size_t)E5_ROW(pe[k]) * (sizeof(Real) * 32);
                            const unsigned dst = ring_addr + (unsigned)(k * CH * 32 * 16);
#pragma unroll
                            for (int c = 0; c < CH; c++) cp_async16(dst + c * 32 * 16, src + c * 16);
                        }
                        cp_async_commit();
                    }
                    cp_async_wait<ODEB6_RING - 2>();
                    load_half(r0, ring);
                }
                PROF_T(t4);
                PROF_ADD(8, t4 - t3);
            }
            if (tr && wdone) {
                // ---- the world is finished: its lanes idle on the parked entries from now on
                e0 = e1 = e2 = e3 = IDLE; si = 0;
                sip = sinfo + IDLE_AT * WPW; ep = oc + IDLE_AT * OST;
                nslots = 0; spos = 0; fy_m = 0;
            }
            PROF_ADD(4, t1 - t0); PROF_ADD(6, t2 - t1); if (tr) PROF_ADD(9, 1);
        }
        __syncwarp();
#if defined(ODEB6_PROF)
        { const long long t_b = clock64(); pf_[2] += 1; if (any_tr) { pf_[1] += t_b - t_a; pf_[3] += 1; } }
#endif
        if (__all_sync(ODEB_FULL, wdone)) break;
        ODEB6_SLOT(r0, r1, e0, e3, 0)
        ODEB6_SLOT(r1, r0, e1, e0, 1)
        ODEB6_SLOT(r0, r1, e2, e1, 2)
        ODEB6_SLOT(r1, r0, e3, e2, 3)
        if (!wdone) { spos += ODEB6_RING; if (P > 1) sip += ODEB6_RING * WPW; else ep += ODEB6_RING * OST; }
        // ---- the next reorder's Fisher-Yates pass advances in the shadow of the sweeps (leader lanes)
#pragma unroll
        for (int t = 0; t < ODEB6_FYSTEPS; t++) if (leader && fy_idx < fy_m) ODEB6_FY_STEP()
    }
    cp_async_wait<0>();
#if defined(ODEB6_PROF)
    pf_[0] = clock64() - pf_start;
    if (lane == 0) { atomicMax(&odeb6_prof[14], (unsigned long long)pf_[0]); atomicAdd(&odeb6_prof[0], (unsigned long long)pf_[0]); atomicAdd(&odeb6_prof[1], (unsigned long long)pf_[1]); atomicAdd(&odeb6_prof[2], (unsigned long long)pf_[2]); atomicAdd(&odeb6_prof[3], (unsigned long long)pf_[3]); atomicAdd(&odeb6_prof[11], 1ull); }
    if (leader) for (int i = 4; i < 14; i++) if (i != 11) atomicAdd(&odeb6_prof[i], (unsigned long long)pf_[i]);
#endif
    if (leader && valid) {
        D.seed[w] = seed;
        unsigned *st = D.stats + 4 * (size_t)w;
        st[0] += st0; st[1] += st1; st[2] += st2; st[3] += st3;
        D.sweeps[2 * (size_t)w] = sweeps; D.sweeps[2 * (size_t)w + 1] = rowsweeps;
    }
}
#endif
