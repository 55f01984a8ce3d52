// odeb_solve_bl.cuh -- k_solve_bl<G>: the SOR-LCP sweeps of dxQuickStepIsland with ONE LANE PER BODY
// (quickstep.cpp:1823-1856 loop, :2329-2355 ReorderPrep, :2578-2611 random reorder via dRandInt,
//  :2917-3033 IterationStep, :3253-3285 dynamic iteration control).
//
// The reference sweeps the rows of an island strictly one after the other.  The only data a row shares with
// other rows are the constraint-force accumulators of its two bodies (and lambda of its own contact), so a row
// may run as soon as every earlier row *on its two bodies* has run: the result is bit-identical to the sequential
// sweep.  This kernel executes that dataflow directly:
//   * a world owns G = 8/16/32 consecutive lanes of a warp, lane b = body b of the island being solved; the
//     body's accumulators (cforce 6 + max-adjustment 2) live in that lane's REGISTERS for the whole solve;
//   * every lane walks its own list of rows (the rows that touch its body, in the current solve order, kept
//     in shared memory).  In each lockstep iteration a lane looks at its next row, asks the partner lane (the
//     row's other body, one shuffle) whether it has arrived at the same row, and if so both execute it:
//     each forms its 6-term dot product in the reference's left-to-right order, the halves meet through one
//     shuffle (delta = ((rhs - lambda*cfm) - s1) - s2, the reference's association), both clamp redundantly
//     and update their own accumulators.  Rows on disjoint bodies execute in the same iteration, so a sweep takes
//     "dataflow depth" iterations instead of m (16-box stack: 136 for the initial order, ~45 after a shuffle,
//     instead of 192);
//   * the row half-records stream from L2/HBM through a per-lane cp.async ring, prefetched along the lane's
//     own (cyclic) list, so the gather never sits on the dependency chain;
//   * the dRand-driven Fisher-Yates reorder at sweeps 8,16,.. is replayed on the world's order array by lane 0,
//     then the per-body lists are rebuilt by all lanes.
// Islands of a world are solved one after the other (the dRand seed threads through them in the reference's
// order); islands beyond the shared-memory row budget take the serial global-memory path.
#ifndef ODEB_SOLVE_BL_CUH
#define ODEB_SOLVE_BL_CUH

#ifndef ODEB_BL_RING
#define ODEB_BL_RING 3
#endif
#define ODEB_BL_CH ODEB_HALF_CHUNKS                        // 16-byte chunks of a half record
#define ODEB_BL_CHQ ((int)(sizeof(Real) * 4 / 16))         // chunks of one Real4
#define ODEB_BL_CHS (ODEB_BL_CH + ODEB_BL_CHQ)             // + the body-2 half's {.., .., lo, hi} vector
#define ODEB_BL_SENT 0xffffu
#define ODEB_BL_MAXROWS 1022

// bytes of shared memory per world / per warp for a row budget of sr rows per world
__host__ __device__ inline size_t odeb_bl_world_bytes(int sr)
{
    const size_t b = (size_t)(sr + 1) * sizeof(Real) + (size_t)sr * sizeof(unsigned) + (size_t)2 * sr * sizeof(unsigned) + (size_t)sr * sizeof(unsigned short);
    return (b + 15) / 16 * 16;
}
__host__ __device__ inline size_t odeb_bl_smem(int G, int sr)
{
    return (size_t)ODEB_BL_RING * ODEB_BL_CHS * 32 * 16 + (size_t)(32 / G) * odeb_bl_world_bytes(sr);
}

// serial global-memory solve of one island (lane 0 of the world's group)
__device__ __noinline__ void solve_island_serial(const DevParams &P, const DevPtrs &D, int w, int bstart, int nb, int rstart, int m,
                                                 unsigned &seed, unsigned &st1, unsigned &st2, unsigned &st3,
                                                 unsigned long long &sweeps, unsigned long long &rowsweeps)
{
    const Real4 *rows = D.rows + (size_t)w * P.MR * 8;
    const int *findex = D.findex + (size_t)w * P.MR;
    const int2 *rbody = D.rbody + (size_t)w * P.MR;
    Real4 *cf_out = D.cforce + (size_t)w * (P.NB + 1) * 2;
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


// list entry: side (bit 0) | row (1..10) | partner lane in the group (11..15) | friction-index row (16..25); bit 31 = no entry.
// (entry & 0x7ff) * HALF_BYTES is the byte offset of the lane's half record inside the island's row block.
#define ODEB_BL_NONE 0x80000000u
#define ODEB_FULLMASK 0xffffffffu

template <int G>
__global__ void __launch_bounds__(32) k_solve_bl(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D, const int SRB)
{
    extern __shared__ __align__(32) unsigned char smem[];
    constexpr int WPW = 32 / G;
    constexpr int RING = ODEB_BL_RING, CH = ODEB_BL_CH, CHQ = ODEB_BL_CHQ, CHS = ODEB_BL_CHS;
    constexpr unsigned HALF_BYTES = sizeof(Real) * 16;
    constexpr int STAGE_BYTES = CHS * 32 * 16;
    constexpr unsigned GBITS = (G == 32) ? 0xffffffffu : ((1u << (G & 31)) - 1u);
    const int lane = threadIdx.x;
    const int gl = lane & (G - 1);                       // lane inside the world's group = body slot inside the island
    const int gbase = lane & ~(G - 1);
    const unsigned below = (1u << gl) - 1u;
    const int wraw = blockIdx.x * WPW + (lane / G);
    const bool valid = wraw < P.W;
    const int w = valid ? wraw : P.W - 1;

    // ---- shared memory: [ring | per world: lam, rmeta, list, order]
    uint4 *ring = (uint4 *)smem + lane;                  // chunk (stage, c) at ring[(stage * CHS + c) * 32]
    const unsigned ring_addr = (unsigned)__cvta_generic_to_shared(ring);
    unsigned char *p = smem + (size_t)RING * CHS * 32 * 16 + odeb_bl_world_bytes(SRB) * (size_t)(lane / G);
    Real *lam = (Real *)p;              p += (size_t)(SRB + 1) * sizeof(Real);
    unsigned *rmeta = (unsigned *)p;    p += (size_t)SRB * sizeof(unsigned);        // b1 | b2 << 5 (32 = none) | friction row << 11
    unsigned *list = (unsigned *)p;     p += (size_t)2 * SRB * sizeof(unsigned);
    unsigned short *order = (unsigned short *)p;

    unsigned seed = D.seed[w];
    unsigned st0 = 0, st1 = 0, st2 = 0, st3 = 0;
    unsigned long long sweeps = 0, rowsweeps = 0;
    const Real4 *rows = D.rows + (size_t)w * P.MR * 8;
    const int *findex = D.findex + (size_t)w * P.MR;
    const int2 *rbody = D.rbody + (size_t)w * P.MR;
    Real4 *cf_out = D.cforce + (size_t)w * (P.NB + 1) * 2;
    const int4 *iinfo = D.island_info + (size_t)w * P.NB;
    const int nis = valid ? D.nislands[w] : 0;
    const int nis_max = __reduce_max_sync(ODEB_FULLMASK, nis);

    // All control flow below is warp-uniform (trip counts are maxima over the worlds of the warp, a world that has
    // nothing to do idles with its stores predicated off), so every shuffle / vote uses the full mask.
    for (int is = 0; is < nis_max; is++) {
        int4 info = make_int4(0, 0, 0, 0);
        if (is < nis) { info = iinfo[is]; st0++; }
        const int bstart = info.x, nb = info.y, rstart = info.z, m = info.w;
        const bool big = m > SRB;
        if (big && gl == 0) solve_island_serial(P, D, w, bstart, nb, rstart, m, seed, st1, st2, st3, sweeps, rowsweeps);
        __syncwarp();
        seed = __shfl_sync(ODEB_FULLMASK, seed, gbase);
        const int m_own = (m > 0 && !big) ? m : 0;
        const int m_max = __reduce_max_sync(ODEB_FULLMASK, m_own);
        if (m_max == 0) continue;

        // ---- island set-up: row metadata, lambda = 0, ReorderPrep (rows without a friction index first, stable)
        int nfree = 0;
        for (int i0 = 0; i0 < m_max; i0 += G) {
            const int i = i0 + gl;
            bool fr = false;
            if (i < m_own) {
                const int fi = findex[rstart + i];
                const int2 rb = rbody[rstart + i];
                const unsigned b1 = (unsigned)(rb.x - bstart), b2 = rb.y == P.NB ? 32u : (unsigned)(rb.y - bstart);
                rmeta[i] = b1 | (b2 << 5) | ((unsigned)(fi == -1 ? i : fi - rstart) << 11);
                lam[i] = 0;
                fr = fi == -1;
            }
            nfree += __popc((__ballot_sync(ODEB_FULLMASK, fr) >> gbase) & GBITS);
        }
        __syncwarp();
        {
            int head = 0, tail = nfree;
            for (int i0 = 0; i0 < m_max; i0 += G) {
                const int i = i0 + gl;
                const bool in = i < m_own;
                const bool fr = in && (rmeta[i] >> 11) == (unsigned)i;
                const unsigned bf = (__ballot_sync(ODEB_FULLMASK, fr) >> gbase) & GBITS;
                const unsigned bo = (__ballot_sync(ODEB_FULLMASK, in && !fr) >> gbase) & GBITS;
                if (fr) order[head + __popc(bf & below)] = (unsigned short)i;
                else if (in) order[tail + __popc(bo & below)] = (unsigned short)i;
                head += __popc(bf); tail += __popc(bo);
            }
        }
        __syncwarp();
        // ---- per-body list extents (static for the island): degree, exclusive scan over the group
        int cnt = 0;
        for (int i = 0; i < m_max; i++) {
            if (i < m_own) {
                const unsigned mt = rmeta[i];
                cnt += ((mt & 31u) == (unsigned)gl || ((mt >> 5) & 63u) == (unsigned)gl) ? 1 : 0;
            }
        }
        int lofs = cnt;
#pragma unroll
        for (int d = 1; d < G; d <<= 1) { const int v = __shfl_up_sync(ODEB_FULLMASK, lofs, d, G); if (gl >= d) lofs += v; }
        lofs -= cnt;
        unsigned *mylist = list + lofs;

        Real f0 = 0, f1 = 0, f2 = 0, f3 = 0, f4 = 0, f5 = 0, fneg = 0, fpos = 0;
        Real exit_delta = P.premature_delta;
        unsigned iteration = 0, extra = 0;
        const char *rec_base = (const char *)(rows + (size_t)rstart * 8);
        const char *rec_lohi = rec_base + HALF_BYTES + sizeof(Real4);        // {.., .., lo, hi} of the body-2 half
        bool done = m_own == 0;
        bool rebuild = !done;
        int pp = 0, stg = 0, stg_prev = (RING - 1) * STAGE_BYTES;   // ring stages as byte offsets
        for (;;) {   // sweeps
            if (__any_sync(ODEB_FULLMASK, rebuild)) {
                // per-body lists in the current solve order
                if (rebuild) cp_async_wait<0>();
                int c = 0;
                for (int i = 0; i < m_max; i++) {
                    if (rebuild && i < m_own) {
                        const unsigned row = order[i];
                        const unsigned mt = rmeta[row];
                        const unsigned b1 = mt & 31u, b2 = (mt >> 5) & 63u, fi16 = (mt >> 11) << 16;
                        if (b1 == (unsigned)gl) mylist[c++] = (row << 1) | ((b2 == 32u ? b1 : b2) << 11) | fi16;
                        else if (b2 == (unsigned)gl) mylist[c++] = (row << 1) | 1u | (b1 << 11) | fi16;
                    }
                }
                // prime the ring along the (cyclic) list
                if (rebuild) {
                    pp = 0; stg = 0; stg_prev = (RING - 1) * STAGE_BYTES;
#pragma unroll
                    for (int k = 0; k < RING - 1; k++) {
                        if (cnt > 0) {
                            const unsigned e = mylist[pp];
                            const char *own = rec_base + (size_t)(e & 0x7ffu) * HALF_BYTES;
                            const char *lohi = rec_lohi + (size_t)(e & 0x7feu) * HALF_BYTES;
                            const unsigned dst = ring_addr + (unsigned)(k * CHS * 32 * 16);
#pragma unroll
                            for (int cc = 0; cc < CH; cc++) cp_async16(dst + cc * 32 * 16, own + cc * 16);
#pragma unroll
                            for (int cc = 0; cc < CHQ; cc++) cp_async16(dst + (CH + cc) * 32 * 16, lohi + cc * 16);
                            pp = (pp + 1 == cnt) ? 0 : pp + 1;
                        }
                        cp_async_commit();
                    }
                }
                rebuild = false;
                __syncwarp();
            }
            // ---- one sweep: dataflow iterations until every lane has walked its list
            const unsigned *lp = mylist, *lend = mylist + (done ? 0 : cnt);
            unsigned e = lp < lend ? lp[0] : ODEB_BL_NONE;
            unsigned en = lp + 1 < lend ? lp[1] : ODEB_BL_NONE;
            while (__any_sync(ODEB_FULLMASK, (int)e >= 0)) {
                const unsigned row = (e >> 1) & 0x3ffu, fi = (e >> 16) & 0x3ffu;
                const int src_lane = gbase + (int)((e >> 11) & 31u);
                const bool side = e & 1u;
                const unsigned pe = __shfl_sync(ODEB_FULLMASK, e, src_lane);
                cp_async_wait<RING - 2>();
                const uint4 *slot = (const uint4 *)((const char *)ring + stg);
                Real4 c0, c1, c2, c3, c4;
#if defined(ODEB_DOUBLE)
                { uint4 t[10];
#pragma unroll
                  for (int cc = 0; cc < 10; cc++) t[cc] = slot[cc * 32];
                  const Real4 *r4 = (const Real4 *)t; c0 = r4[0]; c1 = r4[1]; c2 = r4[2]; c3 = r4[3]; c4 = r4[4]; }
#else
                { const Real4 *s4 = (const Real4 *)slot; c0 = s4[0]; c1 = s4[32]; c2 = s4[64]; c3 = s4[96]; c4 = s4[128]; }
#endif
                const Real old_lambda = lam[row];
                const Real lam_fi = lam[fi];
                const Real s = f0 * c0.x + f1 * c0.y + f2 * c0.z + f3 * c0.w + f4 * c1.x + f5 * c1.y;
                const Real ta = (c1.z - old_lambda * c1.w) - s;            // meaningful on the body-1 side
                const Real mine = side ? s : ta;
                Real other = __shfl_sync(ODEB_FULLMASK, mine, src_lane);
                const bool ex = (((pe ^ e) & 0x7feu) == 0u) && ((int)(pe | e) >= 0);
                if (ex) {
                    if (src_lane == lane) other = 0;                     // one-body row: delta = (rhs - lambda*cfm) - s1
                    Real delta = side ? (other - mine) : (mine - other);
                    const bool hasfi = fi != row;
                    const Real hi_f = RFABS(c4.w * lam_fi);
                    const Real hi_act = hasfi ? hi_f : c4.w;
                    const Real lo_act = hasfi ? -hi_f : c4.z;
                    Real new_lambda = old_lambda + delta;
                    const bool c_lo = new_lambda < lo_act;
                    const bool c_hi = !c_lo && (new_lambda > hi_act);
                    const Real lim = c_lo ? lo_act : hi_act;
                    if (c_lo || c_hi) { delta = lim - old_lambda; new_lambda = lim; }
                    if (!side) lam[row] = new_lambda;
                    f0 += delta * c2.x; f1 += delta * c2.y; f2 += delta * c2.z; f3 += delta * c2.w;
                    f4 += delta * c3.x; f5 += delta * c3.y;
                    {
                        const Real t1 = delta * c3.z;
                        if (delta > 0) fpos += t1; else fneg += t1;
                    }
                    // prefetch along the list into the stage consumed by the previous execution
                    {
                        const unsigned pe2 = mylist[pp];
                        const char *own = rec_base + (size_t)(pe2 & 0x7ffu) * HALF_BYTES;
                        const char *lohi = rec_lohi + (size_t)(pe2 & 0x7feu) * HALF_BYTES;
                        const unsigned dst = ring_addr + (unsigned)stg_prev;
#pragma unroll
                        for (int cc = 0; cc < CH; cc++) cp_async16(dst + cc * 32 * 16, own + cc * 16);
#pragma unroll
                        for (int cc = 0; cc < CHQ; cc++) cp_async16(dst + (CH + cc) * 32 * 16, lohi + cc * 16);
                        cp_async_commit();
                        pp = (pp + 1 == cnt) ? 0 : pp + 1;
                    }
                    stg_prev = stg;
                    stg = (stg + STAGE_BYTES == RING * STAGE_BYTES) ? 0 : stg + STAGE_BYTES;
                    lp++;
                    e = en;
                    en = (lp + 1 < lend) ? lp[1] : ODEB_BL_NONE;
                }
                __syncwarp();
            }
            // ---- iteration control (quickstep.cpp:1832-1855, :3253-3285), evaluated redundantly by every lane of the world
            bool need_shuffle = false;
            if (!done) {
                ++iteration;
                if (gl == 0) { ++sweeps; rowsweeps += m_own; }
                if (iteration - extra == P.num_iter) {
                    if (extra != 0 || P.max_extra == 0) { if (extra != 0 && gl == 0) st3++; done = true; }
                    else { extra = P.max_extra; exit_delta = P.extra_delta; }
                }
            }
            // the dynamic test needs the limit chosen above, so the vote is taken after it
            {
                const bool live = !done && P.dyn_enabled;
                const bool over = live && (gl < nb) && (!(fpos < exit_delta) || !(-fneg < exit_delta));
                const bool any_over = ((__ballot_sync(ODEB_FULLMASK, over) >> gbase) & GBITS) != 0u;
                if (live) {
                    const bool hit = (exit_delta == 0) || any_over;
                    fneg = 0; fpos = 0;
                    if (!hit) {
                        if (gl == 0) { if (iteration < P.num_iter) st1++; else if (iteration > P.num_iter) st2++; }
                        done = true;
                    }
                }
                need_shuffle = !done && iteration >= 8 && (iteration & 7) == 0;
            }
            if (__all_sync(ODEB_FULLMASK, done)) break;
            if (__any_sync(ODEB_FULLMASK, need_shuffle)) {
                // ConstraintsShuffling quickstep.cpp:2578-2611 with dRandInt misc.cpp:78-139, on the world's order array
                if (need_shuffle && gl == 0) {
                    for (int idx = 1; idx < m_own; idx++) {
                        const int sw = odeb_rand_int(&seed, idx + 1);
                        const unsigned short a = order[idx], b = order[sw];
                        order[idx] = b; order[sw] = a;
                    }
                }
                __syncwarp();
                seed = __shfl_sync(ODEB_FULLMASK, seed, gbase);
                rebuild = need_shuffle;
            }
        }
        cp_async_wait<0>();
        if (m_own > 0 && gl < nb) {
            const Real4 a = { f0, f1, f2, f3 }, b = { f4, f5, fneg, fpos };
            cf_out[2 * (bstart + gl)] = a; cf_out[2 * (bstart + gl) + 1] = b;
        }
        if (D.jcopy && m_own > 0) for (int i = gl; i < m_own; i += G) D.lambda[(size_t)w * P.MR + rstart + i] = lam[i];   // joint feedback
        __syncwarp();
    }
    if (gl == 0 && valid) {
        D.seed[w] = seed;
        unsigned *st = D.stats + 4 * (size_t)w;
        st[0] += st0; st[1] += st1; st[2] += st2; st[3] += st3;
        D.sweeps[2 * (size_t)w] = sweeps; D.sweeps[2 * (size_t)w + 1] = rowsweeps;
    }
}
#endif
