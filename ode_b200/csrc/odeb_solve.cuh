// odeb_solve.cuh -- k_solve: the SOR-LCP sweeps of dxQuickStepIsland (quickstep.cpp:1823-1856 loop,
// :2329-2355 ReorderPrep, :2578-2611 random reorder via dRandInt, :2917-3033 IterationStep,
// :3253-3285 dynamic iteration control), one thread per world, islands in the reference's order.
//
// The row update is a strict dependency chain (each row reads the constraint-force accumulators the
// previous row wrote), so the kernel is latency-bound per world and throughput comes from worlds in
// flight.  Data placement on B200:
//   * lambda, the solve order and the per-body accumulators (cforce + max-adjustment pair) live in shared
//     memory, interleaved [element][lane] so that every lane always hits its own bank;
//   * the 32-real row records (J row + iMJ row, 128 B single / 256 B double) stream from L2/HBM through a
//     per-lane ring of cp.async (LDGSTS) stages, issued RING-1 rows ahead in solve order, bypassing L1;
//   * `solver_lanes` (template LS) lanes of each warp carry a world each: fewer lanes = more warps for the
//     four schedulers, more lanes = fewer shared-memory instructions per world (measured optimum: 16).
// Islands too large for the shared-memory budget take the global-memory path (same arithmetic).
#ifndef ODEB_SOLVE_CUH
#define ODEB_SOLVE_CUH

#define ODEB_RING 8
#define ODEB_REC_CHUNKS ((int)(sizeof(Real) * 32 / 16))   // 16-byte chunks per row record

struct RowRegs { Real4 j0, j1, j2, j3, m0, m1, m2, m3; };

__device__ __forceinline__ void cp_async16(unsigned dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// one SOR row update on registers + accumulators reached through LD/ST functors (shared or global)
template <class CF, class LAM>
__device__ __forceinline__ void row_update(const RowRegs &cur, int index, int rbase, CF &cf, LAM &lam)
{
    Real old_lambda = lam.get(index);
    int b1raw = *(const int *)&cur.m3.z, b2 = *(const int *)&cur.m3.w;
    int b1 = b1raw & ~FINDEX_FLAG;
    Real delta = cur.j1.z - old_lambda * cur.j1.w;
    Real4 f1a = cf.get(2 * b1), f1b = cf.get(2 * b1 + 1);
    delta -= f1a.x * cur.j0.x + f1a.y * cur.j0.y + f1a.z * cur.j0.z + f1a.w * cur.j0.w + f1b.x * cur.j1.x + f1b.y * cur.j1.y;
    Real4 f2a = cf.get(2 * b2), f2b = cf.get(2 * b2 + 1);      // one-body rows: b2 = dummy slot NB, J2 = iMJ2 = 0
    delta -= f2a.x * cur.j2.x + f2a.y * cur.j2.y + f2a.z * cur.j2.z + f2a.w * cur.j2.w + f2b.x * cur.j3.x + f2b.y * cur.j3.y;
    Real hi_act, lo_act;
    if (b1raw & FINDEX_FLAG) { int fi = *(const int *)&cur.j3.z; hi_act = RFABS(cur.j3.w * lam.get(fi - rbase)); lo_act = -hi_act; }
    else { hi_act = cur.j3.w; lo_act = cur.j3.z; }
    Real new_lambda = old_lambda + delta;
    if (new_lambda < lo_act) { delta = lo_act - old_lambda; lam.set(index, lo_act); }
    else if (new_lambda > hi_act) { delta = hi_act - old_lambda; lam.set(index, hi_act); }
    else lam.set(index, new_lambda);
    if (delta != 0) {
        f1a.x += delta * cur.m0.x; f1a.y += delta * cur.m0.y; f1a.z += delta * cur.m0.z; f1a.w += delta * cur.m0.w;
        f1b.x += delta * cur.m1.x; f1b.y += delta * cur.m1.y;
        if (delta > 0) f1b.w += delta * cur.m1.z; else f1b.z += delta * cur.m1.z;
        cf.set(2 * b1, f1a); cf.set(2 * b1 + 1, f1b);
        if (delta > 0) f2b.w += delta * cur.m3.y; else f2b.z += delta * cur.m3.y;
        f2a.x += delta * cur.m1.w; f2a.y += delta * cur.m2.x; f2a.z += delta * cur.m2.y; f2a.w += delta * cur.m2.z;
        f2b.x += delta * cur.m2.w; f2b.y += delta * cur.m3.x;
        cf.set(2 * b2, f2a); cf.set(2 * b2 + 1, f2b);
    }
}

struct CfGlobal { Real4 *p; __device__ Real4 get(int i) const { return p[i]; } __device__ void set(int i, const Real4 &v) { p[i] = v; } };
struct LamGlobal { Real *p; __device__ Real get(int i) const { return p[i]; } __device__ void set(int i, Real v) { p[i] = v; } };
struct CfShared { Real4 *p; int ls; __device__ Real4 get(int i) const { return p[i * ls]; } __device__ void set(int i, const Real4 &v) { p[i * ls] = v; } };
struct LamShared { Real *p; int ls; __device__ Real get(int i) const { return p[i * ls]; } __device__ void set(int i, Real v) { p[i * ls] = v; } };

// iteration control shared by both paths; returns true when the island is finished
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

// Shared-memory island solve, software-pipelined for a single in-order warp:
//   per row i:  (1) loads that depend on row i-1's stores (lambda, the two bodies' accumulators),
//               (2) bookkeeping for later rows that fills their latency (cp.async issue for row i+RING-1,
//                   wait + register load of row i+1's record and solve-order entry),
//               (3) the branch-free arithmetic chain, (4) stores.
// One-body rows address a dummy accumulator slot (index NB) with zero J2/iMJ2 instead of branching.
template <int LS>
__device__ __forceinline__ void solve_island_shared(const DevParams &P, unsigned char *smem, int lane,
                                                    const Real4 *rows, const int *findex, Real4 *cf_out,
                                                    int bstart, int nb, int rstart, int m, unsigned &seed,
                                                    unsigned &st1, unsigned &st2, unsigned &st3,
                                                    unsigned long long &sweeps, unsigned long long &rowsweeps)
{
    constexpr int CH = ODEB_REC_CHUNKS;
    uint4 *ring = (uint4 *)smem + lane;                                            // chunk (slot,c) at ring[(slot*CH+c)*LS]
    Real4 *cf = (Real4 *)((uint4 *)smem + ODEB_RING * CH * LS) + lane;             // cf[k*LS], k < 2*(NB+1)
    Real *lam = (Real *)((Real4 *)((uint4 *)smem + ODEB_RING * CH * LS) + 2 * (P.NB + 1) * LS) + lane;   // lam[i*LS]
    unsigned short *ord = (unsigned short *)((Real *)((Real4 *)((uint4 *)smem + ODEB_RING * CH * LS) + 2 * (P.NB + 1) * LS) + (P.SR + 1) * LS) + lane;
    const unsigned ring_addr = (unsigned)__cvta_generic_to_shared(ring);
    const Real4 z4 = { 0, 0, 0, 0 };
    for (int k = 0; k < 2 * nb; k++) cf[(2 * bstart + k) * LS] = z4;
    cf[(2 * P.NB) * LS] = z4; cf[(2 * P.NB + 1) * LS] = z4;
    int nvalid = 0;
    for (int i = 0; i < m; i++) { lam[i * LS] = 0; if (findex[rstart + i] != -1) nvalid++; }
    {   // ReorderPrep quickstep.cpp:2329-2355
        int head = 0, tail = m - nvalid;
        for (int i = 0; i < m; i++) { if (findex[rstart + i] == -1) ord[(head++) * LS] = (unsigned short)i; else ord[(tail++) * LS] = (unsigned short)i; }
    }
    ord[m * LS] = 0;
    const char *rec_base = (const char *)(rows + (size_t)rstart * 8);
    Real exit_delta = P.premature_delta;
    CfShared cfs = { cf, LS };
    for (unsigned iteration = 0, extra = 0;;) {
        if (iteration >= 8 && (iteration & 7) == 0) {
            // ConstraintsShuffling quickstep.cpp:2578-2611 with dRandInt misc.cpp:78-139
            for (int idx = 1; idx < m; idx++) {
                int sw = odeb_rand_int(&seed, idx + 1);
                unsigned short a = ord[idx * LS], b = ord[sw * LS];
                ord[idx * LS] = b; ord[sw * LS] = a;
            }
        }
#pragma unroll
        for (int k = 0; k < ODEB_RING - 1; k++) {
            if (k < m) {
                const char *src = rec_base + (size_t)ord[k * LS] * (sizeof(Real) * 32);
                unsigned dst = ring_addr + (unsigned)(k * CH * LS * 16);
#pragma unroll
                for (int c = 0; c < CH; c++) cp_async16(dst + c * LS * 16, src + c * 16);
            }
            cp_async_commit();
        }
        cp_async_wait<ODEB_RING - 2>();
        RowRegs cur;
        {
            const uint4 *slot = ring;
#if defined(ODEB_DOUBLE)
            uint4 t[16];
#pragma unroll
            for (int c = 0; c < 16; c++) t[c] = slot[c * LS];
            const Real4 *r4 = (const Real4 *)t;
            cur.j0 = r4[0]; cur.j1 = r4[1]; cur.j2 = r4[2]; cur.j3 = r4[3]; cur.m0 = r4[4]; cur.m1 = r4[5]; cur.m2 = r4[6]; cur.m3 = r4[7];
#else
            const Real4 *s4 = (const Real4 *)slot;
            cur.j0 = s4[0]; cur.j1 = s4[LS]; cur.j2 = s4[2 * LS]; cur.j3 = s4[3 * LS];
            cur.m0 = s4[4 * LS]; cur.m1 = s4[5 * LS]; cur.m2 = s4[6 * LS]; cur.m3 = s4[7 * LS];
#endif
        }
        int index = ord[0];
        for (int i = 0; i < m; i++) {
            // (1) loads depending on the previous row's stores
            const int b1raw = *(const int *)&cur.m3.z, b2 = *(const int *)&cur.m3.w;
            const int b1 = b1raw & ~FINDEX_FLAG;
            const bool hasfi = (b1raw & FINDEX_FLAG) != 0;
            const int fi = hasfi ? (*(const int *)&cur.j3.z - rstart) : index;
            const Real old_lambda = lam[index * LS];
            const Real lam_fi = lam[fi * LS];
            Real4 f1a = cf[(2 * b1) * LS], f1b = cf[(2 * b1 + 1) * LS];
            Real4 f2a = cf[(2 * b2) * LS], f2b = cf[(2 * b2 + 1) * LS];
            // (2) bookkeeping for later rows
            {
                const int ahead = i + ODEB_RING - 1;
                if (ahead < m) {
                    const char *src = rec_base + (size_t)ord[ahead * LS] * (sizeof(Real) * 32);
                    unsigned dst = ring_addr + (unsigned)((ahead & (ODEB_RING - 1)) * CH * LS * 16);
#pragma unroll
                    for (int c = 0; c < CH; c++) cp_async16(dst + c * LS * 16, src + c * 16);
                }
                cp_async_commit();
                cp_async_wait<ODEB_RING - 2>();
            }
            RowRegs nxt;
            {
                const uint4 *slot = ring + (size_t)((i + 1) & (ODEB_RING - 1)) * CH * LS;
#if defined(ODEB_DOUBLE)
                uint4 t[16];
#pragma unroll
                for (int c = 0; c < 16; c++) t[c] = slot[c * LS];
                const Real4 *r4 = (const Real4 *)t;
                nxt.j0 = r4[0]; nxt.j1 = r4[1]; nxt.j2 = r4[2]; nxt.j3 = r4[3]; nxt.m0 = r4[4]; nxt.m1 = r4[5]; nxt.m2 = r4[6]; nxt.m3 = r4[7];
#else
                const Real4 *s4 = (const Real4 *)slot;
                nxt.j0 = s4[0]; nxt.j1 = s4[LS]; nxt.j2 = s4[2 * LS]; nxt.j3 = s4[3 * LS];
                nxt.m0 = s4[4 * LS]; nxt.m1 = s4[5 * LS]; nxt.m2 = s4[6 * LS]; nxt.m3 = s4[7 * LS];
#endif
            }
            const int nindex = ord[(i + 1) * LS];
            // (3) IterationStep quickstep.cpp:2917-3033, branch-free
            Real delta = cur.j1.z - old_lambda * cur.j1.w;
            delta -= f1a.x * cur.j0.x + f1a.y * cur.j0.y + f1a.z * cur.j0.z + f1a.w * cur.j0.w + f1b.x * cur.j1.x + f1b.y * cur.j1.y;
            delta -= f2a.x * cur.j2.x + f2a.y * cur.j2.y + f2a.z * cur.j2.z + f2a.w * cur.j2.w + f2b.x * cur.j3.x + f2b.y * cur.j3.y;
            const Real hi_f = RFABS(cur.j3.w * lam_fi);
            const Real hi_act = hasfi ? hi_f : cur.j3.w;
            const Real lo_act = hasfi ? -hi_f : cur.j3.z;
            Real new_lambda = old_lambda + delta;
            const bool c_lo = new_lambda < lo_act;
            const bool c_hi = !c_lo && (new_lambda > hi_act);
            const Real lim = c_lo ? lo_act : hi_act;
            if (c_lo || c_hi) { delta = lim - old_lambda; new_lambda = lim; }
            const bool pos = delta > 0;
            f1a.x += delta * cur.m0.x; f1a.y += delta * cur.m0.y; f1a.z += delta * cur.m0.z; f1a.w += delta * cur.m0.w;
            f1b.x += delta * cur.m1.x; f1b.y += delta * cur.m1.y;
            {
                const Real t1 = delta * cur.m1.z, t2 = delta * cur.m3.y;
                const Real p1v = f1b.w + t1, n1v = f1b.z + t1, p2v = f2b.w + t2, n2v = f2b.z + t2;
                f1b.w = pos ? p1v : f1b.w; f1b.z = pos ? f1b.z : n1v;
                f2b.w = pos ? p2v : f2b.w; f2b.z = pos ? f2b.z : n2v;
            }
            f2a.x += delta * cur.m1.w; f2a.y += delta * cur.m2.x; f2a.z += delta * cur.m2.y; f2a.w += delta * cur.m2.z;
            f2b.x += delta * cur.m2.w; f2b.y += delta * cur.m3.x;
            // (4) stores
            lam[index * LS] = new_lambda;
            cf[(2 * b1) * LS] = f1a; cf[(2 * b1 + 1) * LS] = f1b;
            cf[(2 * b2) * LS] = f2a; cf[(2 * b2 + 1) * LS] = f2b;
            cur = nxt; index = nindex;
        }
        cp_async_wait<0>();
        ++iteration; ++sweeps; rowsweeps += m;
        if (sweep_control(P, cfs, bstart, nb, iteration, extra, exit_delta, st1, st2, st3)) break;
    }
    for (int k = 0; k < 2 * nb; k++) cf_out[2 * bstart + k] = cf[(2 * bstart + k) * LS];
}

template <int LS>
__global__ void __launch_bounds__(32) k_solve(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D)
{
    extern __shared__ __align__(32) unsigned char smem[];
    const int lane = threadIdx.x;
    const int w = blockIdx.x * LS + lane;
    if (lane >= LS || w >= P.W) return;
    unsigned seed = D.seed[w];
    unsigned st0 = 0, st1 = 0, st2 = 0, st3 = 0;
    unsigned long long sweeps = 0, rowsweeps = 0;
    const Real4 *rows = D.rows + (size_t)w * P.MR * 8;
    const int *findex = D.findex + (size_t)w * P.MR;
    Real4 *cf_out = D.cforce + (size_t)w * (P.NB + 1) * 2;
    const int4 *iinfo = D.island_info + (size_t)w * P.NB;
    const int nis = D.nislands[w];
    for (int is = 0; is < nis; is++) {
        const int4 info = iinfo[is];
        const int bstart = info.x, nb = info.y, rstart = info.z, m = info.w;
        if (m > 0 && m <= P.SR) {
            solve_island_shared<LS>(P, smem, lane, rows, findex, cf_out, bstart, nb, rstart, m, seed, st1, st2, st3, sweeps, rowsweeps);
        } else if (m > 0) {
            // ---------------- global-memory path (islands beyond the shared-memory budget)
            CfGlobal cfg = { cf_out };
            LamGlobal lamg = { D.lambda + (size_t)w * P.MR + rstart };
            int *order = D.order + (size_t)w * P.MR + rstart;
            const Real4 z4 = { 0, 0, 0, 0 };
            for (int k = 0; k < 2 * nb; k++) cfg.set(2 * bstart + k, z4);
            cfg.set(2 * P.NB, z4); cfg.set(2 * P.NB + 1, z4);
            int nvalid = 0;
            for (int i = 0; i < m; i++) { lamg.set(i, 0); if (findex[rstart + i] != -1) nvalid++; }
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
                    const Real4 *r4 = rec + (size_t)index * 8;
                    RowRegs cur;
                    cur.j0 = r4[0]; cur.j1 = r4[1]; cur.j2 = r4[2]; cur.j3 = r4[3]; cur.m0 = r4[4]; cur.m1 = r4[5]; cur.m2 = r4[6]; cur.m3 = r4[7];
                    row_update(cur, index, rstart, cfg, lamg);
                }
                ++iteration; ++sweeps; rowsweeps += m;
                if (sweep_control(P, cfg, bstart, nb, iteration, extra, exit_delta, st1, st2, st3)) break;
            }
        }
        st0++;
    }
    D.seed[w] = seed;
    unsigned *st = D.stats + 4 * (size_t)w;
    st[0] += st0; st[1] += st1; st[2] += st2; st[3] += st3;
    D.sweeps[2 * (size_t)w] = sweeps; D.sweeps[2 * (size_t)w + 1] = rowsweeps;
}
#endif
