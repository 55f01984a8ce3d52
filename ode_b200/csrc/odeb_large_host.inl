// odeb_large_host.inl -- host side of the large-world path (kernels in odeb_large.cuh): buffer set-up and the launch
// sequence of one step.  The sequence has a few small device->host reads (pair / contact / row counts, number of
// unfinished islands at each phase boundary) because the radix sorts need their sizes; a step of a 10^5-body world
// is milliseconds of GPU work, so these cost nothing measurable.  No CPU fallback, no host-side physics.

#define LCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { set_err("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); return 0; } } while (0)

static size_t large_cub_bytes(const DevParams &P)
{
    size_t best = 0, b = 0;
    cub::DeviceRadixSort::SortPairs(0, b, (unsigned *)0, (unsigned *)0, (int *)0, (int *)0, P.NG); best = b > best ? b : best;
    cub::DeviceRadixSort::SortKeys(0, b, (u64 *)0, (u64 *)0, P.MP); best = b > best ? b : best;
    cub::DeviceRadixSort::SortPairs(0, b, (u64 *)0, (u64 *)0, (int *)0, (int *)0, P.NB); best = b > best ? b : best;
    cub::DeviceRadixSort::SortPairs(0, b, (u64 *)0, (u64 *)0, (int *)0, (int *)0, P.NJT); best = b > best ? b : best;
    cub::DeviceRadixSort::SortPairs(0, b, (unsigned *)0, (unsigned *)0, (int *)0, (int *)0, P.MR); best = b > best ? b : best;
    cub::DeviceScan::ExclusiveSum(0, b, (int *)0, (int *)0, P.MP); best = b > best ? b : best;
    cub::DeviceScan::ExclusiveSum(0, b, (int *)0, (int *)0, P.NJT); best = b > best ? b : best;
    cub::DeviceScan::InclusiveSum(0, b, (int *)0, (int *)0, P.NB + 2); best = b > best ? b : best;
    return best + 256;
}

static int large_prepare(OdebBatch *B)
{
    if (B->large_ready) return 1;
    const DevParams &P = B->P;
    LargePtrs &L = B->L;
    const size_t NB = P.NB, NG = P.NG > 0 ? P.NG : 1, MP = P.MP, MR = P.MR, NJT = P.NJT;
    bool ok = dev_alloc(B, &L.counters, LWC_COUNT) && dev_alloc(B, &L.draws, 4)
           && dev_alloc(B, &L.bp_key, NG) && dev_alloc(B, &L.bp_key_s, NG) && dev_alloc(B, &L.bp_idx, NG) && dev_alloc(B, &L.bp_idx_s, NG) && dev_alloc(B, &L.bp_big, NG) && dev_alloc(B, &L.bp_yz, NG)
           && dev_alloc(B, &L.pair_key, MP) && dev_alloc(B, &L.pair_key_s, MP) && dev_alloc(B, &L.pc_base, MP)
           && dev_alloc(B, &L.parent, NB) && dev_alloc(B, &L.maxen, NB) && dev_alloc(B, &L.head_scan, NB) && dev_alloc(B, &L.deg, NB)
           && dev_alloc(B, &L.bkey, NB) && dev_alloc(B, &L.bkey_s, NB) && dev_alloc(B, &L.bval, NB)
           && dev_alloc(B, &L.isl_nb, NB + 2) && dev_alloc(B, &L.isl_m, NB + 2) && dev_alloc(B, &L.isl_bstart, NB + 2) && dev_alloc(B, &L.isl_rstart, NB + 2)
           && dev_alloc(B, &L.isl_done, NB + 2) && dev_alloc(B, &L.isl_viol, NB + 2)
           && dev_alloc(B, &L.jkey, NJT) && dev_alloc(B, &L.jkey_s, NJT) && dev_alloc(B, &L.jmv, NJT) && dev_alloc(B, &L.jmv_s, NJT) && dev_alloc(B, &L.jrow, NJT)
           && dev_alloc(B, &L.row_island, MR) && dev_alloc(B, &L.row_group, MR) && dev_alloc(B, &L.gsize, MR) && dev_alloc(B, &L.heads, MR)
           && dev_alloc(B, &L.ginc_ofs, NB + 2) && dev_alloc(B, &L.ginc_cur, NB + 2) && dev_alloc(B, &L.ginc, 2 * MR)
           && dev_alloc(B, &L.gkey, MR) && dev_alloc(B, &L.gcolor, MR) && dev_alloc(B, &L.gwin, MR)
           && dev_alloc(B, &L.clist, MR) && dev_alloc(B, &L.ccount, ODEB_CANON_COLOURS) && dev_alloc(B, &L.cofs, ODEB_CANON_COLOURS + 1) && dev_alloc(B, &L.tstart, ODEB_CANON_COLOURS + 1)
           && dev_alloc(B, &L.skey, MR) && dev_alloc(B, &L.skey_s, MR)
           && dev_alloc(B, &L.theight, MR / 32 + ODEB_CANON_COLOURS + 4) && dev_alloc(B, &L.tbase, MR / 32 + ODEB_CANON_COLOURS + 4)
           && dev_alloc(B, &L.tgroup, MR + 32 * ODEB_CANON_COLOURS + 64) && dev_alloc(B, &L.tginfo, MR + 32 * ODEB_CANON_COLOURS + 64);   // every colour ends with a partial tile
    // tile rows: every row once, plus the padding of tiles whose groups differ in size (sorted by size: a few tile heights per colour)
    L.trcap = (int)((MR + MR / 4 + 32 * ODEB_CANON_COLOURS * 16 + 31) / 32);
    ok = ok && dev_alloc(B, &L.trec, (size_t)L.trcap * LWT_ROW_BYTES) && dev_alloc(B, &L.pinvm, NB + 1);
    if (!ok) return 0;
    L.tmp_bytes = large_cub_bytes(P);
    { unsigned char *t = 0; if (!dev_alloc(B, &t, L.tmp_bytes)) return 0; L.tmp = t; }
    B->D.row_island = L.row_island; B->D.row_group = L.row_group;
    B->large_ready = true;
    return 1;
}

extern "C" int odeb_set_solver_mode(OdebBatch *B, int mode)
{
    if (mode != ODEB_MODE_REPLAY && mode != ODEB_MODE_CANONICAL) { set_err("unknown solver mode %d", mode); return 0; }
    if (mode == ODEB_MODE_CANONICAL) {
        if (B->P.W != 1 || B->P.classic) { set_err("ODEB_MODE_CANONICAL is the single-world path: needs nworlds == 1 and the batch API"); return 0; }
        CK(cudaSetDevice(B->device));
        if (!large_prepare(B)) return 0;
    }
    B->mode = mode;
    return 1;
}

extern "C" uint32_t odeb_canon_key(uint32_t seed, uint32_t island, uint32_t phase, uint32_t row) { return odebi_canon_key(seed, island, phase, row); }

__global__ void k_test_atan2f(const float *y, const float *x, float *out, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = odeb_atan2f_fdlibm(y[i], x[i]);
}
extern "C" int odeb_test_atan2f(const float *y, const float *x, float *out, int n, int on_device)
{
    if (n <= 0) return 1;
    if (!on_device) { for (int i = 0; i < n; i++) out[i] = odeb_atan2f_fdlibm(y[i], x[i]); return 1; }
    float *d = 0;
    CK(cudaMalloc(&d, (size_t)3 * n * sizeof(float)));
    cudaError_t e = cudaMemcpy(d, y, (size_t)n * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d + n, x, (size_t)n * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) { k_test_atan2f<<<(n + 255) / 256, 256>>>(d, d + n, d + 2 * (size_t)n, n); e = cudaGetLastError(); }
    if (e == cudaSuccess) e = cudaMemcpy(out, d + 2 * (size_t)n, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) { set_err("odeb_test_atan2f: %s", cudaGetErrorString(e)); return 0; }
    return 1;
}

static int bits_for(unsigned v) { int b = 1; while ((v >> b) != 0 && b < 32) b++; return b; }

static int large_step(OdebBatch *B)
{
    const DevParams &P = B->P; const DevPtrs &D = B->D; LargePtrs &L = B->L;
    cudaStream_t s = B->stream;
    const int NB = P.NB, NG = P.NG;
    int hc[LWC_COUNT];
    LCK(cudaMemsetAsync(L.counters, 0, LWC_COUNT * sizeof(int), s));
    LCK(cudaMemsetAsync(L.draws, 0, 4 * sizeof(u64), s));
    if (D.jfb) LCK(cudaMemsetAsync(D.jfb, 0, (size_t)P.NJT * 4 * sizeof(Real4), s));      // state 0 = joint not stepped
    // ---------------- collision
    OdebRange nv_collide("dSpaceCollide (dxSAPSpace / dxHashSpace) + nearCallback (dCollide, dJointCreateContact)");
    int np = 0;
    if (NG > 0) {
        k_aabb<<<nblk(NG, 128), 128, 0, s>>>(P, D);
        k_bp_keys<<<nblk(NG, 128), 128, 0, s>>>(P, D, L);
        LCK(cub::DeviceRadixSort::SortPairs(L.tmp, L.tmp_bytes, L.bp_key, L.bp_key_s, L.bp_idx, L.bp_idx_s, NG, 0, 32, s));
        k_bp_gather<<<nblk(NG, 256), 256, 0, s>>>(P, D, L);
        k_bp_sweep<<<nblk(NG, 64), 64, 0, s>>>(P, D, L);
        k_bp_big<<<nblk(NG, 128), 128, 0, s>>>(P, D, L);
        B->launches += 6;
        LCK(cudaMemcpyAsync(hc, L.counters, sizeof(hc), cudaMemcpyDeviceToHost, s));
        LCK(cudaStreamSynchronize(s));
        np = hc[LWC_NPAIRS];
        if (np > P.MP) { set_err("capacity overflow (pairs: %d > %d): raise ODEB_MAX_PAIRS", np, P.MP); return 0; }
        if (np > 0) {
            LCK(cub::DeviceRadixSort::SortKeys(L.tmp, L.tmp_bytes, L.pair_key, L.pair_key_s, np, 0, 32 + bits_for((unsigned)NG), s));
            B->launches++;
        }
        k_bp_unpack<<<nblk(np > 0 ? np : 1, 256), 256, 0, s>>>(np, L.pair_key_s, D.pairs, D.npairs);
        B->launches++;
        if (np > 0) {
            k_narrow<<<nblk(np, 64), 64, 0, s>>>(P, D, 0);
            LCK(cub::DeviceScan::ExclusiveSum(L.tmp, L.tmp_bytes, D.pc_count, L.pc_base, np, s));
            k_lw_contact_fill<<<nblk(np, 128), 128, 0, s>>>(P, D, L, np);
            B->launches += 3;
        } else LCK(cudaMemsetAsync(D.ncontacts, 0, sizeof(int), s));
    } else {
        LCK(cudaMemsetAsync(D.npairs, 0, sizeof(int), s));
        LCK(cudaMemsetAsync(D.ncontacts, 0, sizeof(int), s));
    }
    nv_collide.end();
    // ---------------- joints, auto-disable, islands
    OdebRange nv_islands("dxProcessIslands (auto-disable, island build) + dxQuickStepIsland_Stage0_Joints + Stage1");
    if (P.NJ > 0) { k_joint_info1<<<nblk(P.NJ, 128), 128, 0, s>>>(P, D); B->launches++; }
    LCK(cudaMemcpyAsync(hc, L.counters, sizeof(hc), cudaMemcpyDeviceToHost, s));
    LCK(cudaStreamSynchronize(s));
    const int nc = hc[LWC_NCONTACTS];
    const int nj = P.NJ + nc;
    LCK(cudaMemsetAsync(L.deg, 0, NB * sizeof(int), s));
    k_lw_init_bodies<<<nblk(NB + 1, 256), 256, 0, s>>>(P, D, L);
    B->launches++;
    if (P.adis_samples > 0) {
        if (nc > 0) { k_lw_degree<<<nblk(nc, 256), 256, 0, s>>>(P, D, L); B->launches++; }
        k_lw_autodisable<<<nblk(NB, 128), 128, 0, s>>>(P, D, L);
        B->launches++;
    }
    if (nj > 0) { k_lw_union<<<nblk(nj, 256), 256, 0, s>>>(P, D, L); B->launches++; }
    k_lw_roots<<<nblk(NB, 256), 256, 0, s>>>(P, D, L);
    k_lw_heads<<<nblk(NB, 256), 256, 0, s>>>(P, L);
    LCK(cub::DeviceScan::InclusiveSum(L.tmp, L.tmp_bytes, L.deg, L.head_scan, NB, s));
    k_lw_label<<<nblk(NB, 256), 256, 0, s>>>(P, D, L);
    k_lw_iota<<<nblk(NB, 256), 256, 0, s>>>(NB, L.bval);
    LCK(cub::DeviceRadixSort::SortPairs(L.tmp, L.tmp_bytes, L.bkey, L.bkey_s, L.bval, D.body_order, NB, 0, 32 + bits_for((unsigned)NB), s));   // island numbers are below NB; unordered bodies carry all-ones keys: last either way
    k_lw_body_pos<<<nblk(NB, 256), 256, 0, s>>>(P, D, L);
    k_lw_joint_keys<<<nblk(P.NJT, 256), 256, 0, s>>>(P, D, L);
    B->launches += 8;
    LCK(cudaMemcpyAsync(hc, L.counters, sizeof(hc), cudaMemcpyDeviceToHost, s));
    LCK(cudaStreamSynchronize(s));
    const int T = hc[LWC_NISLANDS], mrows = hc[LWC_MROWS], nordered = hc[LWC_NORDERED];
    if (mrows > P.MR) { set_err("capacity overflow (rows: %d > %d): raise ODEB_MAX_CONTACTS", mrows, P.MR); return 0; }
    LCK(cub::DeviceScan::ExclusiveSum(L.tmp, L.tmp_bytes, L.isl_nb, L.isl_bstart, NB + 1, s));
    LCK(cub::DeviceScan::ExclusiveSum(L.tmp, L.tmp_bytes, L.isl_m, L.isl_rstart, NB + 1, s));
    B->launches += 2;
    if (nj > 0) {
        LCK(cub::DeviceRadixSort::SortPairs(L.tmp, L.tmp_bytes, L.jkey, L.jkey_s, L.jmv, L.jmv_s, nj, 0, 32 + bits_for((unsigned)(T > 0 ? T : 1)), s));   // unordered joints carry all-ones keys: last either way
        LCK(cub::DeviceScan::ExclusiveSum(L.tmp, L.tmp_bytes, L.jmv_s, L.jrow, nj, s));
        B->launches += 2;
    }
    k_lw_joint_final<<<nblk(nj > 0 ? nj : 1, 256), 256, 0, s>>>(P, D, L, nj);
    k_lw_island_info<<<nblk(T > 0 ? T : 1, 256), 256, 0, s>>>(P, D, L);
    B->launches += 2;
    nv_islands.end();
    // ---------------- QuickStep stages 0..3 (the batched path's kernels, W == 1)
    OdebRange nv_("dxQuickStepIsland stages 0-6 (large world: colouring, tiles, Stage4LCP_Iteration phases, dxStepBody)");
    k_body_pre<<<nblk(NB, 128), 128, 0, s>>>(P, D);
    B->launches++;
    if (hc[LWC_NJORD] > 0) {
        if (B->rows_std3) k_rows_t<false, true><<<nblk(hc[LWC_NJORD], 64), 64, 0, s>>>(P, D); else k_rows_t<false><<<nblk(hc[LWC_NJORD], 64), 64, 0, s>>>(P, D);
        B->launches++;
    }
    LCK(cudaMemsetAsync(D.cforce, 0, (size_t)(NB + 1) * 2 * sizeof(Real4), s));
    if (mrows > 0) {
        // groups (rows of one geom pair's contacts / of one joint), their compact list and the body -> groups incidence (CSR): once per step
        LCK(cudaMemsetAsync(L.gsize, 0, (size_t)mrows * sizeof(int), s));
        LCK(cudaMemsetAsync(L.counters + LWC_NGROUPS, 0, 3 * sizeof(int), s));
        k_lwc_groups<<<nblk(mrows, 256), 256, 0, s>>>(P, D, L, mrows);
        LCK(cub::DeviceScan::ExclusiveSum(L.tmp, L.tmp_bytes, L.ginc_cur, L.ginc_ofs, NB + 1, s));
        k_lw_zero_cur<<<nblk(NB + 1, 256), 256, 0, s>>>(NB + 1, L.ginc_cur);
        int ngroups = 0;
        LCK(cudaMemcpyAsync(&ngroups, L.counters + LWC_NGROUPS, sizeof(int), cudaMemcpyDeviceToHost, s));
        LCK(cudaStreamSynchronize(s));
        k_lwc_ginc_fill<<<nblk(ngroups, 256), 256, 0, s>>>(P, D, L);
        B->launches += 4;
        // ---------------- SOR sweeps
        cudaEvent_t e0 = 0, e1 = 0;
        
        // the step's colouring: rounds in batches of 8, until no group is left uncoloured
        LCK(cudaMemsetAsync(L.counters + LWC_UNCOLORED, 0, 2 * sizeof(int), s));
        k_lwc_color_init<<<nblk(ngroups > ODEB_CANON_COLOURS ? ngroups : ODEB_CANON_COLOURS, 256), 256, 0, s>>>(P, D, L);
        B->launches++;
        {
            if (B->lwc_grid == 0) {
                int occ = 0, sms = 0;
                LCK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_lwc_color_rounds, 1024, 0));
                LCK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, B->device));
                B->lwc_grid = occ * sms;
                if (B->lwc_grid <= 0) { set_err("k_lwc_color_rounds does not fit an SM"); return 0; }
            }
            int grid = nblk(ngroups, 1024);                    // few, large blocks: the grid barrier costs one atomic per block
            if (grid > B->lwc_grid) grid = B->lwc_grid;
            if (B->lw_maxgrid > 0 && grid > B->lw_maxgrid) grid = B->lw_maxgrid;
            LCK(cudaMemsetAsync(L.counters + LWC_GBAR, 0, sizeof(int), s));
            void *args[3] = { (void *)&P, (void *)&D, (void *)&L };
            LCK(cudaLaunchCooperativeKernel((const void *)k_lwc_color_rounds, dim3(grid), dim3(1024), args, 0, s));
            B->launches++;
        }
        // tiles: groups by (colour, rows descending), 32 per tile; records and lambdas into the lane-interleaved layout
        int tstart[ODEB_CANON_COLOURS + 1];
        k_lwt_sort_keys<<<nblk(ngroups, 256), 256, 0, s>>>(L);
        LCK(cub::DeviceRadixSort::SortPairs(L.tmp, L.tmp_bytes, L.skey, L.skey_s, L.heads, L.clist, ngroups, 0, 20, s));
        k_lwt_color_scan<<<1, 1, 0, s>>>(D, L);
        LCK(cudaMemcpyAsync(tstart, L.tstart, sizeof(tstart), cudaMemcpyDeviceToHost, s));
        LCK(cudaMemcpyAsync(hc, L.counters, sizeof(hc), cudaMemcpyDeviceToHost, s));
        LCK(cudaMemcpyAsync(B->h_ov, D.overflow, 4 * sizeof(int), cudaMemcpyDeviceToHost, s));
        LCK(cudaStreamSynchronize(s));
        if (B->h_ov[0] == 4) return overflow_check(B);          // a colouring that ran out of colours is not a colouring: stop before any sweep
        const int ntiles = tstart[ODEB_CANON_COLOURS], ncolors = hc[LWC_NCOLORS];
        const unsigned step_seed = (unsigned)hc[LWC_SEED];
        k_lwt_tiles<<<nblk(ntiles * 32, 128), 128, 0, s>>>(P, D, L);
        LCK(cub::DeviceScan::ExclusiveSum(L.tmp, L.tmp_bytes, L.theight, L.tbase, ntiles + 1, s));
        k_lwt_finish<<<ntiles, 128, 0, s>>>(P, D, L);
        B->launches += 6;
        const size_t lw_tma_smem = (size_t)LWT_WARPS * LWT_STAGES * LWT_STAGE_BYTES + (size_t)LWT_WARPS * LWT_STAGES * sizeof(unsigned long long);
        const size_t lw_small_smem = (size_t)LWT_SMALL_WARPS * LWT_SMALL_STAGES * LWT_STAGE_BYTES + (size_t)LWT_SMALL_WARPS * LWT_SMALL_STAGES * sizeof(unsigned long long);
        int corder[ODEB_CANON_COLOURS];
        if (B->timing) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, s); }   // the sweeps proper (what odeb_solver_ms reports)
        {
            // persistent phases: one cooperative launch per 8 sweeps
            if (B->lw_grid == 0) {
                cudaFuncSetAttribute(k_lwt_phase_t<LWT_WARPS, LWT_STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lw_tma_smem);
                cudaFuncSetAttribute(k_lwt_phase_t<LWT_SMALL_WARPS, LWT_SMALL_STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lw_small_smem);
                int occ = 0, sms = 0;
                LCK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_lwt_phase_t<LWT_WARPS, LWT_STAGES>, 32 * LWT_WARPS, lw_tma_smem));
                LCK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, B->device));
                B->lw_grid = occ * sms;
                if (B->lw_grid <= 0) { set_err("k_lwt_phase does not fit an SM"); return 0; }
            }
            int maxnt = 1;
            for (int c = 0; c < ncolors && c < ODEB_CANON_COLOURS; c++) if (tstart[c + 1] - tstart[c] > maxnt) maxnt = tstart[c + 1] - tstart[c];
            int grid = (maxnt + LWT_WARPS - 1) / LWT_WARPS;
            { const int by_bodies = (nordered + 32 * LWT_WARPS - 1) / (32 * LWT_WARPS); if (by_bodies > grid) grid = by_bodies; }
            if (grid > B->lw_grid) grid = B->lw_grid;
            if (B->lw_maxgrid > 0 && grid > B->lw_maxgrid) grid = B->lw_maxgrid;
            // small worlds: one block whose warps hold every tile of the largest colour (block barrier instead of the grid barrier)
            const bool small = maxnt <= LWT_SMALL_WARPS && nordered <= 64 * 32 * LWT_SMALL_WARPS && B->lw_maxgrid == 0;
            LwPhase ph;
            for (int c = 0; c <= ODEB_CANON_COLOURS; c++) ph.tstart[c] = tstart[c];
            ph.nordered = nordered; ph.nislands = T;
            // every phase the solve can need, queued back to back; a launch whose predecessors finished the solve returns at once
            const int nphases = (int)((P.num_iter + P.max_extra + 7) / 8);
            if (nphases > 12) { set_err("ODEB_MODE_CANONICAL: more than 96 sweeps per step are not supported"); return 0; }
            for (int phase = 0; phase < nphases; phase++) {
                // the phase's visiting order (odebi_canon_colour_ranks, restricted to the colours in use: the same relative order)
                std::pair<uint32_t, int> ck[ODEB_CANON_COLOURS];
                const int nc = ncolors < ODEB_CANON_COLOURS ? ncolors : ODEB_CANON_COLOURS;
                for (int c = 0; c < nc; c++) ck[c] = std::make_pair(odebi_canon_key(step_seed, 0xffffffffu, (uint32_t)phase, (uint32_t)c), c);
                std::sort(ck, ck + nc);
                for (int k = 0; k < nc; k++) corder[k] = ck[k].second;
                ph.norder = 0;
                for (int k = 0; k < nc; k++) { const int c = corder[k]; if (tstart[c + 1] > tstart[c]) ph.corder[ph.norder++] = c; }
                ph.phase = phase;
                void *args[4] = { (void *)&P, (void *)&D, (void *)&L, (void *)&ph };
                if (small) LCK(cudaLaunchCooperativeKernel((const void *)k_lwt_phase_t<LWT_SMALL_WARPS, LWT_SMALL_STAGES>, dim3(1), dim3(32 * LWT_SMALL_WARPS), args, lw_small_smem, s));
                else LCK(cudaLaunchCooperativeKernel((const void *)k_lwt_phase_t<LWT_WARPS, LWT_STAGES>, dim3(grid), dim3(32 * LWT_WARPS), args, lw_tma_smem, s));
                B->launches++;
            }

        }
        if (B->timing) { cudaEventRecord(e1, s); B->pending.push_back(std::make_pair(e0, e1)); }
        if (D.jcopy) { k_lwt_lambda_out<<<ntiles, 128, 0, s>>>(D, L); B->launches++; }
    }
    k_lw_finish<<<1, 1, 0, s>>>(P, D, L);
    if (D.jcopy && mrows > 0) { k_feedback<<<nblk(P.NJT, 128), 128, 0, s>>>(P, D); B->launches++; }   // Stage 4b: the lambdas are back in row order (k_lwt_lambda_out)
    k_integrate<<<nblk(NB, 128), 128, 0, s>>>(P, D);
    B->launches += 2;
    LCK(cudaGetLastError());
    return 1;
}
