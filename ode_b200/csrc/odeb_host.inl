// odeb_host.inl -- host side of the C-ABI (include/ode_b200.h): mirrors world/body/geom/joint
// descriptions into the SoA device buffers and launches the step kernels. No CPU fallback: every
// entry point fails loudly when no CUDA device is usable.

static thread_local std::string g_err;
static void set_err(const char *fmt, ...)
{
    char buf[512]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    g_err = buf;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { set_err("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); return 0; } } while (0)

struct OdebBatch {
    DevParams P;
    DevPtrs D;
    int device;
    cudaStream_t stream;
    std::vector<void *> allocs;
    size_t bytes;
    uint64_t launches;
    // optional per-kernel timing of k_solve
    bool timing; cudaEvent_t ev0, ev1; double solver_ms; int solver_launches;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t> > pending;
    // staging
    Real4 *d_stage; size_t stage_elems;
    Real4 *h_stage;
    int *h_ov;                                // [4] page-locked copy of D.overflow (read back by every blocking getter)
    Real4 *fb_jcopy, *fb_jfb;                 // joint-feedback buffers, kept across odeb_enable_feedback(0/1) toggles
    bool own_stream;                          // the stream was created here (odeb_set_stream replaces it with the caller's)
    // cuda graph of one step
    cudaGraphExec_t graph; double graph_h; bool use_graph; int graph_cfg;
    // solver selection: 0 = k_solve (one row at a time per world), 1..3 = k_solve5<2/4/8> (static P-processor schedule),
    // 4 = k_solve_bl (one lane per body). hint_m = largest island (rows) seen since the last sync, read back in odeb_sync.
    int s5_sr[4]; size_t s5_smem[4]; int hint_m; int solver_force; int lw_grid, lwc_grid, lw_maxgrid;   // lw_maxgrid: ODEB_LW_MAXGRID (tests: fewer blocks than tiles, so that warps walk several tiles per colour)
    bool rows_std3;                               // every joint is a contact of exactly three rows (normal + two friction directions): k_rows_t<true, true>
    bool env_no_solve6, env_no_hybrid, env_no_fused_rows; int env_hy_rows;   // experiment / test switches, read once at creation (ODEB_NO_SOLVE6, ODEB_NO_HYBRID, ODEB_NO_FUSED_ROWS, ODEB_TEST_HY_ROWS)   // s5_sr / s5_smem: row budget and shared memory of k_solve5<2^k> for the next launch
    int s6_sr, s6_nbi; size_t s6_smem;        // k_solve6<P> (odeb_solve6.cuh): row / body budget per island and shared memory per warp for the next launch
    int hint_nb, hint_nis;                    // largest island (bodies) and most islands with rows in one world seen so far
    int graph_sr; int graph_nlaunch;          // kernels per replay of the captured step (counted while capturing)
    size_t isl_smem;                         // shared memory of k_islands_t<true> per block, 0 = scratch in global memory
    int isl_one; size_t isl_one_smem;        // k_islands with one world per warp (4 per block) and its shared memory
    size_t solve_smem;
    int bl_G, bl_SR; size_t bl_smem;          // body-lane solver (odeb_solve_bl.cuh): lanes per world (0 = not used), row budget, bytes per warp
    void *flush_buf; size_t flush_bytes;
    // large-world path (ODEB_MODE_CANONICAL, odeb_large_host.inl)
    int mode; LargePtrs L; bool large_ready;
};
static int large_step(OdebBatch *B);

// B->h_ov holds a fresh copy of D.overflow (the stream is idle): refresh the solver hints, report and clear a capacity overflow.
// Device counters are reset on the batch's stream (a legacy-stream memset would not order against it).
static int overflow_check(OdebBatch *B)
{
    const int *ovh = B->h_ov;
    if (ovh[1] > 0) { B->hint_m = ovh[1]; B->hint_nb = ovh[2]; B->hint_nis = ovh[3]; CK(cudaMemsetAsync(B->D.overflow + 1, 0, 3 * sizeof(int), B->stream)); }
    const int ov = ovh[0];
    if (ov) {
        set_err("capacity overflow (%s): raise OdebWorldParams.max_pairs / max_contacts_per_world (or ODEB_MAX_PAIRS / ODEB_MAX_CONTACTS); the body state is the one the last complete step left",
                ov == 1 ? "pairs" : ov == 2 ? "contacts" : ov == 3 ? "rows" : "row groups on one body (large-world colouring: more than 254)");
        CK(cudaMemsetAsync(B->D.overflow, 0, sizeof(int), B->stream));
        CK(cudaStreamSynchronize(B->stream));
        return 0;
    }
    return 1;
}



template <class T> static bool dev_alloc(OdebBatch *B, T **p, size_t n)
{
    if (n == 0) n = 1;
    void *q = 0;
    cudaError_t e = cudaMalloc(&q, n * sizeof(T));
    if (e != cudaSuccess) { set_err("cudaMalloc(%zu bytes) failed: %s", n * sizeof(T), cudaGetErrorString(e)); return false; }
    cudaMemset(q, 0, n * sizeof(T));
    B->allocs.push_back(q); B->bytes += n * sizeof(T);
    *p = (T *)q;
    return true;
}
template <class T> static bool upload(T *dst, const std::vector<T> &v)
{
    if (v.empty()) return true;
    return cudaMemcpy(dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice) == cudaSuccess;
}

// dxFactorCholesky / dxSolveCholesky / dxInvertPDMatrix for n=3 (ode/src/matrix.cpp:107-253): dBodySetMass (ode.cpp:486-503)
static int host_invert_pd3(const Real *A, Real *Ainv)
{
    Real L[12], recip[3];
    memcpy(L, A, sizeof(L));
    for (int i = 0; i < 3; i++) {
        Real *aa = L + 4 * i;
        for (int j = 0; j < i; j++) {
            Real sum = aa[j];
            const Real *bb = L + 4 * j;
            for (int k = 0; k < j; k++) sum -= aa[k] * bb[k];
            aa[j] = sum * recip[j];
        }
        Real sum = aa[i];
        for (int k = 0; k < i; k++) sum -= aa[k] * aa[k];
        if (sum <= R_(0.0)) return 0;
        Real sq = RSQRT(sum);
        aa[i] = sq;
        recip[i] = rrecip(sq);
    }
    memset(Ainv, 0, 12 * sizeof(Real));
    for (int col = 0; col < 3; col++) {
        Real X[3] = { 0, 0, 0 }, y[3];
        X[col] = R_(1.0);
        for (int i = 0; i < 3; i++) {
            Real sum = R_(0.0);
            for (int k = 0; k < i; k++) sum += L[4 * i + k] * y[k];
            y[i] = (X[i] - sum) / L[4 * i + i];
        }
        for (int i = 3; i > 0;) {
            --i;
            Real sum = R_(0.0);
            for (int k = i + 1; k < 3; k++) sum += L[4 * k + i] * X[k];
            X[i] = (y[i] - sum) / L[4 * i + i];
        }
        for (int i = 0; i < 3; i++) Ainv[4 * i + col] = X[i];
    }
    return 1;
}

struct HostBody { Real pos[3], q[4], R[12]; };

// setAnchors joints/joint.cpp:289-320
static void host_set_anchors(const std::vector<HostBody> &hb, DJointT &j, Real x, Real y, Real z)
{
    if (j.b0 >= 0) {
        const HostBody &b0 = hb[j.b0];
        Real q[3] = { x - b0.pos[0], y - b0.pos[1], z - b0.pos[2] };
        mul1_331(j.anchor1, b0.R, q);
        if (j.b1 >= 0) {
            const HostBody &b1 = hb[j.b1];
            Real q2[3] = { x - b1.pos[0], y - b1.pos[1], z - b1.pos[2] };
            mul1_331(j.anchor2, b1.R, q2);
        } else { j.anchor2[0] = x; j.anchor2[1] = y; j.anchor2[2] = z; }
    }
    j.anchor1[3] = 0; j.anchor2[3] = 0;
}
// setAxes joints/joint.cpp:325-369
static void host_set_axes(const std::vector<HostBody> &hb, DJointT &j, Real x, Real y, Real z, Real *axis1, Real *axis2)
{
    if (j.b0 >= 0) {
        Real q[3] = { x, y, z };
        normalize3(q);
        if (axis1) { mul1_331(axis1, hb[j.b0].R, q); axis1[3] = 0; }
        if (axis2) {
            if (j.b1 >= 0) mul1_331(axis2, hb[j.b1].R, q);
            else { axis2[0] = x; axis2[1] = y; axis2[2] = z; }
            axis2[3] = 0;
        }
    }
}
static void host_limot(DLimot &l, Real erp, Real cfm, const OdebJointDesc &d, int a)
{   // dxJointLimitMotor::init + ::set joints/joint.cpp:494-540
    l.vel = 0; l.fmax = 0; l.lostop = -R_INF; l.histop = R_INF; l.fudge_factor = 1;
    l.normal_cfm = cfm; l.stop_erp = erp; l.stop_cfm = cfm; l.bounce = 0;
    l.lostop = (Real)d.lo_stop[a]; l.histop = (Real)d.hi_stop[a];
    l.vel = (Real)d.vel[a];
    if ((Real)d.fmax[a] >= 0) l.fmax = (Real)d.fmax[a];
    if (d.fudge_factor[a] >= 0 && (Real)d.fudge_factor[a] <= 1) l.fudge_factor = (Real)d.fudge_factor[a];
    if (d.bounce[a] >= 0) l.bounce = (Real)d.bounce[a];
    if (d.stop_erp[a] >= 0) l.stop_erp = (Real)d.stop_erp[a];
    if (d.stop_cfm[a] >= 0) l.stop_cfm = (Real)d.stop_cfm[a];
}


// dJointSetLMotorAxis lmotor.cpp:118-160 / dxJointAMotor::setAxisValue amotor.cpp:393-442 (x, y, z in the world frame)
static void host_motor_set_axis(const std::vector<HostBody> &hb, DJointT &j, int anum, int rel, Real x, Real y, Real z)
{
    if (j.type == ODEB_JOINT_LMOTOR) { if (j.b1 < 0 && rel == 2) rel = 1; }
    else if (rel != 0 && j.reverse) rel = 3 - rel;
    j.mrel[anum] = rel;
    Real r[3] = { x, y, z }, *a = j.maxis[anum];
    if (rel == 1) mul1_331(a, hb[j.b0].R, r);
    else if (rel == 2 && j.b1 >= 0) mul1_331(a, hb[j.b1].R, r);
    else { a[0] = x; a[1] = y; a[2] = z; }
    a[3] = 0;
    normalize3(a);
}
// dxJointAMotor::setEulerReferenceVectors amotor.cpp:768-796
static void host_amotor_euler_references(const std::vector<HostBody> &hb, DJointT &j)
{
    const int first = j.reverse ? 2 : 0, second = 2 - first;
    if (j.b1 >= 0) {
        Real r[3];
        mul0_331(r, hb[j.b0].R, j.maxis[first]);
        mul1_331(j.mref[1], hb[j.b1].R, r);
        mul0_331(r, hb[j.b1].R, j.maxis[second]);
        mul1_331(j.mref[0], hb[j.b0].R, r);
    } else {
        mul0_331(j.mref[1], hb[j.b0].R, j.maxis[first]);
        mul1_331(j.mref[0], hb[j.b0].R, j.maxis[second]);
    }
    j.mref[0][3] = j.mref[1][3] = 0;
}

// hinge.cpp:376-393 computeInitialRelativeRotation
static void host_hinge_initial_rotation(const std::vector<HostBody> &hb, DJointT &j)
{
    if (j.b0 < 0) return;
    if (j.b1 >= 0) qmul1(j.qrel, hb[j.b0].q, hb[j.b1].q);
    else { const Real *q = hb[j.b0].q; j.qrel[0] = q[0]; j.qrel[1] = -q[1]; j.qrel[2] = -q[2]; j.qrel[3] = -q[3]; }
}
// dJointSetFixed fixed.cpp:113-142: offset between the bodies in body 1's frame (anchor1) + computeInitialRelativeRotation
static void host_set_fixed(const std::vector<HostBody> &hb, DJointT &j)
{
    if (j.b0 < 0) return;
    const HostBody &b0 = hb[j.b0];
    if (j.b1 >= 0) {
        Real ofs[3] = { b0.pos[0] - hb[j.b1].pos[0], b0.pos[1] - hb[j.b1].pos[1], b0.pos[2] - hb[j.b1].pos[2] };
        mul1_331(j.anchor1, b0.R, ofs);
    } else { j.anchor1[0] = b0.pos[0]; j.anchor1[1] = b0.pos[1]; j.anchor1[2] = b0.pos[2]; }
    j.anchor1[3] = 0;
    host_hinge_initial_rotation(hb, j);          // fixed.cpp:172-192 is the same formula as hinge.cpp:376-393
}
// dJointSetHinge2Axes hinge2.cpp:283-312: getAxisInfo -> s0, c0 (kept in qrel[1], qrel[0]) and makeV1andV2 :214-240 (v1, v2 kept in
// qrel1, qrel2). Both bodies are required (dJOINT_TWOBODIES).
static void host_hinge2_finish(const std::vector<HostBody> &hb, DJointT &j)
{
    if (j.b0 < 0 || j.b1 < 0) return;
    Real ax1[3], ax2[3], ax[3], v[3];
    mul0_331(ax1, hb[j.b0].R, j.axis1);
    mul0_331(ax2, hb[j.b1].R, j.axis2);
    cross3(ax, ax1, ax2);
    j.qrel[1] = RSQRT(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
    j.qrel[0] = dot3(ax1, ax2);
    const Real k = dot3(ax1, ax2);
    ax2[0] = ax2[0] + ax1[0] * (-k); ax2[1] = ax2[1] + ax1[1] * (-k); ax2[2] = ax2[2] + ax1[2] * (-k);
    normalize3(ax2);
    cross3(v, ax1, ax2);
    mul1_331(j.qrel1, hb[j.b0].R, ax2); j.qrel1[3] = 0;
    mul1_331(j.qrel2, hb[j.b0].R, v); j.qrel2[3] = 0;
}
// dJointSetSliderAxis slider.cpp:249-260: setAxes(axis1) + computeOffset (:406-425, centre of body 1 in body 2's frame, kept in
// anchor1) + computeInitialRelativeRotation (:382-401)
static void host_set_slider_axis(const std::vector<HostBody> &hb, DJointT &j, Real x, Real y, Real z)
{
    host_set_axes(hb, j, x, y, z, j.axis1, 0);
    if (j.b0 < 0) return;
    const HostBody &b0 = hb[j.b0];
    if (j.b1 >= 0) {
        Real c[3] = { b0.pos[0] - hb[j.b1].pos[0], b0.pos[1] - hb[j.b1].pos[1], b0.pos[2] - hb[j.b1].pos[2] };
        mul1_331(j.anchor1, hb[j.b1].R, c);
    } else { j.anchor1[0] = b0.pos[0]; j.anchor1[1] = b0.pos[1]; j.anchor1[2] = b0.pos[2]; }
    j.anchor1[3] = 0;
    host_hinge_initial_rotation(hb, j);
}
// universal.cpp:372-401 computeInitialRelativeRotations
static void host_universal_initial_rotations(const std::vector<HostBody> &hb, DJointT &j)
{
    if (j.b0 < 0) return;
    Real ax1[3], ax2[3], R[12] = { 0 }, qcross[4];
    mul0_331(ax1, hb[j.b0].R, j.axis1);
    if (j.b1 >= 0) mul0_331(ax2, hb[j.b1].R, j.axis2); else { ax2[0] = j.axis2[0]; ax2[1] = j.axis2[1]; ax2[2] = j.axis2[2]; }
    odeb_r_from_2axes(R, ax1[0], ax1[1], ax1[2], ax2[0], ax2[1], ax2[2]);
    q_from_r(qcross, R);
    qmul1(j.qrel1, hb[j.b0].q, qcross);
    odeb_r_from_2axes(R, ax2[0], ax2[1], ax2[2], ax1[0], ax1[1], ax1[2]);
    q_from_r(qcross, R);
    if (j.b1 >= 0) qmul1(j.qrel2, hb[j.b1].q, qcross); else for (int k = 0; k < 4; k++) j.qrel2[k] = qcross[k];
}
// make_sure_plane_normal_has_unit_length plane.cpp:48-63
static void host_normalize_plane(Real *p)
{
    Real l = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
    if (l > 0) { l = rrecipsqrt(l); p[0] *= l; p[1] *= l; p[2] *= l; p[3] *= l; }
    else { p[0] = 1; p[1] = 0; p[2] = 0; p[3] = 0; }
}

// world-level parameters -> DevParams (dWorldSet*; defaults ode/src/objects.cpp:37-121). Also used by the classic layer
// before every launch, because the classic setters may change them between steps.
static void apply_world_params(DevParams &P, const OdebWorldParams *wp, bool classic)
{
    DSurface &S = P.surf;
    S.mode = wp->surf_mode;
    S.mu = (Real)wp->mu < 0 ? 0 : (Real)wp->mu;
    S.mu2 = (Real)wp->mu2 < 0 ? 0 : (Real)wp->mu2;
    S.bounce = (Real)wp->bounce; S.bounce_vel = (Real)wp->bounce_vel; S.soft_erp = (Real)wp->soft_erp; S.soft_cfm = (Real)wp->soft_cfm;
    S.motion1 = (Real)wp->motion1; S.motion2 = (Real)wp->motion2; S.motionN = (Real)wp->motionN; S.slip1 = (Real)wp->slip1; S.slip2 = (Real)wp->slip2;
    {   // getInfo1 row count of a contact joint (contact.cpp:48-122); uniform under one policy, per contact in classic mode
        const int m = odeb_contact_rows(S.mode, S.mu, S.mu2, (Real)wp->rho, (Real)wp->rho2, (Real)wp->rhoN);
        S.rho = (Real)wp->rho < 0 ? 0 : (Real)wp->rho; S.rho2 = (Real)wp->rho2 < 0 ? 0 : (Real)wp->rho2; S.rhoN = (Real)wp->rhoN < 0 ? 0 : (Real)wp->rhoN;
        S.the_m = m; P.m_contact = classic ? 6 : m;
    }
    for (int k = 0; k < 3; k++) P.gravity[k] = (Real)wp->gravity[k];
    P.erp = (Real)wp->erp;
#if defined(ODEB_DOUBLE)
    P.cfm = wp->cfm >= 0 ? (Real)wp->cfm : R_(1e-10);
#else
    P.cfm = wp->cfm >= 0 ? (Real)wp->cfm : R_(1e-5);
#endif
    P.sor_w = (Real)wp->sor_w;
    P.num_iter = wp->num_iterations > 1 ? wp->num_iterations : 1;
    P.premature_delta = (Real)wp->premature_exit_delta; P.extra_delta = (Real)wp->extra_iter_delta;
    { Real f = (Real)wp->max_extra_factor; Real ex = P.num_iter * f; P.max_extra = ex < (Real)UINT32_MAX ? (unsigned)ex : UINT32_MAX; }  // objects.h:177-184
    P.dyn_enabled = (P.max_extra != 0 || P.premature_delta != 0) ? 1 : 0;
    P.max_vel = (Real)wp->contact_max_vel; P.min_depth = (Real)wp->contact_surface_layer;
    { Real t = (Real)wp->adis_linear_thr; P.adis_lin = t * t; t = (Real)wp->adis_angular_thr; P.adis_ang = t * t; }
    P.adis_time = (Real)wp->adis_time; P.adis_steps = wp->adis_steps; P.adis_samples = wp->adis_samples;
    P.damp_lin_scale = (Real)wp->linear_damping; P.damp_ang_scale = (Real)wp->angular_damping;
    { Real t = (Real)wp->linear_damping_thr; P.damp_lin_thr = t * t; t = (Real)wp->angular_damping_thr; P.damp_ang_thr = t * t; }
    P.max_ang_speed = (Real)wp->max_angular_speed;
}

static inline unsigned nblk(size_t n, unsigned bs) { return (unsigned)((n + bs - 1) / bs); }

extern "C" {

const char *odeb_last_error(void) { return g_err.c_str(); }

void odeb_destroy(OdebBatch *B)
{
    if (!B) return;
    cudaSetDevice(B->device);
    cudaDeviceSynchronize();
    for (size_t i = 0; i < B->pending.size(); i++) { cudaEventDestroy(B->pending[i].first); cudaEventDestroy(B->pending[i].second); }
    if (B->graph) cudaGraphExecDestroy(B->graph);
    for (size_t i = 0; i < B->allocs.size(); i++) cudaFree(B->allocs[i]);
    if (B->h_stage) cudaFreeHost(B->h_stage);
    if (B->h_ov) cudaFreeHost(B->h_ov);
    if (B->flush_buf) cudaFree(B->flush_buf);
    if (B->stream && B->own_stream) cudaStreamDestroy(B->stream);
    delete B;
}


// Host-side template of one world: what odeb_create derives from the scene description and what the classic
// per-object layer (odeb_classic.inl) maintains incrementally.
struct HostTemplate {
    std::vector<Real> bmass, binvmass, bI, binvI;            // [NB], [NB*12]
    std::vector<HostBody> hb;                                // pose of every body
    std::vector<int> bflags0;                                // BF_* per body
    std::vector<int> gtype, gbody; std::vector<Real> gparam; std::vector<unsigned> gcat, gcol;
    std::vector<Real4> gspose;                               // [NG*4]: position + 3 rotation rows of geoms without a body / offset pose
    std::vector<int> gofs;                                   // [NG]: geom has an offset relative to its body
    std::vector<DJointT> jt;
    std::vector<int> sofs, sj, so;                           // per-body joint adjacency in attach order (CSR)
};

struct BatchCaps { int classic, max_pairs, max_contacts; };  // classic: contacts / surfaces / adjacency come from the host every step

static OdebBatch *batch_build(const OdebWorldParams *wp, const HostTemplate &T, int nworlds, int device, const BatchCaps *caps)
{
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) { set_err("no CUDA device: %s (this library has no CPU path)", cudaGetErrorString(e)); return 0; }
    if (device < 0 || device >= ndev) { set_err("bad device index %d (have %d)", device, ndev); return 0; }
    if (cudaSetDevice(device) != cudaSuccess) { set_err("cudaSetDevice(%d) failed", device); return 0; }
    const int nbody = (int)T.bmass.size(), ngeom = (int)T.gtype.size(), njoint = (int)T.jt.size();
    const bool classic = caps && caps->classic;
    if (nworlds <= 0 || nbody <= 0) { set_err("bad sizes"); return 0; }
    if (!classic && (wp->surf_mode & ODEB_CONTACT_FDIR1)) { set_err("contact mode FDir1 needs a per-contact direction: use the classic API"); return 0; }
    if (wp->max_contacts < 1 || wp->max_contacts > 8) { set_err("max_contacts must be in 1..8"); return 0; }

    OdebBatch *B = new OdebBatch();
    B->device = device; B->bytes = 0; B->launches = 0; B->timing = false; B->solver_ms = 0; B->solver_launches = 0;
    B->mode = ODEB_MODE_REPLAY; B->large_ready = false; memset(&B->L, 0, sizeof(B->L));
    B->graph = 0; B->graph_h = -1; B->graph_cfg = -1; B->graph_sr = -1; B->hint_m = 0; B->hint_nb = 0; B->hint_nis = 0; B->s6_sr = 0; B->s6_nbi = 0; B->s6_smem = 0; B->solver_force = -1; B->lw_grid = 0; B->lwc_grid = 0; B->lw_maxgrid = getenv("ODEB_LW_MAXGRID") ? atoi(getenv("ODEB_LW_MAXGRID")) : 0; B->env_no_solve6 = getenv("ODEB_NO_SOLVE6") != 0; B->env_no_hybrid = getenv("ODEB_NO_HYBRID") != 0; B->env_no_fused_rows = getenv("ODEB_NO_FUSED_ROWS") != 0; B->env_hy_rows = getenv("ODEB_TEST_HY_ROWS") ? atoi(getenv("ODEB_TEST_HY_ROWS")) : 0; B->use_graph = getenv("ODEB_NO_GRAPH") == 0 && !classic; B->h_stage = 0; B->h_ov = 0; B->fb_jcopy = 0; B->fb_jfb = 0; B->own_stream = true; B->d_stage = 0; B->stream = 0; B->flush_buf = 0; B->flush_bytes = 0;
    memset(&B->D, 0, sizeof(B->D));
    DevParams &P = B->P;
    memset(&P, 0, sizeof(P));
    P.W = nworlds; P.NB = nbody; P.NG = ngeom; P.NJ = njoint; P.classic = classic ? 1 : 0;
    P.maxc = wp->max_contacts; P.space_type = wp->space_type; P.skip_connected = wp->skip_connected;
    P.hash_minlevel = wp->hash_levels_set ? wp->hash_minlevel : -3; P.hash_maxlevel = wp->hash_levels_set ? wp->hash_maxlevel : 10;   // dxHashSpace::dxHashSpace collision_space.cpp:383-389
    {
        long long all = (long long)ngeom * (ngeom - 1) / 2;
        long long mp = all < 16LL * ngeom ? all : 16LL * ngeom;
        if (const char *s = getenv("ODEB_MAX_PAIRS")) mp = atoll(s);
        if (wp->max_pairs > 0) mp = wp->max_pairs;
        if (caps && caps->max_pairs > 0) mp = caps->max_pairs;
        P.MP = (int)(mp < 1 ? 1 : mp);
        long long mc = (long long)P.MP * P.maxc;
        long long cap = (nworlds == 1 ? 4LL : 2LL) * ngeom * P.maxc;   // a single (large) world may be a dense wall / pile
        if (mc > cap) mc = cap;
        if (const char *s = getenv("ODEB_MAX_CONTACTS")) mc = atoll(s);
        if (wp->max_contacts_per_world > 0) mc = wp->max_contacts_per_world;
        if (caps && caps->max_contacts > 0) mc = caps->max_contacts;
        P.MC = (int)(mc < 1 ? 1 : mc);
    }
    apply_world_params(P, wp, classic);
    P.NJT = P.NJ + P.MC;
    P.MR = P.MC * P.m_contact + P.NJ * 6;
    B->rows_std3 = !classic && P.NJ == 0 && P.surf.the_m == 3 && !(P.surf.mode & 0x400) && P.surf.mu > 0 && ((P.surf.mode & 0x001) ? P.surf.mu2 : P.surf.mu) > 0
                   && getenv("ODEB_NO_STD3_ROWS") == 0;
    {   // shared-memory budget of k_solve (16 worlds per warp): ring + per-body accumulators + lambda/metadata for SR rows
        int sr = P.MR < 512 ? P.MR : 512;
        if (const char *s = getenv("ODEB_SOLVER_ROWS")) sr = atoi(s);
        if (sr > 1023) sr = 1023;                      // metadata word: 10-bit friction-index row
        if (P.NB + 1 > 2047) sr = 0;                   // metadata word: 11-bit body slots -> global path only
        {
            const size_t budget = 112 * 1024;          // two resident warps' worth per SM (227 KB / 2)
            const size_t fixed = (size_t)ODEB_RING * ODEB_HALF_CHUNKS * 32 * 16 + 2 * (size_t)(P.NB + 1) * ODEB_WPW * sizeof(Real4)
                               + ODEB_WPW * sizeof(Real) + (size_t)ODEB_META_PAD * 32 * sizeof(unsigned);
            const size_t per_row = ODEB_WPW * sizeof(Real) + 32 * sizeof(unsigned);
            if (fixed + per_row > budget) sr = 0;
            else if (fixed + (size_t)sr * per_row > budget) sr = (int)((budget - fixed) / per_row);
            B->solve_smem = ((fixed + (size_t)sr * per_row + 31) / 32) * 32;
        }
        P.SR = sr;
        if (P.SR == 0) B->solve_smem = 0;
    }
    {   // body-lane solver: one lane per body, G lanes per world. Row budget sized so that at least `want` warps are resident per SM.
        B->bl_G = 0; B->bl_SR = 0; B->bl_smem = 0;
        const char *sel = getenv("ODEB_SOLVER");
        if (sel && strcmp(sel, "bl") == 0 && P.NB <= 32) {
            const int G = P.NB <= 8 ? 8 : P.NB <= 16 ? 16 : 32;
            int sr = P.MR < ODEB_BL_MAXROWS ? P.MR : ODEB_BL_MAXROWS;
            const size_t budget = (size_t)(227 * 1024) / (sizeof(Real) == 4 ? 14 : 7) - 1024;   // 14 (single) / 7 (double) resident warps per SM; 1 KB per block is reserved by the system
            while (sr > 64 && odeb_bl_smem(G, sr) > budget) sr -= 8;
            if (const char *s = getenv("ODEB_BL_ROWS")) { sr = atoi(s); if (sr > ODEB_BL_MAXROWS) sr = ODEB_BL_MAXROWS; if (sr < 1) sr = 1; }
            B->bl_G = G; B->bl_SR = sr; B->bl_smem = odeb_bl_smem(G, sr);
        }
    }

    {   // k_solve5<P>: the row budget is chosen per launch from the largest island seen so far (choose_solver)
        for (int k = 0; k <= 3; k++) { B->s5_sr[k] = 0; B->s5_smem[k] = 0; }
        if (const char *sel = getenv("ODEB_SOLVER")) {
            if (!strcmp(sel, "v4")) B->solver_force = 0;
            else if (!strcmp(sel, "p2")) B->solver_force = 1;
            else if (!strcmp(sel, "p4")) B->solver_force = 2;
            else if (!strcmp(sel, "p8")) B->solver_force = 3;
            else if (!strcmp(sel, "bl")) B->solver_force = 4;
            else if (!strcmp(sel, "hy")) B->solver_force = 5;
            else if (!strcmp(sel, "d1")) B->solver_force = 6;
            else if (!strcmp(sel, "d2")) B->solver_force = 7;
            else if (!strcmp(sel, "d4")) B->solver_force = 8;
            else if (!strcmp(sel, "d8")) B->solver_force = 9;
        }
    }

    // ---- device buffers
    DevPtrs &D = B->D;
    const size_t W = nworlds, WB = W * nbody, WG = W * ngeom, WJ = W * njoint;
    const int NS = P.adis_samples > 0 ? P.adis_samples : 1;
    const size_t nadj = classic ? 2 * (size_t)(njoint + P.MC) : T.sj.size();
    bool ok = true;
    ok = ok && dev_alloc(B, &D.pos, WB) && dev_alloc(B, &D.quat, WB) && dev_alloc(B, &D.lvel, WB) && dev_alloc(B, &D.avel, WB)
            && dev_alloc(B, &D.facc, WB) && dev_alloc(B, &D.tacc, WB) && dev_alloc(B, &D.R, 3 * WB)
            && dev_alloc(B, &D.bflags, WB) && dev_alloc(B, &D.adis_steps, WB) && dev_alloc(B, &D.adis_time, WB)
            && dev_alloc(B, &D.avg_buf, WB * 6 * NS) && dev_alloc(B, &D.avg_counter, WB) && dev_alloc(B, &D.avg_ready, WB);
    ok = ok && dev_alloc(B, &D.bmass, nbody) && dev_alloc(B, &D.binvmass, nbody) && dev_alloc(B, &D.bI, 12 * (size_t)nbody) && dev_alloc(B, &D.binvI, 12 * (size_t)nbody)
            && dev_alloc(B, &D.gtype, ngeom) && dev_alloc(B, &D.gbody, ngeom) && dev_alloc(B, &D.gparam, 4 * (size_t)ngeom)
            && dev_alloc(B, &D.gcat, ngeom) && dev_alloc(B, &D.gcol, ngeom) && dev_alloc(B, &D.gspose, 4 * (size_t)ngeom) && dev_alloc(B, &D.gofs, ngeom) && dev_alloc(B, &D.joints, njoint)
            && dev_alloc(B, &D.sadj_ofs, nbody + 1) && dev_alloc(B, &D.sadj_joint, nadj) && dev_alloc(B, &D.sadj_other, nadj);
    ok = ok && dev_alloc(B, &D.aabb, WG * 6) && dev_alloc(B, &D.pair_cnt, WG) && dev_alloc(B, &D.pair_ofs, WG) && dev_alloc(B, &D.npairs, W)
            && dev_alloc(B, &D.pairs, W * P.MP) && dev_alloc(B, &D.pc_count, W * P.MP)
            && (std::find(T.gtype.begin(), T.gtype.end(), (int)ODEB_RAY) == T.gtype.end() || classic
                || (dev_alloc(B, &D.ray_count, W * P.MP) && dev_alloc(B, &D.ray_geom, ngeom) && dev_alloc(B, &D.ray_range, WG) && dev_alloc(B, &D.ray_hit, WG))) && dev_alloc(B, &D.cgeom, W * (classic ? (size_t)P.MC : (size_t)P.MP * P.maxc) * 2)
            && dev_alloc(B, &D.ncontacts, W) && dev_alloc(B, &D.cinfo, W * P.MC) && dev_alloc(B, &D.jm, WJ) && dev_alloc(B, &D.jlimit, WJ);
    if (classic) ok = ok && dev_alloc(B, &D.csurf, (size_t)P.MC);
    ok = ok && dev_alloc(B, &D.c_ofs, W * (nbody + 1)) && dev_alloc(B, &D.c_cur, WB) && dev_alloc(B, &D.c_adj_c, W * 2 * P.MC) && dev_alloc(B, &D.c_adj_o, W * 2 * P.MC)
            && dev_alloc(B, &D.btag, WB) && dev_alloc(B, &D.jtag, W * P.NJT) && dev_alloc(B, &D.stack, WB)
            && dev_alloc(B, &D.body_order, WB) && dev_alloc(B, &D.body_pos, WB) && dev_alloc(B, &D.body_island, WB)
            && dev_alloc(B, &D.joint_order, W * P.NJT) && dev_alloc(B, &D.joint_row, W * P.NJT) && dev_alloc(B, &D.joint_island, W * P.NJT)
            && dev_alloc(B, &D.island_info, WB) && dev_alloc(B, &D.nislands, W) && dev_alloc(B, &D.nordered, W) && dev_alloc(B, &D.njord, W) && dev_alloc(B, &D.mrows, W);
    ok = ok && dev_alloc(B, &D.rows, W * P.MR * 8) && dev_alloc(B, &D.rbody, W * P.MR) && dev_alloc(B, &D.findex, W * P.MR + 8) && dev_alloc(B, &D.order, W * P.MR) && dev_alloc(B, &D.order0, W * P.MR) && dev_alloc(B, &D.wkey, W) && dev_alloc(B, &D.wlist, W)
            && dev_alloc(B, &D.lambda, W * P.MR + 8) && dev_alloc(B, &D.cforce, W * (nbody + 1) * 2) && dev_alloc(B, &D.invIw, WB * 12)
            && dev_alloc(B, &D.stats, W * 4) && dev_alloc(B, &D.seed, W) && dev_alloc(B, &D.sweeps, 2 * W) && dev_alloc(B, &D.overflow, 4) && dev_alloc(B, &D.isl_done, WB) && dev_alloc(B, &D.maxpairs, 1);
    B->stage_elems = 4 * WB;                     // room for the tightly packed state of every body (13 reals) in one transfer
    ok = ok && dev_alloc(B, &B->d_stage, 4 * WB);
    if (ok && cudaMallocHost((void **)&B->h_stage, 4 * WB * sizeof(Real4)) != cudaSuccess) { set_err("cudaMallocHost failed"); ok = false; }
    if (ok && cudaMallocHost((void **)&B->h_ov, 4 * sizeof(int)) != cudaSuccess) { set_err("cudaMallocHost failed"); ok = false; }
    if (ok && cudaStreamCreateWithFlags(&B->stream, cudaStreamNonBlocking) != cudaSuccess) { set_err("cudaStreamCreate failed"); ok = false; }
    {   // island replay scratch in shared memory when 32 worlds fit (and the 16-bit indices hold)
        const size_t need = odeb_islands_smem(P.NB, P.MC, P.NJT);
        B->isl_smem = (need <= 200 * 1024 && 2 * (size_t)P.MC < 65535 && (size_t)P.NB < 65535 && !getenv("ODEB_ISLANDS_GLOBAL")) ? need : 0;
        if (ok && cudaFuncSetAttribute(k_islands_t<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) { set_err("cudaFuncSetAttribute(k_islands) failed"); ok = false; }
        // one world per warp (4 worlds per block) once a world is more than a handful of bodies: see k_islands_t
        B->isl_one = 0; B->isl_one_smem = 0;
        {
            const size_t per_world = (need / 32 + 15) / 16 * 16;
            const char *e = getenv("ODEB_ISLANDS_ONE");
            const bool want = e ? atoi(e) != 0 : P.NB > 24;
            if (want && B->isl_smem && 4 * per_world <= 200 * 1024) {
                B->isl_one = 1; B->isl_one_smem = 4 * per_world;
                if (ok && cudaFuncSetAttribute(k_islands_t<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) { set_err("cudaFuncSetAttribute(k_islands) failed"); ok = false; }
            }
        }
    }
    if (ok) {   // opt every solver kernel into the full 227 KB of shared memory once (the attribute is per function, not per batch)
        const int mx = 227 * 1024;
        cudaError_t ce = cudaFuncSetAttribute(k_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(k_solve5_t<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(k_solve5_t<4, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(k_solve5_t<8, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(k_solve_hy, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(k_solve5_t<4, ODEB_HYBRID_SWEEPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(k_solve6_t<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(k_solve6_t<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(k_solve6_t<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(k_solve6_t<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(k_solve_bl<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(k_solve_bl<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(k_solve_bl<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        if (ce != cudaSuccess) { set_err("cudaFuncSetAttribute(max dynamic shared memory) failed: %s", cudaGetErrorString(ce)); ok = false; }
        // the solvers live on shared memory and never reuse a global line through L1: ask for the largest carveout
        cudaFuncSetAttribute(k_solve, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(k_solve_hy, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(k_solve5_t<4, ODEB_HYBRID_SWEEPS>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(k_solve5_t<2, 0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(k_solve5_t<4, 0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(k_solve5_t<8, 0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(k_solve6_t<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(k_solve6_t<2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(k_solve6_t<4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(k_solve6_t<8>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(k_solve_bl<8>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(k_solve_bl<16>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(k_solve_bl<32>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (getenv("ODEB_DEBUG")) {
            int nb0 = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb0, k_solve, 32, B->solve_smem);
            fprintf(stderr, "[odeb] k_solve: SR=%d smem=%zu blocks/SM=%d\n", P.SR, B->solve_smem, nb0);
        }
    }
    if (!ok) { odeb_destroy(B); return 0; }

    ok = upload(D.bmass, T.bmass) && upload(D.binvmass, T.binvmass) && upload(D.bI, T.bI) && upload(D.binvI, T.binvI)
      && upload(D.gtype, T.gtype) && upload(D.gbody, T.gbody) && upload(D.gparam, T.gparam) && upload(D.gcat, T.gcat) && upload(D.gcol, T.gcol)
      && upload(D.gspose, T.gspose) && upload(D.gofs, T.gofs)
      && upload(D.joints, T.jt) && upload(D.sadj_ofs, T.sofs) && upload(D.sadj_joint, T.sj) && upload(D.sadj_other, T.so);
    if (ok && D.ray_geom) {     // ray geoms in geom order (odeb_get_ray_ranges)
        std::vector<int> rg;
        for (int i = 0; i < ngeom; i++) if (T.gtype[i] == ODEB_RAY) rg.push_back(i);
        D.nray = (int)rg.size();
        ok = cudaMemcpy(D.ray_geom, rg.data(), rg.size() * sizeof(int), cudaMemcpyHostToDevice) == cudaSuccess;
    }
    // initial per-world state = template pose
    {
        std::vector<Real4> pos(WB), quat(WB), R(3 * WB);
        std::vector<int> fl(WB), st(WB); std::vector<Real> tl(WB);
        for (size_t w = 0; w < W; w++) for (int i = 0; i < nbody; i++) {
            size_t k = w * nbody + i;
            const HostBody &h = T.hb[i];
            Real4 p = { h.pos[0], h.pos[1], h.pos[2], 0 }, q = { h.q[0], h.q[1], h.q[2], h.q[3] };
            pos[k] = p; quat[k] = q;
            Real4 r0 = { h.R[0], h.R[1], h.R[2], 0 }, r1 = { h.R[4], h.R[5], h.R[6], 0 }, r2 = { h.R[8], h.R[9], h.R[10], 0 };
            R[3 * k] = r0; R[3 * k + 1] = r1; R[3 * k + 2] = r2;
            fl[k] = T.bflags0[i]; st[k] = P.adis_steps; tl[k] = P.adis_time;
        }
        ok = ok && upload(D.pos, pos) && upload(D.quat, quat) && upload(D.R, R) && upload(D.bflags, fl) && upload(D.adis_steps, st) && upload(D.adis_time, tl);
    }
    if (!ok || cudaDeviceSynchronize() != cudaSuccess) { set_err("template upload failed: %s", cudaGetErrorString(cudaGetLastError())); odeb_destroy(B); return 0; }
    return B;
}

OdebBatch *odeb_create(const OdebWorldParams *wp,
                       int nbody, const OdebBodyDesc *bodies, const double *body_pos, const double *body_quat,
                       int ngeom, const OdebGeomDesc *geoms,
                       int njoint, const OdebJointDesc *joints,
                       int nworlds, int device)
{
    if (nworlds <= 0 || nbody <= 0 || ngeom < 0 || njoint < 0) { set_err("bad sizes"); return 0; }
    HostTemplate T;
    Real erp = (Real)wp->erp;
#if defined(ODEB_DOUBLE)
    Real cfm = wp->cfm >= 0 ? (Real)wp->cfm : R_(1e-10);
#else
    Real cfm = wp->cfm >= 0 ? (Real)wp->cfm : R_(1e-5);
#endif
    // ---- template on the host
    T.bmass.resize(nbody); T.binvmass.resize(nbody); T.bI.assign(12 * (size_t)nbody, 0); T.binvI.assign(12 * (size_t)nbody, 0);
    T.hb.resize(nbody); T.bflags0.resize(nbody);
    std::vector<HostBody> &hb = T.hb;
    for (int i = 0; i < nbody; i++) {
        T.bmass[i] = (Real)bodies[i].mass;
        if (!(T.bmass[i] > 0)) { set_err("body %d: mass must be > 0", i); return 0; }
        T.binvmass[i] = rrecip(T.bmass[i]);
        const double *I = bodies[i].inertia;
        Real *Ib = &T.bI[12 * i];
        Ib[0] = (Real)I[0]; Ib[5] = (Real)I[4]; Ib[10] = (Real)I[8];          // dMassSetParameters mass.cpp:74-93
        Ib[1] = (Real)I[1]; Ib[2] = (Real)I[2]; Ib[6] = (Real)I[5];
        Ib[4] = (Real)I[1]; Ib[8] = (Real)I[2]; Ib[9] = (Real)I[5];
        if (!host_invert_pd3(Ib, &T.binvI[12 * i])) { Real *v = &T.binvI[12 * i]; memset(v, 0, 12 * sizeof(Real)); v[0] = v[5] = v[10] = 1; }
        if (bodies[i].flags & ODEB_BODY_KINEMATIC) { memset(&T.binvI[12 * i], 0, 12 * sizeof(Real)); T.binvmass[i] = 0; }     // dBodySetKinematic ode.cpp:837-842
        HostBody &h = hb[i];
        for (int k = 0; k < 3; k++) h.pos[k] = (Real)body_pos[3 * i + k];
        for (int k = 0; k < 4; k++) h.q[k] = (Real)body_quat[4 * i + k];
        normalize4(h.q); r_from_q(h.R, h.q);
        int fl = BF_GYRO;
        if (wp->auto_disable) fl |= BF_AUTO_DISABLE;
        if ((Real)wp->linear_damping) fl |= BF_LIN_DAMP;
        if ((Real)wp->angular_damping) fl |= BF_ANG_DAMP;
        if ((Real)wp->max_angular_speed < R_INF) fl |= BF_MAX_ANG_SPEED;
        int sf = bodies[i].flags;
        if (sf & ODEB_BODY_NO_GRAVITY) fl |= BF_NO_GRAVITY;
        if (sf & ODEB_BODY_NO_GYRO) fl &= ~BF_GYRO;
        if (sf & ODEB_BODY_FINITE_ROTATION) fl |= BF_FINITE_ROT;
        if (sf & ODEB_BODY_DISABLED) fl |= BF_DISABLED;
        T.bflags0[i] = fl;
    }
    T.gtype.resize(ngeom); T.gbody.resize(ngeom); T.gparam.resize(4 * (size_t)ngeom); T.gcat.resize(ngeom); T.gcol.resize(ngeom);
    T.gspose.resize(4 * (size_t)ngeom); T.gofs.assign(ngeom, 0);
    for (int i = 0; i < ngeom; i++) {
        T.gtype[i] = geoms[i].type; T.gbody[i] = geoms[i].body; T.gcat[i] = geoms[i].category_bits; T.gcol[i] = geoms[i].collide_bits;
        if (T.gbody[i] >= nbody) { set_err("geom %d: bad body index", i); return 0; }
        if (T.gtype[i] != ODEB_SPHERE && T.gtype[i] != ODEB_BOX && T.gtype[i] != ODEB_CAPSULE && T.gtype[i] != ODEB_PLANE && T.gtype[i] != ODEB_CYLINDER && T.gtype[i] != ODEB_RAY) { set_err("geom %d: unsupported class %d", i, T.gtype[i]); return 0; }
        Real *p = &T.gparam[4 * i];
        for (int k = 0; k < 4; k++) p[k] = (Real)geoms[i].p[k];
        if (T.gtype[i] == ODEB_PLANE) host_normalize_plane(p);
        Real4 z = { 0, 0, 0, 0 }, r0 = { 1, 0, 0, 0 }, r1 = { 0, 1, 0, 0 }, r2 = { 0, 0, 1, 0 };
        T.gspose[4 * i] = z; T.gspose[4 * i + 1] = r0; T.gspose[4 * i + 2] = r1; T.gspose[4 * i + 3] = r2;
        if (geoms[i].has_offset && T.gbody[i] >= 0) {      // dGeomSetOffsetPosition / dGeomSetOffsetQuaternion (dRfromQ)
            Real q[4] = { (Real)geoms[i].offset_quat[0], (Real)geoms[i].offset_quat[1], (Real)geoms[i].offset_quat[2], (Real)geoms[i].offset_quat[3] }, oR[12];
            r_from_q(oR, q);
            Real4 op = { (Real)geoms[i].offset_pos[0], (Real)geoms[i].offset_pos[1], (Real)geoms[i].offset_pos[2], 0 };
            Real4 a = { oR[0], oR[1], oR[2], 0 }, b = { oR[4], oR[5], oR[6], 0 }, c = { oR[8], oR[9], oR[10], 0 };
            T.gspose[4 * i] = op; T.gspose[4 * i + 1] = a; T.gspose[4 * i + 2] = b; T.gspose[4 * i + 3] = c;
            T.gofs[i] = 1;
        }
    }
    T.jt.resize(njoint);
    std::vector<std::vector<std::pair<int, int> > > adj(nbody);
    for (int i = 0; i < njoint; i++) {
        const OdebJointDesc &d = joints[i];
        DJointT &j = T.jt[i];
        memset(&j, 0, sizeof(j));
        if (d.type != ODEB_JOINT_BALL && d.type != ODEB_JOINT_HINGE && d.type != ODEB_JOINT_UNIVERSAL && d.type != ODEB_JOINT_FIXED && d.type != ODEB_JOINT_SLIDER && d.type != ODEB_JOINT_HINGE2 && d.type != ODEB_JOINT_AMOTOR && d.type != ODEB_JOINT_LMOTOR) { set_err("joint %d: unsupported type %d", i, d.type); return 0; }
        j.type = d.type; j.erp = erp; j.cfm = cfm;
        int b1 = d.body1, b2 = d.body2;
        if (b1 >= nbody || b2 >= nbody || (b1 < 0 && b2 < 0) || b1 == b2) { set_err("joint %d: bad bodies", i); return 0; }
        if (b1 < 0) { b1 = b2; b2 = -1; j.reverse = 1; }         // dJointAttach ode.cpp:1404-1411
        j.b0 = b1; j.b1 = b2;
        adj[b1].push_back(std::make_pair(i, b2));
        if (b2 >= 0) adj[b2].push_back(std::make_pair(i, b1));
        host_set_anchors(hb, j, (Real)d.anchor[0], (Real)d.anchor[1], (Real)d.anchor[2]);
        host_limot(j.limot1, erp, cfm, d, 0);
        host_limot(j.limot2, erp, cfm, d, 1);
        host_limot(j.limot3, erp, cfm, d, 2);
        if (j.type == ODEB_JOINT_LMOTOR || j.type == ODEB_JOINT_AMOTOR) {
            // dJointSet{L,A}MotorNumAxes, dJointSetAMotorMode (amotor.cpp:354-377: Euler mode has 3 axes, axis 1 derived), axes, user angles
            const bool am = j.type == ODEB_JOINT_AMOTOR;
            if (d.motor_num < 0 || d.motor_num > 3) { set_err("joint %d: motor_num must be 0..3", i); return 0; }
            j.mmode = am ? d.motor_mode : 0;
            j.mnum = (am && j.mmode == 1) ? 3 : d.motor_num;
            for (int k = 0; k < d.motor_num; k++) {
                if (am && j.mmode == 1 && k == 1) continue;
                host_motor_set_axis(hb, j, k, d.motor_rel[k], (Real)d.motor_axis[k][0], (Real)d.motor_axis[k][1], (Real)d.motor_axis[k][2]);
            }
            if (am && j.mmode == 1) host_amotor_euler_references(hb, j);
            for (int k = 0; k < 3; k++) j.mangle[k] = am ? (Real)d.motor_angle[k] : 0;
        } else if (j.type == ODEB_JOINT_HINGE2) {
            if (j.b1 < 0 || j.reverse) { set_err("joint %d: a hinge2 joint needs two bodies", i); return 0; }
            j.axis1[0] = 1; j.axis2[1] = 1;                       // hinge2.cpp:76-100
            host_set_axes(hb, j, (Real)d.axis1[0], (Real)d.axis1[1], (Real)d.axis1[2], j.axis1, 0);
            host_set_axes(hb, j, (Real)d.axis2[0], (Real)d.axis2[1], (Real)d.axis2[2], 0, j.axis2);
            host_hinge2_finish(hb, j);
            j.qrel[2] = d.susp_erp >= 0 ? (Real)d.susp_erp : erp;
            j.qrel[3] = d.susp_cfm >= 0 ? (Real)d.susp_cfm : cfm;
        } else if (j.type == ODEB_JOINT_FIXED) host_set_fixed(hb, j);
        else if (j.type == ODEB_JOINT_SLIDER) { j.axis1[0] = 1; host_set_slider_axis(hb, j, (Real)d.axis1[0], (Real)d.axis1[1], (Real)d.axis1[2]); }
        else if (j.type == ODEB_JOINT_HINGE) {
            j.axis1[0] = 1; j.axis2[0] = 1;
            host_set_axes(hb, j, (Real)d.axis1[0], (Real)d.axis1[1], (Real)d.axis1[2], j.axis1, j.axis2);
            host_hinge_initial_rotation(hb, j);
        } else if (j.type == ODEB_JOINT_UNIVERSAL) {
            j.axis1[0] = 1; j.axis2[1] = 1;
            if (j.reverse) host_set_axes(hb, j, (Real)d.axis1[0], (Real)d.axis1[1], (Real)d.axis1[2], 0, j.axis2);
            else host_set_axes(hb, j, (Real)d.axis1[0], (Real)d.axis1[1], (Real)d.axis1[2], j.axis1, 0);
            if (j.reverse) host_set_axes(hb, j, (Real)d.axis2[0], (Real)d.axis2[1], (Real)d.axis2[2], j.axis1, 0);
            else host_set_axes(hb, j, (Real)d.axis2[0], (Real)d.axis2[1], (Real)d.axis2[2], 0, j.axis2);
            host_universal_initial_rotations(hb, j);
        }
    }
    T.sofs.assign(nbody + 1, 0);
    for (int b = 0; b < nbody; b++) {
        T.sofs[b] = (int)T.sj.size();
        for (size_t k = 0; k < adj[b].size(); k++) { T.sj.push_back(adj[b][k].first); T.so.push_back(adj[b][k].second); }
    }
    T.sofs[nbody] = (int)T.sj.size();
    return batch_build(wp, T, nworlds, device, 0);
}

static int stage_up(OdebBatch *B, const Real *src, int k, Real4 *dst)
{
    const size_t n = (size_t)B->P.W * B->P.NB;
    for (size_t i = 0; i < n; i++) {
        Real4 v = { src[k * i], src[k * i + 1], src[k * i + 2], k == 4 ? src[k * i + 3] : R_(0.0) };
        B->h_stage[i] = v;
    }
    CK(cudaMemcpyAsync(dst, B->h_stage, n * sizeof(Real4), cudaMemcpyHostToDevice, B->stream));
    CK(cudaStreamSynchronize(B->stream));
    return 1;
}
int odeb_set_state(OdebBatch *B, const odeb_real *pos, const odeb_real *quat, const odeb_real *lvel, const odeb_real *avel)
{
    CK(cudaSetDevice(B->device));
    const size_t n = (size_t)B->P.W * B->P.NB;
    if (pos && !stage_up(B, pos, 3, B->D.pos)) return 0;
    if (quat) {
        if (!stage_up(B, quat, 4, B->D.quat)) return 0;
        k_set_quat<<<nblk(n, 128), 128, 0, B->stream>>>((int)n, B->D.quat, B->D.R);
        B->launches++;
        CK(cudaStreamSynchronize(B->stream));
    }
    if (lvel && !stage_up(B, lvel, 3, B->D.lvel)) return 0;
    if (avel && !stage_up(B, avel, 3, B->D.avel)) return 0;
    return 1;
}

// Page-locked host memory for the caller's state / force arrays (cudaHostAlloc): transfers then go straight between the device and
// the caller's arrays, without the pinned staging buffer and the host-side memcpy (3.4 MB per odeb_get_state at 4096 x 16 bodies).
void *odeb_alloc_host(size_t bytes)
{
    void *p = 0;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) { set_err("cudaHostAlloc(%zu) failed", bytes); cudaGetLastError(); return 0; }
    return p;
}
void odeb_free_host(void *p) { if (p) cudaFreeHost(p); }
static bool host_ptr_is_pinned(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

int odeb_get_state(OdebBatch *B, odeb_real *pos, odeb_real *quat, odeb_real *lvel, odeb_real *avel)
{
    // one pack kernel (Real4 SoA -> the caller's tight [body][3|4] layouts), then per array: device -> caller directly when the caller's
    // array is page-locked (odeb_alloc_host / cudaHostRegister), else device -> pinned staging -> memcpy
    CK(cudaSetDevice(B->device));
    const size_t n = (size_t)B->P.W * B->P.NB;
    Real *d = (Real *)B->d_stage, *h = (Real *)B->h_stage;
    k_pack_state<<<nblk(n, 256), 256, 0, B->stream>>>(n, B->D.pos, B->D.quat, B->D.lvel, B->D.avel, d);
    B->launches++;
    odeb_real *dst[4] = { pos, quat, lvel, avel };
    const size_t ofs[4] = { 0, 3 * n, 7 * n, 10 * n }, len[4] = { 3 * n, 4 * n, 3 * n, 3 * n };
    bool staged[4] = { false, false, false, false };
    for (int k = 0; k < 4; k++) {
        if (!dst[k]) continue;
        staged[k] = !host_ptr_is_pinned(dst[k]);
        CK(cudaMemcpyAsync(staged[k] ? h + ofs[k] : dst[k], d + ofs[k], len[k] * sizeof(Real), cudaMemcpyDeviceToHost, B->stream));
    }
    // this is the blocking call of the asynchronous loop (odeb_add_force, odeb_step_async, odeb_get_state): the capacity-overflow flag and
    // the solver hints ride along, so a truncated step is reported here and not only by odeb_sync
    CK(cudaMemcpyAsync(B->h_ov, B->D.overflow, 4 * sizeof(int), cudaMemcpyDeviceToHost, B->stream));
    CK(cudaStreamSynchronize(B->stream));
    for (int k = 0; k < 4; k++) if (dst[k] && staged[k]) memcpy(dst[k], h + ofs[k], len[k] * sizeof(Real));
    return overflow_check(B);
}

/* The packed body state of the whole batch on the DEVICE: pos [W*NB*3], quat [W*NB*4], lvel [W*NB*3], avel [W*NB*3] (odeb_real, in
 * that order) in a buffer owned by the batch, written by one kernel on the batch's stream; valid until the next call.  For consumers
 * that live on the GPU (an NCCL gather of observations, a policy network): no host round trip.  *bytes = size of the buffer. */
int odeb_pack_state_device(OdebBatch *B, void **dev_ptr, size_t *bytes)
{
    CK(cudaSetDevice(B->device));
    const size_t n = (size_t)B->P.W * B->P.NB;
    k_pack_state<<<nblk(n, 256), 256, 0, B->stream>>>(n, B->D.pos, B->D.quat, B->D.lvel, B->D.avel, (Real *)B->d_stage);
    B->launches++;
    if (dev_ptr) *dev_ptr = B->d_stage;
    if (bytes) *bytes = 13 * n * sizeof(Real);
    CK(cudaGetLastError());
    return 1;
}
/* per-world counters of the last step on the device: stats [W*4] uint32 (the four dynamic-iteration counters), seeds [W] uint32 */
int odeb_device_counters(OdebBatch *B, void **stats, void **seeds)
{
    if (stats) *stats = B->D.stats;
    if (seeds) *seeds = B->D.seed;
    return 1;
}
/* Run the batch on the caller's CUDA stream (a cudaStream_t, e.g. torch.cuda.Stream().cuda_stream) instead of its own, so that the
 * caller's events / stream waits order its kernels against the steps.  The stream must outlive the batch or be replaced first. */
int odeb_set_stream(OdebBatch *B, void *stream)
{
    CK(cudaSetDevice(B->device));
    CK(cudaStreamSynchronize(B->stream));
    if (B->own_stream && B->stream) cudaStreamDestroy(B->stream);
    B->stream = (cudaStream_t)stream; B->own_stream = false;
    return 1;
}

int odeb_add_force(OdebBatch *B, const odeb_real *force, const odeb_real *torque)
{
    // tight [body][3] arrays: straight from the caller's array when it is page-locked, else through the pinned staging buffer; the add
    // runs stream-ordered before the next step
    CK(cudaSetDevice(B->device));
    const size_t n = (size_t)B->P.W * B->P.NB;
    if (!force && !torque) return 1;
    Real *d = (Real *)B->d_stage, *h = (Real *)B->h_stage;
    const odeb_real *src[2] = { force, torque };
    bool synced = false;
    for (int k = 0; k < 2; k++) {
        if (!src[k]) continue;
        const Real *from = src[k];
        if (!host_ptr_is_pinned(src[k])) {
            if (!synced) { CK(cudaStreamSynchronize(B->stream)); synced = true; }   // the staging buffer may still feed an earlier transfer
            memcpy(h + 3 * n * k, src[k], 3 * n * sizeof(Real));
            from = h + 3 * n * k;
        }
        CK(cudaMemcpyAsync(d + 3 * n * k, from, 3 * n * sizeof(Real), cudaMemcpyHostToDevice, B->stream));
    }
    k_add_ft<<<nblk(n, 256), 256, 0, B->stream>>>(n, force ? B->D.facc : 0, torque ? B->D.tacc : 0, d);
    B->launches++;
    // pageable arrays went through the staging buffer: wait, so that the caller may reuse its arrays (and we ours) right away.  Page-locked
    // arrays are read by the DMA engine in stream order: no host wait (include/ode_b200.h: leave them alone until the next blocking call)
    if (synced) CK(cudaStreamSynchronize(B->stream));
    return 1;
}

int odeb_set_seeds(OdebBatch *B, const uint32_t *seeds)
{
    CK(cudaSetDevice(B->device));
    // on the batch's own stream, behind whatever odeb_step_async queued (a plain cudaMemcpy runs on the legacy stream, which does
    // not order against this non-blocking stream: an in-flight solver could overwrite the new seeds or consume them mid-step)
    CK(cudaMemcpyAsync(B->D.seed, seeds, B->P.W * sizeof(uint32_t), cudaMemcpyHostToDevice, B->stream));
    CK(cudaStreamSynchronize(B->stream));
    return 1;
}
int odeb_get_seeds(OdebBatch *B, uint32_t *seeds)
{
    CK(cudaSetDevice(B->device));
    CK(cudaStreamSynchronize(B->stream));
    CK(cudaMemcpy(seeds, B->D.seed, B->P.W * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return 1;
}
int odeb_get_enabled(OdebBatch *B, int *enabled)
{
    CK(cudaSetDevice(B->device));
    CK(cudaStreamSynchronize(B->stream));
    const size_t n = (size_t)B->P.W * B->P.NB;
    CK(cudaMemcpy(enabled, B->D.bflags, n * sizeof(int), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < n; i++) enabled[i] = !(enabled[i] & BF_DISABLED);
    return 1;
}

// Host-side NVTX ranges carrying the reference's own stage names (quickstep.cpp:569-575, :648-658, :784-800; util.cpp dxProcessIslands;
// collision_space.cpp dSpaceCollide): a timeline of this library reads like a TIMING build of the reference.
struct OdebRange {
    bool open;
    explicit OdebRange(const char *name) : open(true) { nvtxRangePushA(name); }
    void end() { if (open) { nvtxRangePop(); open = false; } }
    ~OdebRange() { end(); }
};

static void launch_collide(OdebBatch *B, cudaStream_t s, bool narrow)
{
    OdebRange nv_("dSpaceCollide + nearCallback (dCollide, dJointCreateContact)");
    const DevParams &P = B->P; const DevPtrs &D = B->D;
    const size_t W = P.W;
    if (P.NG <= 0) return;
    k_aabb<<<nblk(W * P.NG, 128), 128, 0, s>>>(P, D);
    k_pair_pass<false><<<nblk(W * P.NG, 128), 128, 0, s>>>(P, D);
    k_pair_scan<<<nblk(W, 64), 64, 0, s>>>(P, D);
    k_pair_pass<true><<<nblk(W * P.NG, 128), 128, 0, s>>>(P, D);
    B->launches += 4;
    if (narrow) { k_narrow<<<nblk(W * P.MP, 64), 64, 0, s>>>(P, D, 1); B->launches++; }
}

// Row budget of k_solve5<2^k> for islands of up to `need` rows: 0 when it cannot be launched (bodies / rows beyond the packed
// entry, or one warp's shared memory beyond an SM).
static int solve5_budget(OdebBatch *B, int k, int need)
{
    B->s5_sr[k] = 0; B->s5_smem[k] = 0;
    if (B->P.NB > ODEB5_MAXBODIES) return 0;
    int sr = (need + 31) / 32 * 32;
    if (sr > B->P.MR) sr = B->P.MR;
    if (sr < need || sr > ODEB5_MAXROWS) return 0;
    size_t smem = odeb5_smem(1 << k, B->P.NB, sr);
    if (smem > 226 * 1024) return 0;
    // Wave quantisation: every block holds 16 / P worlds and its shared memory decides how many blocks an SM holds.  4096 worlds x P = 4 are
    // 1024 blocks = 6.9 per SM; with the generous row margin above a block took 32.5 KB, 6 fit, and the kernel ran two waves (1.77 ms
    // instead of ~1.2).  When trimming the margin (never below the largest island seen, hint_m) brings the whole batch into one wave, do so;
    // an island that outgrows the budget meanwhile takes the serial fallback for one step and raises the hint.
    if (B->hint_m > 0) {
        int nsm = 148;
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, B->device);
        const long long blocks = ((long long)B->P.W + (16 >> k) - 1) / (16 >> k);
        const long long per_sm = (blocks + nsm - 1) / nsm;
        const size_t sm_bytes = 228 * 1024, reserve = 1024;
        if (per_sm >= 1 && (smem + reserve) * (size_t)per_sm > sm_bytes) {
            const int floor_sr = ((B->hint_m + 4) + 7) / 8 * 8;
            for (int t = sr - 8; t >= floor_sr && t >= 8; t -= 8) {
                const size_t b = odeb5_smem(1 << k, B->P.NB, t);
                if ((b + reserve) * (size_t)per_sm <= sm_bytes) { sr = t; smem = b; break; }
            }
        }
    }
    B->s5_sr[k] = sr; B->s5_smem[k] = smem;
    return sr;
}

// Budget of k_solve6<2^k> (k = 0..2): rows and bodies per island with a margin over the largest island seen; 0 when it cannot be launched.
static int solve6_budget(OdebBatch *B, int k)
{
    B->s6_sr = 0; B->s6_nbi = 0; B->s6_smem = 0;
    const int hm = B->hint_m > 0 ? B->hint_m : (B->P.MR < 256 ? B->P.MR : 256);
    const int hb = B->hint_nb > 0 ? B->hint_nb : (B->P.NB < 32 ? B->P.NB : 32);
    // Margins over the largest island seen (an island that outgrows the budget before the hints are refreshed takes the serial path for a
    // step): generous by default, trimmed when that lets one more warp live on an SM (shared memory decides how many worlds are resident,
    // and the kernel is latency-bound: 64-body piles 4 -> 5 warps per SM).
    int best_sr = 0, best_nbi = 0; size_t best_smem = 0; long long best_per_sm = -1;
    for (int t = 0; t < 3; t++) {
        const int mr = t == 0 ? hm / 8 + 8 : t == 1 ? hm / 16 + 8 : hm / 32 + 4;
        const int mb = t == 0 ? hb / 4 + 2 : t == 1 ? hb / 8 + 2 : hb / 16 + 1;
        int sr = (hm + mr + 7) / 8 * 8, nbi = hb + mb;
        if (sr > B->P.MR) sr = (B->P.MR + 7) / 8 * 8;
        if (nbi > B->P.NB) nbi = B->P.NB;
        if (sr > ODEB5_MAXROWS) sr = ODEB5_MAXROWS / 8 * 8;
        if (nbi > ODEB5_MAXBODIES - 1) nbi = ODEB5_MAXBODIES - 1;
        if (sr < hm || nbi < hb) continue;
        const size_t smem = odeb6_smem(1 << k, nbi, sr);
        if (smem > 226 * 1024) continue;
        long long per_sm = (long long)((228 * 1024) / (smem + 1024));
        if (per_sm > 10) per_sm = 10;                                   // registers: ~200 per thread
        if (per_sm > best_per_sm) { best_per_sm = per_sm; best_sr = sr; best_nbi = nbi; best_smem = smem; }
    }
    if (best_per_sm < 0) return 0;
    B->s6_sr = best_sr; B->s6_nbi = best_nbi; B->s6_smem = best_smem;
    return best_sr;
}

// Which solver kernel the next step uses. The largest island of the previous call (hint_m) decides, so the first call after
// creation runs k_solve.
static bool hybrid_ok(OdebBatch *B, int need)
{   // both halves must hold the islands (larger ones are completed by k_solve's own fallback), and the hand-over point has to lie
    // before the first reorder and before the end of the regular iterations
    // ... and the row stream should sit in L2 (126 MB): when it streams from HBM instead (65536 x 10-link chains: 350 MB) the one-row-at-a-time
    // kernel alone is the faster one (4.25 against 4.79 ms per step)
    const size_t stream_bytes = (size_t)B->P.W * (size_t)(B->hint_m > 0 ? B->hint_m : B->P.MR) * 32 * sizeof(Real);
    return B->P.SR > 0 && (int)B->P.num_iter > ODEB_HYBRID_SWEEPS && stream_bytes <= ((size_t)128 << 20) && solve5_budget(B, 2, need) > 0;
}
static int choose_solver(OdebBatch *B)
{
    const int f = B->solver_force;
    const int need = B->hint_m > 0 ? B->hint_m + B->hint_m / 8 + 8 : (B->P.MR < 512 ? B->P.MR : 512);
    if (f == 0) return 0;
    if (f == 5) return (B->P.SR > 0 && (int)B->P.num_iter > ODEB_HYBRID_SWEEPS && solve5_budget(B, 2, need) > 0) ? 5 : 0;
    if (f == 4) return B->bl_G ? 4 : 0;
    if (f >= 1 && f <= 3) return solve5_budget(B, f, need) > 0 ? f : 0;
    if (f >= 6 && f <= 9) return solve6_budget(B, f - 6) > 0 ? f : 0;
    if (B->hint_m <= 0) return 0;
    if (B->hint_nis >= 2 && !B->env_no_solve6) {
        // worlds with several islands: independent walkers (odeb_solve6.cuh).  Measured on B200, 64-body piles, solver ms per step for
        // P = 1 / 2 / 4 / 8 processors per world: 4096 worlds with one large island each 12.0 / 11.2 / 8.1 / 8.6 (k_solve5<4>: 9.3),
        // 4096 settled piles (35 small islands per world) 6.6 / 6.9 / 6.4 / - (k_solve5<4>: 8.4), 16384 worlds - / 38.7 / 27.4 / 30.8.
        if (solve6_budget(B, 2) > 0) return 8;
    }
    // Measured on B200 (4096 x 16-box stacks, single; solver ms for W = 1024 / 2048 / 4096 worlds): k_solve 1.19 / 1.19 / 1.26,
    // k_solve5<4> 0.67 / 0.83 / 1.25, k_solve5<8> 0.78 / - / -, k_solve5<2> - / - / 1.39.  With more than ~3.5 warps per SM the
    // extra warps of the P-processor schedule contend for the SM's shared-memory pipe and the gain is gone, so P = 4 is used
    // while the batch needs at most that many warps. It is also used whenever an island does not fit k_solve's own row budget
    // (64-body piles: 45 ms per step through the serial fallback against 3 ms), in several waves if need be.
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, B->device);
    const long long warps4 = ((long long)B->P.W + 3) / 4;
    if (warps4 * 2 <= 7LL * nsm || need > B->P.SR) { if (solve5_budget(B, 2, need) > 0) return 2; }
    if (!B->env_no_hybrid && hybrid_ok(B, need)) return 5;     // large batches: see launch_dynamics, case 5
    return 0;
}

static void launch_dynamics(OdebBatch *B, cudaStream_t s, bool timed, int cfg)
{
    const DevParams &P = B->P; const DevPtrs &D = B->D;
    const size_t W = P.W;
    if (D.jfb) cudaMemsetAsync(D.jfb, 0, W * P.NJT * 4 * sizeof(Real4), s);      // state 0 = joint not stepped
    nvtxRangePushA("dxProcessIslands (auto-disable, island build) + dxQuickStepIsland_Stage0_Joints (getInfo1) + Stage1");
    if (P.NJ > 0) { k_joint_info1<<<nblk(W * P.NJ, 128), 128, 0, s>>>(P, D); B->launches++; }
    if (B->isl_one) k_islands_t<true, true><<<nblk(W * 32, 128), 128, B->isl_one_smem, s>>>(P, D);
    else if (B->isl_smem) k_islands_t<true, false><<<nblk(W, 32), 32, B->isl_smem, s>>>(P, D);
    else k_islands_t<false, false><<<nblk(W, 32), 32, 0, s>>>(P, D);
    nvtxRangePop();
    nvtxRangePushA("dxQuickStepIsland_Stage0_Bodies");
    k_body_pre<<<nblk(W * P.NB, 128), 128, 0, s>>>(P, D);
    nvtxRangePop();
    nvtxRangePushA("dxQuickStepIsland_Stage2a/2b/2c (getInfo2, rhs) + Stage4LCP_iMJ + Stage4LCP_AdComputation");
    if (P.NJ == 0 && !D.row_island && !B->env_no_fused_rows) {
        if (B->rows_std3) k_rows_t<true, true><<<nblk(W * P.NJT, 64), 64, 0, s>>>(P, D); else k_rows_t<true><<<nblk(W * P.NJT, 64), 64, 0, s>>>(P, D);
        B->launches--;
    }
    else {
        k_rows_t<false><<<nblk(W * P.NJT, 64), 64, 0, s>>>(P, D);
        k_rows_finish<<<nblk(W * P.MR, 128), 128, 0, s>>>(P, D);
    }
    nvtxRangePop();
    nvtxRangePushA("dxQuickStepIsland_Stage4LCP_ReorderPrep + Stage4LCP_Iteration (SOR sweeps)");
    cudaEvent_t e0 = 0, e1 = 0;
    if (timed) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, s); }
    switch (cfg) {
    case 1: k_solve5_t<2, 0><<<nblk(W, 8), 32, B->s5_smem[1], s>>>(P, D, B->s5_sr[1]); break;
    case 2: k_solve5_t<4, 0><<<nblk(W, 4), 32, B->s5_smem[2], s>>>(P, D, B->s5_sr[2]); break;
    case 3: k_solve5_t<8, 0><<<nblk(W, 2), 32, B->s5_smem[3], s>>>(P, D, B->s5_sr[3]); break;
    case 5:     // hybrid: the 8 sweeps in the initial order (long dependency chains: 136 schedule slots for 192 rows of a 16-box stack) one
                // row at a time per world, the sweeps after the first dRand reorder (~52 slots) under the 4-processor schedule
        {   // ODEB_TEST_HY_ROWS (tests only): a smaller hand-over threshold, so that worlds mix islands that may and may not be paused
            int srb = B->s5_sr[2];
            if (B->env_hy_rows > 0 && B->env_hy_rows < srb) srb = B->env_hy_rows;
            k_solve_hy<<<nblk(W, ODEB_WPW), 32, B->solve_smem, s>>>(P, D, srb);
        }
        k_solve5_t<4, ODEB_HYBRID_SWEEPS><<<nblk(W, 4), 32, B->s5_smem[2], s>>>(P, D, B->s5_sr[2]);
        B->launches++;
        break;
    case 6: case 7: case 8: case 9:     // independent world walkers (odeb_solve6.cuh); the initial order of every island comes from k_reorder_prep
        k_reorder_prep<<<nblk(W * 32, 128), 128, 0, s>>>(P, D);
        k_world_sort<<<1, 1024, 0, s>>>(P, D);
        B->launches++;
        if (cfg == 6) k_solve6_t<1><<<nblk(W, 16), 32, B->s6_smem, s>>>(P, D, B->s6_sr, B->s6_nbi);
        else if (cfg == 7) k_solve6_t<2><<<nblk(W, 8), 32, B->s6_smem, s>>>(P, D, B->s6_sr, B->s6_nbi);
        else if (cfg == 8) k_solve6_t<4><<<nblk(W, 4), 32, B->s6_smem, s>>>(P, D, B->s6_sr, B->s6_nbi);
        else k_solve6_t<8><<<nblk(W, 2), 32, B->s6_smem, s>>>(P, D, B->s6_sr, B->s6_nbi);
        B->launches++;
        break;
    case 4:
        if (B->bl_G == 8) k_solve_bl<8><<<nblk(W, 4), 32, B->bl_smem, s>>>(P, D, B->bl_SR);
        else if (B->bl_G == 16) k_solve_bl<16><<<nblk(W, 2), 32, B->bl_smem, s>>>(P, D, B->bl_SR);
        else k_solve_bl<32><<<nblk(W, 1), 32, B->bl_smem, s>>>(P, D, B->bl_SR);
        break;
    default: k_solve<<<nblk(W, ODEB_WPW), 32, B->solve_smem, s>>>(P, D, 0);
    }
    if (timed) { cudaEventRecord(e1, s); B->pending.push_back(std::make_pair(e0, e1)); }
    nvtxRangePop();
    OdebRange nv_("dxQuickStepIsland_Stage4b (feedback) + Stage6a/6b (velocity update, dxStepBody)");
    if (D.jcopy) { k_feedback<<<nblk(W * P.NJT, 128), 128, 0, s>>>(P, D); B->launches++; }
    k_integrate<<<nblk(W * P.NB, 128), 128, 0, s>>>(P, D);
    B->launches += 6;
}

static int launch_step(OdebBatch *B, cudaStream_t s, bool timed, int cfg)
{
    launch_collide(B, s, true);
    launch_dynamics(B, s, timed, cfg);
    return 1;
}

int odeb_step_async(OdebBatch *B, double h, int nsteps)
{
    CK(cudaSetDevice(B->device));
    if (!(h > 0)) { set_err("stepsize must be > 0"); return 0; }
    B->P.h = (Real)h; B->P.hrecip = rrecip((Real)h);
    if (B->mode == ODEB_MODE_CANONICAL) {
        for (int s = 0; s < nsteps; s++) if (!large_step(B)) return 0;
        return 1;
    }
    // The solver kernel (and k_solve5's row budget) follow the largest island seen: a fresh batch reads it back once after its
    // first step, long calls re-read it every 64 steps, so that a many-step call does not stay on a kernel its islands outgrew.
    int done = 0;
    while (done < nsteps) {
        const bool graph_ok = B->use_graph && !B->timing;
        const int cfg = choose_solver(B);
        const int cfg_sr = (cfg >= 1 && cfg <= 3) ? B->s5_sr[cfg] : (cfg == 5 ? B->s5_sr[2] : (cfg >= 6 ? B->s6_sr * 256 + B->s6_nbi : 0));
        if (graph_ok && (B->graph == 0 || B->graph_h != h || B->graph_cfg != cfg || B->graph_sr != cfg_sr)) {
            if (B->graph) { cudaGraphExecDestroy(B->graph); B->graph = 0; }
            cudaGraph_t g = 0;
            uint64_t l0 = B->launches;
            CK(cudaStreamBeginCapture(B->stream, cudaStreamCaptureModeThreadLocal));
            launch_step(B, B->stream, false, cfg);
            CK(cudaStreamEndCapture(B->stream, &g));
            B->graph_nlaunch = (int)(B->launches - l0);
            B->launches = l0;
            CK(cudaGraphInstantiate(&B->graph, g, 0));
            cudaGraphDestroy(g);
            B->graph_h = h; B->graph_cfg = cfg; B->graph_sr = cfg_sr;
        }
        int chunk = nsteps - done;
        const bool probe = B->solver_force < 0 && chunk > 1;          // nothing to adapt when a kernel is forced
        if (probe) { const int lim = B->hint_m <= 0 ? 1 : 64; if (chunk > lim) chunk = lim; }
        for (int s = 0; s < chunk; s++) {
            if (graph_ok) {
                CK(cudaGraphLaunch(B->graph, B->stream));
                B->launches += B->graph_nlaunch;
            } else launch_step(B, B->stream, B->timing, cfg);
        }
        done += chunk;
        if (probe && done < nsteps) {
            int m[3] = { 0, 0, 0 };
            CK(cudaMemcpyAsync(m, B->D.overflow + 1, sizeof(m), cudaMemcpyDeviceToHost, B->stream));
            CK(cudaStreamSynchronize(B->stream));
            if (m[0] > B->hint_m) B->hint_m = m[0];
            if (m[1] > B->hint_nb) B->hint_nb = m[1];
            if (m[2] > B->hint_nis) B->hint_nis = m[2];
        }
    }
    CK(cudaGetLastError());
    return 1;
}

int odeb_sync(OdebBatch *B)
{
    CK(cudaSetDevice(B->device));
    CK(cudaStreamSynchronize(B->stream));
    for (size_t i = 0; i < B->pending.size(); i++) {
        float ms = 0;
        cudaEventElapsedTime(&ms, B->pending[i].first, B->pending[i].second);
        B->solver_ms += ms; B->solver_launches++;
        cudaEventDestroy(B->pending[i].first); cudaEventDestroy(B->pending[i].second);
    }
    B->pending.clear();
    CK(cudaMemcpyAsync(B->h_ov, B->D.overflow, 4 * sizeof(int), cudaMemcpyDeviceToHost, B->stream));
    CK(cudaStreamSynchronize(B->stream));
    return overflow_check(B);
}

int odeb_step(OdebBatch *B, double h, int nsteps)
{
    if (!odeb_step_async(B, h, nsteps)) return 0;
    return odeb_sync(B);
}

/* K steps, each bracketed by its own CUDA event pair on the batch's stream; optionally the L2 is flushed
 * (memset of flush_bytes, outside the event pairs) before every step. Returns the summed device time. */
int odeb_timed_steps(OdebBatch *B, double h, int nsteps, size_t flush_bytes, double *total_ms)
{
    CK(cudaSetDevice(B->device));
    if (flush_bytes > B->flush_bytes) {
        if (B->flush_buf) cudaFree(B->flush_buf);
        B->flush_buf = 0; B->flush_bytes = 0;
        CK(cudaMalloc(&B->flush_buf, flush_bytes));
        B->flush_bytes = flush_bytes;
    }
    std::vector<cudaEvent_t> ev(2 * (size_t)nsteps);
    for (size_t i = 0; i < ev.size(); i++) CK(cudaEventCreate(&ev[i]));
    for (int s = 0; s < nsteps; s++) {
        if (flush_bytes) CK(cudaMemsetAsync(B->flush_buf, s & 0xff, flush_bytes, B->stream));
        CK(cudaEventRecord(ev[2 * s], B->stream));
        if (!odeb_step_async(B, h, 1)) return 0;
        CK(cudaEventRecord(ev[2 * s + 1], B->stream));
    }
    if (!odeb_sync(B)) return 0;
    double tot = 0;
    for (int s = 0; s < nsteps; s++) { float ms = 0; CK(cudaEventElapsedTime(&ms, ev[2 * s], ev[2 * s + 1])); tot += ms; }
    for (size_t i = 0; i < ev.size(); i++) cudaEventDestroy(ev[i]);
    *total_ms = tot;
    return 1;
}

int odeb_enable_feedback(OdebBatch *B, int on)
{
    CK(cudaSetDevice(B->device));
    CK(cudaStreamSynchronize(B->stream));
    if (on && !B->D.jcopy) {
        const size_t W = B->P.W;
        // allocated once per batch and kept: the classic dWorldQuickStep toggles feedback whenever "some joint has a dJointFeedback" changes
        if (!B->fb_jcopy && (!dev_alloc(B, &B->fb_jcopy, W * B->P.MR * 3) || !dev_alloc(B, &B->fb_jfb, W * B->P.NJT * 4))) return 0;
        CK(cudaDeviceSynchronize());                       // dev_alloc clears on the legacy stream
        B->D.jcopy = B->fb_jcopy; B->D.jfb = B->fb_jfb;
    } else if (!on) { B->D.jcopy = 0; B->D.jfb = 0; }
    if (B->graph) { cudaGraphExecDestroy(B->graph); B->graph = 0; }
    return 1;
}
int odeb_get_feedback(OdebBatch *B, int world, odeb_real *out12, int *state, int cap)
{
    CK(cudaSetDevice(B->device));
    CK(cudaStreamSynchronize(B->stream));
    if (!B->D.jfb) { set_err("joint feedback is not enabled"); return -1; }
    if (world < 0 || world >= B->P.W) { set_err("bad world index"); return -1; }
    int nc = 0;
    CK(cudaMemcpy(&nc, B->D.ncontacts + world, sizeof(int), cudaMemcpyDeviceToHost));
    if (nc > B->P.MC) nc = B->P.MC;
    const int n = B->P.NJ + nc;
    std::vector<Real4> v(4 * (size_t)n);
    if (n) CK(cudaMemcpy(v.data(), B->D.jfb + (size_t)world * B->P.NJT * 4, v.size() * sizeof(Real4), cudaMemcpyDeviceToHost));
    for (int i = 0; i < n && i < cap; i++) {
        const Real4 *q = &v[4 * (size_t)i];
        odeb_real *o = out12 + 12 * (size_t)i;
        o[0] = q[0].x; o[1] = q[0].y; o[2] = q[0].z; o[3] = q[1].x; o[4] = q[1].y; o[5] = q[1].z;
        o[6] = q[2].x; o[7] = q[2].y; o[8] = q[2].z; o[9] = q[3].x; o[10] = q[3].y; o[11] = q[3].z;
        state[i] = (int)q[0].w;
    }
    return n;
}

// ---- snapshot / restore of everything a step reads from earlier steps (checkpoint / resume; the reference only has the
//      text dump of dWorldExportDIF, ode/src/export-dif.cpp). Layout: header {magic, sizeof(real), W, NB, samples}, then the arrays.
struct SnapItem { void *p; size_t bytes; };
static std::vector<SnapItem> snapshot_items(OdebBatch *B)
{
    const DevPtrs &D = B->D; const size_t WB = (size_t)B->P.W * B->P.NB, W = B->P.W;
    const size_t NS = B->P.adis_samples > 0 ? B->P.adis_samples : 1;
    std::vector<SnapItem> v;
    SnapItem it[] = { { D.pos, WB * sizeof(Real4) }, { D.quat, WB * sizeof(Real4) }, { D.lvel, WB * sizeof(Real4) }, { D.avel, WB * sizeof(Real4) },
                      { D.facc, WB * sizeof(Real4) }, { D.tacc, WB * sizeof(Real4) }, { D.R, 3 * WB * sizeof(Real4) },
                      { D.bflags, WB * sizeof(int) }, { D.adis_steps, WB * sizeof(int) }, { D.adis_time, WB * sizeof(Real) },
                      { D.avg_buf, WB * 6 * NS * sizeof(Real) }, { D.avg_counter, WB * sizeof(int) }, { D.avg_ready, WB * sizeof(int) },
                      { D.seed, W * sizeof(unsigned) }, { D.stats, 4 * W * sizeof(unsigned) } };
    for (size_t i = 0; i < sizeof(it) / sizeof(it[0]); i++) v.push_back(it[i]);
    return v;
}
size_t odeb_snapshot_size(OdebBatch *B)
{
    size_t n = 5 * sizeof(uint64_t);
    std::vector<SnapItem> v = snapshot_items(B);
    for (size_t i = 0; i < v.size(); i++) n += v[i].bytes;
    return n;
}
int odeb_snapshot(OdebBatch *B, void *buf, size_t cap)
{
    CK(cudaSetDevice(B->device));
    CK(cudaStreamSynchronize(B->stream));
    if (cap < odeb_snapshot_size(B)) { set_err("snapshot buffer too small"); return 0; }
    uint64_t hdr[5] = { 0x4f44454232303053ull, sizeof(Real), (uint64_t)B->P.W, (uint64_t)B->P.NB, (uint64_t)(B->P.adis_samples > 0 ? B->P.adis_samples : 1) };
    char *o = (char *)buf; memcpy(o, hdr, sizeof(hdr)); o += sizeof(hdr);
    std::vector<SnapItem> v = snapshot_items(B);
    for (size_t i = 0; i < v.size(); i++) { CK(cudaMemcpy(o, v[i].p, v[i].bytes, cudaMemcpyDeviceToHost)); o += v[i].bytes; }
    return 1;
}
int odeb_restore(OdebBatch *B, const void *buf, size_t bytes)
{
    CK(cudaSetDevice(B->device));
    CK(cudaStreamSynchronize(B->stream));
    uint64_t hdr[5];
    if (bytes < sizeof(hdr) || bytes != odeb_snapshot_size(B)) { set_err("snapshot does not match this batch (size)"); return 0; }
    memcpy(hdr, buf, sizeof(hdr));
    if (hdr[0] != 0x4f44454232303053ull || hdr[1] != sizeof(Real) || hdr[2] != (uint64_t)B->P.W || hdr[3] != (uint64_t)B->P.NB
        || hdr[4] != (uint64_t)(B->P.adis_samples > 0 ? B->P.adis_samples : 1)) { set_err("snapshot does not match this batch (header)"); return 0; }
    const char *o = (const char *)buf + sizeof(hdr);
    std::vector<SnapItem> v = snapshot_items(B);
    for (size_t i = 0; i < v.size(); i++) { CK(cudaMemcpy(v[i].p, o, v[i].bytes, cudaMemcpyHostToDevice)); o += v[i].bytes; }
    return 1;
}

uint64_t odeb_launch_count(const OdebBatch *B) { return B->launches; }
const char *odeb_solver_kernel(OdebBatch *B)
{
    static const char *names[10] = { "k_solve", "k_solve5<2>", "k_solve5<4>", "k_solve5<8>", "k_solve_bl", "k_solve (8 sweeps) + k_solve5<4>", "k_solve6<1>", "k_solve6<2>", "k_solve6<4>", "k_solve6<8>" };
    if (B->mode == ODEB_MODE_CANONICAL) return "k_lwt_phase";
    return names[choose_solver(B)];
}
void odeb_enable_timing(OdebBatch *B, int on) { B->timing = on != 0; }
double odeb_solver_ms(OdebBatch *B, int *launches)
{
    double v = B->solver_ms;
    if (launches) *launches = B->solver_launches;
    B->solver_ms = 0; B->solver_launches = 0;
    return v;
}

int odeb_get_pairs(OdebBatch *B, int world, int *pairs, int cap)
{
    CK(cudaSetDevice(B->device));
    CK(cudaStreamSynchronize(B->stream));
    int n = 0;
    CK(cudaMemcpy(&n, B->D.npairs + world, sizeof(int), cudaMemcpyDeviceToHost));
    int m = n < cap ? n : cap;
    if (m > 0) CK(cudaMemcpy(pairs, B->D.pairs + (size_t)world * B->P.MP, (size_t)m * sizeof(int2), cudaMemcpyDeviceToHost));
    return n;
}

int odeb_get_contacts(OdebBatch *B, int world, odeb_real *geom7, int *g12, int cap)
{
    CK(cudaSetDevice(B->device));
    CK(cudaStreamSynchronize(B->stream));
    const DevParams &P = B->P;
    int n = 0;
    CK(cudaMemcpy(&n, B->D.ncontacts + world, sizeof(int), cudaMemcpyDeviceToHost));
    if (n == 0) return 0;
    std::vector<int4> ci(n);
    CK(cudaMemcpy(ci.data(), B->D.cinfo + (size_t)world * P.MC, n * sizeof(int4), cudaMemcpyDeviceToHost));
    int np = 0;
    CK(cudaMemcpy(&np, B->D.npairs + world, sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<int2> pr(np);
    CK(cudaMemcpy(pr.data(), B->D.pairs + (size_t)world * P.MP, np * sizeof(int2), cudaMemcpyDeviceToHost));
    std::vector<Real4> cg((size_t)np * P.maxc * 2);
    CK(cudaMemcpy(cg.data(), B->D.cgeom + (size_t)world * P.MP * P.maxc * 2, cg.size() * sizeof(Real4), cudaMemcpyDeviceToHost));
    for (int i = 0; i < n && i < cap; i++) {
        int slot = ci[i].x;
        Real4 a = cg[2 * slot], b = cg[2 * slot + 1];
        geom7[7 * i] = a.x; geom7[7 * i + 1] = a.y; geom7[7 * i + 2] = a.z;
        geom7[7 * i + 3] = b.x; geom7[7 * i + 4] = b.y; geom7[7 * i + 5] = b.z; geom7[7 * i + 6] = a.w;
        int2 p = pr[slot / P.maxc];
        g12[2 * i] = p.x; g12[2 * i + 1] = p.y;
    }
    return n;
}

int odeb_get_ray_hits(OdebBatch *B, int world, odeb_real *geom7, int *g12, int cap)
{
    CK(cudaSetDevice(B->device));
    CK(cudaStreamSynchronize(B->stream));
    const DevParams &P = B->P;
    if (!B->D.ray_count) return 0;
    int np = 0;
    CK(cudaMemcpy(&np, B->D.npairs + world, sizeof(int), cudaMemcpyDeviceToHost));
    if (np == 0) return 0;
    std::vector<int2> pr(np); std::vector<int> rc(np);
    CK(cudaMemcpy(pr.data(), B->D.pairs + (size_t)world * P.MP, np * sizeof(int2), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(rc.data(), B->D.ray_count + (size_t)world * P.MP, np * sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<Real4> cg((size_t)np * P.maxc * 2);
    CK(cudaMemcpy(cg.data(), B->D.cgeom + (size_t)world * P.MP * P.maxc * 2, cg.size() * sizeof(Real4), cudaMemcpyDeviceToHost));
    int n = 0;
    for (int p = 0; p < np; p++) for (int k = 0; k < rc[p]; k++, n++) {
        if (n >= cap) continue;
        Real4 a = cg[2 * ((size_t)p * P.maxc + k)], b = cg[2 * ((size_t)p * P.maxc + k) + 1];
        geom7[7 * n] = a.x; geom7[7 * n + 1] = a.y; geom7[7 * n + 2] = a.z;
        geom7[7 * n + 3] = b.x; geom7[7 * n + 4] = b.y; geom7[7 * n + 5] = b.z; geom7[7 * n + 6] = a.w;
        g12[2 * n] = pr[p].x; g12[2 * n + 1] = pr[p].y;
    }
    return n;
}

int odeb_num_rays(OdebBatch *B) { return B->D.ray_count ? B->D.nray : 0; }
int odeb_get_ray_ranges(OdebBatch *B, odeb_real *range, int *hit_geom)
{
    CK(cudaSetDevice(B->device));
    DevPtrs &D = B->D;
    if (!D.ray_count || D.nray == 0) return 0;
    const size_t n = (size_t)B->P.W * D.nray;
    k_ray_ranges<<<nblk(n, 128), 128, 0, B->stream>>>(B->P, D);
    B->launches++;
    if (range) CK(cudaMemcpyAsync(range, D.ray_range, n * sizeof(Real), cudaMemcpyDeviceToHost, B->stream));
    if (hit_geom) CK(cudaMemcpyAsync(hit_geom, D.ray_hit, n * sizeof(int), cudaMemcpyDeviceToHost, B->stream));
    CK(cudaStreamSynchronize(B->stream));
    return D.nray;
}

int odeb_get_islands(OdebBatch *B, int world, int *label_per_body)
{
    CK(cudaSetDevice(B->device));
    CK(cudaStreamSynchronize(B->stream));
    int n = 0;
    CK(cudaMemcpy(&n, B->D.nislands + world, sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(label_per_body, B->D.body_island + (size_t)world * B->P.NB, B->P.NB * sizeof(int), cudaMemcpyDeviceToHost));
    return n;
}

int odeb_get_stats(OdebBatch *B, int world, OdebStats *out)
{
    CK(cudaSetDevice(B->device));
    CK(cudaStreamSynchronize(B->stream));
    CK(cudaMemcpy(out->v, B->D.stats + 4 * (size_t)world, 4 * sizeof(unsigned), cudaMemcpyDeviceToHost));
    return 1;
}

int odeb_get_totals(OdebBatch *B, uint64_t out[6])
{
    CK(cudaSetDevice(B->device));
    CK(cudaStreamSynchronize(B->stream));
    const int W = B->P.W;
    std::vector<int> a(W), b(W), c(W), d(W); std::vector<unsigned long long> s(2 * (size_t)W);
    CK(cudaMemcpy(a.data(), B->D.npairs, W * sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(b.data(), B->D.ncontacts, W * sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(c.data(), B->D.mrows, W * sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(d.data(), B->D.nislands, W * sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(s.data(), B->D.sweeps, 2 * (size_t)W * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    for (int k = 0; k < 6; k++) out[k] = 0;
    for (int w = 0; w < W; w++) { out[0] += a[w]; out[1] += b[w]; out[2] += c[w]; out[3] += d[w]; out[4] += s[2 * (size_t)w]; out[5] += s[2 * (size_t)w + 1]; }
    return 1;
}

} // extern "C"

#if defined(ODEB6_PROF)
// experiment builds only (tools/build_variant.sh): cycle accounting of k_solve6
extern "C" int odeb_debug_prof(unsigned long long *out16, int reset)
{
    if (out16 && cudaMemcpyFromSymbol(out16, odeb6_prof, 16 * sizeof(unsigned long long)) != cudaSuccess) return 0;
    if (reset) { unsigned long long z[16] = { 0 }; if (cudaMemcpyToSymbol(odeb6_prof, z, sizeof(z)) != cudaSuccess) return 0; }
    return 1;
}
#endif
