// odeb_classic.inl -- the reference's classic per-object C API (include/ode_b200_classic.h) on the B200 kernels.
//
// The host side keeps what the reference keeps in its object graph (ode/src/objects.h:221-291, joints/joint.h:122-245,
// collision_kernel.h:103-271): handles, user data, per-body joint lists in dJointAttach order, a mirror of the body
// state (getters hand out pointers into it, ode.cpp:413-484).  Everything the hot path computes runs in CUDA:
//
//   dSpaceCollide   k_aabb + k_pair_pass/k_pair_scan (pair set of the space type) -> pairs come back -> near-callback
//   dCollide        k_collide_req: narrowphase of every broadphase pair in one launch at the first call of a collide
//                   pass (results cached per pair), or of one (o1,o2) request when called outside that pattern
//   dWorldQuickStep contacts created by dJointCreateContact + joint adjacency go up; k_joint_info1, k_islands, k_body_pre,
//                   k_rows, k_rows_finish, k_solve, k_integrate; body state comes back into the mirror
//
// One device context (an OdebBatch with W = 1, classic mode) per world, rebuilt when the topology (bodies, geoms,
// permanent joints, masses) changes.  No CPU implementation of any stage exists here.

#include <map>
#include <algorithm>

struct dxWorld; struct dxBody; struct dxGeom; struct dxSpace; struct dxJoint; struct dxJointGroup;
#define dReal odeb_real
#include "../../include/ode_b200_classic.h"
#undef dReal

// ------------------------------------------------------------------------------------------------ objects

struct AdjNode { dxJoint *joint; dxBody *other; };

struct dxBody {
    dxWorld *world; void *userdata;
    Real pos[4], R[12], q[4], lvel[4], avel[4], facc[4], tacc[4];
    dMass mass; Real invI[12], invMass;
    int flags;                               // BF_*
    std::vector<AdjNode> joints;             // dJointAttach order (oldest first; the reference walks newest first)
    std::vector<dxGeom *> geoms;
    int index;                               // position in world->bodies (creation order)
    // auto-disable mirror (util.cpp:427-561)
    int adis_steps_left; Real adis_time_left; int avg_counter, avg_ready; std::vector<Real> avg_buf;
};

struct dxJoint {
    dxWorld *world; dxJointGroup *group; void *userdata; dJointFeedback *feedback;
    int type;                                // dJointType
    dxBody *body[2];                         // node[0].body, node[1].body after dJointAttach's swap
    int reverse;                             // dJOINT_REVERSE
    DJointT t;                               // ball / hinge / universal parameters (body-relative anchors, axes, limits)
    dContact contact;                        // contact joints
};

struct dxJointGroup { std::vector<dxJoint *> joints; };

struct dxGeom {
    int type; dxSpace *space; dxBody *body; void *userdata;
    Real p[4];                               // class parameters (radius | sides | radius,length | plane)
    Real pos[4], R[12];                      // own placement when no body is attached
    int has_ofs; Real opos[4], oR[12];       // dGeomSetOffset*: pose relative to the body (collision_kernel.cpp:1000-1130)
    Real fpos[4], fR[12];                    // final pose handed out by dGeomGetPosition / dGeomGetRotation for offset geoms
    unsigned long cat, col;
    int index;                               // position in space->geoms
};

struct dxSpace {
    int type;                                // ODEB_SPACE_HASH | ODEB_SPACE_SAP | ODEB_SPACE_SIMPLE
    int minlevel, maxlevel;                  // dxHashSpace::global_minlevel / global_maxlevel
    int cleanup;
    std::vector<dxGeom *> geoms;
    int lock_count;
};

struct CollideReq { int g1, g2, flags; };

struct ClassicCtx {
    OdebBatch *B;
    dxSpace *space;
    int nb, ng, nj;
    int cap_pairs, cap_contacts;
    std::vector<dxJoint *> perm;             // permanent joints in creation order = device joint ids
    // collide pass state
    std::vector<int2> pairs;
    bool in_pass; int cache_flags;           // flags of the cached narrowphase (-1: none yet)
    std::vector<int> pc_count; std::vector<Real> pc_geom;      // per pair: count, maxc x 7 reals
    int cache_maxc;
    CollideReq *d_req; Real *d_out; int *d_cnt; int req_cap;
};

struct dxWorld {
    OdebWorldParams wp;
    std::vector<dxBody *> bodies;            // creation order (world->firstbody walks it backwards, ode.cpp:54-60,264)
    std::vector<dxJoint *> joints;           // every joint of the world in creation order
    ClassicCtx *ctx; bool topo_dirty;
    dWorldQuickStepIterationCount_DynamicAdjustmentStatistics *stats_sink;
    int body_flags_default;
};

static unsigned long g_seed = 0;             // dRand's process-global seed (misc.cpp:33)
static int g_init_count = 0;
static std::vector<dxSpace *> g_spaces;

static void classic_error(const char *fmt, ...)
{   // dError -> default handler prints and exits (error.cpp:86-94); the message is also kept for odeb_last_error
    char buf[512]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    g_err = buf;
    fprintf(stderr, "\nODE-B200 Error: %s\n", buf);
    fflush(stderr);
}

// ------------------------------------------------------------------------------------------------ device kernels of this layer

// dCollide for a list of (g1, g2, flags) requests: one thread per request, same device colliders as k_narrow
__global__ void k_collide_req(const __grid_constant__ DevParams P, const __grid_constant__ DevPtrs D, int n, const CollideReq *req, Real *out, int *cnt, int maxc)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    CollideReq r = req[t];
    DGeom g1, g2;
    load_geom(P, D, 0, r.g1, g1);
    load_geom(P, D, 0, r.g2, g2);
    DContactGeom c[8];
    int k = odeb_collide(g1, g2, r.flags, c);
    Real *o = out + (size_t)t * maxc * 7;
    for (int i = 0; i < k && i < maxc; i++) {
        o[7 * i] = c[i].pos[0]; o[7 * i + 1] = c[i].pos[1]; o[7 * i + 2] = c[i].pos[2];
        o[7 * i + 3] = c[i].normal[0]; o[7 * i + 4] = c[i].normal[1]; o[7 * i + 5] = c[i].normal[2]; o[7 * i + 6] = c[i].depth;
    }
    cnt[t] = k;
}

// ------------------------------------------------------------------------------------------------ context management

static void ctx_free(ClassicCtx *c)
{
    if (!c) return;
    if (c->B) { cudaSetDevice(c->B->device); if (c->d_req) cudaFree(c->d_req); if (c->d_out) cudaFree(c->d_out); if (c->d_cnt) cudaFree(c->d_cnt); odeb_destroy(c->B); }
    delete c;
}

static dxWorld *space_world(dxSpace *s)
{
    for (size_t i = 0; i < s->geoms.size(); i++) if (s->geoms[i]->body) return s->geoms[i]->body->world;
    return 0;
}

static void body_to_host(const dxBody *b, HostBody &h)
{
    for (int k = 0; k < 3; k++) h.pos[k] = b->pos[k];
    for (int k = 0; k < 4; k++) h.q[k] = b->q[k];
    for (int k = 0; k < 12; k++) h.R[k] = b->R[k];
}

// (re)build the device context of a world for the given space (may be NULL: no collision geometry)
static ClassicCtx *ctx_get(dxWorld *w, dxSpace *s, int need_contacts)
{
    ClassicCtx *c = w->ctx;
    if (c && !w->topo_dirty && (s == 0 || c->space == s) && need_contacts <= c->cap_contacts) return c;
    int cap_pairs = 0, cap_contacts = 0;
    if (c) { cap_pairs = c->cap_pairs; cap_contacts = c->cap_contacts; if (!s) s = c->space; }
    ctx_free(c); w->ctx = 0;
    if (w->bodies.empty()) { classic_error("world has no bodies"); return 0; }
    const int nb = (int)w->bodies.size(), ng = s ? (int)s->geoms.size() : 0;
    HostTemplate T;
    T.bmass.resize(nb); T.binvmass.resize(nb); T.bI.assign(12 * (size_t)nb, 0); T.binvI.assign(12 * (size_t)nb, 0);
    T.hb.resize(nb); T.bflags0.resize(nb);
    for (int i = 0; i < nb; i++) {
        dxBody *b = w->bodies[i];
        b->index = i;
        T.bmass[i] = b->mass.mass; T.binvmass[i] = b->invMass;
        for (int k = 0; k < 12; k++) { T.bI[12 * i + k] = b->mass.I[k]; T.binvI[12 * i + k] = b->invI[k]; }
        body_to_host(b, T.hb[i]);
        T.bflags0[i] = b->flags;
    }
    T.gtype.resize(ng); T.gbody.resize(ng); T.gparam.resize(4 * (size_t)ng); T.gcat.resize(ng); T.gcol.resize(ng); T.gspose.resize(4 * (size_t)ng); T.gofs.assign(ng, 0);
    for (int i = 0; i < ng; i++) {
        dxGeom *g = s->geoms[i];
        g->index = i;
        if (g->body && g->body->world != w) { classic_error("geoms of one space must belong to bodies of one world"); return 0; }
        T.gtype[i] = g->type; T.gbody[i] = g->body ? g->body->index : -1;
        T.gcat[i] = (unsigned)g->cat; T.gcol[i] = (unsigned)g->col;
        for (int k = 0; k < 4; k++) T.gparam[4 * i + k] = g->p[k];
        const bool ofs = g->body && g->has_ofs;
        const Real *gp = ofs ? g->opos : g->pos, *gR = ofs ? g->oR : g->R;
        Real4 p = { gp[0], gp[1], gp[2], 0 }, r0 = { gR[0], gR[1], gR[2], 0 }, r1 = { gR[4], gR[5], gR[6], 0 }, r2 = { gR[8], gR[9], gR[10], 0 };
        T.gspose[4 * i] = p; T.gspose[4 * i + 1] = r0; T.gspose[4 * i + 2] = r1; T.gspose[4 * i + 3] = r2;
        T.gofs[i] = ofs ? 1 : 0;
    }
    c = new ClassicCtx();
    c->B = 0; c->space = s; c->nb = nb; c->ng = ng; c->in_pass = false; c->cache_flags = -1; c->cache_maxc = 0;
    c->d_req = 0; c->d_out = 0; c->d_cnt = 0; c->req_cap = 0;
    for (size_t i = 0; i < w->joints.size(); i++) {
        dxJoint *j = w->joints[i];
        if (j->type != dJointTypeContact) c->perm.push_back(j);
    }
    c->nj = (int)c->perm.size();
    T.jt.resize(c->nj);
    T.sofs.assign(nb + 1, 0);
    // capacities: pairs grow on overflow, contacts to what the caller is about to submit
    long long all = (long long)ng * (ng - 1) / 2;
    if (cap_pairs < 64) cap_pairs = 64;
    if ((long long)cap_pairs < 8LL * ng) cap_pairs = (int)std::min<long long>(8LL * ng, std::max<long long>(all, 64));
    if (cap_contacts < 256) cap_contacts = 256;
    if (cap_contacts < 4 * ng) cap_contacts = 4 * ng;
    while (cap_contacts < need_contacts) cap_contacts *= 2;
    c->cap_pairs = cap_pairs; c->cap_contacts = cap_contacts;
    OdebWorldParams wp = w->wp;
    wp.space_type = s ? s->type : ODEB_SPACE_HASH;
    wp.max_contacts = 8;
    BatchCaps caps = { 1, cap_pairs, cap_contacts };
    c->B = batch_build(&wp, T, 1, 0, &caps);
    if (!c->B) { delete c; classic_error("device context: %s", g_err.c_str()); return 0; }
    w->ctx = c; w->topo_dirty = false;
    return c;
}

// host mirror -> device (every collide / step: setters may have touched anything)
static int ctx_upload_state(dxWorld *w, ClassicCtx *c)
{
    OdebBatch *B = c->B;
    const int nb = c->nb;
    const int NS = B->P.adis_samples > 0 ? B->P.adis_samples : 1;
    std::vector<Real4> v(nb * 9);
    std::vector<int> fl(nb), st(nb), ac(nb), ar(nb); std::vector<Real> tl(nb), ab((size_t)nb * 6 * NS, 0);
    for (int i = 0; i < nb; i++) {
        const dxBody *b = w->bodies[i];
        Real4 p = { b->pos[0], b->pos[1], b->pos[2], 0 }, q = { b->q[0], b->q[1], b->q[2], b->q[3] };
        Real4 l = { b->lvel[0], b->lvel[1], b->lvel[2], 0 }, a = { b->avel[0], b->avel[1], b->avel[2], 0 };
        Real4 f = { b->facc[0], b->facc[1], b->facc[2], 0 }, t = { b->tacc[0], b->tacc[1], b->tacc[2], 0 };
        Real4 r0 = { b->R[0], b->R[1], b->R[2], 0 }, r1 = { b->R[4], b->R[5], b->R[6], 0 }, r2 = { b->R[8], b->R[9], b->R[10], 0 };
        v[i] = p; v[nb + i] = q; v[2 * nb + i] = l; v[3 * nb + i] = a; v[4 * nb + i] = f; v[5 * nb + i] = t;
        v[6 * nb + 3 * i] = r0; v[6 * nb + 3 * i + 1] = r1; v[6 * nb + 3 * i + 2] = r2;
        fl[i] = b->flags; st[i] = b->adis_steps_left; tl[i] = b->adis_time_left; ac[i] = b->avg_counter; ar[i] = b->avg_ready;
        for (size_t k = 0; k < b->avg_buf.size() && k < (size_t)6 * NS; k++) ab[(size_t)i * 6 * NS + k] = b->avg_buf[k];
    }
    const DevPtrs &D = B->D;
    CK(cudaMemcpy(D.pos, &v[0], nb * sizeof(Real4), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(D.quat, &v[nb], nb * sizeof(Real4), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(D.lvel, &v[2 * nb], nb * sizeof(Real4), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(D.avel, &v[3 * nb], nb * sizeof(Real4), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(D.facc, &v[4 * nb], nb * sizeof(Real4), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(D.tacc, &v[5 * nb], nb * sizeof(Real4), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(D.R, &v[6 * nb], 3 * nb * sizeof(Real4), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(D.bflags, fl.data(), nb * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(D.adis_steps, st.data(), nb * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(D.adis_time, tl.data(), nb * sizeof(Real), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(D.avg_counter, ac.data(), nb * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(D.avg_ready, ar.data(), nb * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(D.avg_buf, ab.data(), ab.size() * sizeof(Real), cudaMemcpyHostToDevice));
    // geoms without a body can be moved between steps
    if (c->space && c->ng > 0) {
        std::vector<Real4> gs(4 * (size_t)c->ng);
        std::vector<unsigned> gc(c->ng), gl(c->ng);
        std::vector<int> go(c->ng);
        for (int i = 0; i < c->ng; i++) {
            const dxGeom *g = c->space->geoms[i];
            const bool ofs = g->body && g->has_ofs;
            const Real *gp = ofs ? g->opos : g->pos, *gR = ofs ? g->oR : g->R;
            Real4 p = { gp[0], gp[1], gp[2], 0 }, r0 = { gR[0], gR[1], gR[2], 0 }, r1 = { gR[4], gR[5], gR[6], 0 }, r2 = { gR[8], gR[9], gR[10], 0 };
            gs[4 * i] = p; gs[4 * i + 1] = r0; gs[4 * i + 2] = r1; gs[4 * i + 3] = r2;
            gc[i] = (unsigned)g->cat; gl[i] = (unsigned)g->col; go[i] = ofs ? 1 : 0;
        }
        CK(cudaMemcpy(D.gspose, gs.data(), gs.size() * sizeof(Real4), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(D.gofs, go.data(), go.size() * sizeof(int), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(D.gcat, gc.data(), gc.size() * sizeof(unsigned), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(D.gcol, gl.data(), gl.size() * sizeof(unsigned), cudaMemcpyHostToDevice));
    }
    return 1;
}

static int ctx_download_state(dxWorld *w, ClassicCtx *c)
{
    OdebBatch *B = c->B;
    const int nb = c->nb;
    const int NS = B->P.adis_samples > 0 ? B->P.adis_samples : 1;
    std::vector<Real4> v(nb * 9);
    std::vector<int> fl(nb), st(nb), ac(nb), ar(nb); std::vector<Real> tl(nb), ab((size_t)nb * 6 * NS);
    const DevPtrs &D = B->D;
    CK(cudaMemcpy(&v[7 * nb], D.facc, nb * sizeof(Real4), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&v[8 * nb], D.tacc, nb * sizeof(Real4), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&v[0], D.pos, nb * sizeof(Real4), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&v[nb], D.quat, nb * sizeof(Real4), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&v[2 * nb], D.lvel, nb * sizeof(Real4), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&v[3 * nb], D.avel, nb * sizeof(Real4), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&v[4 * nb], D.R, 3 * nb * sizeof(Real4), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(fl.data(), D.bflags, nb * sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(st.data(), D.adis_steps, nb * sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(tl.data(), D.adis_time, nb * sizeof(Real), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ac.data(), D.avg_counter, nb * sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ar.data(), D.avg_ready, nb * sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ab.data(), D.avg_buf, ab.size() * sizeof(Real), cudaMemcpyDeviceToHost));
    for (int i = 0; i < nb; i++) {
        dxBody *b = w->bodies[i];
        Real4 p = v[i], q = v[nb + i], l = v[2 * nb + i], a = v[3 * nb + i];
        b->pos[0] = p.x; b->pos[1] = p.y; b->pos[2] = p.z;
        b->q[0] = q.x; b->q[1] = q.y; b->q[2] = q.z; b->q[3] = q.w;
        b->lvel[0] = l.x; b->lvel[1] = l.y; b->lvel[2] = l.z;
        b->avel[0] = a.x; b->avel[1] = a.y; b->avel[2] = a.z;
        Real4 r0 = v[4 * nb + 3 * i], r1 = v[4 * nb + 3 * i + 1], r2 = v[4 * nb + 3 * i + 2];
        b->R[0] = r0.x; b->R[1] = r0.y; b->R[2] = r0.z; b->R[3] = 0; b->R[4] = r1.x; b->R[5] = r1.y; b->R[6] = r1.z; b->R[7] = 0;
        b->R[8] = r2.x; b->R[9] = r2.y; b->R[10] = r2.z; b->R[11] = 0;
        // accumulators: consumed (zeroed) for the bodies that were stepped, kept for disabled ones (quickstep.cpp:3432-3433)
        Real4 f = v[7 * nb + i], t = v[8 * nb + i];
        b->facc[0] = f.x; b->facc[1] = f.y; b->facc[2] = f.z; b->tacc[0] = t.x; b->tacc[1] = t.y; b->tacc[2] = t.z;
        b->flags = fl[i]; b->adis_steps_left = st[i]; b->adis_time_left = tl[i]; b->avg_counter = ac[i]; b->avg_ready = ar[i];
        b->avg_buf.assign(ab.begin() + (size_t)i * 6 * NS, ab.begin() + (size_t)(i + 1) * 6 * NS);
    }
    return 1;
}

// ------------------------------------------------------------------------------------------------ helpers shared by the setters

static void set_identity_R(Real *R) { for (int k = 0; k < 12; k++) R[k] = 0; R[0] = R[5] = R[10] = 1; }

static bool safe_normalize3_host(Real *a)
{   // dxSafeNormalize3 odemath.cpp:95-163 is what normalize3() restates; it reports failure for the zero vector
    if (a[0] == 0 && a[1] == 0 && a[2] == 0) return false;
    normalize3(a);
    return true;
}

// dxOrthogonalizeR odemath.cpp:260-310
static bool orthogonalize_R(Real *m)
{
    if (m[0] == 0 && m[1] == 0 && m[2] == 0) return false;
    Real n0 = m[0] * m[0] + m[1] * m[1] + m[2] * m[2];
    Real store[3]; Real *row2 = m + 4;
    Real proj = m[0] * m[4] + m[1] * m[5] + m[2] * m[6];
    if (proj != 0) {
        Real pd = proj / n0;
        store[0] = m[4] - pd * m[0]; store[1] = m[5] - pd * m[1]; store[2] = m[6] - pd * m[2];
        row2 = store;
    }
    if (row2[0] == 0 && row2[1] == 0 && row2[2] == 0) return false;
    if (n0 != R_(1.0)) safe_normalize3_host(m);
    Real n1 = row2[0] * row2[0] + row2[1] * row2[1] + row2[2] * row2[2];
    if (n1 != R_(1.0)) safe_normalize3_host(row2);
    cross3(m + 8, m, row2);
    m[3] = m[7] = m[11] = 0;
    return true;
}

static std::vector<HostBody> joint_bodies(const dxJoint *j, DJointT &t)
{   // two-entry body table so that the batch helpers (indexing by t.b0 / t.b1) can be reused
    std::vector<HostBody> hb(2);
    t.b0 = j->body[0] ? 0 : -1; t.b1 = j->body[1] ? 1 : -1;
    if (j->body[0]) body_to_host(j->body[0], hb[0]);
    if (j->body[1]) body_to_host(j->body[1], hb[1]);
    return hb;
}

static void joint_get_anchor(const dxJoint *j, const Real *anchor1, Real *result)
{   // getAnchor joints/joint.cpp:372-383
    if (j->body[0]) { mul0_331(result, j->body[0]->R, anchor1); for (int k = 0; k < 3; k++) result[k] += j->body[0]->pos[k]; }
}
static void joint_get_anchor2(const dxJoint *j, const Real *anchor2, Real *result)
{   // getAnchor2 joints/joint.cpp:386-403
    if (j->body[1]) { mul0_331(result, j->body[1]->R, anchor2); for (int k = 0; k < 3; k++) result[k] += j->body[1]->pos[k]; }
    else for (int k = 0; k < 3; k++) result[k] = anchor2[k];
}
static void joint_get_axis(const dxJoint *j, const Real *axis1, Real *result) { if (j->body[0]) mul0_331(result, j->body[0]->R, axis1); }
static void joint_get_axis2(const dxJoint *j, const Real *axis2, Real *result)
{
    if (j->body[1]) mul0_331(result, j->body[1]->R, axis2); else for (int k = 0; k < 3; k++) result[k] = axis2[k];
}

static void limot_init(DLimot &l, const dxWorld *w)
{   // dxJointLimitMotor::init joints/joint.cpp:494-507
    Real cfm;
#if defined(ODEB_DOUBLE)
    cfm = w->wp.cfm >= 0 ? (Real)w->wp.cfm : R_(1e-10);
#else
    cfm = w->wp.cfm >= 0 ? (Real)w->wp.cfm : R_(1e-5);
#endif
    l.vel = 0; l.fmax = 0; l.lostop = -R_INF; l.histop = R_INF; l.fudge_factor = 1;
    l.normal_cfm = cfm; l.stop_erp = (Real)w->wp.erp; l.stop_cfm = cfm; l.bounce = 0;
}
static void limot_set(DLimot &l, int num, Real value)
{   // dxJointLimitMotor::set joints/joint.cpp:510-541
    switch (num) {
    case dParamLoStop: l.lostop = value; break;
    case dParamHiStop: l.histop = value; break;
    case dParamVel: l.vel = value; break;
    case dParamFMax: if (value >= 0) l.fmax = value; break;
    case dParamFudgeFactor: if (value >= 0 && value <= 1) l.fudge_factor = value; break;
    case dParamBounce: l.bounce = value; break;
    case dParamCFM: l.normal_cfm = value; break;
    case dParamStopERP: l.stop_erp = value; break;
    case dParamStopCFM: l.stop_cfm = value; break;
    }
}

static void joint_set_relative_values(dxJoint *j)
{   // dxJoint*::setRelativeValues (ball.cpp:179-185, hinge.cpp:359-369, universal.cpp:785-808): called by dJointAttach
    DJointT &t = j->t;
    if (j->type == dJointTypeContact || j->type == dJointTypeFixed) return;     // dxJointFixed keeps offset / qrel until dJointSetFixed
    if (j->type == dJointTypeAMotor || j->type == dJointTypeLMotor) return;     // no setRelativeValues override (joints/joint.cpp:68-71)
    if (j->type == dJointTypeHinge2) {                                          // hinge2.cpp:517-536: anchor, both axes, v1/v2 from the new bodies
        Real anchor[3] = { 0, 0, 0 }, a1[3] = { 0, 0, 0 }, a2[3] = { 0, 0, 0 };
        if (!j->body[0] || !j->body[1]) return;
        joint_get_anchor(j, t.anchor1, anchor);
        mul0_331(a1, j->body[0]->R, t.axis1);
        mul0_331(a2, j->body[1]->R, t.axis2);
        std::vector<HostBody> hb = joint_bodies(j, t);
        host_set_anchors(hb, t, anchor[0], anchor[1], anchor[2]);
        host_set_axes(hb, t, a1[0], a1[1], a1[2], t.axis1, 0);
        host_set_axes(hb, t, a2[0], a2[1], a2[2], 0, t.axis2);
        host_hinge2_finish(hb, t);
        return;
    }
    if (j->type == dJointTypeSlider) {                                          // slider.cpp:371-376: computeOffset + computeInitialRelativeRotation
        std::vector<HostBody> hb = joint_bodies(j, t);
        if (t.b0 >= 0) {
            if (t.b1 >= 0) { Real c[3] = { hb[0].pos[0] - hb[1].pos[0], hb[0].pos[1] - hb[1].pos[1], hb[0].pos[2] - hb[1].pos[2] }; mul1_331(t.anchor1, hb[1].R, c); }
            else { t.anchor1[0] = hb[0].pos[0]; t.anchor1[1] = hb[0].pos[1]; t.anchor1[2] = hb[0].pos[2]; }
            host_hinge_initial_rotation(hb, t);
        }
        return;
    }
    Real anchor[3] = { 0, 0, 0 };
    if (j->reverse) joint_get_anchor2(j, t.anchor2, anchor); else joint_get_anchor(j, t.anchor1, anchor);     // dJointGet{Ball,Hinge,Universal}Anchor
    std::vector<HostBody> hb = joint_bodies(j, t);
    host_set_anchors(hb, t, anchor[0], anchor[1], anchor[2]);
    if (j->type == dJointTypeHinge) {
        Real ax[3] = { 0, 0, 0 };
        joint_get_axis(j, t.axis1, ax);                                           // dJointGetHingeAxis hinge.cpp:258-265
        host_set_axes(hb, t, ax[0], ax[1], ax[2], t.axis1, t.axis2);
        host_hinge_initial_rotation(hb, t);
    } else if (j->type == dJointTypeUniversal) {
        Real ax1[3] = { 0, 0, 0 }, ax2[3] = { 0, 0, 0 };
        if (j->reverse) { joint_get_axis2(j, t.axis2, ax1); joint_get_axis(j, t.axis1, ax2); }       // dJointGetUniversalAxis1/2 universal.cpp:610-633
        else { joint_get_axis(j, t.axis1, ax1); joint_get_axis2(j, t.axis2, ax2); }
        if (j->reverse) { host_set_axes(hb, t, ax1[0], ax1[1], ax1[2], 0, t.axis2); host_set_axes(hb, t, ax2[0], ax2[1], ax2[2], t.axis1, 0); }
        else { host_set_axes(hb, t, ax1[0], ax1[1], ax1[2], t.axis1, 0); host_set_axes(hb, t, ax2[0], ax2[1], ax2[2], 0, t.axis2); }
        host_universal_initial_rotations(hb, t);
    }
}

static void detach_joint(dxJoint *j)
{   // removeJointReferencesFromAttachedBodies ode.cpp:77-99
    for (int k = 0; k < 2; k++) {
        dxBody *b = j->body[k];
        if (!b) continue;
        for (size_t i = 0; i < b->joints.size(); i++) if (b->joints[i].joint == j) { b->joints.erase(b->joints.begin() + i); break; }
    }
    j->body[0] = j->body[1] = 0;
}

static void destroy_joint(dxJoint *j)
{
    dxWorld *w = j->world;
    detach_joint(j);
    // searched from the back: a joint group is emptied newest first and a step's contact joints are the newest joints of the world, so
    // dJointGroupEmpty is linear, not quadratic, in the number of contacts
    for (size_t i = w->joints.size(); i > 0;) if (w->joints[--i] == j) { w->joints.erase(w->joints.begin() + (std::ptrdiff_t)i); break; }
    if (j->type != dJointTypeContact) w->topo_dirty = true;
    delete j;
}

static dxJoint *new_joint(dxWorld *w, dxJointGroup *g, int type)
{
    dxJoint *j = new dxJoint();
    j->world = w; j->group = g; j->userdata = 0; j->feedback = 0; j->type = type; j->body[0] = j->body[1] = 0; j->reverse = 0;
    memset(&j->t, 0, sizeof(j->t)); memset(&j->contact, 0, sizeof(j->contact));
    j->t.type = type;
    j->t.erp = (Real)w->wp.erp;
#if defined(ODEB_DOUBLE)
    j->t.cfm = w->wp.cfm >= 0 ? (Real)w->wp.cfm : R_(1e-10);
#else
    j->t.cfm = w->wp.cfm >= 0 ? (Real)w->wp.cfm : R_(1e-5);
#endif
    limot_init(j->t.limot1, w); limot_init(j->t.limot2, w); limot_init(j->t.limot3, w);
    if (type == dJointTypeHinge) { j->t.axis1[0] = 1; j->t.axis2[0] = 1; }
    if (type == dJointTypeSlider) j->t.axis1[0] = 1;
    if (type == dJointTypeHinge2) { j->t.axis1[0] = 1; j->t.axis2[1] = 1; j->t.qrel1[0] = 1; j->t.qrel2[1] = 1; j->t.qrel[2] = j->t.erp; j->t.qrel[3] = j->t.cfm; }   // hinge2.cpp:76-100
    if (type == dJointTypeUniversal) { j->t.axis1[0] = 1; j->t.axis2[1] = 1; }
    w->joints.push_back(j);
    if (g) g->joints.push_back(j);
    if (type != dJointTypeContact) w->topo_dirty = true;
    return j;
}

extern "C" {

// ------------------------------------------------------------------------------------------------ init, rng, mass, rotation

void dInitODE(void) { g_init_count++; }
int dInitODE2(unsigned int) { g_init_count++; return 1; }
int dAllocateODEDataForThread(unsigned int) { return 1; }
void dCloseODE(void) { if (g_init_count > 0) g_init_count--; }
const char *dGetConfiguration(void)
{
#if defined(ODEB_DOUBLE)
    return "ODE ODE_B200_cuda_sm_100a ODE_double_precision";
#else
    return "ODE ODE_B200_cuda_sm_100a ODE_single_precision";
#endif
}
int dCheckConfiguration(const char *token)
{   // ode.cpp:2398-2430: whole-token match inside the configuration string
    const char *cfg = dGetConfiguration();
    size_t n = strlen(token);
    if (!n) return 1;
    for (const char *p = cfg; (p = strstr(p, token)) != 0; p += n)
        if ((p == cfg || p[-1] == ' ') && (p[n] == ' ' || p[n] == 0)) return 1;
    return 0;
}

unsigned long dRand(void) { unsigned s = (unsigned)g_seed; unsigned r = odeb_rand(&s); g_seed = s; return r; }
unsigned long dRandGetSeed(void) { return g_seed; }
void dRandSetSeed(unsigned long s) { g_seed = s; }
int dRandInt(int n) { unsigned s = (unsigned)g_seed; int r = odeb_rand_int(&s, n); g_seed = s; return r; }
odeb_real dRandReal(void) { return (Real)(((double)dRand()) / ((double)0xffffffff)); }

void dMassSetZero(dMass *m) { m->mass = 0; for (int k = 0; k < 4; k++) m->c[k] = 0; for (int k = 0; k < 12; k++) m->I[k] = 0; }
void dMassSetParameters(dMass *m, Real themass, Real cgx, Real cgy, Real cgz, Real I11, Real I22, Real I33, Real I12, Real I13, Real I23)
{   // mass.cpp:74-93
    dMassSetZero(m);
    m->mass = themass; m->c[0] = cgx; m->c[1] = cgy; m->c[2] = cgz;
    m->I[0] = I11; m->I[5] = I22; m->I[10] = I33; m->I[1] = I12; m->I[2] = I13; m->I[6] = I23; m->I[4] = I12; m->I[8] = I13; m->I[9] = I23;
}
void dMassSetSphereTotal(dMass *m, Real total_mass, Real radius)
{   // mass.cpp:104-117
    dMassSetZero(m);
    m->mass = total_mass;
    Real II = R_(0.4) * total_mass * radius * radius;
    m->I[0] = II; m->I[5] = II; m->I[10] = II;
}
void dMassSetSphere(dMass *m, Real density, Real radius)
{   // mass.cpp:96-101: the expression is evaluated in double (M_PI) and cast once
    dMassSetSphereTotal(m, (Real)((R_(4.0) / R_(3.0)) * M_PI * radius * radius * radius * density), radius);
}
void dMassAdjust(dMass *m, Real newmass)
{   // mass.cpp:383-390
    Real scale = newmass / m->mass;
    m->mass = newmass;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m->I[i * 4 + j] *= scale;
}
void dMassSetCapsule(dMass *m, Real density, int direction, Real radius, Real length)
{   // mass.cpp:120-143
    Real M1, M2, Ia, Ib;
    dMassSetZero(m);
    M1 = (Real)(M_PI * radius * radius * length * density);
    M2 = (Real)((R_(4.0) / R_(3.0)) * M_PI * radius * radius * radius * density);
    m->mass = M1 + M2;
    Ia = M1 * (R_(0.25) * radius * radius + (R_(1.0) / R_(12.0)) * length * length) +
         M2 * (R_(0.4) * radius * radius + R_(0.375) * radius * length + R_(0.25) * length * length);
    Ib = (M1 * R_(0.5) + M2 * R_(0.4)) * radius * radius;
    m->I[0] = Ia; m->I[5] = Ia; m->I[10] = Ia;
    m->I[(direction - 1) * 5] = Ib;
}
void dMassSetCylinderTotal(dMass *m, Real total_mass, int direction, Real radius, Real length)
{   // mass.cpp:180-198
    dMassSetZero(m);
    const Real r2 = radius * radius;
    m->mass = total_mass;
    const Real I = total_mass * (R_(0.25) * r2 + (R_(1.0) / R_(12.0)) * length * length);
    m->I[0] = I; m->I[5] = I; m->I[10] = I;
    m->I[(direction - 1) * 5] = total_mass * R_(0.5) * r2;
}
void dMassSetCylinder(dMass *m, Real density, int direction, Real radius, Real length)
{   // mass.cpp:173-178
    dMassSetCylinderTotal(m, (Real)(M_PI * radius * radius * length * density), direction, radius, length);
}
void dMassSetCapsuleTotal(dMass *m, Real total_mass, int direction, Real a, Real b)
{
    dMassSetCapsule(m, 1.0, direction, a, b);
    dMassAdjust(m, total_mass);
}
void dMassSetBoxTotal(dMass *m, Real total_mass, Real lx, Real ly, Real lz)
{   // mass.cpp:198-212
    dMassSetZero(m);
    m->mass = total_mass;
    m->I[0] = total_mass / R_(12.0) * (ly * ly + lz * lz);
    m->I[5] = total_mass / R_(12.0) * (lx * lx + lz * lz);
    m->I[10] = total_mass / R_(12.0) * (lx * lx + ly * ly);
}
void dMassSetBox(dMass *m, Real density, Real lx, Real ly, Real lz) { dMassSetBoxTotal(m, lx * ly * lz * density, lx, ly, lz); }

void dRSetIdentity(Real *R) { set_identity_R(R); }
void dQSetIdentity(Real *q) { q[0] = 1; q[1] = q[2] = q[3] = 0; }
void dQFromAxisAndAngle(Real *q, Real ax, Real ay, Real az, Real angle)
{   // rotation.cpp:169-188
    Real l = ax * ax + ay * ay + az * az;
    if (l > R_(0.0)) {
        angle *= R_(0.5);
        q[0] = RCOS(angle);
        l = RSIN(angle) * rrecipsqrt(l);
        q[1] = ax * l; q[2] = ay * l; q[3] = az * l;
    } else { q[0] = 1; q[1] = 0; q[2] = 0; q[3] = 0; }
}
void dRfromQ(Real *R, const Real *q) { r_from_q(R, q); }
void dQfromR(Real *q, const Real *R) { q_from_r(q, R); }
void dRFromAxisAndAngle(Real *R, Real ax, Real ay, Real az, Real angle)
{   // rotation.cpp:58-65
    Real q[4];
    dQFromAxisAndAngle(q, ax, ay, az, angle);
    r_from_q(R, q);
}

// ------------------------------------------------------------------------------------------------ world

dWorldID dWorldCreate(void)
{   // dxWorld::dxWorld objects.cpp:99-121 + defaults :37-91
    dxWorld *w = new dxWorld();
    OdebWorldParams &p = w->wp;
    memset(&p, 0, sizeof(p));
    p.erp = 0.2; p.cfm = -1; p.num_iterations = 20; p.sor_w = 1.3;
    p.premature_exit_delta = 1e-8; p.max_extra_factor = 1.0; p.extra_iter_delta = 1e-2;
    p.contact_max_vel = INFINITY; p.contact_surface_layer = 0;
    p.auto_disable = 0; p.adis_linear_thr = 0.01; p.adis_angular_thr = 0.01; p.adis_steps = 10; p.adis_time = 0; p.adis_samples = 1;
    p.linear_damping = 0; p.angular_damping = 0; p.linear_damping_thr = 0.01; p.angular_damping_thr = 0.01;
    p.max_angular_speed = INFINITY;
    p.space_type = ODEB_SPACE_HASH; p.max_contacts = 8; p.skip_connected = 0; p.surf_mode = 0; p.mu = 0;
    w->ctx = 0; w->topo_dirty = true; w->stats_sink = 0; w->body_flags_default = 0;
    return w;
}
void dWorldDestroy(dWorldID w)
{   // ode.cpp:1590-1623: the world owns its bodies and ungrouped joints
    ctx_free(w->ctx); w->ctx = 0;
    for (size_t i = 0; i < w->bodies.size(); i++) {
        dxBody *b = w->bodies[i];
        for (size_t k = 0; k < b->geoms.size(); k++) b->geoms[k]->body = 0;
        delete b;
    }
    for (size_t i = 0; i < w->joints.size(); i++) {
        dxJoint *j = w->joints[i];
        if (j->group) { j->world = 0; j->body[0] = j->body[1] = 0; }   // group-owned storage: deactivated, freed with the group (ode.cpp:1608-1618)
        else delete j;
    }
    delete w;
}
void dWorldSetGravity(dWorldID w, Real x, Real y, Real z) { w->wp.gravity[0] = x; w->wp.gravity[1] = y; w->wp.gravity[2] = z; }
void dWorldGetGravity(dWorldID w, Real *g) { for (int k = 0; k < 3; k++) g[k] = (Real)w->wp.gravity[k]; }
void dWorldSetERP(dWorldID w, Real erp) { w->wp.erp = erp; }
odeb_real dWorldGetERP(dWorldID w) { return (Real)w->wp.erp; }
void dWorldSetCFM(dWorldID w, Real cfm) { w->wp.cfm = cfm; }
odeb_real dWorldGetCFM(dWorldID w)
{
#if defined(ODEB_DOUBLE)
    return w->wp.cfm >= 0 ? (Real)w->wp.cfm : R_(1e-10);
#else
    return w->wp.cfm >= 0 ? (Real)w->wp.cfm : R_(1e-5);
#endif
}
void dWorldSetQuickStepNumIterations(dWorldID w, int num) { w->wp.num_iterations = num; }
int dWorldGetQuickStepNumIterations(dWorldID w) { return w->wp.num_iterations; }
void dWorldSetQuickStepW(dWorldID w, Real v) { w->wp.sor_w = v; }
odeb_real dWorldGetQuickStepW(dWorldID w) { return (Real)w->wp.sor_w; }
void dWorldSetQuickStepDynamicIterationParameters(dWorldID w, const Real *a, const Real *b, const Real *c)
{   // ode.cpp:2079-2106
    if (a) w->wp.premature_exit_delta = *a;
    if (b) w->wp.max_extra_factor = *b;
    if (c) w->wp.extra_iter_delta = *c;
}
void dWorldGetQuickStepDynamicIterationParameters(dWorldID w, Real *a, Real *b, Real *c)
{
    if (a) *a = (Real)w->wp.premature_exit_delta;
    if (b) *b = (Real)w->wp.max_extra_factor;
    if (c) *c = (Real)w->wp.extra_iter_delta;
}
int dWorldAttachQuickStepDynamicIterationStatisticsSink(dWorldID w, dWorldQuickStepIterationCount_DynamicAdjustmentStatistics *s)
{   // ode.cpp:2124-2146: the structure size must cover the known fields
    if (s && s->struct_size < sizeof(*s)) return 0;
    w->stats_sink = s;
    return 1;
}
void dWorldSetContactMaxCorrectingVel(dWorldID w, Real v) { w->wp.contact_max_vel = v; }
odeb_real dWorldGetContactMaxCorrectingVel(dWorldID w) { return (Real)w->wp.contact_max_vel; }
void dWorldSetContactSurfaceLayer(dWorldID w, Real d) { w->wp.contact_surface_layer = d; }
odeb_real dWorldGetContactSurfaceLayer(dWorldID w) { return (Real)w->wp.contact_surface_layer; }
void dWorldSetAutoDisableFlag(dWorldID w, int f) { w->wp.auto_disable = f ? 1 : 0; }     // default for bodies created afterwards (ode.cpp:1960-1967)
int dWorldGetAutoDisableFlag(dWorldID w) { return w->wp.auto_disable; }
void dWorldSetAutoDisableLinearThreshold(dWorldID w, Real t) { w->wp.adis_linear_thr = t; }
void dWorldSetAutoDisableAngularThreshold(dWorldID w, Real t) { w->wp.adis_angular_thr = t; }
void dWorldSetAutoDisableSteps(dWorldID w, int s) { w->wp.adis_steps = s; }
void dWorldSetAutoDisableTime(dWorldID w, Real t) { w->wp.adis_time = t; }
void dWorldSetAutoDisableAverageSamplesCount(dWorldID w, unsigned int n) { w->wp.adis_samples = (int)n; w->topo_dirty = true; }
void dWorldSetLinearDampingThreshold(dWorldID w, Real t) { w->wp.linear_damping_thr = t; }
void dWorldSetAngularDampingThreshold(dWorldID w, Real t) { w->wp.angular_damping_thr = t; }
void dWorldSetLinearDamping(dWorldID w, Real s) { w->wp.linear_damping = s; }     // flags of bodies created afterwards (ode.cpp:2233-2241)
void dWorldSetAngularDamping(dWorldID w, Real s) { w->wp.angular_damping = s; }
void dWorldSetDamping(dWorldID w, Real l, Real a) { dWorldSetLinearDamping(w, l); dWorldSetAngularDamping(w, a); }
void dWorldSetMaxAngularSpeed(dWorldID w, Real m) { w->wp.max_angular_speed = m; }

// ------------------------------------------------------------------------------------------------ bodies

dBodyID dBodyCreate(dWorldID w)
{   // ode.cpp:240-286
    dxBody *b = new dxBody();
    b->world = w; b->userdata = 0;
    for (int k = 0; k < 4; k++) { b->pos[k] = 0; b->q[k] = 0; b->lvel[k] = 0; b->avel[k] = 0; b->facc[k] = 0; b->tacc[k] = 0; }
    b->q[0] = 1; set_identity_R(b->R);
    dMassSetParameters(&b->mass, 1, 0, 0, 0, 1, 1, 1, 0, 0, 0);
    set_identity_R(b->invI); b->invMass = 1;
    int fl = BF_GYRO;
    if (w->wp.auto_disable) fl |= BF_AUTO_DISABLE;
    if ((Real)w->wp.linear_damping) fl |= BF_LIN_DAMP;
    if ((Real)w->wp.angular_damping) fl |= BF_ANG_DAMP;
    if ((Real)w->wp.max_angular_speed < R_INF) fl |= BF_MAX_ANG_SPEED;
    b->flags = fl;
    b->adis_steps_left = w->wp.adis_steps; b->adis_time_left = (Real)w->wp.adis_time; b->avg_counter = 0; b->avg_ready = 0;
    b->index = (int)w->bodies.size();
    w->bodies.push_back(b);
    w->topo_dirty = true;
    return b;
}
void dBodyDestroy(dBodyID b)
{   // ode.cpp:289-329: detach geoms and joints, unlink from the world
    dxWorld *w = b->world;
    for (size_t k = 0; k < b->geoms.size(); k++) b->geoms[k]->body = 0;
    while (!b->joints.empty()) {
        dxJoint *j = b->joints.back().joint;
        detach_joint(j);            // the joint stays in the world, in limbo (ode.cpp:305-312)
    }
    w->bodies.erase(std::find(w->bodies.begin(), w->bodies.end(), b));
    for (size_t i = 0; i < w->bodies.size(); i++) w->bodies[i]->index = (int)i;
    w->topo_dirty = true;
    delete b;
}
dWorldID dBodyGetWorld(dBodyID b) { return b->world; }
void dBodySetData(dBodyID b, void *d) { b->userdata = d; }
void *dBodyGetData(dBodyID b) { return b->userdata; }
void dBodySetPosition(dBodyID b, Real x, Real y, Real z) { b->pos[0] = x; b->pos[1] = y; b->pos[2] = z; }
void dBodySetRotation(dBodyID b, const Real *R)
{   // ode.cpp:362-377
    memcpy(b->R, R, 12 * sizeof(Real));
    orthogonalize_R(b->R);
    q_from_r(b->q, R);
    normalize4(b->q);
}
void dBodySetQuaternion(dBodyID b, const Real *q)
{   // ode.cpp:379-392
    for (int k = 0; k < 4; k++) b->q[k] = q[k];
    normalize4(b->q);
    r_from_q(b->R, b->q);
}
void dBodySetLinearVel(dBodyID b, Real x, Real y, Real z) { b->lvel[0] = x; b->lvel[1] = y; b->lvel[2] = z; }
void dBodySetAngularVel(dBodyID b, Real x, Real y, Real z) { b->avel[0] = x; b->avel[1] = y; b->avel[2] = z; }
const odeb_real *dBodyGetPosition(dBodyID b) { return b->pos; }
const odeb_real *dBodyGetRotation(dBodyID b) { return b->R; }
const odeb_real *dBodyGetQuaternion(dBodyID b) { return b->q; }
const odeb_real *dBodyGetLinearVel(dBodyID b) { return b->lvel; }
const odeb_real *dBodyGetAngularVel(dBodyID b) { return b->avel; }
void dBodySetMass(dBodyID b, const dMass *mass)
{   // ode.cpp:486-503
    b->mass = *mass;
    if (!host_invert_pd3(b->mass.I, b->invI)) set_identity_R(b->invI);
    b->invMass = rrecip(b->mass.mass);
    b->world->topo_dirty = true;
}
void dBodyGetMass(dBodyID b, dMass *mass) { *mass = b->mass; }
void dBodySetKinematic(dBodyID b)
{   // ode.cpp:837-842
    memset(b->invI, 0, sizeof(b->invI)); b->invMass = 0;
    b->world->topo_dirty = true;
}
void dBodySetDynamic(dBodyID b) { dBodySetMass(b, &b->mass); }      // ode.cpp:830-835
int dBodyIsKinematic(dBodyID b) { return b->invMass == 0; }
void dBodyAddForce(dBodyID b, Real fx, Real fy, Real fz) { b->facc[0] += fx; b->facc[1] += fy; b->facc[2] += fz; }
void dBodyAddTorque(dBodyID b, Real fx, Real fy, Real fz) { b->tacc[0] += fx; b->tacc[1] += fy; b->tacc[2] += fz; }
const odeb_real *dBodyGetForce(dBodyID b) { return b->facc; }
const odeb_real *dBodyGetTorque(dBodyID b) { return b->tacc; }
void dBodySetForce(dBodyID b, Real x, Real y, Real z) { b->facc[0] = x; b->facc[1] = y; b->facc[2] = z; }
void dBodySetTorque(dBodyID b, Real x, Real y, Real z) { b->tacc[0] = x; b->tacc[1] = y; b->tacc[2] = z; }
void dBodyEnable(dBodyID b)
{   // ode.cpp:1011-1020
    b->flags &= ~BF_DISABLED;
    b->adis_steps_left = b->world->wp.adis_steps; b->adis_time_left = (Real)b->world->wp.adis_time;
}
void dBodyDisable(dBodyID b) { b->flags |= BF_DISABLED; }
int dBodyIsEnabled(dBodyID b) { return (b->flags & BF_DISABLED) == 0; }
void dBodySetGravityMode(dBodyID b, int mode) { if (mode) b->flags &= ~BF_NO_GRAVITY; else b->flags |= BF_NO_GRAVITY; }
int dBodyGetGravityMode(dBodyID b) { return (b->flags & BF_NO_GRAVITY) == 0; }
void dBodySetGyroscopicMode(dBodyID b, int en) { if (en) b->flags |= BF_GYRO; else b->flags &= ~BF_GYRO; }
int dBodyGetGyroscopicMode(dBodyID b) { return (b->flags & BF_GYRO) != 0; }
void dBodySetFiniteRotationMode(dBodyID b, int mode) { if (mode) b->flags |= BF_FINITE_ROT; else b->flags &= ~BF_FINITE_ROT; }
int dBodyGetFiniteRotationMode(dBodyID b) { return (b->flags & BF_FINITE_ROT) != 0; }
void dBodySetAutoDisableFlag(dBodyID b, int f) { if (f) b->flags |= BF_AUTO_DISABLE; else b->flags &= ~BF_AUTO_DISABLE; }
int dBodyGetAutoDisableFlag(dBodyID b) { return (b->flags & BF_AUTO_DISABLE) != 0; }
int dBodyGetNumJoints(dBodyID b) { return (int)b->joints.size(); }

// ------------------------------------------------------------------------------------------------ joints

dJointGroupID dJointGroupCreate(int) { return new dxJointGroup(); }
void dJointGroupEmpty(dJointGroupID g)
{   // ode.cpp:1325-1366: joints are destroyed newest first
    for (size_t i = g->joints.size(); i > 0;) {
        --i;
        dxJoint *j = g->joints[i];
        if (j->world) { j->group = 0; destroy_joint(j); } else delete j;
    }
    g->joints.clear();
}
void dJointGroupDestroy(dJointGroupID g) { dJointGroupEmpty(g); delete g; }
dJointID dJointCreateContact(dWorldID w, dJointGroupID g, const dContact *c)
{   // ode.cpp:1192-1200: the contact is copied into the joint
    dxJoint *j = new_joint(w, g, dJointTypeContact);
    j->contact = *c;
    return j;
}
dJointID dJointCreateBall(dWorldID w, dJointGroupID g) { return new_joint(w, g, dJointTypeBall); }
dJointID dJointCreateHinge(dWorldID w, dJointGroupID g) { return new_joint(w, g, dJointTypeHinge); }
dJointID dJointCreateUniversal(dWorldID w, dJointGroupID g) { return new_joint(w, g, dJointTypeUniversal); }
dJointID dJointCreateFixed(dWorldID w, dJointGroupID g) { return new_joint(w, g, dJointTypeFixed); }
dJointID dJointCreateHinge2(dWorldID w, dJointGroupID g) { return new_joint(w, g, dJointTypeHinge2); }
void dJointSetHinge2Anchor(dJointID j, Real x, Real y, Real z)
{   // hinge2.cpp:268-279
    std::vector<HostBody> hb = joint_bodies(j, j->t);
    host_set_anchors(hb, j->t, x, y, z);
    host_hinge2_finish(hb, j->t);
}
void dJointSetHinge2Axes(dJointID j, const Real *axis1, const Real *axis2)
{   // hinge2.cpp:283-312
    std::vector<HostBody> hb = joint_bodies(j, j->t);
    if (axis1) host_set_axes(hb, j->t, axis1[0], axis1[1], axis1[2], j->t.axis1, 0);
    if (axis2) host_set_axes(hb, j->t, axis2[0], axis2[1], axis2[2], 0, j->t.axis2);
    host_hinge2_finish(hb, j->t);
}
void dJointSetHinge2Axis1(dJointID j, Real x, Real y, Real z) { Real a[4] = { x, y, z, 0 }; dJointSetHinge2Axes(j, a, 0); }
void dJointSetHinge2Axis2(dJointID j, Real x, Real y, Real z) { Real a[4] = { x, y, z, 0 }; dJointSetHinge2Axes(j, 0, a); }
void dJointSetHinge2Param(dJointID j, int parameter, Real value)
{   // hinge2.cpp:333-348; suspension ERP / CFM live in t.qrel[2] / t.qrel[3]
    if ((parameter & 0xff00) == 0x100) limot_set(j->t.limot2, parameter & 0xff, value);
    else if (parameter == dParamSuspensionERP) j->t.qrel[2] = value;
    else if (parameter == dParamSuspensionCFM) j->t.qrel[3] = value;
    else limot_set(j->t.limot1, parameter, value);
}
// ---- linear / angular motors (joints/lmotor.cpp, joints/amotor.cpp)
static DLimot &motor_limot(dxJoint *j, int anum) { return anum <= 0 ? j->t.limot1 : anum == 1 ? j->t.limot2 : j->t.limot3; }
static void motor_set_axis(dxJoint *j, int anum, int rel, Real x, Real y, Real z)
{
    if (anum < 0) anum = 0;
    if (anum > 2) anum = 2;
    std::vector<HostBody> hb = joint_bodies(j, j->t);
    j->t.reverse = j->reverse;
    host_motor_set_axis(hb, j->t, anum, rel, x, y, z);
    if (j->type == dJointTypeAMotor && j->t.mmode == dAMotorEuler) host_amotor_euler_references(hb, j->t);
   
}
static void motor_get_axis(dxJoint *j, int anum, Real *result)
{   // dJointGetLMotorAxis lmotor.cpp:196-205 returns the stored axis; dxJointAMotor::doGetUserAxis amotor.cpp:471-493 rotates it into the world frame
    if (anum < 0) anum = 0;
    if (anum > 2) anum = 2;
    const Real *a = j->t.maxis[anum];
    result[0] = a[0]; result[1] = a[1]; result[2] = a[2];
    if (j->type == dJointTypeAMotor && j->t.mmode == dAMotorUser) {
        if (j->t.mrel[anum] == 1 && j->body[0]) mul0_331(result, j->body[0]->R, a);
        else if (j->t.mrel[anum] == 2 && j->body[1]) mul0_331(result, j->body[1]->R, a);
    }
}
dJointID dJointCreateLMotor(dWorldID w, dJointGroupID g) { return new_joint(w, g, dJointTypeLMotor); }
void dJointSetLMotorNumAxes(dJointID j, int num) { j->t.mnum = num < 0 ? 0 : num > 3 ? 3 : num; }
int dJointGetLMotorNumAxes(dJointID j) { return j->t.mnum; }
void dJointSetLMotorAxis(dJointID j, int anum, int rel, Real x, Real y, Real z) { motor_set_axis(j, anum, rel, x, y, z); }
void dJointGetLMotorAxis(dJointID j, int anum, Real *result) { motor_get_axis(j, anum, result); }
void dJointSetLMotorParam(dJointID j, int parameter, Real value) { limot_set(motor_limot(j, parameter >> 8), parameter & 0xff, value); }
dJointID dJointCreateAMotor(dWorldID w, dJointGroupID g) { return new_joint(w, g, dJointTypeAMotor); }
void dJointSetAMotorMode(dJointID j, int mode)
{   // dxJointAMotor::setOperationMode amotor.cpp:354-363
    j->t.mmode = mode;
    if (mode == dAMotorEuler) {
        j->t.mnum = 3;
        if (j->body[0]) { std::vector<HostBody> hb = joint_bodies(j, j->t); j->t.reverse = j->reverse; host_amotor_euler_references(hb, j->t); }
    }
   
}
int dJointGetAMotorMode(dJointID j) { return j->t.mmode; }
void dJointSetAMotorNumAxes(dJointID j, int num) { j->t.mnum = j->t.mmode == dAMotorEuler ? 3 : (num < 0 ? 0 : num > 3 ? 3 : num); }
int dJointGetAMotorNumAxes(dJointID j) { return j->t.mnum; }
void dJointSetAMotorAxis(dJointID j, int anum, int rel, Real x, Real y, Real z) { motor_set_axis(j, anum, rel, x, y, z); }
void dJointGetAMotorAxis(dJointID j, int anum, Real *result) { motor_get_axis(j, anum, result); }
int dJointGetAMotorAxisRel(dJointID j, int anum)
{   // getAxisBodyRelativity amotor.cpp:380-390
    int rel = j->t.mrel[anum < 0 ? 0 : anum > 2 ? 2 : anum];
    return (rel != 0 && j->reverse) ? 3 - rel : rel;
}
void dJointSetAMotorAngle(dJointID j, int anum, Real angle) { if (j->t.mmode == dAMotorUser) { j->t.mangle[anum < 0 ? 0 : anum > 2 ? 2 : anum] = angle; } }
Real dJointGetAMotorAngle(dJointID j, int anum) { return j->t.mangle[anum < 0 ? 0 : anum > 2 ? 2 : anum]; }     // user mode: the value that was set
void dJointSetAMotorParam(dJointID j, int parameter, Real value) { limot_set(motor_limot(j, parameter >> 8), parameter & 0xff, value); }
dJointID dJointCreateSlider(dWorldID w, dJointGroupID g) { return new_joint(w, g, dJointTypeSlider); }
void dJointSetSliderAxis(dJointID j, Real x, Real y, Real z)
{   // slider.cpp:249-260
    std::vector<HostBody> hb = joint_bodies(j, j->t);
    host_set_slider_axis(hb, j->t, x, y, z);
}
void dJointGetSliderAxis(dJointID j, Real *result) { joint_get_axis(j, j->t.axis1, result); }
void dJointSetSliderParam(dJointID j, int parameter, Real value) { limot_set(j->t.limot1, parameter, value); }
Real dJointGetSliderPosition(dJointID j)
{   // slider.cpp:46-82 on the host mirror of the body state
    const DJointT &t = j->t;
    const dxBody *b0 = j->body[0], *b1 = j->body[1];
    if (!b0) return 0;
    Real ax1[3], q[3];
    mul0_331(ax1, b0->R, t.axis1);
    if (b1) {
        mul0_331(q, b1->R, t.anchor1);
        for (int i = 0; i < 3; i++) q[i] = b0->pos[i] - q[i] - b1->pos[i];
    } else {
        q[0] = b0->pos[0] - t.anchor1[0]; q[1] = b0->pos[1] - t.anchor1[1]; q[2] = b0->pos[2] - t.anchor1[2];
        if (j->reverse) { ax1[0] = -ax1[0]; ax1[1] = -ax1[1]; ax1[2] = -ax1[2]; }
    }
    return dot3(ax1, q);
}
void dJointSetFixed(dJointID j)
{   // fixed.cpp:113-136
    std::vector<HostBody> hb = joint_bodies(j, j->t);
    host_set_fixed(hb, j->t);
}
void dJointSetFixedParam(dJointID j, int parameter, Real value)
{   // fixed.cpp:138-149
    if (parameter == dParamCFM) j->t.cfm = value; else if (parameter == dParamERP) j->t.erp = value;
}
Real dJointGetFixedParam(dJointID j, int parameter)
{
    return parameter == dParamCFM ? j->t.cfm : parameter == dParamERP ? j->t.erp : 0;
}
void dJointDestroy(dJointID j)
{   // ode.cpp:1301-1321: grouped joints are only destroyed through their group
    if (j->group) return;
    destroy_joint(j);
}
void dJointAttach(dJointID j, dBodyID body1, dBodyID body2)
{   // ode.cpp:1383-1439
    if (body1 && body1 == body2) { classic_error("dJointAttach: can't have body1==body2"); return; }
    if (j->body[0] || j->body[1]) detach_joint(j);
    if (body1 == 0) { body1 = body2; body2 = 0; j->reverse = 1; } else j->reverse = 0;
    j->body[0] = body1; j->body[1] = body2;
    j->t.reverse = j->reverse;
    if (body1) { AdjNode n = { j, body2 }; body1->joints.push_back(n); }
    if (body2) { AdjNode n = { j, body1 }; body2->joints.push_back(n); }
    if (body1 || body2) joint_set_relative_values(j);
    if (j->type != dJointTypeContact) j->world->topo_dirty = true;
}
dBodyID dJointGetBody(dJointID j, int index)
{   // ode.cpp:1489-1499
    if (index == 0 || index == 1) { if (j->reverse) return j->body[1 - index]; return j->body[index]; }
    return 0;
}
dJointType dJointGetType(dJointID j) { return (dJointType)j->type; }
void dJointSetData(dJointID j, void *d) { j->userdata = d; }
void *dJointGetData(dJointID j) { return j->userdata; }
void dJointSetFeedback(dJointID j, dJointFeedback *f) { j->feedback = f; }
dJointFeedback *dJointGetFeedback(dJointID j) { return j->feedback; }

void dJointSetBallAnchor(dJointID j, Real x, Real y, Real z)
{
    std::vector<HostBody> hb = joint_bodies(j, j->t);
    host_set_anchors(hb, j->t, x, y, z);
}
void dJointGetBallAnchor(dJointID j, Real *result)
{
    if (j->reverse) joint_get_anchor2(j, j->t.anchor2, result); else joint_get_anchor(j, j->t.anchor1, result);
}
void dJointSetBallParam(dJointID j, int parameter, Real value)
{   // ball.cpp:112-123
    if (parameter == dParamCFM) j->t.cfm = value; else if (parameter == dParamERP) j->t.erp = value;
}
void dJointSetHingeAnchor(dJointID j, Real x, Real y, Real z)
{
    std::vector<HostBody> hb = joint_bodies(j, j->t);
    host_set_anchors(hb, j->t, x, y, z);
    host_hinge_initial_rotation(hb, j->t);
}
void dJointSetHingeAxis(dJointID j, Real x, Real y, Real z)
{
    std::vector<HostBody> hb = joint_bodies(j, j->t);
    host_set_axes(hb, j->t, x, y, z, j->t.axis1, j->t.axis2);
    host_hinge_initial_rotation(hb, j->t);
}
void dJointGetHingeAnchor(dJointID j, Real *result)
{
    if (j->reverse) joint_get_anchor2(j, j->t.anchor2, result); else joint_get_anchor(j, j->t.anchor1, result);
}
void dJointGetHingeAxis(dJointID j, Real *result) { joint_get_axis(j, j->t.axis1, result); }
void dJointSetHingeParam(dJointID j, int parameter, Real value) { limot_set(j->t.limot1, parameter, value); }
void dJointSetUniversalAnchor(dJointID j, Real x, Real y, Real z)
{
    std::vector<HostBody> hb = joint_bodies(j, j->t);
    host_set_anchors(hb, j->t, x, y, z);
    host_universal_initial_rotations(hb, j->t);
}
void dJointSetUniversalAxis1(dJointID j, Real x, Real y, Real z)
{
    std::vector<HostBody> hb = joint_bodies(j, j->t);
    if (j->reverse) host_set_axes(hb, j->t, x, y, z, 0, j->t.axis2); else host_set_axes(hb, j->t, x, y, z, j->t.axis1, 0);
    host_universal_initial_rotations(hb, j->t);
}
void dJointSetUniversalAxis2(dJointID j, Real x, Real y, Real z)
{
    std::vector<HostBody> hb = joint_bodies(j, j->t);
    if (j->reverse) host_set_axes(hb, j->t, x, y, z, j->t.axis1, 0); else host_set_axes(hb, j->t, x, y, z, 0, j->t.axis2);
    host_universal_initial_rotations(hb, j->t);
}
void dJointGetUniversalAnchor(dJointID j, Real *result)
{
    if (j->reverse) joint_get_anchor2(j, j->t.anchor2, result); else joint_get_anchor(j, j->t.anchor1, result);
}
void dJointSetUniversalParam(dJointID j, int parameter, Real value)
{
    if ((parameter & 0xff00) == 0x100) limot_set(j->t.limot2, parameter & 0xff, value); else limot_set(j->t.limot1, parameter, value);
}
int dAreConnected(dBodyID b1, dBodyID b2)
{   // ode.cpp:1557-1566
    for (size_t i = 0; i < b1->joints.size(); i++) if (b1->joints[i].other == b2) return 1;
    return 0;
}
int dAreConnectedExcluding(dBodyID b1, dBodyID b2, int joint_type)
{   // ode.cpp:1569-1577
    for (size_t i = 0; i < b1->joints.size(); i++) if (b1->joints[i].joint->type != joint_type && b1->joints[i].other == b2) return 1;
    return 0;
}

// ------------------------------------------------------------------------------------------------ spaces and geoms

static dxSpace *new_space(int type)
{
    dxSpace *s = new dxSpace();
    s->type = type; s->cleanup = 1; s->lock_count = 0; s->minlevel = -3; s->maxlevel = 10;
    g_spaces.push_back(s);
    return s;
}
static void space_changed(dxSpace *s)
{
    dxWorld *w = s ? space_world(s) : 0;
    if (w) w->topo_dirty = true;
}
dSpaceID dSimpleSpaceCreate(dSpaceID parent) { if (parent) classic_error("nested spaces are outside the supported subset"); return new_space(ODEB_SPACE_SIMPLE); }
dSpaceID dHashSpaceCreate(dSpaceID parent) { if (parent) classic_error("nested spaces are outside the supported subset"); return new_space(ODEB_SPACE_HASH); }
dSpaceID dSweepAndPruneSpaceCreate(dSpaceID parent, int axisorder)
{
    if (parent) classic_error("nested spaces are outside the supported subset");
    if (axisorder != dSAP_AXES_XYZ) classic_error("dSweepAndPruneSpaceCreate: only dSAP_AXES_XYZ is supported");
    return new_space(ODEB_SPACE_SAP);
}
void dHashSpaceSetLevels(dSpaceID s, int minlevel, int maxlevel) { s->minlevel = minlevel; s->maxlevel = maxlevel; }   // dxHashSpace::setLevels collision_space.cpp:392-397
void dHashSpaceGetLevels(dSpaceID s, int *minlevel, int *maxlevel) { if (minlevel) *minlevel = s->minlevel; if (maxlevel) *maxlevel = s->maxlevel; }
void dSpaceSetCleanup(dSpaceID s, int mode) { s->cleanup = mode != 0; }
int dSpaceGetCleanup(dSpaceID s) { return s->cleanup; }
int dSpaceGetNumGeoms(dSpaceID s) { return (int)s->geoms.size(); }
dGeomID dSpaceGetGeom(dSpaceID s, int i) { return (i >= 0 && i < (int)s->geoms.size()) ? s->geoms[i] : 0; }
void dSpaceAdd(dSpaceID s, dGeomID g)
{
    if (g->space == s) return;
    if (g->space) { classic_error("dSpaceAdd: geom is already in a space"); return; }
    g->space = s; g->index = (int)s->geoms.size(); s->geoms.push_back(g);
    space_changed(s);
    if (g->body) g->body->world->topo_dirty = true;
}
void dSpaceRemove(dSpaceID s, dGeomID g)
{
    if (g->space != s) return;
    space_changed(s);
    if (g->body) g->body->world->topo_dirty = true;
    s->geoms.erase(std::find(s->geoms.begin(), s->geoms.end(), g));
    for (size_t i = 0; i < s->geoms.size(); i++) s->geoms[i]->index = (int)i;
    g->space = 0;
}
void dGeomDestroy(dGeomID g)
{
    if (g->space) dSpaceRemove(g->space, g);
    if (g->body) { std::vector<dxGeom *> &v = g->body->geoms; v.erase(std::find(v.begin(), v.end(), g)); }
    delete g;
}
void dSpaceDestroy(dSpaceID s)
{   // collision_space.cpp:95-113: with cleanup set the space destroys its geoms
    space_changed(s);
    std::vector<dxGeom *> gs = s->geoms;
    for (size_t i = 0; i < gs.size(); i++) { if (s->cleanup) dGeomDestroy(gs[i]); else dSpaceRemove(s, gs[i]); }
    for (size_t i = 0; i < g_spaces.size(); i++) if (g_spaces[i] == s) { g_spaces.erase(g_spaces.begin() + i); break; }
    delete s;
}
static dxGeom *new_geom(dxSpace *s, int type, Real p0, Real p1, Real p2, Real p3)
{
    dxGeom *g = new dxGeom();
    g->type = type; g->space = 0; g->body = 0; g->userdata = 0;
    g->p[0] = p0; g->p[1] = p1; g->p[2] = p2; g->p[3] = p3;
    for (int k = 0; k < 4; k++) g->pos[k] = 0;
    set_identity_R(g->R);
    g->has_ofs = 0; for (int k = 0; k < 4; k++) g->opos[k] = g->fpos[k] = 0;
    set_identity_R(g->oR); set_identity_R(g->fR);
    g->cat = ~0UL; g->col = ~0UL; g->index = -1;
    if (s) dSpaceAdd(s, g);
    return g;
}
dGeomID dCreateSphere(dSpaceID s, Real radius) { return new_geom(s, ODEB_SPHERE, radius, 0, 0, 0); }
dGeomID dCreateBox(dSpaceID s, Real lx, Real ly, Real lz) { return new_geom(s, ODEB_BOX, lx, ly, lz, 0); }
dGeomID dCreateCapsule(dSpaceID s, Real radius, Real length) { return new_geom(s, ODEB_CAPSULE, radius, length, 0, 0); }
dGeomID dCreateCylinder(dSpaceID s, Real radius, Real length) { return new_geom(s, ODEB_CYLINDER, radius, length, 0, 0); }   // cylinder.cpp:52-60
void dGeomCylinderSetParams(dGeomID g, Real radius, Real length) { g->p[0] = radius; g->p[1] = length; }
void dGeomCylinderGetParams(dGeomID g, Real *radius, Real *length) { *radius = g->p[0]; *length = g->p[1]; }
// rays (ray.cpp:46-205): length in p[0], direction = the geom's local z axis
dGeomID dCreateRay(dSpaceID s, Real length) { return new_geom(s, ODEB_RAY, length, 0, 0, 0); }
void dGeomRaySetLength(dGeomID g, Real length) { g->p[0] = length; }
odeb_real dGeomRayGetLength(dGeomID g) { return g->p[0]; }
void dGeomRaySet(dGeomID g, Real px, Real py, Real pz, Real dx, Real dy, Real dz)
{   // ray.cpp:113-135: position + third column of the rotation.  The reference writes through final_posr, i.e. into the body's own
    // pose for a ray that sits on a body without an offset; that aliasing is not reproduced: such a ray follows its body.
    if (g->body) { classic_error("dGeomRaySet: the ray is attached to a body and follows it (use dGeomSetOffset* to aim it)"); return; }
    Real n[3] = { dx, dy, dz };
    normalize3(n);
    g->pos[0] = px; g->pos[1] = py; g->pos[2] = pz;
    g->R[2] = n[0]; g->R[6] = n[1]; g->R[10] = n[2];
}
void dGeomRayGet(dGeomID g, Real *start, Real *dir)
{
    const Real *p = dGeomGetPosition(g), *R = dGeomGetRotation(g);
    start[0] = p[0]; start[1] = p[1]; start[2] = p[2];
    dir[0] = R[2]; dir[1] = R[6]; dir[2] = R[10];
}
dGeomID dCreatePlane(dSpaceID s, Real a, Real b, Real c, Real d)
{
    dxGeom *g = new_geom(s, ODEB_PLANE, a, b, c, d);
    host_normalize_plane(g->p);
    return g;
}
void dGeomSetData(dGeomID g, void *d) { g->userdata = d; }
void *dGeomGetData(dGeomID g) { return g->userdata; }
void dGeomSetBody(dGeomID g, dBodyID b)
{   // collision_kernel.cpp:497-533
    if (g->body == b) return;
    if (g->body) { std::vector<dxGeom *> &v = g->body->geoms; v.erase(std::find(v.begin(), v.end(), g)); g->body->world->topo_dirty = true; }
    g->body = b;
    if (b) { b->geoms.push_back(g); b->world->topo_dirty = true; }
    space_changed(g->space);
}
dBodyID dGeomGetBody(dGeomID g) { return g->body; }
void dGeomSetPosition(dGeomID g, Real x, Real y, Real z)
{   // collision_kernel.cpp:542-557: with a body attached this moves the body
    if (g->body) dBodySetPosition(g->body, x, y, z); else { g->pos[0] = x; g->pos[1] = y; g->pos[2] = z; }
}
void dGeomSetRotation(dGeomID g, const Real *R) { if (g->body) dBodySetRotation(g->body, R); else memcpy(g->R, R, 12 * sizeof(Real)); }
void dGeomSetQuaternion(dGeomID g, const Real *q) { if (g->body) dBodySetQuaternion(g->body, q); else r_from_q(g->R, q); }
static void geom_final_pose(dxGeom *g)
{   // dxGeom::computePosr collision_kernel.cpp:455-466
    mul0_331(g->fpos, g->body->R, g->opos);
    g->fpos[0] += g->body->pos[0]; g->fpos[1] += g->body->pos[1]; g->fpos[2] += g->body->pos[2];
    mul0_333(g->fR, g->body->R, g->oR);
}
const odeb_real *dGeomGetPosition(dGeomID g)
{
    if (g->body && g->has_ofs) { geom_final_pose(g); return g->fpos; }
    return g->body ? g->body->pos : g->pos;
}
const odeb_real *dGeomGetRotation(dGeomID g)
{
    if (g->body && g->has_ofs) { geom_final_pose(g); return g->fR; }
    return g->body ? g->body->R : g->R;
}
// geom offsets relative to the body (collision_kernel.cpp:1000-1130); a geom without a body cannot have one
void dGeomSetOffsetPosition(dGeomID g, Real x, Real y, Real z) { if (!g->body) return; g->has_ofs = 1; g->opos[0] = x; g->opos[1] = y; g->opos[2] = z; }
void dGeomSetOffsetRotation(dGeomID g, const Real *R) { if (!g->body) return; g->has_ofs = 1; memcpy(g->oR, R, 12 * sizeof(Real)); }
void dGeomSetOffsetQuaternion(dGeomID g, const Real *q) { if (!g->body) return; g->has_ofs = 1; r_from_q(g->oR, q); }
void dGeomClearOffset(dGeomID g) { g->has_ofs = 0; g->opos[0] = g->opos[1] = g->opos[2] = 0; set_identity_R(g->oR); }
int dGeomIsOffset(dGeomID g) { return g->has_ofs; }
const odeb_real *dGeomGetOffsetPosition(dGeomID g) { return g->opos; }
const odeb_real *dGeomGetOffsetRotation(dGeomID g) { return g->oR; }
int dGeomGetClass(dGeomID g) { return g->type; }
void dGeomSetCategoryBits(dGeomID g, unsigned long bits) { g->cat = bits; }
void dGeomSetCollideBits(dGeomID g, unsigned long bits) { g->col = bits; }
unsigned long dGeomGetCategoryBits(dGeomID g) { return g->cat; }
unsigned long dGeomGetCollideBits(dGeomID g) { return g->col; }
odeb_real dGeomSphereGetRadius(dGeomID g) { return g->p[0]; }
void dGeomBoxGetLengths(dGeomID g, Real *r) { r[0] = g->p[0]; r[1] = g->p[1]; r[2] = g->p[2]; }
void dGeomCapsuleGetParams(dGeomID g, Real *radius, Real *length) { *radius = g->p[0]; *length = g->p[1]; }
void dGeomPlaneGetParams(dGeomID g, Real *r) { for (int k = 0; k < 4; k++) r[k] = g->p[k]; }

// context of a geom's space: world of its bodies, or a body-less private world when the space holds only static geoms
static ClassicCtx *ctx_for_space(dxSpace *s, dxWorld **wout)
{
    dxWorld *w = space_world(s);
    if (!w) { classic_error("dSpaceCollide: the space holds no geom attached to a body"); return 0; }
    ClassicCtx *c = ctx_get(w, s, 0);
    *wout = w;
    return c;
}

void dGeomGetAABB(dGeomID g, Real aabb[6])
{   // computed by k_aabb on the device, like every AABB of the path
    for (int k = 0; k < 6; k++) aabb[k] = 0;
    if (!g->space) { classic_error("dGeomGetAABB: geom is not in a space"); return; }
    dxWorld *w = 0;
    ClassicCtx *c = ctx_for_space(g->space, &w);
    if (!c || !ctx_upload_state(w, c)) return;
    OdebBatch *B = c->B;
    k_aabb<<<nblk(B->P.NG, 128), 128, 0, B->stream>>>(B->P, B->D);
    B->launches++;
    cudaStreamSynchronize(B->stream);
    cudaMemcpy(aabb, B->D.aabb + 6 * (size_t)g->index, 6 * sizeof(Real), cudaMemcpyDeviceToHost);
}

static int ensure_req_buffers(ClassicCtx *c, int n)
{
    if (n <= c->req_cap) return 1;
    if (c->d_req) cudaFree(c->d_req); if (c->d_out) cudaFree(c->d_out); if (c->d_cnt) cudaFree(c->d_cnt);
    c->d_req = 0; c->d_out = 0; c->d_cnt = 0; c->req_cap = 0;
    int cap = n < 256 ? 256 : 2 * n;
    CK(cudaMalloc((void **)&c->d_req, cap * sizeof(CollideReq)));
    CK(cudaMalloc((void **)&c->d_out, (size_t)cap * 8 * 7 * sizeof(Real)));
    CK(cudaMalloc((void **)&c->d_cnt, cap * sizeof(int)));
    c->req_cap = cap;
    return 1;
}

// run the narrowphase kernel over a request list; results: cnt[n], out[n*maxc*7]
static int run_collide_requests(ClassicCtx *c, const std::vector<CollideReq> &req, int maxc, std::vector<int> &cnt, std::vector<Real> &out)
{
    const int n = (int)req.size();
    cnt.assign(n, 0); out.assign((size_t)n * maxc * 7, 0);
    if (n == 0) return 1;
    OdebBatch *B = c->B;
    if (!ensure_req_buffers(c, n)) return 0;
    CK(cudaMemcpyAsync(c->d_req, req.data(), n * sizeof(CollideReq), cudaMemcpyHostToDevice, B->stream));
    k_collide_req<<<nblk(n, 64), 64, 0, B->stream>>>(B->P, B->D, n, c->d_req, c->d_out, c->d_cnt, maxc);
    B->launches++;
    CK(cudaMemcpyAsync(cnt.data(), c->d_cnt, n * sizeof(int), cudaMemcpyDeviceToHost, B->stream));
    CK(cudaMemcpyAsync(out.data(), c->d_out, (size_t)n * maxc * 7 * sizeof(Real), cudaMemcpyDeviceToHost, B->stream));
    CK(cudaStreamSynchronize(B->stream));
    return 1;
}

void dSpaceCollide(dSpaceID s, void *data, dNearCallback *callback)
{   // collision_space.cpp:779-784 -> dxHashSpace::collide :421-614 / dxSAPSpace::collide collision_sapspace.cpp:428-496
    if (s->geoms.empty()) return;
    dxWorld *w = 0;
    ClassicCtx *c = ctx_for_space(s, &w);
    if (!c) return;
    for (int attempt = 0; attempt < 8; attempt++) {
        if (!ctx_upload_state(w, c)) { classic_error("dSpaceCollide: %s", g_err.c_str()); return; }
        OdebBatch *B = c->B;
        OdebWorldParams wp = w->wp; wp.space_type = s->type; wp.max_contacts = 8;
        apply_world_params(B->P, &wp, true);
        B->P.space_type = s->type; B->P.hash_minlevel = s->minlevel; B->P.hash_maxlevel = s->maxlevel;
        launch_collide(B, B->stream, false);
        int np = 0, ov = 0;
        cudaStreamSynchronize(B->stream);
        cudaMemcpy(&ov, B->D.overflow, sizeof(int), cudaMemcpyDeviceToHost);
        if (ov) {   // pair capacity exceeded: grow and redo
            int z = 0; cudaMemcpy(B->D.overflow, &z, sizeof(int), cudaMemcpyHostToDevice);
            c->cap_pairs *= 4; w->topo_dirty = true;
            c = ctx_get(w, s, 0);
            if (!c) return;
            continue;
        }
        cudaMemcpy(&np, B->D.npairs, sizeof(int), cudaMemcpyDeviceToHost);
        c->pairs.resize(np);
        if (np) cudaMemcpy(c->pairs.data(), B->D.pairs, np * sizeof(int2), cudaMemcpyDeviceToHost);
        break;
    }
    c->in_pass = true; c->cache_flags = -1;
    s->lock_count++;
    // the pair stream is handed out in canonical (geomA < geomB, lexicographic) order; every pair of the reference's
    // callback stream for this space type is present exactly once
    std::vector<int2> pairs = c->pairs;
    for (size_t i = 0; i < pairs.size(); i++) callback(data, s->geoms[pairs[i].x], s->geoms[pairs[i].y]);
    s->lock_count--;
    if (w->ctx == c) c->in_pass = false;
}

int dCollide(dGeomID o1, dGeomID o2, int flags, dContactGeom *contact, int skip)
{   // collision_kernel.cpp:292-338
    const int maxc = flags & 0xffff;
    if (maxc < 1 || !contact || skip < (int)sizeof(dContactGeom)) { classic_error("dCollide: bad arguments"); return 0; }
    if (o1 == o2) return 0;
    if (o1->body == o2->body && o1->body) return 0;
    dxSpace *s = o1->space;
    if (!s || o2->space != s) { classic_error("dCollide: both geoms must be in the same space"); return 0; }
    dxWorld *w = space_world(s);
    ClassicCtx *c = w ? w->ctx : 0;
    const int mc = maxc > 8 ? 8 : maxc;
    const Real *res = 0; int n = 0;
    std::vector<int> cnt1; std::vector<Real> out1;
    if (c && c->in_pass && c->space == s && !w->topo_dirty) {
        if (c->cache_flags == -1) {   // first dCollide of this collide pass: narrowphase of every broadphase pair in one launch
            std::vector<CollideReq> req(c->pairs.size());
            for (size_t i = 0; i < req.size(); i++) { req[i].g1 = c->pairs[i].x; req[i].g2 = c->pairs[i].y; req[i].flags = flags; }
            if (!run_collide_requests(c, req, mc, c->pc_count, c->pc_geom)) { classic_error("dCollide: %s", g_err.c_str()); return 0; }
            c->cache_flags = flags; c->cache_maxc = mc;
        }
        if (c->cache_flags == flags && o1->index < o2->index) {
            int2 key = make_int2(o1->index, o2->index);
            std::vector<int2>::const_iterator it = std::lower_bound(c->pairs.begin(), c->pairs.end(), key,
                [](const int2 &a, const int2 &b) { return a.x < b.x || (a.x == b.x && a.y < b.y); });
            if (it != c->pairs.end() && it->x == key.x && it->y == key.y) {
                size_t k = it - c->pairs.begin();
                n = c->pc_count[k]; res = &c->pc_geom[k * c->cache_maxc * 7];
            }
        }
    }
    if (!res) {   // any other use: one request
        if (!w) { classic_error("dCollide: no geom of the space is attached to a body"); return 0; }
        bool was_pass = c && c->in_pass;
        c = ctx_get(w, s, 0);
        if (!c) return 0;
        if (!was_pass && !ctx_upload_state(w, c)) { classic_error("dCollide: %s", g_err.c_str()); return 0; }
        std::vector<CollideReq> req(1);
        req[0].g1 = o1->index; req[0].g2 = o2->index; req[0].flags = flags;
        if (!run_collide_requests(c, req, mc, cnt1, out1)) { classic_error("dCollide: %s", g_err.c_str()); return 0; }
        n = cnt1[0]; res = out1.data();
    }
    if (n > mc) n = mc;
    for (int i = 0; i < n; i++) {
        dContactGeom *cg = (dContactGeom *)((char *)contact + (size_t)i * skip);
        cg->pos[0] = res[7 * i]; cg->pos[1] = res[7 * i + 1]; cg->pos[2] = res[7 * i + 2]; cg->pos[3] = 0;
        cg->normal[0] = res[7 * i + 3]; cg->normal[1] = res[7 * i + 4]; cg->normal[2] = res[7 * i + 5]; cg->normal[3] = 0;
        cg->depth = res[7 * i + 6];
        cg->g1 = o1; cg->g2 = o2; cg->side1 = -1; cg->side2 = -1;
    }
    return n;
}

// ------------------------------------------------------------------------------------------------ the step

static void surface_to_device(const dContact &ct, DSurface &S)
{   // dxJointContact::getInfo1 contact.cpp:48-122 (row count, negative mu clamped to 0)
    const dSurfaceParameters &p = ct.surface;
    memset(&S, 0, sizeof(S));
    S.mode = p.mode;
    S.mu = p.mu < 0 ? 0 : p.mu;
    S.mu2 = p.mu2 < 0 ? 0 : p.mu2;
    S.bounce = p.bounce; S.bounce_vel = p.bounce_vel; S.soft_erp = p.soft_erp; S.soft_cfm = p.soft_cfm;
    S.motion1 = p.motion1; S.motion2 = p.motion2; S.motionN = p.motionN; S.slip1 = p.slip1; S.slip2 = p.slip2;
    S.fdir1[0] = ct.fdir1[0]; S.fdir1[1] = ct.fdir1[1]; S.fdir1[2] = ct.fdir1[2];
    S.the_m = odeb_contact_rows(S.mode, S.mu, S.mu2, p.rho, p.rho2, p.rhoN);
    S.rho = p.rho < 0 ? 0 : p.rho; S.rho2 = p.rho2 < 0 ? 0 : p.rho2; S.rhoN = p.rhoN < 0 ? 0 : p.rhoN;
}

int dWorldQuickStep(dWorldID w, Real stepsize)
{   // ode.cpp:1847-1864
    if (!(stepsize > 0)) { classic_error("dWorldQuickStep: stepsize must be > 0"); return 0; }
    if (w->bodies.empty()) return 1;
    // contact joints of this step, in creation order; permanent joints keep the device ids of the context
    std::vector<dxJoint *> contacts;
    for (size_t i = 0; i < w->joints.size(); i++) {
        dxJoint *j = w->joints[i];
        if (j->type == dJointTypeContact && j->body[0]) contacts.push_back(j);
    }
    const int nc = (int)contacts.size();
    ClassicCtx *c = ctx_get(w, 0, nc);
    if (!c) return 0;
    OdebBatch *B = c->B;
    if (cudaSetDevice(B->device) != cudaSuccess) return 0;
    if (!ctx_upload_state(w, c)) return 0;
    {
        OdebWorldParams wp = w->wp; wp.space_type = c->space ? c->space->type : ODEB_SPACE_HASH; wp.max_contacts = 8;
        int adis_samples = B->P.adis_samples;
        apply_world_params(B->P, &wp, true);
        B->P.adis_samples = adis_samples;
        B->P.h = stepsize; B->P.hrecip = rrecip(stepsize);
    }
    const DevPtrs &D = B->D;
    const int nb = c->nb, NJ = c->nj;
    // permanent joints: current parameters
    std::map<dxJoint *, int> jid;
    {
        std::vector<DJointT> jt(NJ);
        for (int i = 0; i < NJ; i++) {
            dxJoint *j = c->perm[i];
            jid[j] = i;
            jt[i] = j->t;
            jt[i].type = j->type; jt[i].reverse = j->reverse;
            jt[i].b0 = j->body[0] ? j->body[0]->index : -1; jt[i].b1 = j->body[1] ? j->body[1]->index : -1;
            if (jt[i].b0 < 0) { jt[i].b0 = 0; jt[i].type = 0; }     // joint in limbo: never reached through a body list
        }
        if (NJ) CK(cudaMemcpy(D.joints, jt.data(), NJ * sizeof(DJointT), cudaMemcpyHostToDevice));
    }
    // contacts
    {
        std::vector<int4> ci(nc); std::vector<Real4> cg(2 * (size_t)nc); std::vector<DSurface> cs(nc);
        for (int k = 0; k < nc; k++) {
            dxJoint *j = contacts[k];
            jid[j] = NJ + k;
            const dContactGeom &g = j->contact.geom;
            ci[k] = make_int4(k, j->body[0]->index, j->body[1] ? j->body[1]->index : -1, j->reverse);
            Real4 a = { g.pos[0], g.pos[1], g.pos[2], g.depth }, n4 = { g.normal[0], g.normal[1], g.normal[2], 0 };
            cg[2 * k] = a; cg[2 * k + 1] = n4;
            surface_to_device(j->contact, cs[k]);
        }
        if (nc) {
            CK(cudaMemcpy(D.cinfo, ci.data(), nc * sizeof(int4), cudaMemcpyHostToDevice));
            CK(cudaMemcpy(D.cgeom, cg.data(), cg.size() * sizeof(Real4), cudaMemcpyHostToDevice));
            CK(cudaMemcpy(D.csurf, cs.data(), nc * sizeof(DSurface), cudaMemcpyHostToDevice));
        }
        CK(cudaMemcpy(D.ncontacts, &nc, sizeof(int), cudaMemcpyHostToDevice));
    }
    // per-body joint lists in dJointAttach order (the island DFS walks them newest first)
    {
        std::vector<int> sofs(nb + 1, 0), sj, so;
        for (int b = 0; b < nb; b++) {
            sofs[b] = (int)sj.size();
            const dxBody *body = w->bodies[b];
            for (size_t k = 0; k < body->joints.size(); k++) {
                std::map<dxJoint *, int>::const_iterator it = jid.find(body->joints[k].joint);
                if (it == jid.end()) continue;
                sj.push_back(it->second);
                so.push_back(body->joints[k].other ? body->joints[k].other->index : -1);
            }
        }
        sofs[nb] = (int)sj.size();
        CK(cudaMemcpy(D.sadj_ofs, sofs.data(), sofs.size() * sizeof(int), cudaMemcpyHostToDevice));
        if (!sj.empty()) {
            CK(cudaMemcpy(D.sadj_joint, sj.data(), sj.size() * sizeof(int), cudaMemcpyHostToDevice));
            CK(cudaMemcpy(D.sadj_other, so.data(), so.size() * sizeof(int), cudaMemcpyHostToDevice));
        }
    }
    // dJointSetFeedback: any joint with a dJointFeedback attached turns the device-side feedback pass on for this context
    bool want_fb = false;
    for (std::map<dxJoint *, int>::const_iterator it = jid.begin(); it != jid.end(); ++it) if (it->first->feedback) { want_fb = true; break; }
    if (want_fb != (B->D.jcopy != 0) && !odeb_enable_feedback(B, want_fb ? 1 : 0)) return 0;
    unsigned seed = (unsigned)g_seed, st0[4] = { 0, 0, 0, 0 }, st1[4];
    CK(cudaMemcpy(D.seed, &seed, sizeof(unsigned), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(D.stats, st0, sizeof(st0), cudaMemcpyHostToDevice));
    launch_dynamics(B, B->stream, false, choose_solver(B));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(B->stream));
    {
        int ovh[4] = { 0, 0, 0, 0 };  // capacity overflow flag; largest island (rows, bodies) and island count of this step (pick the solver kernel of the next one)
        CK(cudaMemcpy(ovh, D.overflow, sizeof(ovh), cudaMemcpyDeviceToHost));
        const int ov = ovh[0];
        if (ovh[1] > 0) { B->hint_m = ovh[1]; B->hint_nb = ovh[2]; B->hint_nis = ovh[3]; cudaMemset(D.overflow + 1, 0, 3 * sizeof(int)); }
        if (ov) { int z = 0; cudaMemcpy(D.overflow, &z, sizeof(int), cudaMemcpyHostToDevice); classic_error("dWorldQuickStep: device capacity overflow (%d)", ov); return 0; }
    }
    CK(cudaMemcpy(&seed, D.seed, sizeof(unsigned), cudaMemcpyDeviceToHost));
    g_seed = seed;
    CK(cudaMemcpy(st1, D.stats, sizeof(st1), cudaMemcpyDeviceToHost));
    if (w->stats_sink) {
        w->stats_sink->iteration_count += st1[0]; w->stats_sink->premature_exits += st1[1];
        w->stats_sink->prolonged_execs += st1[2]; w->stats_sink->full_extra_execs += st1[3];
    }
    if (want_fb) {   // quickstep.cpp:3108-3182: f1/t1 always, f2/t2 only when the joint has a second body; untouched when not stepped
        const int n = NJ + nc;
        std::vector<Real4> v(4 * (size_t)n);
        if (n) CK(cudaMemcpy(v.data(), B->D.jfb, v.size() * sizeof(Real4), cudaMemcpyDeviceToHost));
        for (std::map<dxJoint *, int>::const_iterator it = jid.begin(); it != jid.end(); ++it) {
            dJointFeedback *fb = it->first->feedback;
            if (!fb || it->second >= n) continue;
            const Real4 *q = &v[4 * (size_t)it->second];
            const int state = (int)q[0].w;
            if (state >= 1) { fb->f1[0] = q[0].x; fb->f1[1] = q[0].y; fb->f1[2] = q[0].z; fb->t1[0] = q[1].x; fb->t1[1] = q[1].y; fb->t1[2] = q[1].z; }
            if (state >= 2) { fb->f2[0] = q[2].x; fb->f2[1] = q[2].y; fb->f2[2] = q[2].z; fb->t2[0] = q[3].x; fb->t2[1] = q[3].y; fb->t2[2] = q[3].z; }
        }
    }
    if (!ctx_download_state(w, c)) return 0;
    return 1;
}

} // extern "C"
