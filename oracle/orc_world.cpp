// orc_world.cpp -- TEST INFRASTRUCTURE ONLY (never linked or loaded by ode_b200/).
//
// CPU restatement ("oracle") of the reference's per-step world update for the scoped feature set
// (sphere/box/capsule/plane, contact/ball/hinge/universal joints, QuickStep), in plain scalar C++ with
// the reference's operation order.  Built with -ffp-contract=off for both precisions
// (liborc_single.so / liborc_double.so, see oracle/Makefile) and pinned bit-exactly against the
// compiled reference (oracle/_ref) by tests/test_oracle_vs_ref.py and against committed golden
// vectors (tests/golden/) generated from the reference by oracle/make_golden.py.
//
// Same C interface as include/ode_b200.h with the prefix orc_.
// Pair order fed to the contact policy is canonical (geom index of o1 < o2, lexicographic) -- the
// same order oracle/ref_driver.cpp imposes on the reference's callback stream.
#include <vector>
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include "../include/ode_b200.h"
#include "orc_collide.h"

namespace {

// body flag values of the reference (ode/src/objects.h:49-57)
enum { BF_FINITE_ROT = 1, BF_FINITE_ROT_AXIS = 2, BF_DISABLED = 4, BF_NO_GRAVITY = 8, BF_AUTO_DISABLE = 16,
       BF_LIN_DAMP = 32, BF_ANG_DAMP = 64, BF_MAX_ANG_SPEED = 128, BF_GYRO = 256 };

struct Adj { int joint; int other; };   // one dxJointNode: (joint, body at the other end or -1)

struct Body {
    Real pos[4], R[12], q[4], lvel[4], avel[4], facc[4], tacc[4];
    Real mass, I[12], invMass, invI[12];
    unsigned flags;
    // auto-disable state (ode/src/objects.h:240-252)
    int adis_stepsleft; Real adis_timeleft;
    std::vector<Real> avg_l, avg_a; unsigned avg_counter; int avg_ready;
    int tag;
    std::vector<Adj> adj;                // attach order; traversed newest-first (ode.cpp:1417-1431)
};

struct Limot {                           // dxJointLimitMotor (joints/joint.h:291-320, joint.cpp:494-512)
    Real vel, fmax, lostop, histop, fudge_factor, normal_cfm, stop_erp, stop_cfm, bounce;
    int limit; Real limit_err;
};

struct Joint {
    int type;                            // ODEB_JOINT_*
    int b0, b1;                          // node[0].body, node[1].body (-1 = none) after the swap of dJointAttach
    int reverse;                         // dJOINT_REVERSE
    int tag;
    // contact
    OrcContactGeom cg; int the_m;
    // ball / hinge / universal
    Real anchor1[4], anchor2[4], axis1[4], axis2[4], qrel[4], qrel1[4], qrel2[4];
    Real erp, cfm;
    Limot limot1, limot2;
    // lmotor / amotor (lmotor.h:30-49, amotor.h:91-101): axes, what they are relative to, Euler reference vectors, angles, third limit motor
    int mnum, mmode, mrel[3];
    Real maxis[3][4], mref[2][4], mangle[3];
    Limot limot3;
    int m, nub;
};

struct World {
    std::vector<Body> bodies;
    std::vector<OrcGeom> geoms;
    std::vector<Real> geom_static_pos;   // identity pose storage for geoms without a body
    std::vector<Joint> pjoints;          // permanent joints (creation order)
    std::vector<Joint> contacts;         // this step's contact joints (creation order)
    unsigned seed;
    unsigned stats[4];
    std::vector<int> pairs;
    std::vector<int> contact_g;
    std::vector<OrcContactGeom> last_cg;  // contact geoms of the last step (kept after dJointGroupEmpty)
    std::vector<OrcContactGeom> ray_cg; std::vector<int> ray_g;   // ray hits of the last collide pass (sensor results: no joints)
    std::vector<int> island_label; int island_count;
    unsigned long long sweeps;
    std::vector<int> isl_log;            // per island of the last step: bodies, rows, sweeps executed (test / profiling read-out)
    unsigned step_seed; int cur_island; unsigned long long draws;   // canonical mode bookkeeping
    // joint feedback of the last step (dJointSetFeedback, quickstep.cpp:3108-3182): per joint id 12 reals f1 t1 f2 t2 and a
    // state (0 = joint not stepped, 1 = body 1 only, 2 = both bodies)
    std::vector<Real> fb; std::vector<int> fb_state;
};

struct Batch {
    OdebWorldParams wp;
    Real gravity[3], erp, cfm, sor_w, premature_delta, extra_delta, extra_factor;
    unsigned num_iter, max_extra;
    bool dyn_enabled;
    Real max_vel, min_depth;
    Real adis_lin, adis_ang, adis_time; int adis_steps; unsigned adis_samples;
    Real damp_lin_scale, damp_ang_scale, damp_lin_thr, damp_ang_thr, max_ang_speed;
    int nbody, ngeom, njoint;
    int canonical;                       // 1: canonical-order mode of the large-world path (see orc_set_solver_mode)
    int feedback;                        // 1: every joint has a dJointFeedback attached
    std::vector<World> worlds;
};

static const Real g_identity[12] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0 };
static const Real g_zero4[4] = { 0, 0, 0, 0 };

Joint &joint_ref(World &W, int id) { return id < (int)W.pjoints.size() ? W.pjoints[id] : W.contacts[id - W.pjoints.size()]; }

// dxFactorCholesky / dxSolveCholesky / dxInvertPDMatrix for n=3, nskip=4 (ode/src/matrix.cpp:107-253)
int invert_pd3(const Real *A, Real *Ainv)
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

void body_set_quat(Body &b, const Real *q)
{
    b.q[0] = q[0]; b.q[1] = q[1]; b.q[2] = q[2]; b.q[3] = q[3];
    normalize4(b.q);            // dBodySetQuaternion ode.cpp:379-392
    r_from_q(b.R, b.q);
}

// ---------------------------------------------------------------------------------------------
// joint set-up (anchors / axes bound at the template pose)

// setAnchors joints/joint.cpp:289-320
void set_anchors(World &W, Joint &j, Real x, Real y, Real z)
{
    if (j.b0 >= 0) {
        Body &b0 = W.bodies[j.b0];
        Real q[3] = { x - b0.pos[0], y - b0.pos[1], z - b0.pos[2] };
        mul1_331(j.anchor1, b0.R, q);
        if (j.b1 >= 0) {
            Body &b1 = W.bodies[j.b1];
            Real q2[3] = { x - b1.pos[0], y - b1.pos[1], z - b1.pos[2] };
            mul1_331(j.anchor2, b1.R, q2);
        } else { j.anchor2[0] = x; j.anchor2[1] = y; j.anchor2[2] = z; }
    }
    j.anchor1[3] = 0; j.anchor2[3] = 0;
}

// setAxes joints/joint.cpp:325-369
void set_axes(World &W, Joint &j, Real x, Real y, Real z, Real *axis1, Real *axis2)
{
    if (j.b0 >= 0) {
        Real q[3] = { x, y, z };
        normalize3(q);
        if (axis1) { mul1_331(axis1, W.bodies[j.b0].R, q); axis1[3] = 0; }
        if (axis2) {
            if (j.b1 >= 0) mul1_331(axis2, W.bodies[j.b1].R, q);
            else { axis2[0] = x; axis2[1] = y; axis2[2] = z; }
            axis2[3] = 0;
        }
    }
}

void limot_init(const Batch &B, Limot &l)
{   // dxJointLimitMotor::init joints/joint.cpp:494-512
    l.vel = 0; l.fmax = 0; l.lostop = -R_INF; l.histop = R_INF; l.fudge_factor = 1;
    l.normal_cfm = B.cfm; l.stop_erp = B.erp; l.stop_cfm = B.cfm; l.bounce = 0; l.limit = 0; l.limit_err = 0;
}

// dxJointLimitMotor::set joints/joint.cpp:515-571 for the parameters a scene can give
void limot_set(Limot &l, const OdebJointDesc &d, int a)
{
    Real lo = (Real)d.lo_stop[a], hi = (Real)d.hi_stop[a];
    l.lostop = lo; l.histop = hi; l.lostop = lo;
    l.vel = (Real)d.vel[a];
    if ((Real)d.fmax[a] >= 0) l.fmax = (Real)d.fmax[a];
    if (d.fudge_factor[a] >= 0 && (Real)d.fudge_factor[a] <= 1) l.fudge_factor = (Real)d.fudge_factor[a];
    if (d.bounce[a] >= 0) l.bounce = (Real)d.bounce[a];
    if (d.stop_erp[a] >= 0) l.stop_erp = (Real)d.stop_erp[a];
    if (d.stop_cfm[a] >= 0) l.stop_cfm = (Real)d.stop_cfm[a];
}

#include "orc_joints.h"

// ---------------------------------------------------------------------------------------------
// collision: dSpaceCollide pair set + the contact policy (near-callback)

inline bool aabb_overlap(const Real *b1, const Real *b2)
{   // collideAABBs collision_space_internal.h:60-67
    return !(b1[0] > b2[1] || b1[1] < b2[0] || b1[2] > b2[3] || b1[3] < b2[2] || b1[4] > b2[5] || b1[5] < b2[4]);
}

inline bool pair_filter(const OrcGeom &g1, const OrcGeom &g2)
{   // collision_space_internal.h:50-57 / collision_sapspace.cpp:238-246
    if (g1.body == g2.body && g1.body >= 0) return false;
    if (((g1.cat & g2.col) || (g2.cat & g1.col)) == 0) return false;
    return true;
}

// dxHashSpace::collide collision_space.cpp:421-614, the part that is not a plain AABB test: whether the cell walk brings two AABBs together.
// Each AABB gets level = exponent of frexp(largest extent) (findLevel :329-349) clamped up to minlevel, and integer cell bounds
// floor(aabb / 2^level) (:448-462); levels above maxlevel (planes: MAXINT) go to the big list, which is tested against everything
// (:590-607).  An AABB is inserted in the cells of its own level and probes its own and every higher level with bounds >>= 1 (:521-582).
// Per (x,y) column the hash index starts at (level*1000UL + x*100UL + y*10UL + zbegin) % sz with level/x/y converted through unsigned int
// and zbegin an int added to an unsigned long (:358-361, :499, :533), and is then incremented per z.  For zbegin < 0 and a base smaller
// than -zbegin the sum wraps around 2^64, and since sz (a prime >= 13) does not divide 2^64 the cells of that column land 2^64 % sz slots
// away from where an AABB with a non-wrapping zbegin looks for them.  Two overlapping AABBs therefore meet iff some shared cell column
// is addressed with the same wrap state from both sides.
static int hash_level(const Real *a, int minlevel)
{
    if (a[0] <= -R_INF || a[1] >= R_INF || a[2] <= -R_INF || a[3] >= R_INF || a[4] <= -R_INF || a[5] >= R_INF) return 0x7fffffff;
    Real q = a[1] - a[0], q2 = a[3] - a[2];
    if (q2 > q) q = q2;
    q2 = a[5] - a[4];
    if (q2 > q) q = q2;
    int level;
    frexp(q, &level);
    return level < minlevel ? minlevel : level;
}
static bool hash_column_wraps(int level, int x, int y, int zbegin)
{
    unsigned long base = (unsigned int)level * 1000UL + (unsigned int)x * 100UL + (unsigned int)y * 10UL;
    return zbegin < 0 && base < (unsigned long)(-(long)zbegin);
}
bool hash_space_meets(const Real *a, const Real *b, int minlevel, int maxlevel)
{
    int la = hash_level(a, minlevel), lb = hash_level(b, minlevel);
    if (la > maxlevel || lb > maxlevel) return true;
    if (la > lb) { const Real *t = a; a = b; b = t; int tl = la; la = lb; lb = tl; }
    int da[6], db[6];
    const Real ra = (Real)1 / (Real)ldexp(1.0, la), rb = (Real)1 / (Real)ldexp(1.0, lb);
    for (int i = 0; i < 6; i++) { da[i] = (int)floor(a[i] * ra); db[i] = (int)floor(b[i] * rb); }
    for (int l = la; l < lb; l++) for (int i = 0; i < 6; i++) da[i] >>= 1;
    if (da[4] >= 0 && db[4] >= 0) return true;                  // nothing wraps: overlapping AABBs always share a cell
    const int x0 = da[0] > db[0] ? da[0] : db[0], x1 = da[1] < db[1] ? da[1] : db[1];
    const int y0 = da[2] > db[2] ? da[2] : db[2], y1 = da[3] < db[3] ? da[3] : db[3];
    const int z0 = da[4] > db[4] ? da[4] : db[4], z1 = da[5] < db[5] ? da[5] : db[5];
    if (z0 > z1) return false;
    for (int x = x0; x <= x1; x++) for (int y = y0; y <= y1; y++)
        if (hash_column_wraps(lb, x, y, da[4]) == hash_column_wraps(lb, x, y, db[4])) return true;
    return false;
}

void find_pairs(const Batch &B, World &W)
{
    int ng = (int)W.geoms.size();
    for (int i = 0; i < ng; i++) orc_compute_aabb(W.geoms[i]);
    W.pairs.clear();
    if (B.wp.space_type == ODEB_SPACE_SAP) {
        // dxSAPSpace::collide collision_sapspace.cpp:428-496 + BoxPruning :521-582, as a SET:
        // geoms whose axis-0 max is +inf are "infinite": tested pairwise among themselves with the
        // AABB test, and paired with EVERY normal geom without an AABB test (:478-493); normal
        // geoms pair when float-cast axis-0 intervals overlap (min of the later <= max of the earlier
        // in float-sorted order) and axes 1,2 overlap inclusively in dReal (:545-573).
        std::vector<char> inf(ng);
        for (int i = 0; i < ng; i++) inf[i] = (W.geoms[i].aabb[1] == R_INF);
        for (int i = 0; i < ng; i++) for (int j = i + 1; j < ng; j++) {
            const OrcGeom &g1 = W.geoms[i], &g2 = W.geoms[j];
            if (!pair_filter(g1, g2)) continue;
            bool hit;
            if (inf[i] && inf[j]) hit = aabb_overlap(g1.aabb, g2.aabb);
            else if (inf[i] || inf[j]) hit = true;
            else {
                float min1 = (float)g1.aabb[0], max1 = (float)g1.aabb[1], min2 = (float)g2.aabb[0], max2 = (float)g2.aabb[1];
                // sorted by float min; the earlier one's max must reach the later one's min
                bool ax0 = (min1 <= min2) ? (min2 <= max1) : (min1 <= max2);
                hit = ax0 && !(g1.aabb[3] < g2.aabb[2] || g2.aabb[3] < g1.aabb[2]) && !(g1.aabb[5] < g2.aabb[4] || g2.aabb[5] < g1.aabb[4]);
            }
            if (hit) { W.pairs.push_back(i); W.pairs.push_back(j); }
        }
    } else {
        // dxSimpleSpace::collide collision_space.cpp:245-267 as a SET == all pairs passing collideAABBs;
        // dxHashSpace::collide :421-614 == those of them that its cell walk brings together
        const bool hash = (B.wp.space_type == ODEB_SPACE_HASH);
        const int minl = B.wp.hash_levels_set ? B.wp.hash_minlevel : -3, maxl = B.wp.hash_levels_set ? B.wp.hash_maxlevel : 10;
        for (int i = 0; i < ng; i++) for (int j = i + 1; j < ng; j++) {
            const OrcGeom &g1 = W.geoms[i], &g2 = W.geoms[j];
            if (!pair_filter(g1, g2)) continue;
            if (!aabb_overlap(g1.aabb, g2.aabb)) continue;
            if (hash && !hash_space_meets(g1.aabb, g2.aabb, minl, maxl)) continue;
            W.pairs.push_back(i); W.pairs.push_back(j);
        }
    }
}

// dAreConnectedExcluding ode.cpp:1569-1577 (joint_type = contact)
bool connected_excluding_contacts(World &W, int b1, int b2)
{
    Body &b = W.bodies[b1];
    for (size_t k = 0; k < b.adj.size(); k++)
        if (joint_ref(W, b.adj[k].joint).type != ODEB_JOINT_CONTACT && b.adj[k].other == b2) return true;
    return false;
}

// dJointAttach ode.cpp:1383-1439 (list mechanics only)
void attach(World &W, int jid, Joint &j, int body1, int body2)
{
    if (body1 < 0) { body1 = body2; body2 = -1; j.reverse = 1; } else j.reverse = 0;
    j.b0 = body1; j.b1 = body2;
    if (body1 >= 0) { Adj a = { jid, body2 }; W.bodies[body1].adj.push_back(a); }
    if (body2 >= 0) { Adj a = { jid, body1 }; W.bodies[body2].adj.push_back(a); }
}

void collide_world(const Batch &B, World &W)
{
    for (size_t i = 0; i < W.geoms.size(); i++) {       // recomputePosr of geoms with an offset
        OrcGeom &g = W.geoms[i];
        if (!g.has_ofs) continue;
        const Body &b = W.bodies[g.body];
        mul0_331(g.fpos, b.R, g.opos);
        g.fpos[0] += b.pos[0]; g.fpos[1] += b.pos[1]; g.fpos[2] += b.pos[2];
        mul0_333(g.fR, b.R, g.oR);
        g.pos = g.fpos; g.R = g.fR;
    }
    find_pairs(B, W);
    W.contacts.clear(); W.contact_g.clear(); W.ray_cg.clear(); W.ray_g.clear();
    const OdebWorldParams &p = B.wp;
    OrcContactGeom cg[16];
    int npj = (int)W.pjoints.size();
    for (size_t k = 0; k + 1 < W.pairs.size(); k += 2) {
        int i1 = W.pairs[k], i2 = W.pairs[k + 1];
        const OrcGeom &o1 = W.geoms[i1], &o2 = W.geoms[i2];
        int b1 = o1.body, b2 = o2.body;
        if (p.skip_connected && b1 >= 0 && b2 >= 0 && connected_excluding_contacts(W, b1, b2)) continue;
        if (b1 < 0 && b2 < 0) continue;
        int n = orc_collide(o1, o2, p.max_contacts, cg);
        if (o1.type == ODEB_RAY || o2.type == ODEB_RAY) {      // sensor policy: the hit is recorded, no contact joint
            for (int i = 0; i < n; i++) { W.ray_cg.push_back(cg[i]); W.ray_g.push_back(i1); W.ray_g.push_back(i2); }
            continue;
        }
        for (int i = 0; i < n; i++) {
            Joint j;
            memset(&j, 0, sizeof(j));
            j.type = ODEB_JOINT_CONTACT;
            j.cg = cg[i];
            int jid = npj + (int)W.contacts.size();
            W.contacts.push_back(j);
            attach(W, jid, W.contacts.back(), b1, b2);
            W.contact_g.push_back(i1); W.contact_g.push_back(i2);
        }
    }
}

// dJointGroupEmpty ode.cpp:1325-1366: contact joints leave every body's list
void remove_contacts(World &W)
{
    int npj = (int)W.pjoints.size();
    for (size_t i = 0; i < W.bodies.size(); i++) {
        std::vector<Adj> &a = W.bodies[i].adj;
        size_t k = 0;
        for (size_t t = 0; t < a.size(); t++) if (a[t].joint < npj) a[k++] = a[t];
        a.resize(k);
    }
    W.contacts.clear();
}

// ---------------------------------------------------------------------------------------------
// dInternalHandleAutoDisabling util.cpp:427-561
void auto_disable(const Batch &B, World &W, Real h)
{
    for (int bi = (int)W.bodies.size() - 1; bi >= 0; bi--) {
        Body &bb = W.bodies[bi];
        if (bb.adj.empty()) continue;
        if ((bb.flags & (BF_AUTO_DISABLE | BF_DISABLED)) != BF_AUTO_DISABLE) continue;
        if (B.adis_samples == 0) continue;
        unsigned c = bb.avg_counter;
        for (int k = 0; k < 3; k++) { bb.avg_l[3 * c + k] = bb.lvel[k]; bb.avg_a[3 * c + k] = bb.avel[k]; }
        bb.avg_counter++;
        if (bb.avg_counter >= B.adis_samples) { bb.avg_counter = 0; bb.avg_ready = 1; }
        int idle = 0;
        if (bb.avg_ready) {
            idle = 1;
            Real al[3], aa[3];
            for (int k = 0; k < 3; k++) { al[k] = bb.avg_l[k]; aa[k] = bb.avg_a[k]; }
            if (B.adis_samples > 1) {
                for (unsigned i = 1; i < B.adis_samples; i++)
                    for (int k = 0; k < 3; k++) { al[k] += bb.avg_l[3 * i + k]; aa[k] += bb.avg_a[3 * i + k]; }
                Real r1 = R_(1.0) / (Real)B.adis_samples;
                for (int k = 0; k < 3; k++) { al[k] *= r1; aa[k] *= r1; }
            }
            Real ls = dot3(al, al);
            if (ls > B.adis_lin) idle = 0;
            else { Real as = dot3(aa, aa); if (as > B.adis_ang) idle = 0; }
        }
        if (idle) { bb.adis_stepsleft--; bb.adis_timeleft -= h; }
        else { bb.adis_stepsleft = B.adis_steps; bb.adis_timeleft = B.adis_time; }
        if (bb.adis_stepsleft <= 0 && bb.adis_timeleft <= 0) {
            bb.flags |= BF_DISABLED;
            for (int k = 0; k < 3; k++) { bb.lvel[k] = 0; bb.avel[k] = 0; }
        }
    }
}

// dxJoint::isEnabled joints/joint.cpp:73-78 (no joint-disable flag in this scope)
bool joint_enabled(World &W, const Joint &j)
{
    return W.bodies[j.b0].invMass > 0 || (j.b1 >= 0 && W.bodies[j.b1].invMass > 0);
}

struct Islands { std::vector<int> body, joint, sizes; };

// BuildIslandsAndEstimateStepperMemoryRequirements util.cpp:724-860
void build_islands(const Batch &B, World &W, Real h, Islands &out)
{
    auto_disable(B, W, h);
    int nb = (int)W.bodies.size();
    for (int i = 0; i < nb; i++) W.bodies[i].tag = 0;
    for (size_t i = 0; i < W.pjoints.size(); i++) W.pjoints[i].tag = 0;
    for (size_t i = 0; i < W.contacts.size(); i++) W.contacts[i].tag = 0;
    std::vector<int> stack;
    for (int bi = nb - 1; bi >= 0; bi--) {       // world->firstbody order = reverse creation (ode.cpp:54-60,264)
        Body &bb = W.bodies[bi];
        if (bb.tag) continue;
        if (bb.flags & BF_DISABLED) { bb.tag = -1; continue; }
        bb.tag = 1;
        size_t bstart = out.body.size(), jstart = out.joint.size();
        out.body.push_back(bi);
        stack.clear();
        int b = bi;
        while (true) {
            std::vector<Adj> &adj = W.bodies[b].adj;
            for (int k = (int)adj.size() - 1; k >= 0; k--) {   // newest attachment first
                Joint &nj = joint_ref(W, adj[k].joint);
                if (!nj.tag) {
                    if (joint_enabled(W, nj)) {
                        nj.tag = 1;
                        out.joint.push_back(adj[k].joint);
                        int nbody = adj[k].other;
                        if (nbody >= 0 && W.bodies[nbody].tag <= 0) {
                            W.bodies[nbody].tag = 1;
                            W.bodies[nbody].flags &= ~BF_DISABLED;
                            stack.push_back(nbody);
                        }
                    } else nj.tag = -1;
                }
            }
            if (stack.empty()) break;
            b = stack.back(); stack.pop_back();
            out.body.push_back(b);
        }
        out.sizes.push_back((int)(out.body.size() - bstart));
        out.sizes.push_back((int)(out.joint.size() - jstart));
    }
}

// ---------------------------------------------------------------------------------------------
// dxStepBody util.cpp:583-692
Real sinc(Real x)
{
    if (RFABS(x) < 1.0e-4) return R_(1.0) - x * x * R_(0.166666666666666666667);
    return RSIN(x) / x;
}

void step_body(const Batch &B, Body &b, Real h)
{
    if (b.flags & BF_MAX_ANG_SPEED) {
        const Real mas = B.max_ang_speed;
        const Real asp = dot3(b.avel, b.avel);
        if (asp > mas * mas) { const Real coef = mas / RSQRT(asp); b.avel[0] *= coef; b.avel[1] *= coef; b.avel[2] *= coef; }
    }
    for (int j = 0; j < 3; j++) b.pos[j] += h * b.lvel[j];
    if (b.flags & BF_FINITE_ROT) {
        Real q[4];
        Real wlen = RSQRT(b.avel[0] * b.avel[0] + b.avel[1] * b.avel[1] + b.avel[2] * b.avel[2]);
        h *= R_(0.5);
        Real theta = wlen * h;
        q[0] = RCOS(theta);
        Real s = sinc(theta) * h;
        q[1] = b.avel[0] * s; q[2] = b.avel[1] * s; q[3] = b.avel[2] * s;
        Real q2[4];
        qmul0(q2, q, b.q);
        for (int j = 0; j < 4; j++) b.q[j] = q2[j];
    } else {
        Real dq[4];
        dq_from_w(dq, b.avel, b.q);
        for (int j = 0; j < 4; j++) b.q[j] += h * dq[j];
    }
    normalize4(b.q);
    r_from_q(b.R, b.q);
    if (b.flags & BF_LIN_DAMP) {
        const Real ls = dot3(b.lvel, b.lvel);
        if (ls > B.damp_lin_thr) { const Real k = 1 - B.damp_lin_scale; b.lvel[0] *= k; b.lvel[1] *= k; b.lvel[2] *= k; }
    }
    if (b.flags & BF_ANG_DAMP) {
        const Real as = dot3(b.avel, b.avel);
        if (as > B.damp_ang_thr) { const Real k = 1 - B.damp_ang_scale; b.avel[0] *= k; b.avel[1] *= k; b.avel[2] *= k; }
    }
}

// ---------------------------------------------------------------------------------------------
// contact joint rows: dxJointContact::getInfo1 / getInfo2 joints/contact.cpp:48-347 (no rolling / AxisDep)
void contact_info1(const Batch &B, Joint &j)
{
    int m = 1, nub = 0;
    Real mu = (Real)B.wp.mu;
    if (B.wp.surf_mode & ODEB_CONTACT_MU2) {
        // dContactAxisDep branch (contact.cpp:56-70): mu and mu2 counted separately
        if (mu > 0) { if (mu == R_INF) nub++; m++; }
        Real mu2 = (Real)B.wp.mu2;
        if (mu2 > 0) { if (mu2 == R_INF) nub++; m++; }
        if (B.wp.surf_mode & ODEB_CONTACT_ROLLING) {          // contact.cpp:73-98: rho == 0 still counts a row
            const Real r[3] = { (Real)B.wp.rho, (Real)B.wp.rho2, (Real)B.wp.rhoN };
            for (int i = 0; i < 3; i++) if (!(r[i] < 0)) { if (r[i] == R_INF) nub++; m++; }
        }
    } else {
        if (mu > 0) { if (mu == R_INF) nub += 2; m += 2; }
        if (B.wp.surf_mode & ODEB_CONTACT_ROLLING) {          // contact.cpp:108-116
            const Real r = (Real)B.wp.rho;
            if (!(r < 0)) { if (r == R_INF) nub += 3; m += 3; }
        }
    }
    j.the_m = m; j.m = m; j.nub = nub;
}

// row layout of quickstep.cpp:267-322: [J1l(3) J1a(3) rhs cfm J2l(3) J2a(3) lo hi]
enum { J1L = 0, J1A = 3, RHS = 6, CFM = 7, J2L = 8, J2A = 11, LO = 14, HI = 15, ROW = 16 };

void contact_info2(const Batch &B, World &W, Joint &j, Real fps, Real worldERP, Real *row, int *findex, const Real *fdir1 = 0)
{
    const OdebWorldParams &p = B.wp;
    const int mode = p.surf_mode;
    Real mu = (Real)p.mu < 0 ? 0 : (Real)p.mu, mu2cfg = (Real)p.mu2 < 0 ? 0 : (Real)p.mu2;
    Real erp = (mode & ODEB_CONTACT_SOFT_ERP) ? (Real)p.soft_erp : worldERP;
    Real k = fps * erp;
    Real depth = j.cg.depth - B.min_depth;
    if (depth < 0) depth = 0;
    Real motionN = (mode & ODEB_CONTACT_MOTIONN) ? (Real)p.motionN : R_(0.0);
    const Real pushout = k * depth + motionN;
    bool apply_bounce = (mode & ODEB_CONTACT_BOUNCE) != 0 && (Real)p.bounce_vel >= 0;
    Real outgoing = 0;
    const Real maxvel = B.max_vel;
    Real c = pushout > maxvel ? maxvel : pushout;
    Real c1[3], c2[3] = { 0, 0, 0 }, normal[3];
    if (j.reverse) { normal[0] = -j.cg.normal[0]; normal[1] = -j.cg.normal[1]; normal[2] = -j.cg.normal[2]; }
    else { normal[0] = j.cg.normal[0]; normal[1] = j.cg.normal[1]; normal[2] = j.cg.normal[2]; }
    Real *J1 = row, *J2 = row + J2L;
    if (j.b1 >= 0) {
        Body &b1 = W.bodies[j.b1];
        for (int i = 0; i < 3; i++) c2[i] = j.cg.pos[i] - b1.pos[i];
        J2[0] = -normal[0]; J2[1] = -normal[1]; J2[2] = -normal[2];
        cross3(J2 + 3, normal, c2);
        if (apply_bounce) outgoing = dot3(J2 + 3, b1.avel) - dot3(normal, b1.lvel);
    }
    Body &b0 = W.bodies[j.b0];
    for (int i = 0; i < 3; i++) c1[i] = j.cg.pos[i] - b0.pos[i];
    J1[0] = normal[0]; J1[1] = normal[1]; J1[2] = normal[2];
    cross3(J1 + 3, c1, normal);
    if (apply_bounce) {
        outgoing += dot3(J1 + 3, b0.avel) + dot3(normal, b0.lvel);
        Real neg_out = motionN - outgoing;
        if (neg_out > (Real)p.bounce_vel) {
            const Real newc = (Real)p.bounce * neg_out + motionN;
            if (newc > c) c = newc;
        }
    }
    row[RHS] = c;
    if (mode & ODEB_CONTACT_SOFT_CFM) row[CFM] = (Real)p.soft_cfm;
    row[LO] = 0; row[HI] = R_INF;
    if (j.the_m > 1) {
        Real t1[3], t2[3];
        if (fdir1) { t1[0] = fdir1[0]; t1[1] = fdir1[1]; t1[2] = fdir1[2]; cross3(t2, normal, t1); }   // dContactFDir1 contact.cpp:203-206
        else plane_space(normal, t1, t2);
        int r = 1;
        if (mu > 0) {
            Real *q = row + r * ROW;
            q[J1L] = t1[0]; q[J1L + 1] = t1[1]; q[J1L + 2] = t1[2];
            cross3(q + J1A, c1, t1);
            if (j.b1 >= 0) { q[J2L] = -t1[0]; q[J2L + 1] = -t1[1]; q[J2L + 2] = -t1[2]; cross3(q + J2A, t1, c2); }
            if (mode & ODEB_CONTACT_MOTION1) q[RHS] = (Real)p.motion1;
            if (mode & ODEB_CONTACT_SLIP1) q[CFM] = (Real)p.slip1;
            q[LO] = -mu; q[HI] = mu;
            if (mode & ODEB_CONTACT_APPROX1_1) findex[r] = 0;
            r++;
        }
        const Real mu2 = (mode & ODEB_CONTACT_MU2) ? mu2cfg : mu;
        if (mu2 > 0) {
            Real *q = row + r * ROW;
            q[J1L] = t2[0]; q[J1L + 1] = t2[1]; q[J1L + 2] = t2[2];
            cross3(q + J1A, c1, t2);
            if (j.b1 >= 0) { q[J2L] = -t2[0]; q[J2L + 1] = -t2[1]; q[J2L + 2] = -t2[2]; cross3(q + J2A, t2, c2); }
            if (mode & ODEB_CONTACT_MOTION2) q[RHS] = (Real)p.motion2;
            if (mode & ODEB_CONTACT_SLIP2) q[CFM] = (Real)p.slip2;
            q[LO] = -mu2; q[HI] = mu2;
            if (mode & ODEB_CONTACT_APPROX1_2) findex[r] = 0;
            r++;
        }
        if (mode & ODEB_CONTACT_ROLLING) {   // contact.cpp:299-343
            const Real *ax[3] = { t1, t2, normal };
            const int approx_bits[3] = { ODEB_CONTACT_APPROX1_1, ODEB_CONTACT_APPROX1_2, ODEB_CONTACT_APPROX1_N };
            Real rho[3];
            rho[0] = (Real)p.rho < 0 ? 0 : (Real)p.rho;
            if (mode & ODEB_CONTACT_MU2) { rho[1] = (Real)p.rho2 < 0 ? 0 : (Real)p.rho2; rho[2] = (Real)p.rhoN < 0 ? 0 : (Real)p.rhoN; }
            else { rho[1] = rho[0]; rho[2] = rho[0]; }
            for (int i = 0; i < 3; i++) {
                if (rho[i] > 0) {
                    Real *q = row + r * ROW;
                    q[J1A] = ax[i][0]; q[J1A + 1] = ax[i][1]; q[J1A + 2] = ax[i][2];
                    if (j.b1 >= 0) { q[J2A] = -ax[i][0]; q[J2A + 1] = -ax[i][1]; q[J2A + 2] = -ax[i][2]; }
                    q[LO] = -rho[i]; q[HI] = rho[i];
                    if (mode & approx_bits[i]) findex[r] = 0;
                    r++;
                }
            }
        }
    }
}

void joint_info1(const Batch &B, World &W, Joint &j)
{
    switch (j.type) {
    case ODEB_JOINT_CONTACT: contact_info1(B, j); break;
    case ODEB_JOINT_BALL: j.m = 3; j.nub = 3; break;       // ball.cpp:49-54
    case ODEB_JOINT_FIXED: j.m = 6; j.nub = 6; break;      // fixed.cpp:52-57
    case ODEB_JOINT_SLIDER: slider_info1(W, j); break;
    case ODEB_JOINT_HINGE2: hinge2_info1(W, j); break;
    case ODEB_JOINT_HINGE: hinge_info1(W, j); break;
    case ODEB_JOINT_UNIVERSAL: universal_info1(W, j); break;
    case ODEB_JOINT_LMOTOR: lmotor_info1(W, j); break;
    case ODEB_JOINT_AMOTOR: amotor_info1(W, j); break;
    }
}

void joint_info2(const Batch &B, World &W, Joint &j, Real fps, Real worldERP, Real *row, int *findex)
{
    switch (j.type) {
    case ODEB_JOINT_CONTACT: contact_info2(B, W, j, fps, worldERP, row, findex); break;
    case ODEB_JOINT_BALL:
        row[CFM] = j.cfm; row[ROW + CFM] = j.cfm; row[2 * ROW + CFM] = j.cfm;   // ball.cpp:57-67
        set_ball(W, j, fps, j.erp, row, j.anchor1, j.anchor2);
        break;
    case ODEB_JOINT_FIXED: fixed_info2(W, j, fps, worldERP, row); break;
    case ODEB_JOINT_SLIDER: slider_info2(W, j, fps, worldERP, row); break;
    case ODEB_JOINT_HINGE2: hinge2_info2(W, j, fps, worldERP, row); break;
    case ODEB_JOINT_HINGE: hinge_info2(W, j, fps, worldERP, row, findex); break;
    case ODEB_JOINT_UNIVERSAL: universal_info2(W, j, fps, worldERP, row, findex); break;
    case ODEB_JOINT_LMOTOR: lmotor_info2(W, j, fps, row); break;
    case ODEB_JOINT_AMOTOR: amotor_info2(W, j, fps, row); break;
    }
}

// ---------------------------------------------------------------------------------------------
// dxQuickStepIsland quickstep.cpp:1112-3439, single-threaded form
enum { IMJ1 = 0, IMJ1_MAX = 6, IMJ2 = 7, IMJ2_MAX = 13, IMJ_ROW = 14 };

Real modulo_max6(const Real *v)
{   // dxCalculateModuloMaximum matrix.h:92-104
    Real r = RFABS(v[0]);
    for (int i = 1; i < 6; i++) { Real a = RFABS(v[i]); if (a > r) r = a; }
    return r;
}

// Canonical row order of the large-world path, ODEB_MODE_CANONICAL (include/ode_b200.h).
// Groups that share a body conflict.  ONCE per step and island the groups are coloured by rounds (Jones-Plassmann with first fit): in
// every round each still uncoloured group whose (key, first row) is larger than that of all its still uncoloured neighbours takes the
// smallest colour none of its already coloured neighbours holds; key = odeb_canon_key(seed, island, 0, first row of the group).  Winners
// of a round are never neighbours, so the result does not depend on any evaluation order (the CUDA path runs the rounds as two kernels,
// mark and assign).  Groups of one colour touch disjoint bodies: the CUDA path relaxes them side by side with the same result.
static void canonical_colours(unsigned seed, unsigned island, unsigned m, int nb, const std::vector<int> &grp, const std::vector<int> &jb, std::vector<int> &colour)
{
    std::vector<int> heads;
    for (unsigned r = 0; r < m; r++) if (grp[r] == (int)r) heads.push_back((int)r);
    std::vector<std::vector<int> > on_body(nb);
    for (size_t k = 0; k < heads.size(); k++) for (int side = 0; side < 2; side++) { const int b = jb[2 * heads[k] + side]; if (b >= 0) on_body[b].push_back(heads[k]); }
    std::vector<unsigned> key(m, 0);
    colour.assign(m, -1);
    for (size_t k = 0; k < heads.size(); k++) key[heads[k]] = odebi_canon_key(seed, island, 0, (unsigned)heads[k]);
    size_t left = heads.size();
    while (left > 0) {
        std::vector<int> winners;
        for (size_t k = 0; k < heads.size(); k++) {
            const int g = heads[k];
            if (colour[g] >= 0) continue;
            bool top = true;
            for (int side = 0; side < 2 && top; side++) {
                const int b = jb[2 * g + side];
                if (b < 0) continue;
                for (size_t t = 0; t < on_body[b].size(); t++) {
                    const int h = on_body[b][t];
                    if (h != g && colour[h] < 0 && (key[h] > key[g] || (key[h] == key[g] && h > g))) { top = false; break; }
                }
            }
            if (top) winners.push_back(g);
        }
        std::vector<int> picked(winners.size());
        for (size_t k = 0; k < winners.size(); k++) {
            const int g = winners[k];
            bool used[ODEB_CANON_COLOURS] = { false };         // colours held by coloured neighbours
            for (int side = 0; side < 2; side++) {
                const int b = jb[2 * g + side];
                if (b < 0) continue;
                for (size_t t = 0; t < on_body[b].size(); t++) { const int c = colour[on_body[b][t]]; if (c >= 0 && c < ODEB_CANON_COLOURS) used[c] = true; }
            }
            int c = 0;
            while (c < ODEB_CANON_COLOURS - 1 && used[c]) c++;
            picked[k] = c;
        }
        for (size_t k = 0; k < winners.size(); k++) colour[winners[k]] = picked[k];
        left -= winners.size();
    }
}
// Sweep order of phase k (the 8 sweeps from sweep 8k on): the colours are visited in ascending (odeb_canon_key(seed, ~0, k, colour), colour)
// -- one permutation of the colour numbers per phase, the same for every island --, the groups of a colour by ascending first row, the
// rows of a group in row order (a contact's normal row right before its friction rows).
static void canonical_order(unsigned seed, unsigned phase, unsigned m, const std::vector<int> &grp, const std::vector<int> &colour, std::vector<int> &order)
{
    int rank[ODEB_CANON_COLOURS];
    odebi_canon_colour_ranks(seed, phase, rank);
    for (unsigned r = 0; r < m; r++) order[r] = (int)r;
    std::sort(order.begin(), order.begin() + m, [&](int a, int b) {
        const int ca = rank[colour[grp[a]]], cb = rank[colour[grp[b]]];
        if (ca != cb) return ca < cb;
        return a < b;                                          // rows of a group are contiguous: (group, row) order = row order
    });
}

void quickstep_island(const Batch &B, World &W, const int *bodies, int nb, const int *joints, int nj_all, Real h)
{
    std::vector<Real> invI(12 * nb);
    // Stage0_Bodies :1176-1306
    for (int i = 0; i < nb; i++) W.bodies[bodies[i]].tag = i;
    for (int ax = 0; ax < 3; ax++) {
        Real g = B.gravity[ax];
        if (g) for (int i = 0; i < nb; i++) { Body &b = W.bodies[bodies[i]]; if (!(b.flags & BF_NO_GRAVITY)) b.facc[ax] += b.mass * g; }
    }
    for (int i = 0; i < nb; i++) {
        Body &b = W.bodies[bodies[i]];
        Real tmp[12], *ii = &invI[12 * i];
        mul2_333(tmp, b.invI, b.R);
        mul0_333(ii, b.R, tmp);
        ii[3] = ii[7] = ii[11] = 0;
        if ((b.flags & BF_GYRO) && b.invMass > 0) {
            Real I[12], L[3], Itild[12] = { 0 }, itInv[12];
            mul2_333(tmp, b.I, b.R);
            mul0_333(I, b.R, tmp);
            I[3] = I[7] = I[11] = 0;
            mul0_331(L, I, b.avel);
            Itild[1] = +L[2]; Itild[2] = -L[1]; Itild[4] = -L[2]; Itild[6] = +L[0]; Itild[8] = +L[1]; Itild[9] = -L[0];
            for (int k = 0; k < 12; k++) Itild[k] = Itild[k] * h + I[k];
            Real hr = rrecip(h);
            L[0] *= hr; L[1] *= hr; L[2] *= hr;
            if (invert3(itInv, Itild) != 0) {
                itInv[3] = itInv[7] = itInv[11] = 0;
                mul0_333(Itild, I, itInv);
                Itild[0] -= 1; Itild[5] -= 1; Itild[10] -= 1;
                Real tau0[3];
                mul0_331(tau0, Itild, L);
                b.tacc[0] += tau0[0]; b.tacc[1] += tau0[1]; b.tacc[2] += tau0[2];
            }
        }
    }
    // Stage0_Joints :1319-1351
    std::vector<int> jl; jl.reserve(nj_all);
    unsigned m = 0;
    for (int k = 0; k < nj_all; k++) {
        Joint &j = joint_ref(W, joints[k]);
        joint_info1(B, W, j);
        if (j.m != 0) { m += j.m; jl.push_back(joints[k]); }
    }
    int nj = (int)jl.size();
    const Real hrecip = rrecip(h);
    std::vector<Real> J, iMJ, lambda, cforce(6 * nb, 0), fa(2 * nb, 0), rhs_tmp(6 * nb);
    std::vector<int> findex, jb, order, mindex(nj + 1), grp, gcolour;
    int isl_sweeps = 0;
    if (m > 0) {
        // Stage1 :1364-1472
        mindex[0] = 0;
        for (int k = 0; k < nj; k++) mindex[k + 1] = mindex[k] + joint_ref(W, jl[k]).m;
        J.assign((size_t)m * ROW, 0); findex.assign(m, -1); jb.assign(2 * m, -1);
        // Stage2a :1486-1610
        unsigned valid_findices = 0;
        for (int k = 0; k < nj; k++) {
            Joint &j = joint_ref(W, jl[k]);
            int ofs = mindex[k], infom = j.m;
            Real *row = &J[(size_t)ofs * ROW];
            for (int r = 0; r < infom; r++) {
                Real *q = row + r * ROW;
                for (int t = 0; t < ROW; t++) q[t] = 0;
                q[CFM] = B.cfm; q[LO] = -R_INF; q[HI] = R_INF;
            }
            joint_info2(B, W, j, hrecip, B.erp, row, &findex[ofs]);
            for (int r = 0; r < infom; r++) if (findex[ofs + r] != -1) { findex[ofs + r] += ofs; valid_findices++; }
            for (int r = 0; r < infom; r++) { row[r * ROW + RHS] *= hrecip; row[r * ROW + CFM] *= hrecip; }
            int t0 = W.bodies[j.b0].tag, t1 = j.b1 >= 0 ? W.bodies[j.b1].tag : -1;
            for (int r = 0; r < infom; r++) { jb[2 * (ofs + r)] = t0; jb[2 * (ofs + r) + 1] = t1; }
        }
        // Jcopy :1590-1606: the J blocks of joints with feedback, before any scaling
        std::vector<Real> Jcopy;
        if (B.feedback) Jcopy = J;
        // Stage2b :1644-1690
        for (int i = 0; i < nb; i++) {
            Body &b = W.bodies[bodies[i]];
            Real *rc = &rhs_tmp[6 * i];
            for (int k = 0; k < 3; k++) rc[k] = -(b.facc[k] * b.invMass + b.lvel[k] * hrecip);
            mul0_331(rc + 3, &invI[12 * i], b.tacc);
            for (int k = 0; k < 3; k++) rc[3 + k] = -(b.avel[k] * hrecip) - rc[3 + k];
        }
        // Stage2c / multiplyAdd_J :1023-1055
        for (unsigned i = 0; i < m; i++) {
            Real *q = &J[(size_t)i * ROW];
            int b1 = jb[2 * i], b2 = jb[2 * i + 1];
            Real sum = R_(0.0);
            const Real *in = &rhs_tmp[6 * b1];
            for (int k = 0; k < 6; k++) sum += q[J1L + k] * in[k];
            if (b2 != -1) { in = &rhs_tmp[6 * b2]; for (int k = 0; k < 6; k++) sum += q[J2L + k] * in[k]; }
            q[RHS] += sum;
        }
        // Stage3/4: lambda = 0; iMJ (compute_invM_JT :859-897)
        lambda.assign(m, 0); iMJ.assign((size_t)m * IMJ_ROW, 0); order.resize(m);
        for (unsigned i = 0; i < m; i++) {
            Real *im = &iMJ[(size_t)i * IMJ_ROW]; const Real *q = &J[(size_t)i * ROW];
            int b1 = jb[2 * i], b2 = jb[2 * i + 1];
            Real k1 = W.bodies[bodies[b1]].invMass;
            for (int k = 0; k < 3; k++) im[IMJ1 + k] = k1 * q[J1L + k];
            mul0_331(im + IMJ1 + 3, &invI[12 * b1], q + J1A);
            im[IMJ1_MAX] = B.dyn_enabled ? modulo_max6(im + IMJ1) : R_(0.0);
            if (b2 != -1) {
                Real k2 = W.bodies[bodies[b2]].invMass;
                for (int k = 0; k < 3; k++) im[IMJ2 + k] = k2 * q[J2L + k];
                mul0_331(im + IMJ2 + 3, &invI[12 * b2], q + J2A);
                im[IMJ2_MAX] = B.dyn_enabled ? modulo_max6(im + IMJ2) : R_(0.0);
            }
        }
        // AdComputation :2251-2316
        for (unsigned i = 0; i < m; i++) {
            Real *im = &iMJ[(size_t)i * IMJ_ROW]; Real *q = &J[(size_t)i * ROW];
            Real sum = R_(0.0);
            for (int k = 0; k < 6; k++) sum += im[IMJ1 + k] * q[J1L + k];
            int b2 = jb[2 * i + 1];
            if (b2 != -1) for (int k = 0; k < 6; k++) sum += im[IMJ2 + k] * q[J2L + k];
            Real cfm_i = q[CFM];
            Real Ad = B.sor_w / (sum + cfm_i);
            q[CFM] = cfm_i * Ad;
            q[RHS] *= Ad;
            for (int k = 0; k < 6; k++) q[J1L + k] *= Ad;
            if (b2 != -1) for (int k = 0; k < 6; k++) q[J2L + k] *= Ad;
        }
        // ReorderPrep :2329-2355
        {
            unsigned head = 0, tail = m - valid_findices;
            for (unsigned i = 0; i < m; i++) { if (findex[i] == -1) order[head++] = i; else order[tail++] = i; }
            if (B.canonical) {
                // canonical mode (include/ode_b200.h): row groups = the rows of the contacts of one geom pair / of one permanent
                // joint (all rows of a group act on the same two bodies); the order of every phase of 8 sweeps is colour-major, see
                // canonical_order below
                grp.resize(m);
                const int npj = (int)W.pjoints.size();
                for (int k = 0; k < nj; k++) {
                    int first = mindex[k];
                    if (k > 0 && jl[k] >= npj && jl[k - 1] >= npj) {
                        int c = jl[k] - npj, cp = jl[k - 1] - npj;
                        if (W.contact_g[2 * c] == W.contact_g[2 * cp] && W.contact_g[2 * c + 1] == W.contact_g[2 * cp + 1]) first = grp[mindex[k - 1]];
                    }
                    for (int r = mindex[k]; r < mindex[k + 1]; r++) grp[r] = first;
                }
                canonical_colours(W.step_seed, (unsigned)W.cur_island, m, nb, grp, jb, gcolour);
                canonical_order(W.step_seed, 0, m, grp, gcolour, order);
            }
        }
        // iteration loop :1823-1856
        Real exit_delta = B.premature_delta;
        const unsigned num_iterations = B.num_iter;
        for (unsigned iteration = 0, extra = 0;;) {
            // IsSORConstraintsReorderRequiredForIteration :1080-1109 + ConstraintsShuffling :2578-2611
            if (iteration >= 8 && (iteration % 8) == 0) {
                if (B.canonical) {
                    // canonical mode: the colour-major order of phase k (same colouring, the colours visited in the phase's own order); the dRand stream is advanced by the m-1 draws the reference's
                    // Fisher-Yates pass would have consumed (world_step)
                    canonical_order(W.step_seed, iteration / 8, m, grp, gcolour, order);
                    W.draws += m - 1;
                } else
                for (unsigned idx = 1; idx < m; idx++) {
                    int sw = orc_rand_int(&W.seed, idx + 1);
                    int t = order[idx]; order[idx] = order[sw]; order[sw] = t;
                }
            }
            // STIteration / IterationStep :2893-3033
            for (unsigned i = 0; i < m; i++) {
                unsigned index = order[i];
                Real old_lambda = lambda[index];
                const Real *q = &J[(size_t)index * ROW];
                Real delta = q[RHS] - old_lambda * q[CFM];
                int b1 = jb[2 * index], b2 = jb[2 * index + 1];
                Real *fc1 = &cforce[6 * b1], *fc2 = 0;
                delta -= fc1[0] * q[J1L] + fc1[1] * q[J1L + 1] + fc1[2] * q[J1L + 2] + fc1[3] * q[J1A] + fc1[4] * q[J1A + 1] + fc1[5] * q[J1A + 2];
                if (b2 != -1) {
                    fc2 = &cforce[6 * b2];
                    delta -= fc2[0] * q[J2L] + fc2[1] * q[J2L + 1] + fc2[2] * q[J2L + 2] + fc2[3] * q[J2A] + fc2[4] * q[J2A + 1] + fc2[5] * q[J2A + 2];
                }
                Real hi_act, lo_act;
                if (findex[index] != -1) { hi_act = RFABS(q[HI] * lambda[findex[index]]); lo_act = -hi_act; }
                else { hi_act = q[HI]; lo_act = q[LO]; }
                Real new_lambda = old_lambda + delta;
                if (new_lambda < lo_act) { delta = lo_act - old_lambda; lambda[index] = lo_act; }
                else if (new_lambda > hi_act) { delta = hi_act - old_lambda; lambda[index] = hi_act; }
                else lambda[index] = new_lambda;
                if (delta != 0) {
                    const Real *im = &iMJ[(size_t)index * IMJ_ROW];
                    int pos = (delta > 0) ? 1 : 0;     // FAE_NEGATIVE = 0, FAE_POSITIVE = 1 (quickstep.cpp:447-470)
                    for (int k = 0; k < 6; k++) fc1[k] += delta * im[IMJ1 + k];
                    fa[2 * b1 + pos] += delta * im[IMJ1_MAX];
                    if (fc2) {
                        fa[2 * b2 + pos] += delta * im[IMJ2_MAX];
                        for (int k = 0; k < 6; k++) fc2[k] += delta * im[IMJ2 + k];
                    }
                }
            }
            ++iteration;
            W.sweeps++; isl_sweeps++;
            if (iteration - extra == num_iterations) {
                if (extra != 0 || B.max_extra == 0) { if (extra != 0) W.stats[3]++; break; }
                extra = B.max_extra;
                exit_delta = B.extra_delta;
            }
            if (B.dyn_enabled) {
                // CheckForMaximumToBeLessThanLimitAndResetMaxAdjustments :3253-3285
                bool hit = false;
                if (exit_delta == 0) { std::fill(fa.begin(), fa.end(), R_(0.0)); hit = true; }
                if (!hit) for (int i = 0; i < nb; i++) {
                    if (!(fa[2 * i + 1] < exit_delta) || !(-fa[2 * i] < exit_delta)) {
                        for (int k = 2 * i; k < 2 * nb; k++) fa[k] = 0;
                        hit = true; break;
                    }
                    fa[2 * i] = 0; fa[2 * i + 1] = 0;
                }
                if (!hit) {
                    if (iteration < num_iterations) W.stats[1]++;
                    else if (iteration > num_iterations) W.stats[2]++;
                    break;
                }
            }
        }
        // Stage4b :3082-3108
        for (int i = 0; i < nb; i++) {
            Body &b = W.bodies[bodies[i]];
            for (int k = 0; k < 3; k++) { b.lvel[k] += h * cforce[6 * i + k]; b.avel[k] += h * cforce[6 * i + 3 + k]; }
        }
        // joint feedback :3108-3182 (Multiply1_12q1 :153-184: one running sum per component over the joint's rows)
        if (B.feedback) for (int k = 0; k < nj; k++) {
            const Joint &j = joint_ref(W, jl[k]);
            Real *out = &W.fb[12 * (size_t)jl[k]];
            for (int side = 0; side < 2; side++) {
                if (side == 1 && j.b1 < 0) break;
                Real acc[6] = { 0, 0, 0, 0, 0, 0 };
                for (int r = 0; r < j.m; r++) {
                    const Real *q = &Jcopy[(size_t)(mindex[k] + r) * ROW] + (side ? J2L : J1L);
                    const Real sl = lambda[mindex[k] + r];
                    for (int t = 0; t < 6; t++) acc[t] += q[t] * sl;
                }
                for (int t = 0; t < 6; t++) out[6 * side + t] = acc[t];
            }
            W.fb_state[jl[k]] = j.b1 < 0 ? 1 : 2;
        }
    }
    W.stats[0]++;     // Stage5 :3201-3203
    W.isl_log.push_back(nb); W.isl_log.push_back((int)m); W.isl_log.push_back(isl_sweeps);
    // Stage6a :3299-3337
    for (int i = 0; i < nb; i++) {
        Body &b = W.bodies[bodies[i]];
        Real k = h * b.invMass;
        for (int t = 0; t < 3; t++) { b.lvel[t] += k * b.facc[t]; b.tacc[t] *= h; }
        Real tmp[3];
        mul0_331(tmp, &invI[12 * i], b.tacc);
        b.avel[0] += tmp[0]; b.avel[1] += tmp[1]; b.avel[2] += tmp[2];
    }
    // Stage6b :3409-3439
    for (int i = 0; i < nb; i++) {
        Body &b = W.bodies[bodies[i]];
        step_body(B, b, h);
        b.facc[0] = b.facc[1] = b.facc[2] = 0; b.tacc[0] = b.tacc[1] = b.tacc[2] = 0;
    }
}

void world_step(const Batch &B, World &W, Real h)
{
    collide_world(B, W);
    if (B.feedback) {
        const size_t n = W.pjoints.size() + W.contacts.size();
        W.fb.assign(12 * n, 0); W.fb_state.assign(n, 0);
    }
    Islands isl;
    build_islands(B, W, h, isl);
    W.island_count = (int)isl.sizes.size() / 2;
    W.island_label.assign(W.bodies.size(), -1);
    {
        size_t bo = 0;
        for (int is = 0; is < W.island_count; is++) { for (int k = 0; k < isl.sizes[2 * is]; k++) W.island_label[isl.body[bo + k]] = is; bo += isl.sizes[2 * is]; }
    }
    size_t bo = 0, jo = 0;
    W.isl_log.clear();
    if (B.canonical) {
        // canonical order of the large-world path: island membership and numbering as the reference finds them, but inside an
        // island bodies in descending creation index and joints in ascending id (permanent joints, then contacts in creation order)
        for (int is = 0; is < W.island_count; is++) {
            int nb = isl.sizes[2 * is], nj = isl.sizes[2 * is + 1];
            std::sort(isl.body.begin() + bo, isl.body.begin() + bo + nb, [](int a, int b) { return a > b; });
            std::sort(isl.joint.begin() + jo, isl.joint.begin() + jo + nj);
            bo += nb; jo += nj;
        }
        bo = 0; jo = 0;
        W.step_seed = W.seed; W.draws = 0;
    }
    for (int is = 0; is < W.island_count; is++) {
        int nb = isl.sizes[2 * is], nj = isl.sizes[2 * is + 1];
        W.cur_island = is;
        quickstep_island(B, W, &isl.body[bo], nb, nj ? &isl.joint[jo] : 0, nj, h);
        bo += nb; jo += nj;
    }
    if (B.canonical) for (unsigned long long d = 0; d < W.draws; d++) orc_rand(&W.seed);
    W.last_cg.resize(W.contacts.size());
    for (size_t i = 0; i < W.contacts.size(); i++) W.last_cg[i] = W.contacts[i].cg;
    remove_contacts(W);
}

} // namespace

extern "C" {

void *orc_create(const OdebWorldParams *wp, int nbody, const OdebBodyDesc *bodies, const double *body_pos, const double *body_quat,
                 int ngeom, const OdebGeomDesc *geoms, int njoint, const OdebJointDesc *joints, int nworlds, int /*device*/)
{
    if (wp->surf_mode & ODEB_CONTACT_FDIR1) return 0;
    Batch *B = new Batch;
    B->wp = *wp; B->nbody = nbody; B->ngeom = ngeom; B->njoint = njoint; B->canonical = 0; B->feedback = 0;
    for (int k = 0; k < 3; k++) B->gravity[k] = (Real)wp->gravity[k];
    B->erp = (Real)wp->erp;
#if defined(ODEB_DOUBLE)
    B->cfm = wp->cfm >= 0 ? (Real)wp->cfm : R_(1e-10);
#else
    B->cfm = wp->cfm >= 0 ? (Real)wp->cfm : R_(1e-5);
#endif
    B->sor_w = (Real)wp->sor_w;
    B->num_iter = wp->num_iterations > 1 ? wp->num_iterations : 1;
    B->premature_delta = (Real)wp->premature_exit_delta; B->extra_delta = (Real)wp->extra_iter_delta;
    B->extra_factor = (Real)wp->max_extra_factor;
    { Real e = B->num_iter * B->extra_factor; B->max_extra = e < (Real)UINT32_MAX ? (unsigned)e : UINT32_MAX; }  // objects.h:177-184
    B->dyn_enabled = B->max_extra != 0 || B->premature_delta != 0;
    B->max_vel = (Real)wp->contact_max_vel; B->min_depth = (Real)wp->contact_surface_layer;
    { Real t = (Real)wp->adis_linear_thr; B->adis_lin = t * t; t = (Real)wp->adis_angular_thr; B->adis_ang = t * t; }
    B->adis_time = (Real)wp->adis_time; B->adis_steps = wp->adis_steps; B->adis_samples = (unsigned)wp->adis_samples;
    B->damp_lin_scale = (Real)wp->linear_damping; B->damp_ang_scale = (Real)wp->angular_damping;
    { Real t = (Real)wp->linear_damping_thr; B->damp_lin_thr = t * t; t = (Real)wp->angular_damping_thr; B->damp_ang_thr = t * t; }
    B->max_ang_speed = (Real)wp->max_angular_speed;
    B->worlds.resize(nworlds);
    for (int wi = 0; wi < nworlds; wi++) {
        World &W = B->worlds[wi];
        W.seed = 0; W.stats[0] = W.stats[1] = W.stats[2] = W.stats[3] = 0; W.island_count = 0; W.sweeps = 0;
        W.bodies.resize(nbody);
        for (int i = 0; i < nbody; i++) {
            Body &b = W.bodies[i];
            memset(b.pos, 0, sizeof(Real) * (4 + 12 + 4 + 4 + 4 + 4 + 4));
            b.mass = (Real)bodies[i].mass;
            memset(b.I, 0, sizeof(b.I));
            const double *I = bodies[i].inertia;
            // dMassSetParameters mass.cpp:74-93: I11 I22 I33 I12 I13 I23, symmetric fill
            b.I[0] = (Real)I[0]; b.I[5] = (Real)I[4]; b.I[10] = (Real)I[8];
            b.I[1] = (Real)I[1]; b.I[2] = (Real)I[2]; b.I[6] = (Real)I[5];
            b.I[4] = (Real)I[1]; b.I[8] = (Real)I[2]; b.I[9] = (Real)I[5];
            if (!invert_pd3(b.I, b.invI)) memcpy(b.invI, g_identity, sizeof(g_identity));
            b.invMass = rrecip(b.mass);
            if (bodies[i].flags & ODEB_BODY_KINEMATIC) { memset(b.invI, 0, sizeof(b.invI)); b.invMass = 0; }     // dBodySetKinematic ode.cpp:837-842
            b.pos[0] = (Real)body_pos[3 * i]; b.pos[1] = (Real)body_pos[3 * i + 1]; b.pos[2] = (Real)body_pos[3 * i + 2];
            Real q[4] = { (Real)body_quat[4 * i], (Real)body_quat[4 * i + 1], (Real)body_quat[4 * i + 2], (Real)body_quat[4 * i + 3] };
            body_set_quat(b, q);
            unsigned fl = BF_GYRO;
            if (wp->auto_disable) fl |= BF_AUTO_DISABLE;
            if (B->damp_lin_scale) fl |= BF_LIN_DAMP;
            if (B->damp_ang_scale) fl |= BF_ANG_DAMP;
            if (B->max_ang_speed < R_INF) fl |= BF_MAX_ANG_SPEED;
            int sf = bodies[i].flags;
            if (sf & ODEB_BODY_NO_GRAVITY) fl |= BF_NO_GRAVITY;
            if (sf & ODEB_BODY_NO_GYRO) fl &= ~BF_GYRO;
            if (sf & ODEB_BODY_FINITE_ROTATION) fl |= BF_FINITE_ROT;
            if (sf & ODEB_BODY_DISABLED) fl |= BF_DISABLED;
            b.flags = fl;
            b.adis_stepsleft = B->adis_steps; b.adis_timeleft = B->adis_time;
            b.avg_counter = 0; b.avg_ready = 0;
            b.avg_l.assign(3 * (B->adis_samples ? B->adis_samples : 1), 0); b.avg_a = b.avg_l;
            b.tag = 0;
        }
        W.geoms.resize(ngeom);
        for (int i = 0; i < ngeom; i++) {
            OrcGeom &g = W.geoms[i];
            g.type = geoms[i].type; g.body = geoms[i].body;
            for (int k = 0; k < 4; k++) g.p[k] = (Real)geoms[i].p[k];
            if (g.type == ODEB_PLANE) {   // make_sure_plane_normal_has_unit_length plane.cpp:48-63
                Real l = g.p[0] * g.p[0] + g.p[1] * g.p[1] + g.p[2] * g.p[2];
                if (l > 0) { l = rrecipsqrt(l); g.p[0] *= l; g.p[1] *= l; g.p[2] *= l; g.p[3] *= l; }
                else { g.p[0] = 1; g.p[1] = 0; g.p[2] = 0; g.p[3] = 0; }
            }
            g.cat = geoms[i].category_bits; g.col = geoms[i].collide_bits;
            if (g.body >= 0) { g.pos = W.bodies[g.body].pos; g.R = W.bodies[g.body].R; }
            else { g.pos = g_zero4; g.R = g_identity; }
            g.has_ofs = 0;
            if (geoms[i].has_offset && g.body >= 0) {
                g.has_ofs = 1;
                Real q[4] = { (Real)geoms[i].offset_quat[0], (Real)geoms[i].offset_quat[1], (Real)geoms[i].offset_quat[2], (Real)geoms[i].offset_quat[3] };
                memset(g.oR, 0, sizeof(g.oR)); memset(g.fR, 0, sizeof(g.fR));
                r_from_q(g.oR, q);
                g.opos[0] = (Real)geoms[i].offset_pos[0]; g.opos[1] = (Real)geoms[i].offset_pos[1]; g.opos[2] = (Real)geoms[i].offset_pos[2]; g.opos[3] = 0;
                g.fpos[3] = 0;
            }
        }
        W.pjoints.resize(njoint);
        for (int i = 0; i < njoint; i++) {
            Joint &j = W.pjoints[i];
            memset(&j, 0, sizeof(j));
            const OdebJointDesc &d = joints[i];
            j.type = d.type;
            j.erp = B->erp; j.cfm = B->cfm;
            limot_init(*B, j.limot1); limot_init(*B, j.limot2); limot_init(*B, j.limot3);
            attach(W, i, j, d.body1, d.body2);
            joint_setup(*B, W, j, d);
        }
    }
    return B;
}

void orc_destroy(void *h) { delete (Batch *)h; }

int orc_set_state(void *h, const Real *pos, const Real *quat, const Real *lvel, const Real *avel)
{
    Batch *B = (Batch *)h;
    for (size_t w = 0; w < B->worlds.size(); w++) for (int i = 0; i < B->nbody; i++) {
        Body &b = B->worlds[w].bodies[i];
        size_t k = w * B->nbody + i;
        if (pos) { b.pos[0] = pos[3 * k]; b.pos[1] = pos[3 * k + 1]; b.pos[2] = pos[3 * k + 2]; }
        if (quat) body_set_quat(b, quat + 4 * k);
        if (lvel) { b.lvel[0] = lvel[3 * k]; b.lvel[1] = lvel[3 * k + 1]; b.lvel[2] = lvel[3 * k + 2]; }
        if (avel) { b.avel[0] = avel[3 * k]; b.avel[1] = avel[3 * k + 1]; b.avel[2] = avel[3 * k + 2]; }
    }
    return 1;
}

int orc_get_state(void *h, Real *pos, Real *quat, Real *lvel, Real *avel)
{
    Batch *B = (Batch *)h;
    for (size_t w = 0; w < B->worlds.size(); w++) for (int i = 0; i < B->nbody; i++) {
        Body &b = B->worlds[w].bodies[i];
        size_t k = w * B->nbody + i;
        if (pos) memcpy(pos + 3 * k, b.pos, 3 * sizeof(Real));
        if (quat) memcpy(quat + 4 * k, b.q, 4 * sizeof(Real));
        if (lvel) memcpy(lvel + 3 * k, b.lvel, 3 * sizeof(Real));
        if (avel) memcpy(avel + 3 * k, b.avel, 3 * sizeof(Real));
    }
    return 1;
}

int orc_add_force(void *h, const Real *force, const Real *torque)
{
    Batch *B = (Batch *)h;
    for (size_t w = 0; w < B->worlds.size(); w++) for (int i = 0; i < B->nbody; i++) {
        Body &b = B->worlds[w].bodies[i];
        size_t k = w * B->nbody + i;
        if (force) for (int t = 0; t < 3; t++) b.facc[t] += force[3 * k + t];
        if (torque) for (int t = 0; t < 3; t++) b.tacc[t] += torque[3 * k + t];
    }
    return 1;
}

int orc_set_solver_mode(void *h, int mode) { ((Batch *)h)->canonical = mode == ODEB_MODE_CANONICAL; return 1; }
int orc_enable_feedback(void *h, int on) { ((Batch *)h)->feedback = on != 0; return 1; }
int orc_get_feedback(void *h, int world, odeb_real *out12, int *state, int cap)
{
    Batch *B = (Batch *)h;
    const World &W = B->worlds[world];
    const int n = (int)W.fb_state.size();
    for (int i = 0; i < n && i < cap; i++) {
        for (int k = 0; k < 12; k++) out12[12 * i + k] = W.fb[12 * (size_t)i + k];
        state[i] = W.fb_state[i];
    }
    return n;
}
int orc_set_seeds(void *h, const uint32_t *s) { Batch *B = (Batch *)h; for (size_t w = 0; w < B->worlds.size(); w++) B->worlds[w].seed = s[w]; return 1; }
int orc_get_seeds(void *h, uint32_t *s) { Batch *B = (Batch *)h; for (size_t w = 0; w < B->worlds.size(); w++) s[w] = B->worlds[w].seed; return 1; }
int orc_get_enabled(void *h, int *en)
{
    Batch *B = (Batch *)h;
    for (size_t w = 0; w < B->worlds.size(); w++) for (int i = 0; i < B->nbody; i++) en[w * B->nbody + i] = !(B->worlds[w].bodies[i].flags & BF_DISABLED);
    return 1;
}

int orc_step(void *h, double hstep, int nsteps)
{
    Batch *B = (Batch *)h;
    for (int s = 0; s < nsteps; s++)
        for (size_t w = 0; w < B->worlds.size(); w++) world_step(*B, B->worlds[w], (Real)hstep);
    return 1;
}

int orc_step_range(void *h, double hstep, int nsteps, int w0, int w1)
{
    Batch *B = (Batch *)h;
    for (int s = 0; s < nsteps; s++) for (int w = w0; w < w1; w++) world_step(*B, B->worlds[w], (Real)hstep);
    return 1;
}

int orc_get_pairs(void *h, int world, int *pairs, int cap)
{
    World &W = ((Batch *)h)->worlds[world];
    int n = (int)W.pairs.size() / 2;
    for (int i = 0; i < n && i < cap; i++) { pairs[2 * i] = W.pairs[2 * i]; pairs[2 * i + 1] = W.pairs[2 * i + 1]; }
    return n;
}

int orc_get_contacts(void *h, int world, Real *geom7, int *g12, int cap)
{
    World &W = ((Batch *)h)->worlds[world];
    int n = (int)W.last_cg.size();
    for (int i = 0; i < n && i < cap; i++) {
        const OrcContactGeom &c = W.last_cg[i];
        for (int k = 0; k < 3; k++) { geom7[7 * i + k] = c.pos[k]; geom7[7 * i + 3 + k] = c.normal[k]; }
        geom7[7 * i + 6] = c.depth;
        g12[2 * i] = W.contact_g[2 * i]; g12[2 * i + 1] = W.contact_g[2 * i + 1];
    }
    return n;
}
int orc_get_ray_hits(void *h, int world, Real *geom7, int *g12, int cap)
{
    World &W = ((Batch *)h)->worlds[world];
    int n = (int)W.ray_cg.size();
    for (int i = 0; i < n && i < cap; i++) {
        const OrcContactGeom &c = W.ray_cg[i];
        for (int k = 0; k < 3; k++) { geom7[7 * i + k] = c.pos[k]; geom7[7 * i + 3 + k] = c.normal[k]; }
        geom7[7 * i + 6] = c.depth;
        g12[2 * i] = W.ray_g[2 * i]; g12[2 * i + 1] = W.ray_g[2 * i + 1];
    }
    return n;
}
int orc_get_islands(void *h, int world, int *label)
{
    Batch *B = (Batch *)h; World &W = B->worlds[world];
    for (int i = 0; i < B->nbody; i++) label[i] = i < (int)W.island_label.size() ? W.island_label[i] : -1;
    return W.island_count;
}
// per island of the world's last step, in solve order: bodies, rows, sweeps executed; returns the number of islands
int orc_get_island_log(void *h, int world, int *out3, int cap)
{
    World &W = ((Batch *)h)->worlds[world];
    const int n = (int)W.isl_log.size() / 3;
    for (int i = 0; i < n && i < cap; i++) for (int k = 0; k < 3; k++) out3[3 * i + k] = W.isl_log[3 * i + k];
    return n;
}
int orc_get_stats(void *h, int world, OdebStats *out)
{
    World &W = ((Batch *)h)->worlds[world];
    for (int k = 0; k < 4; k++) out->v[k] = W.stats[k];
    return 1;
}

} // extern "C"

#include "orc_extra.h"
