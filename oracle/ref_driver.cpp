// ref_driver.cpp -- TEST INFRASTRUCTURE ONLY (never linked or loaded by ode_b200/).
//
// Headless driver around the UNMODIFIED reference library: it is compiled together with the
// reference's own sources (oracle/Makefile) into oracle/_ref/libode_ref_{single,double}.so and
// exposes the same scene-description C interface as include/ode_b200.h with the prefix ref_, so
// the parity tests can feed byte-identical scenes to the reference and to the CUDA path.
//
// Per step and per world it does what a user of the reference does (ode/demo/demo_boxstack.cpp:510-513):
//     dRandSetSeed(seed[w]); dSpaceCollide(space, ctx, cb); dWorldQuickStep(world, h); dJointGroupEmpty(group)
// The near-callback only buffers the pairs it is handed; after dSpaceCollide returns they are put
// in canonical order (geom index of o1 < geom index of o2, sorted lexicographically) and only then
// dCollide + dJointCreateContact + dJointAttach run (the callback is user code).
// dWorldQuickStep is spelled out from its two internal calls (ode/src/ode.cpp:1847-1864) so that the
// island arrays (ode/src/util.h:259-279) can be read between them.
#include <ode/ode.h>
#include "config.h"
#include "objects.h"
#include "joints/joints.h"
#include "util.h"
#include "quickstep.h"
#include "collision_kernel.h"
#include <vector>
#include <algorithm>
#include <cstring>
#include <cstdint>
#include <cmath>

#if defined(dDOUBLE)
#define ODEB_DOUBLE 1
#endif
#include "../include/ode_b200.h"

namespace {

struct RefWorld {
    dWorldID world;
    dSpaceID space;
    dJointGroupID group;
    std::vector<dBodyID> bodies;
    std::vector<dGeomID> geoms;
    std::vector<dJointID> joints;
    uint32_t seed;
    dWorldQuickStepIterationCount_DynamicAdjustmentStatistics stats;
    // observables of the last step
    std::vector<int> pairs;              // 2 per pair
    std::vector<dContactGeom> ray_hits; std::vector<int> ray_g;   // hits of ray geoms (no joints are made from them)
    std::vector<dContactGeom> contacts;
    std::vector<int> contact_g;          // 2 per contact
    std::vector<int> island_label;       // per body
    int island_count;
    // joint feedback (dJointSetFeedback): storage for the permanent joints and for this step's contact joints
    std::vector<dJointFeedback> fb_perm, fb_contact;
    std::vector<dJointID> cjoints;
    std::vector<dReal> fb; std::vector<int> fb_state;   // last step: 12 reals per joint id + 0/1/2 (not stepped / body 1 / both)
};

struct RefBatch {
    OdebWorldParams wp;
    int nbody, ngeom, njoint, nworlds;
    int feedback;
    std::vector<RefWorld> worlds;
};

struct CbCtx { RefWorld *w; std::vector<std::pair<int,int> > buf; };

int g_init = 0;

void near_cb(void *data, dGeomID o1, dGeomID o2)
{
    CbCtx *c = (CbCtx *)data;
    int i1 = (int)(intptr_t)dGeomGetData(o1), i2 = (int)(intptr_t)dGeomGetData(o2);
    if (i1 > i2) std::swap(i1, i2);
    c->buf.push_back(std::make_pair(i1, i2));
}

void apply_world_params(dWorldID w, const OdebWorldParams &p)
{
    dWorldSetGravity(w, (dReal)p.gravity[0], (dReal)p.gravity[1], (dReal)p.gravity[2]);
    dWorldSetERP(w, (dReal)p.erp);
    if (p.cfm >= 0) dWorldSetCFM(w, (dReal)p.cfm);
    dWorldSetQuickStepNumIterations(w, p.num_iterations);
    dWorldSetQuickStepW(w, (dReal)p.sor_w);
    dReal d1 = (dReal)p.premature_exit_delta, f = (dReal)p.max_extra_factor, d2 = (dReal)p.extra_iter_delta;
    dWorldSetQuickStepDynamicIterationParameters(w, &d1, &f, &d2);
    dWorldSetContactMaxCorrectingVel(w, (dReal)p.contact_max_vel);
    dWorldSetContactSurfaceLayer(w, (dReal)p.contact_surface_layer);
    dWorldSetAutoDisableFlag(w, p.auto_disable);
    dWorldSetAutoDisableLinearThreshold(w, (dReal)p.adis_linear_thr);
    dWorldSetAutoDisableAngularThreshold(w, (dReal)p.adis_angular_thr);
    dWorldSetAutoDisableSteps(w, p.adis_steps);
    dWorldSetAutoDisableTime(w, (dReal)p.adis_time);
    dWorldSetAutoDisableAverageSamplesCount(w, (unsigned)p.adis_samples);
    dWorldSetLinearDampingThreshold(w, (dReal)p.linear_damping_thr);
    dWorldSetAngularDampingThreshold(w, (dReal)p.angular_damping_thr);
    dWorldSetLinearDamping(w, (dReal)p.linear_damping);
    dWorldSetAngularDamping(w, (dReal)p.angular_damping);
    dWorldSetMaxAngularSpeed(w, (dReal)p.max_angular_speed);
}

void set_limot(dJointID j, int type, const OdebJointDesc &d)
{
    for (int a = 0; a < ((type == ODEB_JOINT_AMOTOR || type == ODEB_JOINT_LMOTOR) ? 3 : (type == ODEB_JOINT_UNIVERSAL || type == ODEB_JOINT_HINGE2) ? 2 : 1); a++) {
        int grp = a * dParamGroup;
        void (*setp)(dJointID, int, dReal) = (type == ODEB_JOINT_AMOTOR) ? dJointSetAMotorParam : (type == ODEB_JOINT_LMOTOR) ? dJointSetLMotorParam : (type == ODEB_JOINT_HINGE) ? dJointSetHingeParam : (type == ODEB_JOINT_SLIDER) ? dJointSetSliderParam : (type == ODEB_JOINT_HINGE2) ? dJointSetHinge2Param : dJointSetUniversalParam;
        // the reference documents setting lo, hi, lo again when lo > hi may be transiently true
        setp(j, dParamLoStop + grp, (dReal)d.lo_stop[a]);
        setp(j, dParamHiStop + grp, (dReal)d.hi_stop[a]);
        setp(j, dParamLoStop + grp, (dReal)d.lo_stop[a]);
        setp(j, dParamVel + grp, (dReal)d.vel[a]);
        setp(j, dParamFMax + grp, (dReal)d.fmax[a]);
        if (d.fudge_factor[a] >= 0) setp(j, dParamFudgeFactor + grp, (dReal)d.fudge_factor[a]);
        if (d.bounce[a] >= 0) setp(j, dParamBounce + grp, (dReal)d.bounce[a]);
        if (d.stop_erp[a] >= 0) setp(j, dParamStopERP + grp, (dReal)d.stop_erp[a]);
        if (d.stop_cfm[a] >= 0) setp(j, dParamStopCFM + grp, (dReal)d.stop_cfm[a]);
    }
}

} // namespace

extern "C" {

void *ref_create(const OdebWorldParams *wp,
                 int nbody, const OdebBodyDesc *bodies, const double *body_pos, const double *body_quat,
                 int ngeom, const OdebGeomDesc *geoms,
                 int njoint, const OdebJointDesc *joints,
                 int nworlds, int /*device*/)
{
    if (!g_init) { dInitODE2(0); dAllocateODEDataForThread(dAllocateMaskAll); g_init = 1; }
    RefBatch *B = new RefBatch;
    B->wp = *wp; B->nbody = nbody; B->ngeom = ngeom; B->njoint = njoint; B->nworlds = nworlds; B->feedback = 0;
    B->worlds.resize(nworlds);
    for (int wi = 0; wi < nworlds; wi++) {
        RefWorld &W = B->worlds[wi];
        W.world = dWorldCreate();
        W.space = (wp->space_type == ODEB_SPACE_SAP) ? dSweepAndPruneSpaceCreate(0, dSAP_AXES_XYZ)
                : (wp->space_type == ODEB_SPACE_SIMPLE) ? dSimpleSpaceCreate(0) : dHashSpaceCreate(0);
        if (wp->space_type == ODEB_SPACE_HASH && wp->hash_levels_set) dHashSpaceSetLevels(W.space, wp->hash_minlevel, wp->hash_maxlevel);
        W.group = dJointGroupCreate(0);
        W.seed = 0; W.island_count = 0;
        memset(&W.stats, 0, sizeof(W.stats));
        W.stats.struct_size = sizeof(W.stats);
        dWorldAttachQuickStepDynamicIterationStatisticsSink(W.world, &W.stats);
        apply_world_params(W.world, *wp);
        for (int i = 0; i < nbody; i++) {
            dBodyID b = dBodyCreate(W.world);
            dMass m;
            const double *I = bodies[i].inertia;
            dMassSetParameters(&m, (dReal)bodies[i].mass, 0, 0, 0, (dReal)I[0], (dReal)I[4], (dReal)I[8],
                               (dReal)I[1], (dReal)I[2], (dReal)I[5]);
            dBodySetMass(b, &m);
            dBodySetPosition(b, (dReal)body_pos[3*i], (dReal)body_pos[3*i+1], (dReal)body_pos[3*i+2]);
            dQuaternion q = { (dReal)body_quat[4*i], (dReal)body_quat[4*i+1], (dReal)body_quat[4*i+2], (dReal)body_quat[4*i+3] };
            dBodySetQuaternion(b, q);
            int fl = bodies[i].flags;
            if (fl & ODEB_BODY_NO_GRAVITY) dBodySetGravityMode(b, 0);
            if (fl & ODEB_BODY_NO_GYRO) dBodySetGyroscopicMode(b, 0);
            if (fl & ODEB_BODY_FINITE_ROTATION) dBodySetFiniteRotationMode(b, 1);
            if (fl & ODEB_BODY_DISABLED) dBodyDisable(b);
            if (fl & ODEB_BODY_KINEMATIC) dBodySetKinematic(b);
            W.bodies.push_back(b);
        }
        for (int i = 0; i < ngeom; i++) {
            const OdebGeomDesc &g = geoms[i];
            dGeomID id = 0;
            switch (g.type) {
            case ODEB_SPHERE: id = dCreateSphere(W.space, (dReal)g.p[0]); break;
            case ODEB_BOX: id = dCreateBox(W.space, (dReal)g.p[0], (dReal)g.p[1], (dReal)g.p[2]); break;
            case ODEB_CAPSULE: id = dCreateCapsule(W.space, (dReal)g.p[0], (dReal)g.p[1]); break;
            case ODEB_PLANE: id = dCreatePlane(W.space, (dReal)g.p[0], (dReal)g.p[1], (dReal)g.p[2], (dReal)g.p[3]); break;
            case ODEB_CYLINDER: id = dCreateCylinder(W.space, (dReal)g.p[0], (dReal)g.p[1]); break;
            case ODEB_RAY: id = dCreateRay(W.space, (dReal)g.p[0]); break;
            default: return 0;
            }
            if (g.body >= 0) dGeomSetBody(id, W.bodies[g.body]);
            if (g.body >= 0 && g.has_offset) {
                dGeomSetOffsetPosition(id, (dReal)g.offset_pos[0], (dReal)g.offset_pos[1], (dReal)g.offset_pos[2]);
                dQuaternion q = { (dReal)g.offset_quat[0], (dReal)g.offset_quat[1], (dReal)g.offset_quat[2], (dReal)g.offset_quat[3] };
                dGeomSetOffsetQuaternion(id, q);
            }
            dGeomSetCategoryBits(id, g.category_bits);
            dGeomSetCollideBits(id, g.collide_bits);
            dGeomSetData(id, (void *)(intptr_t)i);
            W.geoms.push_back(id);
        }
        for (int i = 0; i < njoint; i++) {
            const OdebJointDesc &d = joints[i];
            dBodyID b1 = d.body1 >= 0 ? W.bodies[d.body1] : 0, b2 = d.body2 >= 0 ? W.bodies[d.body2] : 0;
            dJointID j = 0;
            if (d.type == ODEB_JOINT_BALL) {
                j = dJointCreateBall(W.world, 0); dJointAttach(j, b1, b2);
                dJointSetBallAnchor(j, (dReal)d.anchor[0], (dReal)d.anchor[1], (dReal)d.anchor[2]);
            } else if (d.type == ODEB_JOINT_HINGE2) {
                if (!b1 || !b2) return 0;
                j = dJointCreateHinge2(W.world, 0); dJointAttach(j, b1, b2);
                dJointSetHinge2Anchor(j, (dReal)d.anchor[0], (dReal)d.anchor[1], (dReal)d.anchor[2]);
                dVector3 a1 = { (dReal)d.axis1[0], (dReal)d.axis1[1], (dReal)d.axis1[2] }, a2 = { (dReal)d.axis2[0], (dReal)d.axis2[1], (dReal)d.axis2[2] };
                dJointSetHinge2Axes(j, a1, a2);
                set_limot(j, d.type, d);
                if (d.susp_erp >= 0) dJointSetHinge2Param(j, dParamSuspensionERP, (dReal)d.susp_erp);
                if (d.susp_cfm >= 0) dJointSetHinge2Param(j, dParamSuspensionCFM, (dReal)d.susp_cfm);
            } else if (d.type == ODEB_JOINT_SLIDER) {
                j = dJointCreateSlider(W.world, 0); dJointAttach(j, b1, b2);
                dJointSetSliderAxis(j, (dReal)d.axis1[0], (dReal)d.axis1[1], (dReal)d.axis1[2]);
                set_limot(j, d.type, d);
            } else if (d.type == ODEB_JOINT_LMOTOR) {
                j = dJointCreateLMotor(W.world, 0); dJointAttach(j, b1, b2);
                dJointSetLMotorNumAxes(j, d.motor_num);
                for (int k = 0; k < d.motor_num; k++) dJointSetLMotorAxis(j, k, d.motor_rel[k], (dReal)d.motor_axis[k][0], (dReal)d.motor_axis[k][1], (dReal)d.motor_axis[k][2]);
                set_limot(j, d.type, d);
            } else if (d.type == ODEB_JOINT_AMOTOR) {
                j = dJointCreateAMotor(W.world, 0); dJointAttach(j, b1, b2);
                dJointSetAMotorMode(j, d.motor_mode);
                dJointSetAMotorNumAxes(j, d.motor_num);
                for (int k = 0; k < d.motor_num; k++) {
                    if (d.motor_mode == dAMotorEuler && k == 1) continue;
                    dJointSetAMotorAxis(j, k, d.motor_rel[k], (dReal)d.motor_axis[k][0], (dReal)d.motor_axis[k][1], (dReal)d.motor_axis[k][2]);
                }
                if (d.motor_mode == dAMotorUser) for (int k = 0; k < d.motor_num; k++) dJointSetAMotorAngle(j, k, (dReal)d.motor_angle[k]);
                set_limot(j, d.type, d);
            } else if (d.type == ODEB_JOINT_FIXED) {
                j = dJointCreateFixed(W.world, 0); dJointAttach(j, b1, b2);
                dJointSetFixed(j);
            } else if (d.type == ODEB_JOINT_HINGE) {
                j = dJointCreateHinge(W.world, 0); dJointAttach(j, b1, b2);
                dJointSetHingeAnchor(j, (dReal)d.anchor[0], (dReal)d.anchor[1], (dReal)d.anchor[2]);
                dJointSetHingeAxis(j, (dReal)d.axis1[0], (dReal)d.axis1[1], (dReal)d.axis1[2]);
                set_limot(j, d.type, d);
            } else if (d.type == ODEB_JOINT_UNIVERSAL) {
                j = dJointCreateUniversal(W.world, 0); dJointAttach(j, b1, b2);
                dJointSetUniversalAnchor(j, (dReal)d.anchor[0], (dReal)d.anchor[1], (dReal)d.anchor[2]);
                dJointSetUniversalAxis1(j, (dReal)d.axis1[0], (dReal)d.axis1[1], (dReal)d.axis1[2]);
                dJointSetUniversalAxis2(j, (dReal)d.axis2[0], (dReal)d.axis2[1], (dReal)d.axis2[2]);
                set_limot(j, d.type, d);
            } else return 0;
            W.joints.push_back(j);
        }
        W.island_label.assign(nbody, -1);
    }
    return B;
}

void ref_destroy(void *h)
{
    RefBatch *B = (RefBatch *)h;
    for (size_t i = 0; i < B->worlds.size(); i++) {
        RefWorld &W = B->worlds[i];
        dJointGroupDestroy(W.group);
        dSpaceDestroy(W.space);
        dWorldDestroy(W.world);
    }
    delete B;
}

int ref_set_state(void *h, const dReal *pos, const dReal *quat, const dReal *lvel, const dReal *avel)
{
    RefBatch *B = (RefBatch *)h;
    for (int w = 0; w < B->nworlds; w++) for (int i = 0; i < B->nbody; i++) {
        dBodyID b = B->worlds[w].bodies[i];
        size_t k = (size_t)w * B->nbody + i;
        if (pos) dBodySetPosition(b, pos[3*k], pos[3*k+1], pos[3*k+2]);
        if (quat) { dQuaternion q = { quat[4*k], quat[4*k+1], quat[4*k+2], quat[4*k+3] }; dBodySetQuaternion(b, q); }
        if (lvel) dBodySetLinearVel(b, lvel[3*k], lvel[3*k+1], lvel[3*k+2]);
        if (avel) dBodySetAngularVel(b, avel[3*k], avel[3*k+1], avel[3*k+2]);
    }
    return 1;
}

int ref_get_state(void *h, dReal *pos, dReal *quat, dReal *lvel, dReal *avel)
{
    RefBatch *B = (RefBatch *)h;
    for (int w = 0; w < B->nworlds; w++) for (int i = 0; i < B->nbody; i++) {
        dBodyID b = B->worlds[w].bodies[i];
        size_t k = (size_t)w * B->nbody + i;
        if (pos) memcpy(pos + 3*k, dBodyGetPosition(b), 3 * sizeof(dReal));
        if (quat) memcpy(quat + 4*k, dBodyGetQuaternion(b), 4 * sizeof(dReal));
        if (lvel) memcpy(lvel + 3*k, dBodyGetLinearVel(b), 3 * sizeof(dReal));
        if (avel) memcpy(avel + 3*k, dBodyGetAngularVel(b), 3 * sizeof(dReal));
    }
    return 1;
}

int ref_add_force(void *h, const dReal *force, const dReal *torque)
{
    RefBatch *B = (RefBatch *)h;
    for (int w = 0; w < B->nworlds; w++) for (int i = 0; i < B->nbody; i++) {
        dBodyID b = B->worlds[w].bodies[i];
        size_t k = (size_t)w * B->nbody + i;
        if (force) dBodyAddForce(b, force[3*k], force[3*k+1], force[3*k+2]);
        if (torque) dBodyAddTorque(b, torque[3*k], torque[3*k+1], torque[3*k+2]);
    }
    return 1;
}

int ref_enable_feedback(void *h, int on)
{
    RefBatch *B = (RefBatch *)h; B->feedback = on != 0;
    if (!on) for (int w = 0; w < B->nworlds; w++) for (size_t k = 0; k < B->worlds[w].joints.size(); k++) dJointSetFeedback(B->worlds[w].joints[k], 0);
    return 1;
}
int ref_get_feedback(void *h, int world, dReal *out12, int *state, int cap)
{
    RefBatch *B = (RefBatch *)h; const RefWorld &W = B->worlds[world];
    const int n = (int)W.fb_state.size();
    for (int i = 0; i < n && i < cap; i++) { for (int k = 0; k < 12; k++) out12[12 * i + k] = W.fb[12 * (size_t)i + k]; state[i] = W.fb_state[i]; }
    return n;
}
int ref_set_seeds(void *h, const uint32_t *s) { RefBatch *B = (RefBatch *)h; for (int w = 0; w < B->nworlds; w++) B->worlds[w].seed = s[w]; return 1; }
int ref_get_seeds(void *h, uint32_t *s) { RefBatch *B = (RefBatch *)h; for (int w = 0; w < B->nworlds; w++) s[w] = B->worlds[w].seed; return 1; }
int ref_get_enabled(void *h, int *en)
{
    RefBatch *B = (RefBatch *)h;
    for (int w = 0; w < B->nworlds; w++) for (int i = 0; i < B->nbody; i++) en[(size_t)w * B->nbody + i] = dBodyIsEnabled(B->worlds[w].bodies[i]);
    return 1;
}

static void ref_collide_world(RefBatch *B, RefWorld &W)
{
    const OdebWorldParams &p = B->wp;
    CbCtx ctx; ctx.w = &W;
    dSpaceCollide(W.space, &ctx, &near_cb);
    std::sort(ctx.buf.begin(), ctx.buf.end());
    W.pairs.clear(); W.contacts.clear(); W.contact_g.clear(); W.cjoints.clear(); W.ray_hits.clear(); W.ray_g.clear();
    dContact contact[9];
    for (size_t k = 0; k < ctx.buf.size(); k++) {
        int i1 = ctx.buf[k].first, i2 = ctx.buf[k].second;
        W.pairs.push_back(i1); W.pairs.push_back(i2);
        dGeomID o1 = W.geoms[i1], o2 = W.geoms[i2];
        dBodyID b1 = dGeomGetBody(o1), b2 = dGeomGetBody(o2);
        if (p.skip_connected && b1 && b2 && dAreConnectedExcluding(b1, b2, dJointTypeContact)) continue;
        if (!b1 && !b2) continue;
        int n = dCollide(o1, o2, p.max_contacts, &contact[0].geom, sizeof(dContact));
        if (dGeomGetClass(o1) == dRayClass || dGeomGetClass(o2) == dRayClass) {      // sensor policy: record the hit, no contact joint
            for (int i = 0; i < n; i++) { W.ray_hits.push_back(contact[i].geom); W.ray_g.push_back(i1); W.ray_g.push_back(i2); }
            continue;
        }
        for (int i = 0; i < n; i++) {
            dSurfaceParameters &s = contact[i].surface;
            memset(&s, 0, sizeof(s));
            s.mode = p.surf_mode; s.mu = (dReal)p.mu; s.mu2 = (dReal)p.mu2;
            s.bounce = (dReal)p.bounce; s.bounce_vel = (dReal)p.bounce_vel;
            s.soft_erp = (dReal)p.soft_erp; s.soft_cfm = (dReal)p.soft_cfm;
            s.motion1 = (dReal)p.motion1; s.motion2 = (dReal)p.motion2; s.motionN = (dReal)p.motionN;
            s.slip1 = (dReal)p.slip1; s.slip2 = (dReal)p.slip2;
            s.rho = (dReal)p.rho; s.rho2 = (dReal)p.rho2; s.rhoN = (dReal)p.rhoN;
            dJointID c = dJointCreateContact(W.world, W.group, &contact[i]);
            dJointAttach(c, b1, b2);
            W.cjoints.push_back(c);
            W.contacts.push_back(contact[i].geom);
            W.contact_g.push_back(i1); W.contact_g.push_back(i2);
        }
    }
}

static int ref_quickstep_world(RefBatch *B, RefWorld &W, dReal h)
{
    // == dWorldQuickStep (ode/src/ode.cpp:1847-1864), with the island arrays read in between
    dxWorldProcessIslandsInfo islandsinfo;
    if (!dxReallocateWorldProcessContext(W.world, islandsinfo, h, &dxEstimateQuickStepMemoryRequirements)) return 0;
    W.island_label.assign(B->nbody, -1);
    W.island_count = (int)islandsinfo.GetIslandsCount();
    {
        const unsigned *sizes = islandsinfo.GetIslandSizes();
        dxBody *const *bodies = islandsinfo.GetBodiesArray();
        for (int is = 0; is < W.island_count; is++) {
            unsigned nb = sizes[2 * is];
            for (unsigned k = 0; k < nb; k++, bodies++) {
                for (int i = 0; i < B->nbody; i++) if (W.bodies[i] == *bodies) { W.island_label[i] = is; break; }
            }
        }
    }
    if (!dxProcessIslands(W.world, islandsinfo, h, &dxQuickStepIsland, &dxEstimateQuickStepMaxCallCount)) return 0;
    return 1;
}

int ref_step(void *h, double hstep, int nsteps)
{
    RefBatch *B = (RefBatch *)h;
    for (int s = 0; s < nsteps; s++) {
        for (int w = 0; w < B->nworlds; w++) {
            RefWorld &W = B->worlds[w];
            dRandSetSeed(W.seed);
            ref_collide_world(B, W);
            const dReal NOTSET = (dReal)-12345.678;
            if (B->feedback) {   // a dJointFeedback on every joint, pre-filled with a marker the step overwrites
                dJointFeedback mark;
                for (int k = 0; k < 4; k++) { mark.f1[k] = mark.t1[k] = mark.f2[k] = mark.t2[k] = NOTSET; }
                W.fb_perm.assign(W.joints.size(), mark); W.fb_contact.assign(W.cjoints.size(), mark);
                for (size_t k = 0; k < W.joints.size(); k++) dJointSetFeedback(W.joints[k], &W.fb_perm[k]);
                for (size_t k = 0; k < W.cjoints.size(); k++) dJointSetFeedback(W.cjoints[k], &W.fb_contact[k]);
            }
            if (!ref_quickstep_world(B, W, (dReal)hstep)) return 0;
            if (B->feedback) {
                const size_t np = W.joints.size(), n = np + W.cjoints.size();
                W.fb.assign(12 * n, 0); W.fb_state.assign(n, 0);
                for (size_t i = 0; i < n; i++) {
                    const dJointFeedback &f = i < np ? W.fb_perm[i] : W.fb_contact[i - np];
                    if (f.f1[0] == NOTSET) continue;
                    const bool two = f.f2[0] != NOTSET;
                    W.fb_state[i] = two ? 2 : 1;
                    for (int k = 0; k < 3; k++) {
                        W.fb[12 * i + k] = f.f1[k]; W.fb[12 * i + 3 + k] = f.t1[k];
                        if (two) { W.fb[12 * i + 6 + k] = f.f2[k]; W.fb[12 * i + 9 + k] = f.t2[k]; }
                    }
                }
            }
            dJointGroupEmpty(W.group);
            W.seed = (uint32_t)dRandGetSeed();
        }
    }
    return 1;
}

/* timing-only variant: the plain public-API loop (natural callback order, dWorldQuickStep itself) */
static void near_cb_direct(void *data, dGeomID o1, dGeomID o2)
{
    std::pair<RefBatch *, RefWorld *> *c = (std::pair<RefBatch *, RefWorld *> *)data;
    const OdebWorldParams &p = c->first->wp;
    RefWorld &W = *c->second;
    dBodyID b1 = dGeomGetBody(o1), b2 = dGeomGetBody(o2);
    if (p.skip_connected && b1 && b2 && dAreConnectedExcluding(b1, b2, dJointTypeContact)) return;
    if (!b1 && !b2) return;
    dContact contact[8];
    int n = dCollide(o1, o2, p.max_contacts, &contact[0].geom, sizeof(dContact));
    for (int i = 0; i < n; i++) {
        dSurfaceParameters &s = contact[i].surface;
        memset(&s, 0, sizeof(s));
        s.mode = p.surf_mode; s.mu = (dReal)p.mu; s.mu2 = (dReal)p.mu2;
        s.bounce = (dReal)p.bounce; s.bounce_vel = (dReal)p.bounce_vel;
        s.soft_erp = (dReal)p.soft_erp; s.soft_cfm = (dReal)p.soft_cfm;
        s.motion1 = (dReal)p.motion1; s.motion2 = (dReal)p.motion2; s.motionN = (dReal)p.motionN;
        s.slip1 = (dReal)p.slip1; s.slip2 = (dReal)p.slip2;
        s.rho = (dReal)p.rho; s.rho2 = (dReal)p.rho2; s.rhoN = (dReal)p.rhoN;
        dJointID cj = dJointCreateContact(W.world, W.group, &contact[i]);
        dJointAttach(cj, b1, b2);
    }
}

/* The reference's threaded stepper, set up exactly like ode/demo/demo_crash.cpp:275-279 / :636-640: one
 * multi-threaded implementation served by a pool of k threads, assigned to every world of the batch.  Timing baseline only: this
 * mode changes the SOR sweep order (quickstep.cpp's multi-threaded LCP iteration), so it is never used as a parity reference.
 * k <= 0 returns the worlds to the self-threaded default and frees the pool. */
static dThreadingImplementationID g_threading = 0;
static dThreadingThreadPoolID g_pool = 0;
int ref_set_threads(void *h, int k)
{
    RefBatch *B = (RefBatch *)h;
    if (g_threading) {
        for (size_t w = 0; w < B->worlds.size(); w++) dWorldSetStepThreadingImplementation(B->worlds[w].world, NULL, NULL);
        dThreadingImplementationShutdownProcessing(g_threading);
        dThreadingFreeThreadPool(g_pool);
        dThreadingFreeImplementation(g_threading);
        g_threading = 0; g_pool = 0;
    }
    if (k <= 0) return 1;
    g_threading = dThreadingAllocateMultiThreadedImplementation();
    if (!g_threading) return 0;
    g_pool = dThreadingAllocateThreadPool((unsigned)k, 0, dAllocateFlagBasicData, NULL);
    if (!g_pool) return 0;
    dThreadingThreadPoolServeMultiThreadedImplementation(g_pool, g_threading);
    for (size_t w = 0; w < B->worlds.size(); w++)
        dWorldSetStepThreadingImplementation(B->worlds[w].world, dThreadingImplementationGetFunctions(g_threading), g_threading);
    return 1;
}

int ref_step_plain(void *h, double hstep, int nsteps, int world_begin, int world_end)
{
    RefBatch *B = (RefBatch *)h;
    for (int s = 0; s < nsteps; s++) {
        for (int w = world_begin; w < world_end; w++) {
            RefWorld &W = B->worlds[w];
            dRandSetSeed(W.seed);
            std::pair<RefBatch *, RefWorld *> ctx(B, &W);
            dSpaceCollide(W.space, &ctx, &near_cb_direct);
            if (!dWorldQuickStep(W.world, (dReal)hstep)) return 0;
            dJointGroupEmpty(W.group);
            W.seed = (uint32_t)dRandGetSeed();
        }
    }
    return 1;
}

int ref_get_pairs(void *h, int world, int *pairs, int cap)
{
    RefWorld &W = ((RefBatch *)h)->worlds[world];
    int n = (int)W.pairs.size() / 2;
    for (int i = 0; i < n && i < cap; i++) { pairs[2*i] = W.pairs[2*i]; pairs[2*i+1] = W.pairs[2*i+1]; }
    return n;
}

int ref_get_ray_hits(void *h, int world, dReal *geom7, int *g12, int cap)
{
    RefWorld &W = ((RefBatch *)h)->worlds[world];
    int n = (int)W.ray_hits.size();
    for (int i = 0; i < n && i < cap; i++) {
        const dContactGeom &c = W.ray_hits[i];
        for (int k = 0; k < 3; k++) { geom7[7 * i + k] = c.pos[k]; geom7[7 * i + 3 + k] = c.normal[k]; }
        geom7[7 * i + 6] = c.depth;
        g12[2 * i] = W.ray_g[2 * i]; g12[2 * i + 1] = W.ray_g[2 * i + 1];
    }
    return n;
}
int ref_get_contacts(void *h, int world, dReal *geom7, int *g12, int cap)
{
    RefWorld &W = ((RefBatch *)h)->worlds[world];
    int n = (int)W.contacts.size();
    for (int i = 0; i < n && i < cap; i++) {
        const dContactGeom &c = W.contacts[i];
        geom7[7*i+0] = c.pos[0]; geom7[7*i+1] = c.pos[1]; geom7[7*i+2] = c.pos[2];
        geom7[7*i+3] = c.normal[0]; geom7[7*i+4] = c.normal[1]; geom7[7*i+5] = c.normal[2];
        geom7[7*i+6] = c.depth;
        g12[2*i] = W.contact_g[2*i]; g12[2*i+1] = W.contact_g[2*i+1];
    }
    return n;
}

int ref_get_islands(void *h, int world, int *label)
{
    RefBatch *B = (RefBatch *)h; RefWorld &W = B->worlds[world];
    for (int i = 0; i < B->nbody; i++) label[i] = W.island_label[i];
    return W.island_count;
}

int ref_get_stats(void *h, int world, OdebStats *out)
{
    RefWorld &W = ((RefBatch *)h)->worlds[world];
    out->v[0] = W.stats.iteration_count; out->v[1] = W.stats.premature_exits;
    out->v[2] = W.stats.prolonged_execs; out->v[3] = W.stats.full_extra_execs;
    return 1;
}

/* --- golden-vector helpers: individual reference functions on explicit inputs ------------------ */

/* dCollide on two free-standing geoms (no space, no bodies): type/params as OdebGeomDesc, pose = pos3 + R(3x4) */
int ref_collide_pair(int type1, const dReal *p1, const dReal *pos1, const dReal *R1,
                     int type2, const dReal *p2, const dReal *pos2, const dReal *R2,
                     int flags, dReal *geom7, int cap)
{
    if (!g_init) { dInitODE2(0); dAllocateODEDataForThread(dAllocateMaskAll); g_init = 1; }
    dGeomID g[2];
    const int types[2] = { type1, type2 };
    const dReal *pp[2] = { p1, p2 }, *ps[2] = { pos1, pos2 }, *Rs[2] = { R1, R2 };
    for (int k = 0; k < 2; k++) {
        switch (types[k]) {
        case ODEB_SPHERE: g[k] = dCreateSphere(0, pp[k][0]); break;
        case ODEB_BOX: g[k] = dCreateBox(0, pp[k][0], pp[k][1], pp[k][2]); break;
        case ODEB_CAPSULE: g[k] = dCreateCapsule(0, pp[k][0], pp[k][1]); break;
        case ODEB_PLANE: g[k] = dCreatePlane(0, pp[k][0], pp[k][1], pp[k][2], pp[k][3]); break;
        default: return -1;
        }
        if (types[k] != ODEB_PLANE) {
            dGeomSetPosition(g[k], ps[k][0], ps[k][1], ps[k][2]);
            dGeomSetRotation(g[k], Rs[k]);
        }
    }
    dContactGeom c[16];
    int n = dCollide(g[0], g[1], flags, c, sizeof(dContactGeom));
    for (int i = 0; i < n && i < cap; i++) {
        geom7[7*i+0] = c[i].pos[0]; geom7[7*i+1] = c[i].pos[1]; geom7[7*i+2] = c[i].pos[2];
        geom7[7*i+3] = c[i].normal[0]; geom7[7*i+4] = c[i].normal[1]; geom7[7*i+5] = c[i].normal[2];
        geom7[7*i+6] = c[i].depth;
    }
    dGeomDestroy(g[0]); dGeomDestroy(g[1]);
    return n;
}

void ref_geom_aabb(int type, const dReal *p, const dReal *pos, const dReal *R, dReal *aabb6)
{
    if (!g_init) { dInitODE2(0); dAllocateODEDataForThread(dAllocateMaskAll); g_init = 1; }
    dGeomID g;
    switch (type) {
    case ODEB_SPHERE: g = dCreateSphere(0, p[0]); break;
    case ODEB_BOX: g = dCreateBox(0, p[0], p[1], p[2]); break;
    case ODEB_CAPSULE: g = dCreateCapsule(0, p[0], p[1]); break;
    default: g = dCreatePlane(0, p[0], p[1], p[2], p[3]); break;
    }
    if (type != ODEB_PLANE) { dGeomSetPosition(g, pos[0], pos[1], pos[2]); dGeomSetRotation(g, R); }
    dGeomGetAABB(g, aabb6);
    dGeomDestroy(g);
}

/* the reference's own contact row builder, called the way tests/friction.cpp:69-173 calls it */
int ref_contact_rows(int mode, dReal mu, dReal mu2, const dReal *cpos, const dReal *cnormal, dReal depth, const dReal *fdir1,
                     const dReal *pos1, const dReal *pos2, dReal fps, dReal erp, dReal *rows48, int *findex3)
{
    if (!g_init) { dInitODE2(0); dAllocateODEDataForThread(dAllocateMaskAll); g_init = 1; }
    dWorldID world = dWorldCreate();
    dWorldSetCFM(world, 0);
    dBodyID b1 = dBodyCreate(world), b2 = dBodyCreate(world);
    dBodySetPosition(b1, pos1[0], pos1[1], pos1[2]);
    dBodySetPosition(b2, pos2[0], pos2[1], pos2[2]);
    dContact c;
    memset(&c, 0, sizeof(c));
    c.surface.mode = mode; c.surface.mu = mu; c.surface.mu2 = mu2;
    for (int k = 0; k < 3; k++) { c.geom.pos[k] = cpos[k]; c.geom.normal[k] = cnormal[k]; c.fdir1[k] = fdir1[k]; }
    c.geom.depth = depth;
    dJointID j = dJointCreateContact(world, 0, &c);
    dJointAttach(j, b1, b2);
    dxJoint::Info1 info1;
    j->getInfo1(&info1);
    for (int k = 0; k < 48; k++) rows48[k] = 0;
    for (int k = 0; k < 3; k++) findex3[k] = -1;
    j->getInfo2(fps, erp, 16, rows48, rows48 + 8, 16, rows48 + 6, rows48 + 14, findex3);
    int m = info1.m;
    dJointDestroy(j);
    dWorldDestroy(world);
    return m;
}

unsigned long ref_rand_next(uint32_t *seed) { dRandSetSeed(*seed); unsigned long r = dRand(); *seed = (uint32_t)dRandGetSeed(); return r; }
int ref_rand_int(uint32_t *seed, int n) { dRandSetSeed(*seed); int r = dRandInt(n); *seed = (uint32_t)dRandGetSeed(); return r; }
int ref_test_rand(void) { return dTestRand(); }

/* dMassSet* + dBodySetMass: returns mass, I(3x4), invMass, invI(3x4) as the reference computes them */
void ref_mass(int kind, dReal density_or_total, int total, const dReal *dims, dReal *mass_out, dReal *I12, dReal *invMass, dReal *invI12)
{
    if (!g_init) { dInitODE2(0); dAllocateODEDataForThread(dAllocateMaskAll); g_init = 1; }
    dMass m;
    if (kind == ODEB_BOX) { if (total) dMassSetBoxTotal(&m, density_or_total, dims[0], dims[1], dims[2]); else dMassSetBox(&m, density_or_total, dims[0], dims[1], dims[2]); }
    else if (kind == ODEB_SPHERE) { if (total) dMassSetSphereTotal(&m, density_or_total, dims[0]); else dMassSetSphere(&m, density_or_total, dims[0]); }
    else { if (total) dMassSetCapsuleTotal(&m, density_or_total, 3, dims[0], dims[1]); else dMassSetCapsule(&m, density_or_total, 3, dims[0], dims[1]); }
    dWorldID w = dWorldCreate(); dBodyID b = dBodyCreate(w);
    dBodySetMass(b, &m);
    *mass_out = m.mass; memcpy(I12, m.I, 12 * sizeof(dReal));
    *invMass = b->invMass; memcpy(invI12, b->invI, 12 * sizeof(dReal));
    dWorldDestroy(w);
}

int ref_sizeof_real(void) { return (int)sizeof(dReal); }

} // extern "C"
