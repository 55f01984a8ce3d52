/*
 * ode_b200_classic.h -- the classic per-object C API of the reference (ODE 0.16, the headers under include/ode) for the
 * hot path, implemented on the B200 kernels of libode_b200_{single,double}.so.
 *
 * Same symbol names, argument meaning, struct layouts and handle semantics as the reference, so a program written
 * against <ode/ode.h> that stays inside this subset links against libode_b200_* unchanged:
 *
 *     dSpaceCollide(space, data, &nearCallback)     include/ode/collision_space.h:49-64, ode/src/collision_space.cpp:779
 *        nearCallback: dCollide(o1, o2, N, &contact[0].geom, sizeof(dContact))      ode/src/collision_kernel.cpp:292
 *                      dJointCreateContact(world, group, &contact[i]); dJointAttach  ode/src/ode.cpp:1192, :1383
 *     dWorldQuickStep(world, h)                      include/ode/objects.h:419-425, ode/src/ode.cpp:1847
 *     dJointGroupEmpty(group)                        ode/src/ode.cpp:1325
 *
 * Where the work runs: broadphase (AABBs + pair set of the space's type), narrowphase (every dCollide result), island
 * building, constraint rows, the SOR-LCP sweeps and the integration all execute in CUDA kernels; the host keeps the
 * object graph (handles, joint attachment order, user data) and a mirror of the body state, because the reference's
 * getters return pointers into library storage (ode.cpp:413-484).  There is no CPU implementation of any of these
 * stages: without a CUDA device dWorldQuickStep returns 0 and dSpaceCollide reports through the error handler.
 *
 * Outside the subset (calls are not exported): dWorldStep, joints other than contact/ball/hinge/slider/universal/hinge2/fixed/amotor/lmotor, geoms other
 * than sphere/box/capsule/cylinder/plane/ray, nested spaces, per-body
 * auto-disable thresholds (the world's are used), SAP axis orders other than dSAP_AXES_XYZ.
 */
#ifndef ODE_B200_CLASSIC_H
#define ODE_B200_CLASSIC_H

#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include <math.h>

#ifdef __cplusplus
extern "C" {
#endif

/* include/ode/common.h:56-65, :270-275 */
#if defined(ODEB_DOUBLE) || defined(dDOUBLE)
typedef double dReal;
#else
typedef float dReal;
#endif
typedef dReal dVector3[4];
typedef dReal dVector4[4];
typedef dReal dMatrix3[4 * 3];
typedef dReal dQuaternion[4];
#ifndef dInfinity
#define dInfinity ((dReal)INFINITY)
#endif

/* include/ode/common.h:364-369 */
typedef struct dxWorld *dWorldID;
typedef struct dxSpace *dSpaceID;
typedef struct dxBody *dBodyID;
typedef struct dxGeom *dGeomID;
typedef struct dxJoint *dJointID;
typedef struct dxJointGroup *dJointGroupID;

/* include/ode/common.h:406-426 */
typedef enum {
    dJointTypeNone = 0, dJointTypeBall, dJointTypeHinge, dJointTypeSlider, dJointTypeContact, dJointTypeUniversal,
    dJointTypeHinge2, dJointTypeFixed, dJointTypeNull, dJointTypeAMotor, dJointTypeLMotor, dJointTypePlane2D,
    dJointTypePR, dJointTypePU, dJointTypePiston, dJointTypeDBall, dJointTypeDHinge, dJointTypeTransmission
} dJointType;

/* include/ode/common.h:439-487: joint parameter names, groups 1..3 */
enum {
    dParamLoStop = 0, dParamHiStop, dParamVel, dParamLoVel, dParamHiVel, dParamFMax, dParamFudgeFactor, dParamBounce,
    dParamCFM, dParamStopERP, dParamStopCFM, dParamSuspensionERP, dParamSuspensionCFM, dParamERP,
    dParamsInGroup,
    dParamGroup1 = 0x000, dParamLoStop1 = 0x000, dParamHiStop1, dParamVel1, dParamLoVel1, dParamHiVel1, dParamFMax1,
    dParamFudgeFactor1, dParamBounce1, dParamCFM1, dParamStopERP1, dParamStopCFM1, dParamSuspensionERP1, dParamSuspensionCFM1, dParamERP1,
    dParamGroup2 = 0x100, dParamLoStop2 = 0x100, dParamHiStop2, dParamVel2, dParamLoVel2, dParamHiVel2, dParamFMax2,
    dParamFudgeFactor2, dParamBounce2, dParamCFM2, dParamStopERP2, dParamStopCFM2, dParamSuspensionERP2, dParamSuspensionCFM2, dParamERP2,
    dParamGroup3 = 0x200, dParamLoStop3 = 0x200, dParamHiStop3, dParamVel3, dParamLoVel3, dParamHiVel3, dParamFMax3,
    dParamFudgeFactor3, dParamBounce3, dParamCFM3, dParamStopERP3, dParamStopCFM3, dParamSuspensionERP3, dParamSuspensionCFM3, dParamERP3,
    dParamGroup = 0x100
};

/* include/ode/contact.h:34-103 */
enum {
    dContactMu2 = 0x001, dContactAxisDep = 0x001, dContactFDir1 = 0x002, dContactBounce = 0x004, dContactSoftERP = 0x008,
    dContactSoftCFM = 0x010, dContactMotion1 = 0x020, dContactMotion2 = 0x040, dContactMotionN = 0x080, dContactSlip1 = 0x100,
    dContactSlip2 = 0x200, dContactRolling = 0x400,
    dContactApprox0 = 0x0000, dContactApprox1_1 = 0x1000, dContactApprox1_2 = 0x2000, dContactApprox1_N = 0x4000, dContactApprox1 = 0x7000
};
typedef struct dSurfaceParameters {
    int mode;
    dReal mu;
    dReal mu2;
    dReal rho, rho2, rhoN;
    dReal bounce, bounce_vel;
    dReal soft_erp, soft_cfm;
    dReal motion1, motion2, motionN;
    dReal slip1, slip2;
} dSurfaceParameters;
typedef struct dContactGeom {
    dVector3 pos;
    dVector3 normal;
    dReal depth;
    dGeomID g1, g2;
    int side1, side2;
} dContactGeom;
typedef struct dContact {
    dSurfaceParameters surface;
    dContactGeom geom;
    dVector3 fdir1;
} dContact;

/* include/ode/mass.h:100-140 (C view of struct dMass) */
typedef struct dMass {
    dReal mass;
    dVector3 c;
    dMatrix3 I;
} dMass;

/* include/ode/common.h:512-517 */
typedef struct dJointFeedback { dVector3 f1, t1, f2, t2; } dJointFeedback;

/* include/ode/collision.h:881-902, :743; include/ode/collision_space.h:49-64 */
enum { dSphereClass = 0, dBoxClass, dCapsuleClass, dCylinderClass, dPlaneClass, dRayClass };
#define CONTACTS_UNIMPORTANT 0x80000000
typedef void dNearCallback(void *data, dGeomID o1, dGeomID o2);
#define dSAP_AXES_XYZ ((0) | (1 << 2) | (2 << 4))

/* include/ode/objects.h:538-576 */
typedef struct {
    unsigned struct_size;
    uint32_t iteration_count, premature_exits, prolonged_execs, full_extra_execs;
} dWorldQuickStepIterationCount_DynamicAdjustmentStatistics;

/* ---- init (include/ode/odeinit.h:119,177,227; ode/src/ode.cpp:2289-2396) */
enum { dAllocateFlagBasicData = 0, dAllocateFlagCollisionData = 1, dAllocateMaskAll = ~0 };
void dInitODE(void);
int dInitODE2(unsigned int uiInitFlags);
int dAllocateODEDataForThread(unsigned int uiAllocateFlags);
void dCloseODE(void);
const char *dGetConfiguration(void);
int dCheckConfiguration(const char *token);

/* ---- random numbers (include/ode/misc.h; ode/src/misc.cpp:35-139) */
unsigned long dRand(void);
unsigned long dRandGetSeed(void);
void dRandSetSeed(unsigned long s);
int dRandInt(int n);
dReal dRandReal(void);

/* ---- mass (include/ode/mass.h; ode/src/mass.cpp) */
void dMassSetZero(dMass *);
void dMassSetParameters(dMass *, dReal themass, dReal cgx, dReal cgy, dReal cgz, dReal I11, dReal I22, dReal I33, dReal I12, dReal I13, dReal I23);
void dMassSetSphere(dMass *, dReal density, dReal radius);
void dMassSetSphereTotal(dMass *, dReal total_mass, dReal radius);
void dMassSetCapsule(dMass *, dReal density, int direction, dReal radius, dReal length);
void dMassSetCapsuleTotal(dMass *, dReal total_mass, int direction, dReal radius, dReal length);
void dMassSetCylinder(dMass *, dReal density, int direction, dReal radius, dReal length);
void dMassSetCylinderTotal(dMass *, dReal total_mass, int direction, dReal radius, dReal length);
void dMassSetBox(dMass *, dReal density, dReal lx, dReal ly, dReal lz);
void dMassSetBoxTotal(dMass *, dReal total_mass, dReal lx, dReal ly, dReal lz);
void dMassAdjust(dMass *, dReal newmass);

/* ---- rotation helpers (include/ode/rotation.h; ode/src/rotation.cpp) */
void dRSetIdentity(dMatrix3 R);
void dRFromAxisAndAngle(dMatrix3 R, dReal ax, dReal ay, dReal az, dReal angle);
void dQSetIdentity(dQuaternion q);
void dQFromAxisAndAngle(dQuaternion q, dReal ax, dReal ay, dReal az, dReal angle);
void dRfromQ(dMatrix3 R, const dQuaternion q);
void dQfromR(dQuaternion q, const dMatrix3 R);

/* ---- world (include/ode/objects.h; ode/src/ode.cpp:1582-2174) */
dWorldID dWorldCreate(void);
void dWorldDestroy(dWorldID);
void dWorldSetGravity(dWorldID, dReal x, dReal y, dReal z);
void dWorldGetGravity(dWorldID, dVector3 gravity);
void dWorldSetERP(dWorldID, dReal erp);
dReal dWorldGetERP(dWorldID);
void dWorldSetCFM(dWorldID, dReal cfm);
dReal dWorldGetCFM(dWorldID);
void dWorldSetQuickStepNumIterations(dWorldID, int num);
int dWorldGetQuickStepNumIterations(dWorldID);
void dWorldSetQuickStepW(dWorldID, dReal over_relaxation);
dReal dWorldGetQuickStepW(dWorldID);
void dWorldSetQuickStepDynamicIterationParameters(dWorldID, const dReal *ptr_iteration_premature_exit_delta,
                                                  const dReal *ptr_max_num_extra_factor, const dReal *ptr_extra_iteration_requirement_delta);
void dWorldGetQuickStepDynamicIterationParameters(dWorldID, dReal *out_iteration_premature_exit_delta,
                                                  dReal *out_max_num_extra_factor, dReal *out_extra_iteration_requirement_delta);
int dWorldAttachQuickStepDynamicIterationStatisticsSink(dWorldID, dWorldQuickStepIterationCount_DynamicAdjustmentStatistics *var_stats);
void dWorldSetContactMaxCorrectingVel(dWorldID, dReal vel);
dReal dWorldGetContactMaxCorrectingVel(dWorldID);
void dWorldSetContactSurfaceLayer(dWorldID, dReal depth);
dReal dWorldGetContactSurfaceLayer(dWorldID);
/* LIMITATION: auto-disable thresholds / steps / time, damping scales / thresholds and the maximum angular speed are WORLD-wide
 * here and apply to every body at step time.  The reference copies the world's current defaults into each body at dBodyCreate
 * (b->adis, b->dampingp), so there a later dWorldSet* call only affects bodies created afterwards; programs that change these
 * defaults between body creations must set the per-body values explicitly (dBodySetAutoDisable*, dBodySet*Damping). */
void dWorldSetAutoDisableFlag(dWorldID, int do_auto_disable);
int dWorldGetAutoDisableFlag(dWorldID);
void dWorldSetAutoDisableLinearThreshold(dWorldID, dReal linear_threshold);
void dWorldSetAutoDisableAngularThreshold(dWorldID, dReal angular_threshold);
void dWorldSetAutoDisableSteps(dWorldID, int steps);
void dWorldSetAutoDisableTime(dWorldID, dReal time);
void dWorldSetAutoDisableAverageSamplesCount(dWorldID, unsigned int average_samples_count);
void dWorldSetLinearDamping(dWorldID, dReal scale);
void dWorldSetAngularDamping(dWorldID, dReal scale);
void dWorldSetDamping(dWorldID, dReal linear_scale, dReal angular_scale);
void dWorldSetLinearDampingThreshold(dWorldID, dReal threshold);
void dWorldSetAngularDampingThreshold(dWorldID, dReal threshold);
void dWorldSetMaxAngularSpeed(dWorldID, dReal max_speed);
int dWorldQuickStep(dWorldID, dReal stepsize);

/* ---- bodies (ode/src/ode.cpp:240-1150) */
dBodyID dBodyCreate(dWorldID);
void dBodyDestroy(dBodyID);
dWorldID dBodyGetWorld(dBodyID);
void dBodySetData(dBodyID, void *data);
void *dBodyGetData(dBodyID);
void dBodySetPosition(dBodyID, dReal x, dReal y, dReal z);
void dBodySetRotation(dBodyID, const dMatrix3 R);
void dBodySetQuaternion(dBodyID, const dQuaternion q);
void dBodySetLinearVel(dBodyID, dReal x, dReal y, dReal z);
void dBodySetAngularVel(dBodyID, dReal x, dReal y, dReal z);
const dReal *dBodyGetPosition(dBodyID);
const dReal *dBodyGetRotation(dBodyID);
const dReal *dBodyGetQuaternion(dBodyID);
const dReal *dBodyGetLinearVel(dBodyID);
const dReal *dBodyGetAngularVel(dBodyID);
void dBodySetMass(dBodyID, const dMass *mass);
void dBodyGetMass(dBodyID, dMass *mass);
void dBodySetKinematic(dBodyID);
void dBodySetDynamic(dBodyID);
int dBodyIsKinematic(dBodyID);
void dBodyAddForce(dBodyID, dReal fx, dReal fy, dReal fz);
void dBodyAddTorque(dBodyID, dReal fx, dReal fy, dReal fz);
const dReal *dBodyGetForce(dBodyID);
const dReal *dBodyGetTorque(dBodyID);
void dBodySetForce(dBodyID, dReal x, dReal y, dReal z);
void dBodySetTorque(dBodyID, dReal x, dReal y, dReal z);
void dBodyEnable(dBodyID);
void dBodyDisable(dBodyID);
int dBodyIsEnabled(dBodyID);
void dBodySetGravityMode(dBodyID, int mode);
int dBodyGetGravityMode(dBodyID);
void dBodySetGyroscopicMode(dBodyID, int enabled);
int dBodyGetGyroscopicMode(dBodyID);
void dBodySetFiniteRotationMode(dBodyID, int mode);
int dBodyGetFiniteRotationMode(dBodyID);
void dBodySetAutoDisableFlag(dBodyID, int do_auto_disable);
int dBodyGetAutoDisableFlag(dBodyID);
int dBodyGetNumJoints(dBodyID);

/* ---- joints (ode/src/ode.cpp:1158-1577, ode/src/joints/{ball,hinge,universal,contact}.cpp) */
dJointGroupID dJointGroupCreate(int max_size);
void dJointGroupDestroy(dJointGroupID);
void dJointGroupEmpty(dJointGroupID);
dJointID dJointCreateContact(dWorldID, dJointGroupID, const dContact *);
dJointID dJointCreateBall(dWorldID, dJointGroupID);
dJointID dJointCreateHinge(dWorldID, dJointGroupID);
dJointID dJointCreateHinge2(dWorldID, dJointGroupID);           /* joints/hinge2.cpp; needs two bodies */
void dJointSetHinge2Anchor(dJointID, dReal x, dReal y, dReal z);
void dJointSetHinge2Axes(dJointID, const dReal *axis1, const dReal *axis2);
void dJointSetHinge2Axis1(dJointID, dReal x, dReal y, dReal z);
void dJointSetHinge2Axis2(dJointID, dReal x, dReal y, dReal z);
void dJointSetHinge2Param(dJointID, int parameter, dReal value);
enum { dAMotorUser = 0, dAMotorEuler = 1 };                      /* include/ode/common.h:500-503 */
dJointID dJointCreateLMotor(dWorldID, dJointGroupID);           /* joints/lmotor.cpp */
void dJointSetLMotorNumAxes(dJointID, int num);
int dJointGetLMotorNumAxes(dJointID);
void dJointSetLMotorAxis(dJointID, int anum, int rel, dReal x, dReal y, dReal z);
void dJointGetLMotorAxis(dJointID, int anum, dVector3 result);
void dJointSetLMotorParam(dJointID, int parameter, dReal value);
dJointID dJointCreateAMotor(dWorldID, dJointGroupID);           /* joints/amotor.cpp */
void dJointSetAMotorMode(dJointID, int mode);
int dJointGetAMotorMode(dJointID);
void dJointSetAMotorNumAxes(dJointID, int num);
int dJointGetAMotorNumAxes(dJointID);
void dJointSetAMotorAxis(dJointID, int anum, int rel, dReal x, dReal y, dReal z);
void dJointGetAMotorAxis(dJointID, int anum, dVector3 result);
int dJointGetAMotorAxisRel(dJointID, int anum);
void dJointSetAMotorAngle(dJointID, int anum, dReal angle);
dReal dJointGetAMotorAngle(dJointID, int anum);                 /* user mode: the angle that was set */
void dJointSetAMotorParam(dJointID, int parameter, dReal value);
dJointID dJointCreateSlider(dWorldID, dJointGroupID);           /* joints/slider.cpp */
void dJointSetSliderAxis(dJointID, dReal x, dReal y, dReal z);
void dJointGetSliderAxis(dJointID, dVector3 result);
void dJointSetSliderParam(dJointID, int parameter, dReal value);
dReal dJointGetSliderPosition(dJointID);
dJointID dJointCreateFixed(dWorldID, dJointGroupID);            /* joints/fixed.cpp */
void dJointSetFixed(dJointID);
void dJointSetFixedParam(dJointID, int parameter, dReal value);
dReal dJointGetFixedParam(dJointID, int parameter);
dJointID dJointCreateUniversal(dWorldID, dJointGroupID);
void dJointDestroy(dJointID);
void dJointAttach(dJointID, dBodyID body1, dBodyID body2);
dBodyID dJointGetBody(dJointID, int index);
dJointType dJointGetType(dJointID);
void dJointSetData(dJointID, void *data);
void *dJointGetData(dJointID);
void dJointSetFeedback(dJointID, dJointFeedback *);
dJointFeedback *dJointGetFeedback(dJointID);
void dJointSetBallAnchor(dJointID, dReal x, dReal y, dReal z);
void dJointGetBallAnchor(dJointID, dVector3 result);
void dJointSetBallParam(dJointID, int parameter, dReal value);
void dJointSetHingeAnchor(dJointID, dReal x, dReal y, dReal z);
void dJointSetHingeAxis(dJointID, dReal x, dReal y, dReal z);
void dJointGetHingeAnchor(dJointID, dVector3 result);
void dJointGetHingeAxis(dJointID, dVector3 result);
void dJointSetHingeParam(dJointID, int parameter, dReal value);
void dJointSetUniversalAnchor(dJointID, dReal x, dReal y, dReal z);
void dJointSetUniversalAxis1(dJointID, dReal x, dReal y, dReal z);
void dJointSetUniversalAxis2(dJointID, dReal x, dReal y, dReal z);
void dJointGetUniversalAnchor(dJointID, dVector3 result);
void dJointSetUniversalParam(dJointID, int parameter, dReal value);
int dAreConnected(dBodyID, dBodyID);
int dAreConnectedExcluding(dBodyID, dBodyID, int joint_type);

/* ---- spaces and geoms (include/ode/collision_space.h, collision.h; ode/src/collision_kernel.cpp, collision_space.cpp) */
dSpaceID dSimpleSpaceCreate(dSpaceID space);
dSpaceID dHashSpaceCreate(dSpaceID space);
dSpaceID dSweepAndPruneSpaceCreate(dSpaceID space, int axisorder);
void dHashSpaceSetLevels(dSpaceID space, int minlevel, int maxlevel);
void dHashSpaceGetLevels(dSpaceID space, int *minlevel, int *maxlevel);
void dSpaceDestroy(dSpaceID);
void dSpaceSetCleanup(dSpaceID space, int mode);
int dSpaceGetCleanup(dSpaceID space);
void dSpaceAdd(dSpaceID, dGeomID);
void dSpaceRemove(dSpaceID, dGeomID);
int dSpaceGetNumGeoms(dSpaceID);
dGeomID dSpaceGetGeom(dSpaceID, int i);
/* CALLBACK ORDER: the reference hands pairs to the callback in the traversal order of its space (dGeomMoved re-inserts a moved
 * geom at the head of the space's list, collision_space.cpp:205-209; dxHashSpace walks hash chains, :504-584; dxSAPSpace walks its
 * sorted axis).  This library reports exactly the same SET of pairs (incl. the hash space's wrapped-address misses) but always in
 * ascending (index of o1, index of o2) order, o1 created before o2.  Constraint order follows contact-creation order, so an
 * application that creates its contacts directly in callback order sees a different (equally valid) SOR row order than on the
 * reference; an application that buffers the pairs and creates contacts in a canonical order (tests/classic_app.py,
 * oracle/ref_driver.cpp: sorted by geom index) gets identical contacts, islands and trajectories on both libraries.
 * profiles/r2_callback_order.txt holds the reference-vs-reference yardstick (natural against sorted callback order). */
void dSpaceCollide(dSpaceID space, void *data, dNearCallback *callback);
int dCollide(dGeomID o1, dGeomID o2, int flags, dContactGeom *contact, int skip);
dGeomID dCreateSphere(dSpaceID space, dReal radius);
dGeomID dCreateBox(dSpaceID space, dReal lx, dReal ly, dReal lz);
dGeomID dCreateCapsule(dSpaceID space, dReal radius, dReal length);
dGeomID dCreatePlane(dSpaceID space, dReal a, dReal b, dReal c, dReal d);
void dGeomDestroy(dGeomID);
void dGeomSetData(dGeomID, void *data);
void *dGeomGetData(dGeomID);
void dGeomSetBody(dGeomID, dBodyID);
dBodyID dGeomGetBody(dGeomID);
void dGeomSetPosition(dGeomID, dReal x, dReal y, dReal z);
void dGeomSetRotation(dGeomID, const dMatrix3 R);
void dGeomSetQuaternion(dGeomID, const dQuaternion Q);
const dReal *dGeomGetPosition(dGeomID);
const dReal *dGeomGetRotation(dGeomID);
void dGeomSetOffsetPosition(dGeomID, dReal x, dReal y, dReal z);      /* composite bodies: pose relative to the body */
void dGeomSetOffsetRotation(dGeomID, const dMatrix3 R);
void dGeomSetOffsetQuaternion(dGeomID, const dQuaternion q);
void dGeomClearOffset(dGeomID);
int dGeomIsOffset(dGeomID);
const dReal *dGeomGetOffsetPosition(dGeomID);
const dReal *dGeomGetOffsetRotation(dGeomID);
void dGeomGetAABB(dGeomID, dReal aabb[6]);
int dGeomGetClass(dGeomID);
void dGeomSetCategoryBits(dGeomID, unsigned long bits);
void dGeomSetCollideBits(dGeomID, unsigned long bits);
unsigned long dGeomGetCategoryBits(dGeomID);
unsigned long dGeomGetCollideBits(dGeomID);
dReal dGeomSphereGetRadius(dGeomID);
void dGeomBoxGetLengths(dGeomID, dVector3 result);
void dGeomCapsuleGetParams(dGeomID, dReal *radius, dReal *length);
dGeomID dCreateCylinder(dSpaceID, dReal radius, dReal length);   /* cylinder.cpp; colliders: plane, sphere, box, ray */
void dGeomCylinderSetParams(dGeomID, dReal radius, dReal length);
void dGeomCylinderGetParams(dGeomID, dReal *radius, dReal *length);
dGeomID dCreateRay(dSpaceID, dReal length);                      /* ray.cpp; colliders: sphere, box, capsule, plane, cylinder */
void dGeomRaySetLength(dGeomID, dReal length);
dReal dGeomRayGetLength(dGeomID);
void dGeomRaySet(dGeomID, dReal px, dReal py, dReal pz, dReal dx, dReal dy, dReal dz);   /* rays without a body */
void dGeomRayGet(dGeomID, dVector3 start, dVector3 dir);
void dGeomPlaneGetParams(dGeomID, dVector4 result);

#ifdef __cplusplus
}
#endif
#endif
