/*
 * ode_b200.h -- C-ABI of the B200-native ODE step path (batch entry points).
 *
 * Boundary: plain C, pointers and sizes only, no torch / CUDA types.  One shared library per
 * precision, like the reference (libode_b200_single.so: odeb_real == float,
 * libode_b200_double.so: odeb_real == double; reference: include/ode/precision.h.in:9-15,
 * include/ode/common.h:56-65).
 *
 * The reference steps ONE world through
 *     dSpaceCollide(space, data, &nearCallback)      include/ode/collision_space.h:49-64
 *     dWorldQuickStep(world, h)                      include/ode/objects.h:419-425
 *     dJointGroupEmpty(contactgroup)                 include/ode/objects.h:1720 (ode.cpp:1325)
 * with user code (the near-callback, ode/demo/demo_boxstack.cpp:131-176) in the middle.  These entry
 * points step W independent worlds of identical topology with that whole sequence resident on the
 * GPU; the near-callback is replaced by the declarative contact policy in OdebWorldParams (the same
 * fields a callback fills into dSurfaceParameters, include/ode/contact.h:55-103).
 *
 * The classic per-object API (dWorldCreate/dBodyCreate/dJointCreateContact/dSpaceCollide/
 * dWorldQuickStep ...) is declared in include/ode_b200_classic.h and layered on the same kernels.
 *
 * The structs below are also read by oracle/ (test infrastructure) so that reference, oracle and
 * CUDA path are fed byte-identical scene descriptions.
 */
#ifndef ODE_B200_H
#define ODE_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(ODEB_DOUBLE)
typedef double odeb_real;
#else
typedef float odeb_real;
#endif

/* geom classes: numbering of the reference (include/ode/collision.h:881-902) */
enum { ODEB_SPHERE = 0, ODEB_BOX = 1, ODEB_CAPSULE = 2, ODEB_CYLINDER = 3, ODEB_PLANE = 4, ODEB_RAY = 5 };
/* joint types: numbering of the reference dJointType (include/ode/common.h:406-426) */
enum { ODEB_JOINT_BALL = 1, ODEB_JOINT_HINGE = 2, ODEB_JOINT_SLIDER = 3, ODEB_JOINT_CONTACT = 4, ODEB_JOINT_UNIVERSAL = 5, ODEB_JOINT_HINGE2 = 6, ODEB_JOINT_FIXED = 7, ODEB_JOINT_AMOTOR = 9, ODEB_JOINT_LMOTOR = 10 };
/* broadphase flavours: which reference space's callback stream is reproduced (as a set).
 * HASH: dxHashSpace::collide collision_space.cpp:421-614 -- AABB-overlapping pairs that the cell walk brings together (a pair whose
 *       shared cells are all reached through differently wrapped hash addresses is not reported, see DESIGN.md "hash space");
 * SAP:  dxSAPSpace::collide collision_sapspace.cpp:428-496;  SIMPLE: dxSimpleSpace::collide collision_space.cpp:245-267, every AABB-overlapping pair. */
enum { ODEB_SPACE_HASH = 0, ODEB_SPACE_SAP = 1, ODEB_SPACE_SIMPLE = 2 };

/* contact surface mode bits: include/ode/contact.h:34-52 */
enum {
    ODEB_CONTACT_MU2 = 0x001, ODEB_CONTACT_FDIR1 = 0x002, ODEB_CONTACT_BOUNCE = 0x004,
    ODEB_CONTACT_SOFT_ERP = 0x008, ODEB_CONTACT_SOFT_CFM = 0x010, ODEB_CONTACT_MOTION1 = 0x020,
    ODEB_CONTACT_MOTION2 = 0x040, ODEB_CONTACT_MOTIONN = 0x080, ODEB_CONTACT_SLIP1 = 0x100,
    ODEB_CONTACT_SLIP2 = 0x200, ODEB_CONTACT_ROLLING = 0x400, ODEB_CONTACT_APPROX1_1 = 0x1000, ODEB_CONTACT_APPROX1_2 = 0x2000,
    ODEB_CONTACT_APPROX1_N = 0x4000, ODEB_CONTACT_APPROX1 = 0x7000
};

/* body flag bits a scene may set (ode/src/objects.h:72-87 keeps these internal; meaning identical) */
enum {
    ODEB_BODY_NO_GRAVITY = 1, ODEB_BODY_NO_GYRO = 2, ODEB_BODY_DISABLED = 4,
    ODEB_BODY_FINITE_ROTATION = 8,
    ODEB_BODY_KINEMATIC = 16    /* dBodySetKinematic (ode.cpp:837-842): inverse mass and inverse inertia are zero */
};

/* World-level parameters (dWorldSet*; defaults ode/src/objects.cpp:37-121) + the contact policy. */
typedef struct OdebWorldParams {
    double gravity[3];          /* dWorldSetGravity */
    double erp;                 /* dWorldSetERP, default 0.2 */
    double cfm;                 /* dWorldSetCFM; < 0 selects the precision default 1e-5 / 1e-10 */
    int    num_iterations;      /* dWorldSetQuickStepNumIterations, default 20 */
    double sor_w;               /* dWorldSetQuickStepW, default 1.3 */
    double premature_exit_delta;/* dWorldSetQuickStepDynamicIterationParameters, default 1e-8 */
    double max_extra_factor;    /*   "   default 1.0 */
    double extra_iter_delta;    /*   "   default 1e-2 */
    double contact_max_vel;     /* dWorldSetContactMaxCorrectingVel, default +inf */
    double contact_surface_layer;/* dWorldSetContactSurfaceLayer, default 0 */
    int    auto_disable;        /* dWorldSetAutoDisableFlag */
    double adis_linear_thr;     /* dWorldSetAutoDisableLinearThreshold (speed, not squared), default 0.01 */
    double adis_angular_thr;    /* dWorldSetAutoDisableAngularThreshold, default 0.01 */
    int    adis_steps;          /* dWorldSetAutoDisableSteps, default 10 */
    double adis_time;           /* dWorldSetAutoDisableTime, default 0 */
    int    adis_samples;        /* dWorldSetAutoDisableAverageSamplesCount, default 1 */
    double linear_damping;      /* dWorldSetLinearDamping, default 0 */
    double angular_damping;     /* dWorldSetAngularDamping, default 0 */
    double linear_damping_thr;  /* dWorldSetLinearDampingThreshold, default 0.01 */
    double angular_damping_thr; /* dWorldSetAngularDampingThreshold, default 0.01 */
    double max_angular_speed;   /* dWorldSetMaxAngularSpeed, default +inf */
    /* --- contact policy: what the reference's near-callback does (demo_boxstack.cpp:131-176) */
    int    space_type;          /* ODEB_SPACE_HASH | ODEB_SPACE_SAP */
    int    max_contacts;        /* dCollide flags & NUMC_MASK, 1..8 */
    int    skip_connected;      /* if both bodies: dAreConnectedExcluding(b1,b2,dJointTypeContact) -> skip */
    int    surf_mode;           /* dSurfaceParameters.mode */
    double mu, mu2, bounce, bounce_vel, soft_erp, soft_cfm;
    double motion1, motion2, motionN, slip1, slip2;
    double rho, rho2, rhoN;     /* rolling / spinning friction, used with ODEB_CONTACT_ROLLING (contact.h:64-66) */
    int    hash_levels_set;     /* 0: the hash space's default levels -3..10 (collision_space.cpp:387-388); 1: dHashSpaceSetLevels values below */
    int    hash_minlevel, hash_maxlevel;
    /* --- device capacities per world; 0 = automatic (pairs: min(all pairs, 16 per geom); contacts: pairs * max_contacts, capped at
     *     2..4 per geom and max_contacts).  A step that needs more fails with "capacity overflow" (odeb_step / odeb_sync /
     *     odeb_get_state return 0) and leaves the body state as the last complete step wrote it. */
    int    max_pairs, max_contacts_per_world;
} OdebWorldParams;

typedef struct OdebBodyDesc {
    double mass;                /* dMass.mass */
    double inertia[9];          /* dMass.I, 3x3 row-major about the centre of mass (c == 0) */
    int    flags;               /* ODEB_BODY_* */
} OdebBodyDesc;

typedef struct OdebGeomDesc {
    int    type;                /* ODEB_SPHERE: p[0]=radius; ODEB_BOX: p[0..2]=side lengths;
                                   ODEB_CAPSULE: p[0]=radius, p[1]=length; ODEB_PLANE: p[0..3]=a,b,c,d;
                                   ODEB_CYLINDER: p[0]=radius, p[1]=length (collides with planes, spheres and boxes:
                                   collision_cylinder_plane.cpp, collision_cylinder_sphere.cpp, collision_cylinder_box.cpp; capsules and other
                                   cylinders pass through it as in the reference's default build, which has no collider for them);
                                   ODEB_RAY: p[0]=length, along the geom's local z axis (ray.cpp).  Rays are sensors: their hits are
                                   reported by odeb_get_ray_hits and never become contact joints */
    int    body;                /* body index in the world, -1 = static (dGeomSetBody not called) */
    double p[4];
    uint32_t category_bits, collide_bits; /* dGeomSetCategoryBits / dGeomSetCollideBits */
    int    has_offset;          /* dGeomSetOffsetPosition / dGeomSetOffsetQuaternion (collision_kernel.cpp:455-466): pose of the geom */
    double offset_pos[3];       /*   relative to its body; composite bodies. Ignored for geoms without a body.                      */
    double offset_quat[4];
} OdebGeomDesc;

typedef struct OdebJointDesc {
    int    type;                /* ODEB_JOINT_BALL | ODEB_JOINT_HINGE | ODEB_JOINT_UNIVERSAL | ODEB_JOINT_FIXED (dJointSetFixed at the template pose)
                                   | ODEB_JOINT_SLIDER (dJointSetSliderAxis(axis1) at the template pose; stops / motor = axis-1 entries, lengths) */
    int    body1, body2;        /* dJointAttach(j, body1, body2); -1 = the static environment */
    double anchor[3];           /* dJointSet*Anchor, world frame at the template pose */
    double axis1[3], axis2[3];  /* dJointSetHingeAxis / dJointSetUniversalAxis1,2 */
    double lo_stop[3], hi_stop[3]; /* dParamLoStop/HiStop (axis 1, axis 2, axis 3: the dParam*, dParam*2, dParam*3 groups); defaults -inf/+inf */
    double vel[3], fmax[3];     /* dParamVel, dParamFMax; default 0 */
    double fudge_factor[3], bounce[3], stop_erp[3], stop_cfm[3]; /* <0 = keep defaults */
    double susp_erp, susp_cfm;  /* ODEB_JOINT_HINGE2: dParamSuspensionERP / dParamSuspensionCFM, <0 = the world's ERP / CFM.
                                   Hinge2 = dJointSetHinge2Anchor(anchor) + dJointSetHinge2Axes(axis1, axis2), needs both bodies;
                                   stops / motor of axis 1 and the motor of axis 2 are the [0] / [1] entries above */
    /* ODEB_JOINT_LMOTOR (lmotor.cpp) / ODEB_JOINT_AMOTOR (amotor.cpp): dJointSet{L,A}MotorNumAxes, dJointSetAMotorMode (0 = dAMotorUser,
     * 1 = dAMotorEuler: 3 axes, axis 0 relative to body 1 and axis 2 relative to body 2), dJointSet{L,A}MotorAxis(anum, rel, x, y, z) with the
     * axis given in the world frame at the template pose and rel = 0 global / 1 body1 / 2 body2, dJointSetAMotorAngle (user mode).
     * Motor velocity / force and the stops of axis k are the [k] entries above (stops only act on AMotor axes). */
    int    motor_num, motor_mode, motor_rel[3];
    double motor_axis[3][3];
    double motor_angle[3];
} OdebJointDesc;

/* per-world counters of dWorldQuickStepIterationCount_DynamicAdjustmentStatistics
 * (include/ode/objects.h:538-576): iteration_count, premature_exits, prolonged_execs, full_extra_execs */
typedef struct OdebStats { uint32_t v[4]; } OdebStats;

typedef struct OdebBatch OdebBatch;

/* Create W worlds from one template. The template pose is body_pos/body_quat ([nbody][3], [nbody][4],
 * doubles); joint anchors/axes are bound at that pose exactly as dJointSet*Anchor/Axis would.
 * Returns NULL on failure (message via odeb_last_error). Fails loudly when no CUDA device is present. */
OdebBatch *odeb_create(const OdebWorldParams *wp,
                       int nbody, const OdebBodyDesc *bodies, const double *body_pos, const double *body_quat,
                       int ngeom, const OdebGeomDesc *geoms,
                       int njoint, const OdebJointDesc *joints,
                       int nworlds, int device);
void odeb_destroy(OdebBatch *);
const char *odeb_last_error(void);

/* bulk state, host buffers laid out [world][body][3|4] in odeb_real. NULL pointers are skipped.
 * Setting a quaternion normalises it and rebuilds R like dBodySetQuaternion (ode.cpp:330-343). */
int odeb_set_state(OdebBatch *, const odeb_real *pos, const odeb_real *quat, const odeb_real *lvel, const odeb_real *avel);
int odeb_get_state(OdebBatch *, odeb_real *pos, odeb_real *quat, odeb_real *lvel, odeb_real *avel);
/* odeb_get_state is a blocking call: like odeb_sync it returns 0 ("capacity overflow ...") when a step queued by odeb_step_async
 * truncated pairs / contacts / rows; no body is moved by such a step or by the steps queued behind it. */
/* The packed body state on the DEVICE (pos [W*NB*3], quat [W*NB*4], lvel [W*NB*3], avel [W*NB*3], odeb_real) in a buffer owned by
 * the batch, written on the batch's stream; valid until the next call.  For GPU-side consumers (NCCL gather of observations). */
int odeb_pack_state_device(OdebBatch *, void **dev_ptr, size_t *bytes);
/* device addresses of the per-world counters: stats [W*4] uint32 (dynamic-iteration counters), seeds [W] uint32 */
int odeb_device_counters(OdebBatch *, void **stats, void **seeds);
/* run the batch on the caller's CUDA stream (cudaStream_t) so that the caller's events order against the steps */
int odeb_set_stream(OdebBatch *, void *stream);
/* Page-locked host memory for the arrays handed to odeb_set/get_state and odeb_add_force: with it the transfers run straight between the
 * device and the caller's arrays (pageable arrays work too, through an internal pinned staging buffer and one extra host copy). */
void *odeb_alloc_host(size_t bytes);
void odeb_free_host(void *);
/* dBodyAddForce / dBodyAddTorque for every body: added to the accumulators consumed by the next step.  Arrays from odeb_alloc_host are
 * read asynchronously, in stream order: do not overwrite them before the next blocking call (odeb_get_state, odeb_sync, odeb_step);
 * pageable arrays are copied before the call returns. */
int odeb_add_force(OdebBatch *, const odeb_real *force, const odeb_real *torque);
/* per-world dRandSetSeed / dRandGetSeed (ode/src/misc.cpp:52-61) */
int odeb_set_seeds(OdebBatch *, const uint32_t *seeds);
int odeb_get_seeds(OdebBatch *, uint32_t *seeds);
int odeb_get_enabled(OdebBatch *, int *enabled /* [world][body] */);

/* Solver order mode.
 *   ODEB_MODE_REPLAY (default): constraint order, dRand-driven reorders and island sequencing of the reference are replayed
 *     exactly (list mechanics of dJointAttach / dxProcessIslands); results are bit-comparable with the reference.  One world
 *     is inherently serial in this mode, so it is meant for batches of small worlds.
 *   ODEB_MODE_CANONICAL: the large-world path (single worlds of 10^3..10^5 bodies).  Pair set, contacts, island membership
 *     and island numbering are those of the reference; inside an island bodies are ordered by descending creation index and
 *     joints by ascending id (permanent joints, then contacts in creation order).  Rows are taken in GROUPS (the rows of the contacts
 *     of one geom pair, or of one permanent joint: they act on the same two bodies).  Once per step the groups of an island are
 *     coloured so that groups of one colour touch disjoint bodies -- rounds of "every uncoloured group whose
 *     (odeb_canon_key(seed, island, 0, first row), first row) exceeds that of all its uncoloured neighbours takes the smallest colour
 *     none of its coloured neighbours holds".  In phase k (the 8 sweeps from sweep 8k on, where the reference reorders) the colours are
 *     visited in ascending (odeb_canon_key(seed, ~0, k, colour), colour) order (odebi_canon_colour_ranks), the groups of a colour by
 *     ascending first row, the rows of a group in row order (a contact's normal row right before its friction rows).  The world's
 *     dRand seed is advanced by the draws the reference's Fisher-Yates reorders would have consumed.  Groups of one colour are
 *     relaxed side by side on the GPU with exactly the result of the sequential sweep in that order, so the CUDA path is
 *     bit-identical to the oracle run in the same mode (oracle: orc_set_solver_mode).  The order is this library's own (the
 *     reference's is the attach order, then random permutations): trajectories differ from the reference the way two random
 *     reorderings of the reference differ from each other (profiles/r2_callback_order.txt); everything a step decides before the
 *     solver runs (pair set, contacts, islands) is the reference's.
 *     Requires nworlds == 1.  At most ODEB_CANON_COLOURS - 2 = 254 row groups (contact pairs / permanent joints) may act on one body (the static environment has no body
 *     and does not count); a step that meets more fails like a capacity overflow, with the state of the last complete step kept. */
enum { ODEB_MODE_REPLAY = 0, ODEB_MODE_CANONICAL = 1 };
int odeb_set_solver_mode(OdebBatch *, int mode);

/* Joint feedback (dJointSetFeedback / dJointFeedback, include/ode/common.h:439-447; quickstep.cpp:3108-3182): when enabled,
 * every joint of every world behaves as if a dJointFeedback were attached. odeb_get_feedback returns, for one world, the
 * joints in id order (the njoint permanent joints of the template, then the contact joints of the last step in creation
 * order): out12 = f1[3] t1[3] f2[3] t2[3], state = 0 (joint not stepped: its island was disabled), 1 (body 1 only: f2/t2
 * are not written by the reference either), 2 (both bodies). Returns the number of joints (may exceed cap), -1 on error. */
int odeb_enable_feedback(OdebBatch *, int on);
int odeb_get_feedback(OdebBatch *, int world, odeb_real *out12, int *state, int cap);

/* Checkpoint / resume: everything a step reads from earlier steps (body state, rotation matrices, force accumulators, enable
 * flags, auto-disable history, per-world dRand seeds, iteration statistics) as one host blob. Restoring into a batch created
 * from the same template continues bit-identically. (The reference offers only the text dump dWorldExportDIF,
 * include/ode/export-dif.h:33.) */
size_t odeb_snapshot_size(OdebBatch *);
int odeb_snapshot(OdebBatch *, void *buf, size_t cap);
int odeb_restore(OdebBatch *, const void *buf, size_t bytes);

#if defined(__CUDACC__)
#define ODEB_HD __host__ __device__
#else
#define ODEB_HD
#endif
/* key of row `row` of island `island` for the reorder at sweep 8*phase in ODEB_MODE_CANONICAL (murmur3 finaliser);
 * `seed` = the world's dRand seed at the start of the step. Exported as odeb_canon_key; the inline body is shared with
 * the kernels and with oracle/. */
uint32_t odeb_canon_key(uint32_t seed, uint32_t island, uint32_t phase, uint32_t row);
/* Test hook: out[i] = the library's single-precision atan2 of (y[i], x[i]) -- the fdlibm algorithm the reference's host libm (glibc <= 2.40)
 * uses, restated operation by operation (odeb_math.cuh) -- evaluated on the host (on_device = 0) or by a kernel on the current device
 * (on_device = 1).  Tests compare it bit for bit with the host's atan2f.  Returns 1 on success. */
int odeb_test_atan2f(const float *y, const float *x, float *out, int n, int on_device);
static inline ODEB_HD uint32_t odebi_canon_key(uint32_t seed, uint32_t island, uint32_t phase, uint32_t row)
{
    uint32_t x = seed ^ (island * 0x9E3779B9u) ^ (phase * 0x85EBCA6Bu) ^ (row * 0xC2B2AE35u);
    x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
    return x;
}
/* rank[c] = position of colour c in the visiting order of phase `phase`: colours by ascending (odeb_canon_key(seed, ~0, phase, c), c).
 * ODEB_CANON_COLOURS = the most colours a colouring may use: one more than the most row groups that can meet on one body. */
#define ODEB_CANON_COLOURS 256
static inline ODEB_HD void odebi_canon_colour_ranks(uint32_t seed, uint32_t phase, int rank[ODEB_CANON_COLOURS])
{
    uint32_t key[ODEB_CANON_COLOURS];
    for (int c = 0; c < ODEB_CANON_COLOURS; c++) key[c] = odebi_canon_key(seed, 0xffffffffu, phase, (uint32_t)c);
    for (int c = 0; c < ODEB_CANON_COLOURS; c++) {
        int r = 0;
        for (int d = 0; d < ODEB_CANON_COLOURS; d++) if (key[d] < key[c] || (key[d] == key[c] && d < c)) r++;
        rank[c] = r;
    }
}

/* nsteps x { dSpaceCollide + contact policy ; dWorldQuickStep(h) ; dJointGroupEmpty }.
 * Returns 1 on success, 0 on failure (like dWorldQuickStep).  On a capacity overflow (more pairs / contacts / rows than the
 * batch was created for) the failing step and every step queued behind it leave positions, orientations and velocities as
 * the last complete step wrote them (dRand seeds, iteration counters and auto-disable history of those steps are undefined):
 * raise OdebWorldParams.max_pairs / max_contacts_per_world, recreate, restore. */
int odeb_step(OdebBatch *, double h, int nsteps);
/* same, asynchronous on the batch's stream; pair with odeb_sync */
int odeb_step_async(OdebBatch *, double h, int nsteps);
int odeb_sync(OdebBatch *);
/* nsteps steps, each timed by its own CUDA event pair on the batch's stream; when flush_bytes > 0 the L2 is
 * flushed by a memset of that size before every step, outside the timed intervals. *total_ms = summed time. */
int odeb_timed_steps(OdebBatch *, double h, int nsteps, size_t flush_bytes, double *total_ms);
/* number of kernel launches issued by this batch so far */
uint64_t odeb_launch_count(const OdebBatch *);
/* name of the SOR kernel the next step will launch (the library picks one of several bit-identical kernels per batch:
 * k_solve, k_solve5<P>, k_solve_bl; ODEB_SOLVER=v4|p2|p4|p8|bl in the environment forces one) */
const char *odeb_solver_kernel(OdebBatch *);
/* device-side duration (ms) of the solver kernel launches accumulated since the last call; resets */
double odeb_solver_ms(OdebBatch *, int *launches);
void   odeb_enable_timing(OdebBatch *, int on);

/* observables of the most recent step, one world at a time (parity tests).
 * pairs: (geomA, geomB) in callback order; contacts: [pos3, normal3, depth] + (g1, g2);
 * islands: label per body (-1 = not stepped), labels numbered in processing order. */
int odeb_get_pairs(OdebBatch *, int world, int *pairs, int cap);
int odeb_get_contacts(OdebBatch *, int world, odeb_real *geom7, int *g12, int cap);
/* hits of the world's ray geoms (ODEB_RAY) in the last step's collide pass, in pair order: [pos3, normal3, depth = distance along the ray]
 * + (g1, g2); what a near-callback that treats rays as sensors collects with dCollide (ray.cpp) */
int odeb_get_ray_hits(OdebBatch *, int world, odeb_real *geom7, int *g12, int cap);
/* Range-sensor read-out for every world at once: range[world][ray] = distance to the nearest hit of the ray (ray geoms numbered in geom
 * order, odeb_num_rays of them), +inf when it saw nothing; hit_geom[world][ray] = the geom it hit or -1 (either pointer may be NULL).
 * Returns the number of rays per world. */
int odeb_num_rays(OdebBatch *);
int odeb_get_ray_ranges(OdebBatch *, odeb_real *range, int *hit_geom);
int odeb_get_islands(OdebBatch *, int world, int *label_per_body);
int odeb_get_stats(OdebBatch *, int world, OdebStats *out);
/* totals over all worlds for the most recent step:
 * [pairs, contacts, rows, islands, sweeps (sum over islands), row-sweeps (sum over islands of rows x sweeps)] */
int odeb_get_totals(OdebBatch *, uint64_t out[6]);

#ifdef __cplusplus
}
#endif
#endif
