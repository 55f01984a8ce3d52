// orc_joints.h -- TEST INFRASTRUCTURE (oracle). Ball / hinge / universal joint row builders restated
// from ode/src/joints/{joint,ball,hinge,universal}.cpp. Included inside orc_world.cpp's namespace
// (uses World, Body, Joint, Limot and the row layout constants).
#ifndef ORC_JOINTS_H
#define ORC_JOINTS_H

static inline void qmul3(Real *qa, const Real *qb, const Real *qc)
{   // dQMultiply3 rotation.cpp:221-228
    qa[0] = qb[0] * qc[0] - qb[1] * qc[1] - qb[2] * qc[2] - qb[3] * qc[3];
    qa[1] = -qb[0] * qc[1] - qb[1] * qc[0] + qb[2] * qc[3] - qb[3] * qc[2];
    qa[2] = -qb[0] * qc[2] - qb[2] * qc[0] + qb[3] * qc[1] - qb[1] * qc[3];
    qa[3] = -qb[0] * qc[3] - qb[3] * qc[0] + qb[1] * qc[2] - qb[2] * qc[1];
}

// dRFrom2Axes rotation.cpp:94-133; returns false when the reference returns without writing R
static inline bool r_from_2axes(Real *R, Real ax, Real ay, Real az, Real bx, Real by, Real bz)
{
    Real l = RSQRT(ax * ax + ay * ay + az * az), k;
    if (l <= R_(0.0)) return false;
    l = rrecip(l); ax *= l; ay *= l; az *= l;
    k = ax * bx + ay * by + az * bz;
    bx -= k * ax; by -= k * ay; bz -= k * az;
    l = RSQRT(bx * bx + by * by + bz * bz);
    if (l <= R_(0.0)) return false;
    l = rrecip(l); bx *= l; by *= l; bz *= l;
    R[0] = ax; R[4] = ay; R[8] = az;
    R[1] = bx; R[5] = by; R[9] = bz;
    R[2] = -by * az + ay * bz; R[6] = -bz * ax + az * bx; R[10] = -bx * ay + ax * by;
    R[3] = R[7] = R[11] = R_(0.0);
    return true;
}

enum { ROWLEN = 16, C_J1L = 0, C_J1A = 3, C_RHS = 6, C_CFM = 7, C_J2L = 8, C_J2A = 11, C_LO = 14, C_HI = 15 };

// setBall joints/joint.cpp:111-157
static void set_ball(World &W, Joint &j, Real fps, Real erp, Real *row, const Real *anchor1, const Real *anchor2)
{
    Real a1[3], a2[3];
    Body &b0 = W.bodies[j.b0];
    row[C_J1L + 0] = 1; row[ROWLEN + C_J1L + 1] = 1; row[2 * ROWLEN + C_J1L + 2] = 1;
    mul0_331(a1, b0.R, anchor1);
    // dSetCrossMatrixMinus(J1 + JA, a1, rowskip) odemath.h:286-296
    row[C_J1A + 1] = +a1[2]; row[C_J1A + 2] = -a1[1];
    row[ROWLEN + C_J1A + 0] = -a1[2]; row[ROWLEN + C_J1A + 2] = +a1[0];
    row[2 * ROWLEN + C_J1A + 0] = +a1[1]; row[2 * ROWLEN + C_J1A + 1] = -a1[0];
    if (j.b1 >= 0) {
        Body &b1 = W.bodies[j.b1];
        row[C_J2L + 0] = -1; row[ROWLEN + C_J2L + 1] = -1; row[2 * ROWLEN + C_J2L + 2] = -1;
        mul0_331(a2, b1.R, anchor2);
        // dSetCrossMatrixPlus odemath.h:274-284
        row[C_J2A + 1] = -a2[2]; row[C_J2A + 2] = +a2[1];
        row[ROWLEN + C_J2A + 0] = +a2[2]; row[ROWLEN + C_J2A + 2] = -a2[0];
        row[2 * ROWLEN + C_J2A + 0] = -a2[1]; row[2 * ROWLEN + C_J2A + 1] = +a2[0];
    }
    Real k = fps * erp;
    if (j.b1 >= 0) {
        Body &b1 = W.bodies[j.b1];
        for (int t = 0; t < 3; t++) row[t * ROWLEN + C_RHS] = k * (a2[t] + b1.pos[t] - a1[t] - b0.pos[t]);
    } else {
        for (int t = 0; t < 3; t++) row[t * ROWLEN + C_RHS] = k * (anchor2[t] - a1[t] - b0.pos[t]);
    }
}

// getHingeAngleFromRelativeQuat joints/joint.cpp:421-458 (double compare with M_PI)
static Real hinge_angle_from_relq(const Real *qrel, const Real *axis)
{
    Real cost2 = qrel[0];
    Real sint2 = RSQRT(qrel[1] * qrel[1] + qrel[2] * qrel[2] + qrel[3] * qrel[3]);
    Real theta = (dot3(qrel + 1, axis) >= 0) ? (2 * RATAN2(sint2, cost2)) : (2 * RATAN2(sint2, -cost2));
    if (theta > M_PI) theta -= (Real)(2 * M_PI);
    theta = -theta;
    return theta;
}

// getHingeAngle joints/joint.cpp:470-490
static Real hinge_angle(World &W, int body1, int body2, const Real *axis, const Real *q_initial)
{
    Real qrel[4];
    if (body2 >= 0) { Real qq[4]; qmul1(qq, W.bodies[body1].q, W.bodies[body2].q); qmul2(qrel, qq, q_initial); }
    else qmul3(qrel, W.bodies[body1].q, q_initial);
    return hinge_angle_from_relq(qrel, axis);
}

// dxJointLimitMotor::testRotationalLimit joints/joint.cpp:574-593
static bool limot_test(Limot &l, Real angle)
{
    if (angle <= l.lostop) { l.limit = 1; l.limit_err = angle - l.lostop; return true; }
    if (angle >= l.histop) { l.limit = 2; l.limit_err = angle - l.histop; return true; }
    l.limit = 0;
    return false;
}

// dxJointLimitMotor::addLimot joints/joint.cpp:596-780, rotational form
static bool add_limot(World &W, Joint &j, Limot &l, Real fps, Real *row, const Real *ax1)
{
    int powered = l.fmax > 0;
    if (!(powered || l.limit)) return false;
    row[C_J1A] = ax1[0]; row[C_J1A + 1] = ax1[1]; row[C_J1A + 2] = ax1[2];
    if (j.b1 >= 0) { row[C_J2A] = -ax1[0]; row[C_J2A + 1] = -ax1[1]; row[C_J2A + 2] = -ax1[2]; }
    if (l.limit && (l.lostop == l.histop)) powered = 0;
    if (powered) {
        row[C_CFM] = l.normal_cfm;
        if (!l.limit) { row[C_RHS] = l.vel; row[C_LO] = -l.fmax; row[C_HI] = l.fmax; }
        else {
            Real fm = l.fmax;
            if ((l.vel > 0) || (l.vel == 0 && l.limit == 2)) fm = -fm;
            if ((l.limit == 1 && l.vel > 0) || (l.limit == 2 && l.vel < 0)) fm *= l.fudge_factor;
            Real f0 = fm * ax1[0], f1 = fm * ax1[1], f2 = fm * ax1[2];
            if (j.b1 >= 0) { Body &b1 = W.bodies[j.b1]; b1.tacc[0] += f0; b1.tacc[1] += f1; b1.tacc[2] += f2; }
            Body &b0 = W.bodies[j.b0]; b0.tacc[0] += -f0; b0.tacc[1] += -f1; b0.tacc[2] += -f2;
        }
    }
    if (l.limit) {
        Real k = fps * l.stop_erp;
        row[C_RHS] = -k * l.limit_err;
        row[C_CFM] = l.stop_cfm;
        if (l.lostop == l.histop) { row[C_LO] = -R_INF; row[C_HI] = R_INF; }
        else {
            if (l.limit == 1) { row[C_LO] = 0; row[C_HI] = R_INF; } else { row[C_LO] = -R_INF; row[C_HI] = 0; }
            if (l.bounce > 0) {
                Real vel = dot3(W.bodies[j.b0].avel, ax1);
                if (j.b1 >= 0) vel -= dot3(W.bodies[j.b1].avel, ax1);
                if (l.limit == 1) { if (vel < 0) { Real newc = -l.bounce * vel; if (newc > row[C_RHS]) row[C_RHS] = newc; } }
                else { if (vel > 0) { Real newc = -l.bounce * vel; if (newc < row[C_RHS]) row[C_RHS] = newc; } }
            }
        }
    }
    return true;
}

// addLimot joints/joint.cpp:596-780, linear form (rotational = 0) incl. the linear torque decoupling step :617-640
static bool add_limot_linear(World &W, Joint &j, Limot &l, Real fps, Real *row, const Real *ax1)
{
    int powered = l.fmax > 0;
    if (!(powered || l.limit)) return false;
    Body &b0 = W.bodies[j.b0];
    row[C_J1L] = ax1[0]; row[C_J1L + 1] = ax1[1]; row[C_J1L + 2] = ax1[2];
    Real ltd[3] = { 0, 0, 0 };
    if (j.b1 >= 0) {
        const Body &b1 = W.bodies[j.b1];
        row[C_J2L] = -ax1[0]; row[C_J2L + 1] = -ax1[1]; row[C_J2L + 2] = -ax1[2];
        Real c[3] = { R_(0.5) * (b1.pos[0] - b0.pos[0]), R_(0.5) * (b1.pos[1] - b0.pos[1]), R_(0.5) * (b1.pos[2] - b0.pos[2]) };
        cross3(ltd, c, ax1);
        row[C_J1A] = ltd[0]; row[C_J1A + 1] = ltd[1]; row[C_J1A + 2] = ltd[2];
        row[C_J2A] = ltd[0]; row[C_J2A + 1] = ltd[1]; row[C_J2A + 2] = ltd[2];
    }
    if (l.limit && (l.lostop == l.histop)) powered = 0;
    if (powered) {
        row[C_CFM] = l.normal_cfm;
        if (!l.limit) { row[C_RHS] = l.vel; row[C_LO] = -l.fmax; row[C_HI] = l.fmax; }
        else {
            Real fm = l.fmax;
            if ((l.vel > 0) || (l.vel == 0 && l.limit == 2)) fm = -fm;
            if ((l.limit == 1 && l.vel > 0) || (l.limit == 2 && l.vel < 0)) fm *= l.fudge_factor;
            Real f0 = fm * ax1[0], f1 = fm * ax1[1], f2 = fm * ax1[2];
            if (j.b1 >= 0) {
                Body &b1 = W.bodies[j.b1];
                Real t0 = -fm * ltd[0], t1 = -fm * ltd[1], t2 = -fm * ltd[2];
                b0.tacc[0] += t0; b0.tacc[1] += t1; b0.tacc[2] += t2;
                b1.tacc[0] += t0; b1.tacc[1] += t1; b1.tacc[2] += t2;
                b1.facc[0] += f0; b1.facc[1] += f1; b1.facc[2] += f2;
            }
            b0.facc[0] += -f0; b0.facc[1] += -f1; b0.facc[2] += -f2;
        }
    }
    if (l.limit) {
        Real k = fps * l.stop_erp;
        row[C_RHS] = -k * l.limit_err;
        row[C_CFM] = l.stop_cfm;
        if (l.lostop == l.histop) { row[C_LO] = -R_INF; row[C_HI] = R_INF; }
        else {
            if (l.limit == 1) { row[C_LO] = 0; row[C_HI] = R_INF; } else { row[C_LO] = -R_INF; row[C_HI] = 0; }
            if (l.bounce > 0) {
                Real vel = dot3(b0.lvel, ax1);
                if (j.b1 >= 0) vel -= dot3(W.bodies[j.b1].lvel, ax1);
                if (l.limit == 1) { if (vel < 0) { Real newc = -l.bounce * vel; if (newc > row[C_RHS]) row[C_RHS] = newc; } }
                else { if (vel > 0) { Real newc = -l.bounce * vel; if (newc < row[C_RHS]) row[C_RHS] = newc; } }
            }
        }
    }
    return true;
}

// dxJointHinge2::getInfo1 hinge2.cpp:110-130 with measureAngle1 :33-52. Storage: qrel = {c0, s0, susp_erp, susp_cfm}, qrel1 = v1, qrel2 = v2
static void hinge2_info1(World &W, Joint &j)
{
    j.m = 4; j.nub = 4;
    j.limot1.limit = 0;
    if ((j.limot1.lostop >= -M_PI || j.limot1.histop <= M_PI) && j.limot1.lostop <= j.limot1.histop) {
        Real p[3], q[3];
        mul0_331(p, W.bodies[j.b1].R, j.axis2);
        mul1_331(q, W.bodies[j.b0].R, p);
        Real x = dot3(j.qrel1, q), y = dot3(j.qrel2, q);
        limot_test(j.limot1, -RATAN2(y, x));
    }
    if (j.limot1.limit || j.limot1.fmax > 0) j.m++;
    j.limot2.limit = 0;
    if (j.limot2.fmax > 0) j.m++;
}

// dxJointHinge2::getInfo2 hinge2.cpp:155-209 with setBall2 joints/joint.cpp:165-213 (two-body form; hinge2 is dJOINT_TWOBODIES)
static void hinge2_info2(World &W, Joint &j, Real fps, Real worldERP, Real *row)
{
    const Body &b0 = W.bodies[j.b0], &b1 = W.bodies[j.b1];
    Real ax1[3], ax2[3], q[3];
    mul0_331(ax1, b0.R, j.axis1);
    mul0_331(ax2, b1.R, j.axis2);
    cross3(q, ax1, ax2);
    Real sn = RSQRT(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]), cs = dot3(ax1, ax2);
    normalize3(q);
    {
        Real q1[3], q2[3], a1[3], a2[3];
        plane_space(ax1, q1, q2);
        Real *r0 = row, *r1 = row + ROWLEN, *r2 = row + 2 * ROWLEN;
        for (int t = 0; t < 3; t++) { r0[C_J1L + t] = ax1[t]; r1[C_J1L + t] = q1[t]; r2[C_J1L + t] = q2[t]; }
        mul0_331(a1, b0.R, j.anchor1);
        cross3(r0 + C_J1A, a1, ax1); cross3(r1 + C_J1A, a1, q1); cross3(r2 + C_J1A, a1, q2);
        a1[0] = a1[0] + b0.pos[0]; a1[1] = a1[1] + b0.pos[1]; a1[2] = a1[2] + b0.pos[2];
        Real k1 = fps * j.qrel[2], k = fps * worldERP;
        for (int t = 0; t < 3; t++) { r0[C_J2L + t] = -ax1[t]; r1[C_J2L + t] = -q1[t]; r2[C_J2L + t] = -q2[t]; }
        mul0_331(a2, b1.R, j.anchor2);
        cross3(r0 + C_J2A, ax1, a2); cross3(r1 + C_J2A, q1, a2); cross3(r2 + C_J2A, q2, a2);
        a2[0] = a2[0] + b1.pos[0]; a2[1] = a2[1] + b1.pos[1]; a2[2] = a2[2] + b1.pos[2];
        Real d[3] = { a2[0] - a1[0], a2[1] - a1[1], a2[2] - a1[2] };
        r0[C_RHS] = k1 * dot3(ax1, d);
        r1[C_RHS] = k * dot3(q1, d);
        r2[C_RHS] = k * dot3(q2, d);
        r0[C_CFM] = j.qrel[3];
    }
    Real *r3 = row + 3 * ROWLEN;
    r3[C_J1A] = q[0]; r3[C_J1A + 1] = q[1]; r3[C_J1A + 2] = q[2];
    r3[C_J2A] = -q[0]; r3[C_J2A + 1] = -q[1]; r3[C_J2A + 2] = -q[2];
    r3[C_RHS] = fps * worldERP * (j.qrel[0] * sn - j.qrel[1] * cs);
    int r = 4;
    if (add_limot(W, j, j.limot1, fps, row + r * ROWLEN, ax1)) r++;
    add_limot(W, j, j.limot2, fps, row + r * ROWLEN, ax2);
}

// dJointGetSliderPosition slider.cpp:46-82 (offset kept in anchor1)
static Real slider_position(World &W, Joint &j)
{
    const Body &b0 = W.bodies[j.b0];
    Real ax1[3], q[3];
    mul0_331(ax1, b0.R, j.axis1);
    if (j.b1 >= 0) {
        const Body &b1 = W.bodies[j.b1];
        mul0_331(q, b1.R, j.anchor1);
        for (int i = 0; i < 3; i++) q[i] = b0.pos[i] - q[i] - b1.pos[i];
    } else {
        q[0] = b0.pos[0] - j.anchor1[0]; q[1] = b0.pos[1] - j.anchor1[1]; q[2] = b0.pos[2] - j.anchor1[2];
        if (j.reverse) { ax1[0] = -ax1[0]; ax1[1] = -ax1[1]; ax1[2] = -ax1[2]; }
    }
    return dot3(ax1, q);
}

// dxJointSlider::getInfo1 slider.cpp:115-145
static void slider_info1(World &W, Joint &j)
{
    j.nub = 5;
    j.m = (j.limot1.fmax > 0) ? 6 : 5;
    j.limot1.limit = 0;
    if ((j.limot1.lostop > -R_INF || j.limot1.histop < R_INF) && j.limot1.lostop <= j.limot1.histop) {
        Real pos = slider_position(W, j);
        if (limot_test(j.limot1, pos)) j.m = 6;
    }
}

// dxJointHinge::getInfo1 hinge.cpp:54-74
static void hinge_info1(World &W, Joint &j)
{
    j.nub = 5;
    j.m = (j.limot1.fmax > 0) ? 6 : 5;
    if ((j.limot1.lostop >= -M_PI || j.limot1.histop <= M_PI) && j.limot1.lostop <= j.limot1.histop) {
        Real angle = hinge_angle(W, j.b0, j.b1, j.axis1, j.qrel);
        if (limot_test(j.limot1, angle)) j.m = 6;
    }
}

// dxJointHinge::getInfo2 hinge.cpp:77-147
static void hinge_info2(World &W, Joint &j, Real fps, Real worldERP, Real *row, int * /*findex*/)
{
    set_ball(W, j, fps, worldERP, row, j.anchor1, j.anchor2);
    Real ax1[3], p[3], q[3];
    mul0_331(ax1, W.bodies[j.b0].R, j.axis1);
    plane_space(ax1, p, q);
    Real *r3 = row + 3 * ROWLEN, *r4 = row + 4 * ROWLEN;
    r3[C_J1A] = p[0]; r3[C_J1A + 1] = p[1]; r3[C_J1A + 2] = p[2];
    if (j.b1 >= 0) { r3[C_J2A] = -p[0]; r3[C_J2A + 1] = -p[1]; r3[C_J2A + 2] = -p[2]; }
    r4[C_J1A] = q[0]; r4[C_J1A + 1] = q[1]; r4[C_J1A + 2] = q[2];
    if (j.b1 >= 0) { r4[C_J2A] = -q[0]; r4[C_J2A + 1] = -q[1]; r4[C_J2A + 2] = -q[2]; }
    Real b[3];
    if (j.b1 >= 0) { Real ax2[3]; mul0_331(ax2, W.bodies[j.b1].R, j.axis2); cross3(b, ax1, ax2); }
    else cross3(b, ax1, j.axis2);
    Real k = fps * worldERP;
    r3[C_RHS] = k * dot3(b, p);
    r4[C_RHS] = k * dot3(b, q);
    add_limot(W, j, j.limot1, fps, row + 5 * ROWLEN, ax1);
}

// dxJointUniversal::getAxes universal.cpp:55-71
static void universal_axes(World &W, Joint &j, Real *ax1, Real *ax2)
{
    mul0_331(ax1, W.bodies[j.b0].R, j.axis1);
    if (j.b1 >= 0) mul0_331(ax2, W.bodies[j.b1].R, j.axis2);
    else { ax2[0] = j.axis2[0]; ax2[1] = j.axis2[1]; ax2[2] = j.axis2[2]; }
}

// dxJointUniversal::getAngles universal.cpp:73-166
static void universal_angles(World &W, Joint &j, Real *angle1, Real *angle2)
{
    Real ax1[3], ax2[3], R[12], qcross[4], qq[4], qrel[4];
    universal_axes(W, j, ax1, ax2);
    r_from_2axes(R, ax1[0], ax1[1], ax1[2], ax2[0], ax2[1], ax2[2]);
    q_from_r(qcross, R);
    qmul1(qq, W.bodies[j.b0].q, qcross);
    qmul2(qrel, qq, j.qrel1);
    *angle1 = hinge_angle_from_relq(qrel, j.axis1);
    Real qcross2[4];
    qrel[0] = 0; qrel[1] = ax1[0] + ax2[0]; qrel[2] = ax1[1] + ax2[1]; qrel[3] = ax1[2] + ax2[2];
    Real l = rrecip(sqrt(qrel[1] * qrel[1] + qrel[2] * qrel[2] + qrel[3] * qrel[3]));
    qrel[1] *= l; qrel[2] *= l; qrel[3] *= l;
    qmul0(qcross2, qrel, qcross);
    if (j.b1 >= 0) { qmul1(qq, W.bodies[j.b1].q, qcross2); qmul2(qrel, qq, j.qrel2); }
    else qmul2(qrel, qcross2, j.qrel2);
    *angle2 = -hinge_angle_from_relq(qrel, j.axis2);
}

// dxJointUniversal::getInfo1 universal.cpp:266-293
static void universal_info1(World &W, Joint &j)
{
    j.nub = 4; j.m = 4;
    bool lim1 = (j.limot1.lostop >= -M_PI || j.limot1.histop <= M_PI) && j.limot1.lostop <= j.limot1.histop;
    bool lim2 = (j.limot2.lostop >= -M_PI || j.limot2.histop <= M_PI) && j.limot2.lostop <= j.limot2.histop;
    j.limot1.limit = 0; j.limot2.limit = 0;
    if (lim1 || lim2) {
        Real a1, a2;
        universal_angles(W, j, &a1, &a2);
        if (lim1) limot_test(j.limot1, a1);
        if (lim2) limot_test(j.limot2, a2);
    }
    if (j.limot1.limit || j.limot1.fmax > 0) j.m++;
    if (j.limot2.limit || j.limot2.fmax > 0) j.m++;
}

// dxJointUniversal::getInfo2 universal.cpp:297-369
static void universal_info2(World &W, Joint &j, Real fps, Real worldERP, Real *row, int * /*findex*/)
{
    set_ball(W, j, fps, worldERP, row, j.anchor1, j.anchor2);
    Real ax1[3], ax2[3], p[3];
    universal_axes(W, j, ax1, ax2);
    Real k = dot3(ax1, ax2);
    Real ax2t[3] = { ax2[0] + (-k) * ax1[0], ax2[1] + (-k) * ax1[1], ax2[2] + (-k) * ax1[2] };  // dAddVectorScaledVector3
    cross3(p, ax1, ax2t);
    normalize3(p);
    Real *r3 = row + 3 * ROWLEN;
    r3[C_J1A] = p[0]; r3[C_J1A + 1] = p[1]; r3[C_J1A + 2] = p[2];
    if (j.b1 >= 0) { r3[C_J2A] = -p[0]; r3[C_J2A + 1] = -p[1]; r3[C_J2A + 2] = -p[2]; }
    r3[C_RHS] = fps * worldERP * (-k);
    int r = 4;
    if (add_limot(W, j, j.limot1, fps, row + r * ROWLEN, ax1)) r++;
    add_limot(W, j, j.limot2, fps, row + r * ROWLEN, ax2);
}

// setFixedOrientation joints/joint.cpp:228-284
static void set_fixed_orientation(World &W, Joint &j, Real fps, Real erp, Real *row, const Real *qrel)
{
    const Body &b0 = W.bodies[j.b0];
    row[C_J1A] = 1; row[ROWLEN + C_J1A + 1] = 1; row[2 * ROWLEN + C_J1A + 2] = 1;
    if (j.b1 >= 0) { row[C_J2A] = -1; row[ROWLEN + C_J2A + 1] = -1; row[2 * ROWLEN + C_J2A + 2] = -1; }
    Real qerr[4], e[3];
    if (j.b1 >= 0) { Real qq[4]; qmul1(qq, b0.q, W.bodies[j.b1].q); qmul2(qerr, qq, qrel); }
    else qmul3(qerr, b0.q, qrel);
    if (qerr[0] < 0) { qerr[1] = -qerr[1]; qerr[2] = -qerr[2]; qerr[3] = -qerr[3]; }
    mul0_331(e, b0.R, qerr + 1);
    Real k2 = fps * erp * R_(2.0);
    row[C_RHS] = k2 * e[0]; row[ROWLEN + C_RHS] = k2 * e[1]; row[2 * ROWLEN + C_RHS] = k2 * e[2];
}

// dxJointFixed::getInfo2 fixed.cpp:60-110 (offset kept in anchor1)
static void fixed_info2(World &W, Joint &j, Real fps, Real worldERP, Real *row)
{
    set_fixed_orientation(W, j, fps, worldERP, row + 3 * ROWLEN, j.qrel);
    row[C_J1L] = 1; row[ROWLEN + C_J1L + 1] = 1; row[2 * ROWLEN + C_J1L + 2] = 1;
    Real k = fps * j.erp;
    const Body &b0 = W.bodies[j.b0];
    Real ofs[3];
    mul0_331(ofs, b0.R, j.anchor1);
    if (j.b1 >= 0) {
        const Body &b1 = W.bodies[j.b1];
        row[C_J1A + 1] = -ofs[2]; row[C_J1A + 2] = +ofs[1];
        row[ROWLEN + C_J1A + 0] = +ofs[2]; row[ROWLEN + C_J1A + 2] = -ofs[0];
        row[2 * ROWLEN + C_J1A + 0] = -ofs[1]; row[2 * ROWLEN + C_J1A + 1] = +ofs[0];
        row[C_J2L] = -1; row[ROWLEN + C_J2L + 1] = -1; row[2 * ROWLEN + C_J2L + 2] = -1;
        for (int t = 0; t < 3; t++) row[t * ROWLEN + C_RHS] = k * (b1.pos[t] - b0.pos[t] + ofs[t]);
    } else {
        for (int t = 0; t < 3; t++) row[t * ROWLEN + C_RHS] = k * (j.anchor1[t] - b0.pos[t]);
    }
    row[C_CFM] = j.cfm; row[ROWLEN + C_CFM] = j.cfm; row[2 * ROWLEN + C_CFM] = j.cfm;
}

// dxJointSlider::getInfo2 slider.cpp:148-246
static void slider_info2(World &W, Joint &j, Real fps, Real worldERP, Real *row)
{
    set_fixed_orientation(W, j, fps, worldERP, row, j.qrel);
    const Body &b0 = W.bodies[j.b0];
    Real ax1[3], p[3], q[3], c[3] = { 0, 0, 0 };
    mul0_331(ax1, b0.R, j.axis1);
    plane_space(ax1, p, q);
    if (j.b1 >= 0) { const Body &b1 = W.bodies[j.b1]; c[0] = b1.pos[0] - b0.pos[0]; c[1] = b1.pos[1] - b0.pos[1]; c[2] = b1.pos[2] - b0.pos[2]; }
    Real *r3 = row + 3 * ROWLEN, *r4 = row + 4 * ROWLEN;
    r3[C_J1L] = p[0]; r3[C_J1L + 1] = p[1]; r3[C_J1L + 2] = p[2];
    r4[C_J1L] = q[0]; r4[C_J1L + 1] = q[1]; r4[C_J1L + 2] = q[2];
    if (j.b1 >= 0) {
        Real tmp[3];
        r3[C_J2L] = -p[0]; r3[C_J2L + 1] = -p[1]; r3[C_J2L + 2] = -p[2];
        cross3(tmp, c, p);
        for (int t = 0; t < 3; t++) { r3[C_J1A + t] = tmp[t] * R_(0.5); r3[C_J2A + t] = r3[C_J1A + t]; }
        r4[C_J2L] = -q[0]; r4[C_J2L + 1] = -q[1]; r4[C_J2L + 2] = -q[2];
        cross3(tmp, c, q);
        for (int t = 0; t < 3; t++) { r4[C_J1A + t] = tmp[t] * R_(0.5); r4[C_J2A + t] = r4[C_J1A + t]; }
    }
    Real k = fps * worldERP;
    if (j.b1 >= 0) {
        Real ofs[3];
        mul0_331(ofs, W.bodies[j.b1].R, j.anchor1);
        c[0] = c[0] + ofs[0]; c[1] = c[1] + ofs[1]; c[2] = c[2] + ofs[2];
        r3[C_RHS] = k * dot3(p, c);
        r4[C_RHS] = k * dot3(q, c);
    } else {
        Real ofs[3] = { j.anchor1[0] - b0.pos[0], j.anchor1[1] - b0.pos[1], j.anchor1[2] - b0.pos[2] };
        r3[C_RHS] = k * dot3(p, ofs);
        r4[C_RHS] = k * dot3(q, ofs);
        if (j.reverse) { ax1[0] = -ax1[0]; ax1[1] = -ax1[1]; ax1[2] = -ax1[2]; }
    }
    add_limot_linear(W, j, j.limot1, fps, row + 5 * ROWLEN, ax1);
}

// ---- linear motor (lmotor.cpp) and angular motor (amotor.cpp)
static Limot &motor_limot(Joint &j, int i) { return i == 0 ? j.limot1 : i == 1 ? j.limot2 : j.limot3; }

// dxJointLMotor::computeGlobalAxes lmotor.cpp:51-76
static void lmotor_global_axes(World &W, Joint &j, Real ax[3][3])
{
    for (int i = 0; i < j.mnum; i++) {
        if (j.mrel[i] == 1) mul0_331(ax[i], W.bodies[j.b0].R, j.maxis[i]);
        else if (j.mrel[i] == 2) { if (j.b1 >= 0) mul0_331(ax[i], W.bodies[j.b1].R, j.maxis[i]); }
        else { ax[i][0] = j.maxis[i][0]; ax[i][1] = j.maxis[i][1]; ax[i][2] = j.maxis[i][2]; }
    }
}
// dxJointLMotor::getInfo1 lmotor.cpp:84-97
static void lmotor_info1(World &W, Joint &j)
{
    j.m = 0; j.nub = 0;
    for (int i = 0; i < j.mnum; i++) if (motor_limot(j, i).fmax > 0) j.m++;
}
// dxJointLMotor::getInfo2 lmotor.cpp:99-116
static void lmotor_info2(World &W, Joint &j, Real fps, Real *row)
{
    Real ax[3][3] = { { 0 } };
    lmotor_global_axes(W, j, ax);
    int r = 0;
    for (int i = 0; i < j.mnum; i++) if (add_limot_linear(W, j, motor_limot(j, i), fps, row + r * ROWLEN, ax[i])) r++;
}
// dJointSetLMotorAxis lmotor.cpp:118-160
static void lmotor_set_axis(World &W, Joint &j, int anum, int rel, Real x, Real y, Real z)
{
    if (j.b1 < 0 && rel == 2) rel = 1;
    j.mrel[anum] = rel;
    Real r[3] = { x, y, z };
    if (rel == 1) mul1_331(j.maxis[anum], W.bodies[j.b0].R, r);
    else if (rel == 2) mul1_331(j.maxis[anum], W.bodies[j.b1].R, r);
    else { j.maxis[anum][0] = x; j.maxis[anum][1] = y; j.maxis[anum][2] = z; }
    normalize3(j.maxis[anum]);
}

// dxJointAMotor::doComputeGlobalUserAxes / doComputeGlobalEulerAxes amotor.cpp:662-713
static void amotor_global_axes(World &W, Joint &j, Real ax[3][3])
{
    if (j.mmode == 0) {
        for (int i = 0; i < j.mnum; i++) {
            bool assigned = false;
            if (j.mrel[i] == 1) { mul0_331(ax[i], W.bodies[j.b0].R, j.maxis[i]); assigned = true; }
            else if (j.mrel[i] == 2 && j.b1 >= 0) { mul0_331(ax[i], W.bodies[j.b1].R, j.maxis[i]); assigned = true; }
            if (!assigned) { ax[i][0] = j.maxis[i][0]; ax[i][1] = j.maxis[i][1]; ax[i][2] = j.maxis[i][2]; }
        }
    } else {
        const int first = j.reverse ? 2 : 0, second = 2 - first;     // BuildFirstBodyEulerAxis :798-807
        mul0_331(ax[first], W.bodies[j.b0].R, j.maxis[first]);
        if (j.b1 >= 0) mul0_331(ax[second], W.bodies[j.b1].R, j.maxis[second]);
        else { ax[second][0] = j.maxis[second][0]; ax[second][1] = j.maxis[second][1]; ax[second][2] = j.maxis[second][2]; }
        cross3(ax[1], ax[2], ax[0]);
        normalize3(ax[1]);
    }
}
// dxJointAMotor::computeEulerAngles amotor.cpp:716-758
static void amotor_euler_angles(World &W, Joint &j, Real ax[3][3])
{
    Real refs[2][3], q[3];
    mul0_331(refs[0], W.bodies[j.b0].R, j.mref[0]);
    if (j.b1 >= 0) mul0_331(refs[1], W.bodies[j.b1].R, j.mref[1]);
    else { refs[1][0] = j.mref[1][0]; refs[1][1] = j.mref[1][1]; refs[1][2] = j.mref[1][2]; }
    const int fb = j.reverse ? 1 : 0, sb = 1 - fb;
    cross3(q, ax[0], refs[fb]);
    j.mangle[0] = -RATAN2(dot3(ax[2], q), dot3(ax[2], refs[fb]));
    cross3(q, ax[0], ax[1]);
    j.mangle[1] = -RATAN2(dot3(ax[2], ax[0]), dot3(ax[2], q));
    cross3(q, ax[1], ax[2]);
    j.mangle[2] = -RATAN2(dot3(refs[sb], ax[1]), dot3(refs[sb], q));
}
// dxJointAMotor::getInfo1 amotor.cpp:264-287
static void amotor_info1(World &W, Joint &j)
{
    j.m = 0; j.nub = 0;
    if (j.mmode == 1) {
        Real ax[3][3] = { { 0 } };
        amotor_global_axes(W, j, ax);
        amotor_euler_angles(W, j, ax);
    }
    for (int i = 0; i < j.mnum; i++) {
        Limot &l = motor_limot(j, i);
        if (limot_test(l, j.mangle[i]) || l.fmax > 0) j.m++;
    }
}
// dxJointAMotor::getInfo2 amotor.cpp:290-339
static void amotor_info2(World &W, Joint &j, Real fps, Real *row)
{
    Real ax[3][3] = { { 0 } }, c01[3], c12[3];
    amotor_global_axes(W, j, ax);
    const Real *axp[3] = { ax[0], ax[1], ax[2] };
    if (j.mmode == 1) {
        cross3(c01, ax[0], ax[1]); axp[2] = c01;
        cross3(c12, ax[1], ax[2]); axp[0] = c12;
    }
    int r = 0;
    for (int i = 0; i < j.mnum; i++) if (add_limot(W, j, motor_limot(j, i), fps, row + r * ROWLEN, axp[i])) r++;
}
// dxJointAMotor::setAxisValue amotor.cpp:393-442
static void amotor_set_axis(World &W, Joint &j, int anum, int rel, Real x, Real y, Real z)
{
    if (rel != 0 && j.reverse) rel = 3 - rel;
    j.mrel[anum] = rel;
    Real r[3] = { x, y, z };
    bool assigned = false;
    if (rel == 1) { mul1_331(j.maxis[anum], W.bodies[j.b0].R, r); assigned = true; }
    else if (rel == 2 && j.b1 >= 0) { mul1_331(j.maxis[anum], W.bodies[j.b1].R, r); assigned = true; }
    if (!assigned) { j.maxis[anum][0] = x; j.maxis[anum][1] = y; j.maxis[anum][2] = z; }
    normalize3(j.maxis[anum]);
}
// dxJointAMotor::setEulerReferenceVectors amotor.cpp:768-796
static void amotor_set_euler_references(World &W, Joint &j)
{
    const int first = j.reverse ? 2 : 0, second = 2 - first;
    if (j.b1 >= 0) {
        Real r[3];
        mul0_331(r, W.bodies[j.b0].R, j.maxis[first]);
        mul1_331(j.mref[1], W.bodies[j.b1].R, r);
        mul0_331(r, W.bodies[j.b1].R, j.maxis[second]);
        mul1_331(j.mref[0], W.bodies[j.b0].R, r);
    } else {
        mul0_331(j.mref[1], W.bodies[j.b0].R, j.maxis[first]);
        mul1_331(j.mref[0], W.bodies[j.b0].R, j.maxis[second]);
    }
}

// dJointSet{Ball,Hinge,Universal}Anchor/Axis at the template pose
static void joint_setup(const Batch &B, World &W, Joint &j, const OdebJointDesc &d)
{
    if (j.type == ODEB_JOINT_LMOTOR || j.type == ODEB_JOINT_AMOTOR) {
        // dJointSet{L,A}MotorNumAxes, dJointSetAMotorMode (amotor.cpp:354-377: Euler mode forces 3 axes), dJointSet{L,A}MotorAxis, dJointSetAMotorAngle
        const bool am = j.type == ODEB_JOINT_AMOTOR;
        j.mmode = am ? d.motor_mode : 0;
        j.mnum = (am && j.mmode == 1) ? 3 : d.motor_num;
        for (int i = 0; i < d.motor_num; i++) {
            if (am && j.mmode == 1 && i == 1) continue;                 // Euler mode: axis 1 is derived (ax[2] x ax[0])
            if (am) amotor_set_axis(W, j, i, d.motor_rel[i], (Real)d.motor_axis[i][0], (Real)d.motor_axis[i][1], (Real)d.motor_axis[i][2]);
            else lmotor_set_axis(W, j, i, d.motor_rel[i], (Real)d.motor_axis[i][0], (Real)d.motor_axis[i][1], (Real)d.motor_axis[i][2]);
        }
        if (am && j.mmode == 1) amotor_set_euler_references(W, j);
        for (int i = 0; i < 3; i++) { j.mangle[i] = am ? (Real)d.motor_angle[i] : 0; limot_set(motor_limot(j, i), d, i); }
        return;
    }
    set_anchors(W, j, (Real)d.anchor[0], (Real)d.anchor[1], (Real)d.anchor[2]);
    if (j.type == ODEB_JOINT_FIXED) {
        // dJointSetFixed fixed.cpp:113-142
        const Body &b0 = W.bodies[j.b0];
        if (j.b1 >= 0) {
            Real ofs[3] = { b0.pos[0] - W.bodies[j.b1].pos[0], b0.pos[1] - W.bodies[j.b1].pos[1], b0.pos[2] - W.bodies[j.b1].pos[2] };
            mul1_331(j.anchor1, b0.R, ofs);
        } else { j.anchor1[0] = b0.pos[0]; j.anchor1[1] = b0.pos[1]; j.anchor1[2] = b0.pos[2]; }
        if (j.b1 >= 0) qmul1(j.qrel, b0.q, W.bodies[j.b1].q);
        else { const Real *q = b0.q; j.qrel[0] = q[0]; j.qrel[1] = -q[1]; j.qrel[2] = -q[2]; j.qrel[3] = -q[3]; }
    } else if (j.type == ODEB_JOINT_HINGE2) {
        // dJointSetHinge2Anchor + dJointSetHinge2Axes hinge2.cpp:268-312, makeV1andV2 :214-240
        j.axis1[0] = 1; j.axis2[1] = 1;
        set_axes(W, j, (Real)d.axis1[0], (Real)d.axis1[1], (Real)d.axis1[2], j.axis1, 0);
        set_axes(W, j, (Real)d.axis2[0], (Real)d.axis2[1], (Real)d.axis2[2], 0, j.axis2);
        const Body &b0 = W.bodies[j.b0], &b1 = W.bodies[j.b1];
        Real ax1[3], ax2[3], ax[3], v[3];
        mul0_331(ax1, b0.R, j.axis1);
        mul0_331(ax2, b1.R, j.axis2);
        cross3(ax, ax1, ax2);
        j.qrel[1] = RSQRT(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
        j.qrel[0] = dot3(ax1, ax2);
        Real k = dot3(ax1, ax2);
        ax2[0] = ax2[0] + ax1[0] * (-k); ax2[1] = ax2[1] + ax1[1] * (-k); ax2[2] = ax2[2] + ax1[2] * (-k);
        normalize3(ax2);
        cross3(v, ax1, ax2);
        mul1_331(j.qrel1, b0.R, ax2);
        mul1_331(j.qrel2, b0.R, v);
        j.qrel[2] = d.susp_erp >= 0 ? (Real)d.susp_erp : B.erp;
        j.qrel[3] = d.susp_cfm >= 0 ? (Real)d.susp_cfm : B.cfm;
        limot_set(j.limot1, d, 0);
        limot_set(j.limot2, d, 1);
    } else if (j.type == ODEB_JOINT_SLIDER) {
        // dJointSetSliderAxis slider.cpp:249-260: setAxes(axis1), computeOffset :406-425, computeInitialRelativeRotation :382-401
        j.axis1[0] = 1;
        set_axes(W, j, (Real)d.axis1[0], (Real)d.axis1[1], (Real)d.axis1[2], j.axis1, 0);
        const Body &b0 = W.bodies[j.b0];
        if (j.b1 >= 0) {
            Real c[3] = { b0.pos[0] - W.bodies[j.b1].pos[0], b0.pos[1] - W.bodies[j.b1].pos[1], b0.pos[2] - W.bodies[j.b1].pos[2] };
            mul1_331(j.anchor1, W.bodies[j.b1].R, c);
        } else { j.anchor1[0] = b0.pos[0]; j.anchor1[1] = b0.pos[1]; j.anchor1[2] = b0.pos[2]; }
        if (j.b1 >= 0) qmul1(j.qrel, b0.q, W.bodies[j.b1].q);
        else { const Real *q = b0.q; j.qrel[0] = q[0]; j.qrel[1] = -q[1]; j.qrel[2] = -q[2]; j.qrel[3] = -q[3]; }
        limot_set(j.limot1, d, 0);
    } else if (j.type == ODEB_JOINT_HINGE) {
        j.axis1[0] = 1; j.axis2[0] = 1;   // constructor defaults hinge.cpp:36-43
        set_axes(W, j, (Real)d.axis1[0], (Real)d.axis1[1], (Real)d.axis1[2], j.axis1, j.axis2);
        // computeInitialRelativeRotation hinge.cpp:376-393
        if (j.b1 >= 0) qmul1(j.qrel, W.bodies[j.b0].q, W.bodies[j.b1].q);
        else { const Real *q = W.bodies[j.b0].q; j.qrel[0] = q[0]; j.qrel[1] = -q[1]; j.qrel[2] = -q[2]; j.qrel[3] = -q[3]; }
        limot_set(j.limot1, d, 0);
    } else if (j.type == ODEB_JOINT_UNIVERSAL) {
        j.axis1[0] = 1; j.axis2[1] = 1;   // universal.cpp:40-52
        if (j.reverse) set_axes(W, j, (Real)d.axis1[0], (Real)d.axis1[1], (Real)d.axis1[2], 0, j.axis2);
        else set_axes(W, j, (Real)d.axis1[0], (Real)d.axis1[1], (Real)d.axis1[2], j.axis1, 0);
        if (j.reverse) set_axes(W, j, (Real)d.axis2[0], (Real)d.axis2[1], (Real)d.axis2[2], j.axis1, 0);
        else set_axes(W, j, (Real)d.axis2[0], (Real)d.axis2[1], (Real)d.axis2[2], 0, j.axis2);
        // computeInitialRelativeRotations universal.cpp:372-401
        Real ax1[3], ax2[3], R[12], qcross[4];
        universal_axes(W, j, ax1, ax2);
        r_from_2axes(R, ax1[0], ax1[1], ax1[2], ax2[0], ax2[1], ax2[2]);
        q_from_r(qcross, R);
        qmul1(j.qrel1, W.bodies[j.b0].q, qcross);
        r_from_2axes(R, ax2[0], ax2[1], ax2[2], ax1[0], ax1[1], ax1[2]);
        q_from_r(qcross, R);
        if (j.b1 >= 0) qmul1(j.qrel2, W.bodies[j.b1].q, qcross);
        else for (int i = 0; i < 4; i++) j.qrel2[i] = qcross[i];
        limot_set(j.limot1, d, 0);
        limot_set(j.limot2, d, 1);
    }
}
#endif
