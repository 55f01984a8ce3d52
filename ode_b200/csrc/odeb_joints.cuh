// odeb_joints.cuh -- device-side constraint-row builders (the reference's getInfo1/getInfo2 contract)
// for contact, ball, hinge and universal joints. One thread builds all rows of one joint in the
// reference's 16-wide row layout [J1l(3) J1a(3) rhs cfm J2l(3) J2a(3) lo hi] (quickstep.cpp:267-322).
// Operation order follows the cited reference functions.
#ifndef ODEB_JOINTS_CUH
#define ODEB_JOINTS_CUH
#include "odeb_math.cuh"

enum { ROWLEN = 16, C_J1L = 0, C_J1A = 3, C_RHS = 6, C_CFM = 7, C_J2L = 8, C_J2A = 11, C_LO = 14, C_HI = 15 };

struct DBody { Real pos[3], q[4], lvel[3], avel[3], R[12]; };

struct DLimot {     // dxJointLimitMotor, static part (joints/joint.h:291-320)
    Real vel, fmax, lostop, histop, fudge_factor, normal_cfm, stop_erp, stop_cfm, bounce;
};

struct DJointT {    // template (per-batch) description of a permanent joint, after dJointAttach's swap
    int type, b0, b1, reverse;
    Real anchor1[4], anchor2[4], axis1[4], axis2[4], qrel[4], qrel1[4], qrel2[4];
    // fixed / slider: anchor1 = offset. hinge2: qrel = {c0, s0, susp_erp, susp_cfm}, qrel1 = v1, qrel2 = v2 (hinge2.cpp:76-100)
    Real erp, cfm;
    DLimot limot1, limot2;
    // lmotor / amotor (lmotor.h:30-49, amotor.h:91-101): number of axes, mode (0 user, 1 Euler), what each axis is relative to (after the
    // reverse swap of amotor.cpp:405-409), axes, Euler reference vectors, user-mode angles, third limit motor
    int mnum, mmode, mrel[3];
    Real maxis[3][4], mref[2][4], mangle[3];
    DLimot limot3;
};

struct DLimitState { int limit1, limit2; Real err1, err2; int limit3; Real err3; };

__host__ __device__ __forceinline__ void odeb_qmul3(Real *qa, const Real *qb, const Real *qc)
{   // dQMultiply3 rotation.cpp:221-228
    qa[0] = qb[0] * qc[0] - qb[1] * qc[1] - qb[2] * qc[2] - qb[3] * qc[3];
    qa[1] = -qb[0] * qc[1] - qb[1] * qc[0] + qb[2] * qc[3] - qb[3] * qc[2];
    qa[2] = -qb[0] * qc[2] - qb[2] * qc[0] + qb[3] * qc[1] - qb[1] * qc[3];
    qa[3] = -qb[0] * qc[3] - qb[3] * qc[0] + qb[1] * qc[2] - qb[2] * qc[1];
}

// dRFrom2Axes rotation.cpp:94-133
__host__ __device__ __forceinline__ bool odeb_r_from_2axes(Real *R, Real ax, Real ay, Real az, Real bx, Real by, Real bz)
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

// setBall joints/joint.cpp:111-157
__device__ void odeb_set_ball(const DBody &b0, const DBody *b1, Real fps, Real erp, Real *row, const Real *anchor1, const Real *anchor2)
{
    Real a1[3], a2[3];
    row[C_J1L + 0] = 1; row[ROWLEN + C_J1L + 1] = 1; row[2 * ROWLEN + C_J1L + 2] = 1;
    mul0_331(a1, b0.R, anchor1);
    row[C_J1A + 1] = +a1[2]; row[C_J1A + 2] = -a1[1];
    row[ROWLEN + C_J1A + 0] = -a1[2]; row[ROWLEN + C_J1A + 2] = +a1[0];
    row[2 * ROWLEN + C_J1A + 0] = +a1[1]; row[2 * ROWLEN + C_J1A + 1] = -a1[0];
    if (b1) {
        row[C_J2L + 0] = -1; row[ROWLEN + C_J2L + 1] = -1; row[2 * ROWLEN + C_J2L + 2] = -1;
        mul0_331(a2, b1->R, anchor2);
        row[C_J2A + 1] = -a2[2]; row[C_J2A + 2] = +a2[1];
        row[ROWLEN + C_J2A + 0] = +a2[2]; row[ROWLEN + C_J2A + 2] = -a2[0];
        row[2 * ROWLEN + C_J2A + 0] = -a2[1]; row[2 * ROWLEN + C_J2A + 1] = +a2[0];
    }
    Real k = fps * erp;
    if (b1) { for (int t = 0; t < 3; t++) row[t * ROWLEN + C_RHS] = k * (a2[t] + b1->pos[t] - a1[t] - b0.pos[t]); }
    else { for (int t = 0; t < 3; t++) row[t * ROWLEN + C_RHS] = k * (anchor2[t] - a1[t] - b0.pos[t]); }
}

// getHingeAngleFromRelativeQuat joints/joint.cpp:421-458
__device__ Real odeb_hinge_angle_from_relq(const Real *qrel, const Real *axis)
{
    Real cost2 = qrel[0];
    Real sint2 = RSQRT(qrel[1] * qrel[1] + qrel[2] * qrel[2] + qrel[3] * qrel[3]);
    Real theta = (dot3(qrel + 1, axis) >= 0) ? (2 * RATAN2(sint2, cost2)) : (2 * RATAN2(sint2, -cost2));
    if (theta > M_PI) theta -= (Real)(2 * M_PI);
    theta = -theta;
    return theta;
}

// getHingeAngle joints/joint.cpp:470-490
__device__ Real odeb_hinge_angle(const DBody &b0, const DBody *b1, const Real *axis, const Real *q_initial)
{
    Real qrel[4];
    if (b1) { Real qq[4]; qmul1(qq, b0.q, b1->q); qmul2(qrel, qq, q_initial); }
    else odeb_qmul3(qrel, b0.q, q_initial);
    return odeb_hinge_angle_from_relq(qrel, axis);
}

// dxJointLimitMotor::testRotationalLimit joints/joint.cpp:574-593
__device__ __forceinline__ bool odeb_limot_test(const DLimot &l, Real angle, int *limit, Real *err)
{
    if (angle <= l.lostop) { *limit = 1; *err = angle - l.lostop; return true; }
    if (angle >= l.histop) { *limit = 2; *err = angle - l.histop; return true; }
    *limit = 0;
    return false;
}

// dxJointLimitMotor::addLimot joints/joint.cpp:596-780 (rotational). Torque applied to the bodies when
// powered at a limit is returned in tq (added to b1.tacc, subtracted from b0.tacc by the caller).
__device__ bool odeb_add_limot(const DLimot &l, int limit, Real limit_err, const DBody &b0, const DBody *b1,
                               Real fps, Real *row, const Real *ax1, Real *tq, bool *has_tq)
{
    int powered = l.fmax > 0;
    if (!(powered || limit)) return false;
    row[C_J1A] = ax1[0]; row[C_J1A + 1] = ax1[1]; row[C_J1A + 2] = ax1[2];
    if (b1) { row[C_J2A] = -ax1[0]; row[C_J2A + 1] = -ax1[1]; row[C_J2A + 2] = -ax1[2]; }
    if (limit && (l.lostop == l.histop)) powered = 0;
    if (powered) {
        row[C_CFM] = l.normal_cfm;
        if (!limit) { row[C_RHS] = l.vel; row[C_LO] = -l.fmax; row[C_HI] = l.fmax; }
        else {
            Real fm = l.fmax;
            if ((l.vel > 0) || (l.vel == 0 && limit == 2)) fm = -fm;
            if ((limit == 1 && l.vel > 0) || (limit == 2 && l.vel < 0)) fm *= l.fudge_factor;
            tq[0] = fm * ax1[0]; tq[1] = fm * ax1[1]; tq[2] = fm * ax1[2];
            *has_tq = true;
        }
    }
    if (limit) {
        Real k = fps * l.stop_erp;
        row[C_RHS] = -k * limit_err;
        row[C_CFM] = l.stop_cfm;
        if (l.lostop == l.histop) { row[C_LO] = -R_INF; row[C_HI] = R_INF; }
        else {
            if (limit == 1) { row[C_LO] = 0; row[C_HI] = R_INF; } else { row[C_LO] = -R_INF; row[C_HI] = 0; }
            if (l.bounce > 0) {
                Real vel = dot3(b0.avel, ax1);
                if (b1) vel -= dot3(b1->avel, ax1);
                if (limit == 1) { if (vel < 0) { Real newc = -l.bounce * vel; if (newc > row[C_RHS]) row[C_RHS] = newc; } }
                else { if (vel > 0) { Real newc = -l.bounce * vel; if (newc < row[C_RHS]) row[C_RHS] = newc; } }
            }
        }
    }
    return true;
}

// addLimot, linear form (rotational = 0: the slider's stop / motor row), incl. the linear torque decoupling of joint.cpp:617-640.
// Forces of a motor working against a stop (joint.cpp:655-704): fq is added to body 1 and subtracted from body 0, tboth is
// added to both bodies' torques.
__device__ bool odeb_add_limot_linear(const DLimot &l, int limit, Real limit_err, const DBody &b0, const DBody *b1,
                                      Real fps, Real *row, const Real *ax1, Real *fq, Real *tboth, bool *has_f)
{
    int powered = l.fmax > 0;
    if (!(powered || limit)) return false;
    row[C_J1L] = ax1[0]; row[C_J1L + 1] = ax1[1]; row[C_J1L + 2] = ax1[2];
    Real ltd[3] = { 0, 0, 0 };
    if (b1) {
        row[C_J2L] = -ax1[0]; row[C_J2L + 1] = -ax1[1]; row[C_J2L + 2] = -ax1[2];
        Real c[3] = { R_(0.5) * (b1->pos[0] - b0.pos[0]), R_(0.5) * (b1->pos[1] - b0.pos[1]), R_(0.5) * (b1->pos[2] - b0.pos[2]) };
        cross3(ltd, c, ax1);
        row[C_J1A] = ltd[0]; row[C_J1A + 1] = ltd[1]; row[C_J1A + 2] = ltd[2];
        row[C_J2A] = ltd[0]; row[C_J2A + 1] = ltd[1]; row[C_J2A + 2] = ltd[2];
    }
    if (limit && (l.lostop == l.histop)) powered = 0;
    if (powered) {
        row[C_CFM] = l.normal_cfm;
        if (!limit) { row[C_RHS] = l.vel; row[C_LO] = -l.fmax; row[C_HI] = l.fmax; }
        else {
            Real fm = l.fmax;
            if ((l.vel > 0) || (l.vel == 0 && limit == 2)) fm = -fm;
            if ((limit == 1 && l.vel > 0) || (limit == 2 && l.vel < 0)) fm *= l.fudge_factor;
            fq[0] = fm * ax1[0]; fq[1] = fm * ax1[1]; fq[2] = fm * ax1[2];
            tboth[0] = -fm * ltd[0]; tboth[1] = -fm * ltd[1]; tboth[2] = -fm * ltd[2];
            *has_f = true;
        }
    }
    if (limit) {
        Real k = fps * l.stop_erp;
        row[C_RHS] = -k * limit_err;
        row[C_CFM] = l.stop_cfm;
        if (l.lostop == l.histop) { row[C_LO] = -R_INF; row[C_HI] = R_INF; }
        else {
            if (limit == 1) { row[C_LO] = 0; row[C_HI] = R_INF; } else { row[C_LO] = -R_INF; row[C_HI] = 0; }
            if (l.bounce > 0) {
                Real vel = dot3(b0.lvel, ax1);
                if (b1) vel -= dot3(b1->lvel, ax1);
                if (limit == 1) { if (vel < 0) { Real newc = -l.bounce * vel; if (newc > row[C_RHS]) row[C_RHS] = newc; } }
                else { if (vel > 0) { Real newc = -l.bounce * vel; if (newc < row[C_RHS]) row[C_RHS] = newc; } }
            }
        }
    }
    return true;
}

// dJointGetSliderPosition slider.cpp:46-82 (offset of dxJointSlider::computeOffset travels in anchor1)
__device__ Real odeb_slider_position(const DJointT &j, const DBody &b0, const DBody *b1)
{
    Real ax1[3], q[3];
    mul0_331(ax1, b0.R, j.axis1);
    if (b1) {
        mul0_331(q, b1->R, j.anchor1);
        for (int i = 0; i < 3; i++) q[i] = b0.pos[i] - q[i] - b1->pos[i];
    } else {
        q[0] = b0.pos[0] - j.anchor1[0]; q[1] = b0.pos[1] - j.anchor1[1]; q[2] = b0.pos[2] - j.anchor1[2];
        if (j.reverse) { ax1[0] = -ax1[0]; ax1[1] = -ax1[1]; ax1[2] = -ax1[2]; }
    }
    return dot3(ax1, q);
}

// dxJointUniversal::getAxes universal.cpp:55-71
__device__ __forceinline__ void odeb_universal_axes(const DJointT &j, const DBody &b0, const DBody *b1, Real *ax1, Real *ax2)
{
    mul0_331(ax1, b0.R, j.axis1);
    if (b1) mul0_331(ax2, b1->R, j.axis2);
    else { ax2[0] = j.axis2[0]; ax2[1] = j.axis2[1]; ax2[2] = j.axis2[2]; }
}

// dxJointUniversal::getAngles universal.cpp:73-166
__device__ void odeb_universal_angles(const DJointT &j, const DBody &b0, const DBody *b1, Real *angle1, Real *angle2)
{
    Real ax1[3], ax2[3], R[12], qcross[4], qq[4], qrel[4];
    odeb_universal_axes(j, b0, b1, ax1, ax2);
    odeb_r_from_2axes(R, ax1[0], ax1[1], ax1[2], ax2[0], ax2[1], ax2[2]);
    q_from_r(qcross, R);
    qmul1(qq, b0.q, qcross);
    qmul2(qrel, qq, j.qrel1);
    *angle1 = odeb_hinge_angle_from_relq(qrel, j.axis1);
    Real qcross2[4];
    qrel[0] = 0; qrel[1] = ax1[0] + ax2[0]; qrel[2] = ax1[1] + ax2[1]; qrel[3] = ax1[2] + ax2[2];
    Real l = rrecip(RSQRT(qrel[1] * qrel[1] + qrel[2] * qrel[2] + qrel[3] * qrel[3]));
    qrel[1] *= l; qrel[2] *= l; qrel[3] *= l;
    qmul0(qcross2, qrel, qcross);
    if (b1) { qmul1(qq, b1->q, qcross2); qmul2(qrel, qq, j.qrel2); }
    else qmul2(qrel, qcross2, j.qrel2);
    *angle2 = -odeb_hinge_angle_from_relq(qrel, j.axis2);
}

// dxJointLMotor::computeGlobalAxes lmotor.cpp:51-76 / dxJointAMotor::doComputeGlobalUserAxes, doComputeGlobalEulerAxes amotor.cpp:662-713
__device__ void odeb_motor_global_axes(const DJointT &j, const DBody &b0, const DBody *b1, Real ax[3][3])
{
    for (int i = 0; i < 3; i++) ax[i][0] = ax[i][1] = ax[i][2] = 0;
    if (j.type == 9 && j.mmode == 1) {
        const int first = j.reverse ? 2 : 0, second = 2 - first;     // BuildFirstBodyEulerAxis amotor.cpp:798-807
        mul0_331(ax[first], b0.R, j.maxis[first]);
        if (b1) mul0_331(ax[second], b1->R, j.maxis[second]);
        else { ax[second][0] = j.maxis[second][0]; ax[second][1] = j.maxis[second][1]; ax[second][2] = j.maxis[second][2]; }
        cross3(ax[1], ax[2], ax[0]);
        normalize3(ax[1]);
        return;
    }
    for (int i = 0; i < j.mnum; i++) {
        if (j.mrel[i] == 1) mul0_331(ax[i], b0.R, j.maxis[i]);
        else if (j.mrel[i] == 2 && b1) mul0_331(ax[i], b1->R, j.maxis[i]);
        else if (j.mrel[i] == 2 && j.type == 10) { }                  // lmotor.cpp:61-66: left as it is
        else { ax[i][0] = j.maxis[i][0]; ax[i][1] = j.maxis[i][1]; ax[i][2] = j.maxis[i][2]; }
    }
}

// getInfo1 of hinge (hinge.cpp:54-74) and universal (universal.cpp:266-293): row count + limit state
__device__ void odeb_joint_info1(const DJointT &j, const DBody &b0, const DBody *b1, int *m, DLimitState *ls)
{
    ls->limit1 = ls->limit2 = ls->limit3 = 0; ls->err1 = ls->err2 = ls->err3 = 0;
    if (j.type == 10) {                            // lmotor.cpp:84-97
        int mm = 0;
        if (j.mnum > 0 && j.limot1.fmax > 0) mm++;
        if (j.mnum > 1 && j.limot2.fmax > 0) mm++;
        if (j.mnum > 2 && j.limot3.fmax > 0) mm++;
        *m = mm;
        return;
    }
    if (j.type == 9) {                             // amotor.cpp:264-287, computeEulerAngles :716-758
        Real ang[3] = { j.mangle[0], j.mangle[1], j.mangle[2] };
        if (j.mmode == 1) {
            Real ax[3][3], refs[2][3], q[3];
            odeb_motor_global_axes(j, b0, b1, ax);
            mul0_331(refs[0], b0.R, j.mref[0]);
            if (b1) mul0_331(refs[1], b1->R, j.mref[1]);
            else { refs[1][0] = j.mref[1][0]; refs[1][1] = j.mref[1][1]; refs[1][2] = j.mref[1][2]; }
            const int fb = j.reverse ? 1 : 0, sb = 1 - fb;
            cross3(q, ax[0], refs[fb]);
            ang[0] = -RATAN2(dot3(ax[2], q), dot3(ax[2], refs[fb]));
            cross3(q, ax[0], ax[1]);
            ang[1] = -RATAN2(dot3(ax[2], ax[0]), dot3(ax[2], q));
            cross3(q, ax[1], ax[2]);
            ang[2] = -RATAN2(dot3(refs[sb], ax[1]), dot3(refs[sb], q));
        }
        int mm = 0;
        if (j.mnum > 0 && (odeb_limot_test(j.limot1, ang[0], &ls->limit1, &ls->err1) || j.limot1.fmax > 0)) mm++;
        if (j.mnum > 1 && (odeb_limot_test(j.limot2, ang[1], &ls->limit2, &ls->err2) || j.limot2.fmax > 0)) mm++;
        if (j.mnum > 2 && (odeb_limot_test(j.limot3, ang[2], &ls->limit3, &ls->err3) || j.limot3.fmax > 0)) mm++;
        *m = mm;
        return;
    }
    if (j.type == 1) { *m = 3; return; }
    if (j.type == 7) { *m = 6; return; }           // fixed.cpp:52-57
    if (j.type == 6) {                             // hinge2.cpp:110-130 (needs both bodies)
        int mm = 4;
        if ((j.limot1.lostop >= -M_PI || j.limot1.histop <= M_PI) && j.limot1.lostop <= j.limot1.histop) {
            Real p[3], q[3];                       // measureAngle1 hinge2.cpp:33-52
            mul0_331(p, b1->R, j.axis2);
            mul1_331(q, b0.R, p);
            Real x = dot3(j.qrel1, q), y = dot3(j.qrel2, q);
            odeb_limot_test(j.limot1, -RATAN2(y, x), &ls->limit1, &ls->err1);
        }
        if (ls->limit1 || j.limot1.fmax > 0) mm++;
        if (j.limot2.fmax > 0) mm++;
        *m = mm;
        return;
    }
    if (j.type == 3) {                             // slider.cpp:115-145
        int mm = (j.limot1.fmax > 0) ? 6 : 5;
        if ((j.limot1.lostop > -R_INF || j.limot1.histop < R_INF) && j.limot1.lostop <= j.limot1.histop) {
            Real pos = odeb_slider_position(j, b0, b1);
            if (odeb_limot_test(j.limot1, pos, &ls->limit1, &ls->err1)) mm = 6;
        }
        *m = mm;
        return;
    }
    if (j.type == 2) {
        int mm = (j.limot1.fmax > 0) ? 6 : 5;
        if ((j.limot1.lostop >= -M_PI || j.limot1.histop <= M_PI) && j.limot1.lostop <= j.limot1.histop) {
            Real angle = odeb_hinge_angle(b0, b1, j.axis1, j.qrel);
            if (odeb_limot_test(j.limot1, angle, &ls->limit1, &ls->err1)) mm = 6;
        }
        *m = mm;
        return;
    }
    int mm = 4;
    bool lim1 = (j.limot1.lostop >= -M_PI || j.limot1.histop <= M_PI) && j.limot1.lostop <= j.limot1.histop;
    bool lim2 = (j.limot2.lostop >= -M_PI || j.limot2.histop <= M_PI) && j.limot2.lostop <= j.limot2.histop;
    if (lim1 || lim2) {
        Real a1, a2;
        odeb_universal_angles(j, b0, b1, &a1, &a2);
        if (lim1) odeb_limot_test(j.limot1, a1, &ls->limit1, &ls->err1);
        if (lim2) odeb_limot_test(j.limot2, a2, &ls->limit2, &ls->err2);
    }
    if (ls->limit1 || j.limot1.fmax > 0) mm++;
    if (ls->limit2 || j.limot2.fmax > 0) mm++;
    *m = mm;
}

// setFixedOrientation joints/joint.cpp:228-284: three rows that keep the relative rotation at qrel
__device__ void odeb_set_fixed_orientation(const DBody &b0, const DBody *b1, Real fps, Real erp, Real *row, const Real *qrel)
{
    row[C_J1A] = 1; row[ROWLEN + C_J1A + 1] = 1; row[2 * ROWLEN + C_J1A + 2] = 1;
    if (b1) { row[C_J2A] = -1; row[ROWLEN + C_J2A + 1] = -1; row[2 * ROWLEN + C_J2A + 2] = -1; }
    Real qerr[4], e[3];
    if (b1) { Real qq[4]; qmul1(qq, b0.q, b1->q); qmul2(qerr, qq, qrel); }
    else odeb_qmul3(qerr, b0.q, qrel);
    if (qerr[0] < 0) { qerr[1] = -qerr[1]; qerr[2] = -qerr[2]; qerr[3] = -qerr[3]; }
    mul0_331(e, b0.R, qerr + 1);
    const Real k2 = fps * erp * R_(2.0);
    row[C_RHS] = k2 * e[0]; row[ROWLEN + C_RHS] = k2 * e[1]; row[2 * ROWLEN + C_RHS] = k2 * e[2];
}

// getInfo2 of ball (ball.cpp:57-67), hinge (hinge.cpp:77-147), universal (universal.cpp:297-369), fixed (fixed.cpp:60-110).
// tq0 accumulates the torque the limit motors add to body0 (negated) / body1.
__device__ void odeb_joint_info2(const DJointT &j, const DLimitState &ls, const DBody &b0, const DBody *b1,
                                 Real fps, Real worldERP, Real *row, Real *tq, bool *has_tq, Real *fq, Real *tboth, bool *has_f)
{
    if (j.type == 10) {  // lmotor.cpp:99-116: one linear limit-motor row per powered axis (no stops: the limit state is never set)
        Real ax[3][3];
        odeb_motor_global_axes(j, b0, b1, ax);
        int r = 0;
        for (int i = 0; i < j.mnum; i++) {
            const DLimot &l = i == 0 ? j.limot1 : i == 1 ? j.limot2 : j.limot3;
            Real f[3], tb[3]; bool hf = false;
            if (odeb_add_limot_linear(l, 0, 0, b0, b1, fps, row + r * ROWLEN, ax[i], f, tb, &hf)) r++;
        }
        return;
    }
    if (j.type == 9) {   // amotor.cpp:290-339: Euler mode constrains along ax1 x ax2, ax1, ax0 x ax1
        Real ax[3][3], c01[3], c12[3];
        odeb_motor_global_axes(j, b0, b1, ax);
        const Real *axp[3] = { ax[0], ax[1], ax[2] };
        if (j.mmode == 1) {
            cross3(c01, ax[0], ax[1]); axp[2] = c01;
            cross3(c12, ax[1], ax[2]); axp[0] = c12;
        }
        int r = 0;
        for (int i = 0; i < j.mnum; i++) {
            const DLimot &l = i == 0 ? j.limot1 : i == 1 ? j.limot2 : j.limot3;
            const int lim = i == 0 ? ls.limit1 : i == 1 ? ls.limit2 : ls.limit3;
            const Real err = i == 0 ? ls.err1 : i == 1 ? ls.err2 : ls.err3;
            Real t[3]; bool ht = false;
            if (odeb_add_limot(l, lim, err, b0, b1, fps, row + r * ROWLEN, axp[i], t, &ht)) r++;
            if (ht) { tq[0] += t[0]; tq[1] += t[1]; tq[2] += t[2]; *has_tq = true; }
        }
        return;
    }
    if (j.type == 6) {   // hinge2.cpp:155-209 with setBall2 joints/joint.cpp:165-213 (two-body form)
        Real ax1[3], ax2[3], q[3];
        mul0_331(ax1, b0.R, j.axis1);
        mul0_331(ax2, b1->R, j.axis2);
        cross3(q, ax1, ax2);
        const Real sn = RSQRT(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]), cs = dot3(ax1, ax2);
        normalize3(q);
        {   // setBall2: rows along (ax1, q1, q2); row 0 is the suspension (its own ERP / CFM)
            Real q1[3], q2[3], a1[3], a2[3];
            plane_space(ax1, q1, q2);
            Real *r0 = row, *r1 = row + ROWLEN, *r2 = row + 2 * ROWLEN;
            for (int t = 0; t < 3; t++) { r0[C_J1L + t] = ax1[t]; r1[C_J1L + t] = q1[t]; r2[C_J1L + t] = q2[t]; }
            mul0_331(a1, b0.R, j.anchor1);
            cross3(r0 + C_J1A, a1, ax1); cross3(r1 + C_J1A, a1, q1); cross3(r2 + C_J1A, a1, q2);
            a1[0] = a1[0] + b0.pos[0]; a1[1] = a1[1] + b0.pos[1]; a1[2] = a1[2] + b0.pos[2];
            const Real k1 = fps * j.qrel[2], k = fps * worldERP;
            for (int t = 0; t < 3; t++) { r0[C_J2L + t] = -ax1[t]; r1[C_J2L + t] = -q1[t]; r2[C_J2L + t] = -q2[t]; }
            mul0_331(a2, b1->R, j.anchor2);
            cross3(r0 + C_J2A, ax1, a2); cross3(r1 + C_J2A, q1, a2); cross3(r2 + C_J2A, q2, a2);
            a2[0] = a2[0] + b1->pos[0]; a2[1] = a2[1] + b1->pos[1]; a2[2] = a2[2] + b1->pos[2];
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
        Real t[3];
        bool ht = false;
        if (odeb_add_limot(j.limot1, ls.limit1, ls.err1, b0, b1, fps, row + r * ROWLEN, ax1, t, &ht)) r++;
        if (ht) { tq[0] += t[0]; tq[1] += t[1]; tq[2] += t[2]; *has_tq = true; }
        ht = false;
        odeb_add_limot(j.limot2, 0, 0, b0, b1, fps, row + r * ROWLEN, ax2, t, &ht);
        return;
    }
    if (j.type == 3) {   // slider.cpp:148-246
        odeb_set_fixed_orientation(b0, b1, fps, worldERP, row, j.qrel);
        Real ax1[3], p[3], q[3], c[3] = { 0, 0, 0 };
        mul0_331(ax1, b0.R, j.axis1);
        plane_space(ax1, p, q);
        if (b1) { c[0] = b1->pos[0] - b0.pos[0]; c[1] = b1->pos[1] - b0.pos[1]; c[2] = b1->pos[2] - b0.pos[2]; }
        Real *r3 = row + 3 * ROWLEN, *r4 = row + 4 * ROWLEN;
        r3[C_J1L] = p[0]; r3[C_J1L + 1] = p[1]; r3[C_J1L + 2] = p[2];
        r4[C_J1L] = q[0]; r4[C_J1L + 1] = q[1]; r4[C_J1L + 2] = q[2];
        if (b1) {
            Real tmp[3];
            r3[C_J2L] = -p[0]; r3[C_J2L + 1] = -p[1]; r3[C_J2L + 2] = -p[2];
            cross3(tmp, c, p);
            for (int t = 0; t < 3; t++) { r3[C_J1A + t] = tmp[t] * R_(0.5); r3[C_J2A + t] = r3[C_J1A + t]; }
            r4[C_J2L] = -q[0]; r4[C_J2L + 1] = -q[1]; r4[C_J2L + 2] = -q[2];
            cross3(tmp, c, q);
            for (int t = 0; t < 3; t++) { r4[C_J1A + t] = tmp[t] * R_(0.5); r4[C_J2A + t] = r4[C_J1A + t]; }
        }
        const Real k = fps * worldERP;
        if (b1) {
            Real ofs[3];
            mul0_331(ofs, b1->R, j.anchor1);
            c[0] = c[0] + ofs[0]; c[1] = c[1] + ofs[1]; c[2] = c[2] + ofs[2];
            r3[C_RHS] = k * dot3(p, c);
            r4[C_RHS] = k * dot3(q, c);
        } else {
            Real ofs[3] = { j.anchor1[0] - b0.pos[0], j.anchor1[1] - b0.pos[1], j.anchor1[2] - b0.pos[2] };
            r3[C_RHS] = k * dot3(p, ofs);
            r4[C_RHS] = k * dot3(q, ofs);
            if (j.reverse) { ax1[0] = -ax1[0]; ax1[1] = -ax1[1]; ax1[2] = -ax1[2]; }
        }
        odeb_add_limot_linear(j.limot1, ls.limit1, ls.err1, b0, b1, fps, row + 5 * ROWLEN, ax1, fq, tboth, has_f);
        return;
    }
    if (j.type == 1) {
        row[C_CFM] = j.cfm; row[ROWLEN + C_CFM] = j.cfm; row[2 * ROWLEN + C_CFM] = j.cfm;
        odeb_set_ball(b0, b1, fps, j.erp, row, j.anchor1, j.anchor2);
        return;
    }
    if (j.type == 7) {   // fixed.cpp:60-110; the body-1 relative offset of dJointSetFixed travels in anchor1
        odeb_set_fixed_orientation(b0, b1, fps, worldERP, row + 3 * ROWLEN, j.qrel);
        row[C_J1L] = 1; row[ROWLEN + C_J1L + 1] = 1; row[2 * ROWLEN + C_J1L + 2] = 1;
        const Real k = fps * j.erp;
        Real ofs[3];
        mul0_331(ofs, b0.R, j.anchor1);
        if (b1) {
            row[C_J1A + 1] = -ofs[2]; row[C_J1A + 2] = +ofs[1];                                  // dSetCrossMatrixPlus odemath.h:277-286
            row[ROWLEN + C_J1A + 0] = +ofs[2]; row[ROWLEN + C_J1A + 2] = -ofs[0];
            row[2 * ROWLEN + C_J1A + 0] = -ofs[1]; row[2 * ROWLEN + C_J1A + 1] = +ofs[0];
            row[C_J2L] = -1; row[ROWLEN + C_J2L + 1] = -1; row[2 * ROWLEN + C_J2L + 2] = -1;
            for (int t = 0; t < 3; t++) row[t * ROWLEN + C_RHS] = k * (b1->pos[t] - b0.pos[t] + ofs[t]);
        } else {
            for (int t = 0; t < 3; t++) row[t * ROWLEN + C_RHS] = k * (j.anchor1[t] - b0.pos[t]);
        }
        row[C_CFM] = j.cfm; row[ROWLEN + C_CFM] = j.cfm; row[2 * ROWLEN + C_CFM] = j.cfm;
        return;
    }
    odeb_set_ball(b0, b1, fps, worldERP, row, j.anchor1, j.anchor2);
    if (j.type == 2) {
        Real ax1[3], p[3], q[3];
        mul0_331(ax1, b0.R, j.axis1);
        plane_space(ax1, p, q);
        Real *r3 = row + 3 * ROWLEN, *r4 = row + 4 * ROWLEN;
        r3[C_J1A] = p[0]; r3[C_J1A + 1] = p[1]; r3[C_J1A + 2] = p[2];
        if (b1) { r3[C_J2A] = -p[0]; r3[C_J2A + 1] = -p[1]; r3[C_J2A + 2] = -p[2]; }
        r4[C_J1A] = q[0]; r4[C_J1A + 1] = q[1]; r4[C_J1A + 2] = q[2];
        if (b1) { r4[C_J2A] = -q[0]; r4[C_J2A + 1] = -q[1]; r4[C_J2A + 2] = -q[2]; }
        Real b[3];
        if (b1) { Real ax2[3]; mul0_331(ax2, b1->R, j.axis2); cross3(b, ax1, ax2); }
        else cross3(b, ax1, j.axis2);
        Real k = fps * worldERP;
        r3[C_RHS] = k * dot3(b, p);
        r4[C_RHS] = k * dot3(b, q);
        Real t[3];
        bool ht = false;
        odeb_add_limot(j.limot1, ls.limit1, ls.err1, b0, b1, fps, row + 5 * ROWLEN, ax1, t, &ht);
        if (ht) { tq[0] += t[0]; tq[1] += t[1]; tq[2] += t[2]; *has_tq = true; }
        return;
    }
    Real ax1[3], ax2[3], p[3];
    odeb_universal_axes(j, b0, b1, ax1, ax2);
    Real k = dot3(ax1, ax2);
    Real ax2t[3] = { ax2[0] + (-k) * ax1[0], ax2[1] + (-k) * ax1[1], ax2[2] + (-k) * ax1[2] };
    cross3(p, ax1, ax2t);
    normalize3(p);
    Real *r3 = row + 3 * ROWLEN;
    r3[C_J1A] = p[0]; r3[C_J1A + 1] = p[1]; r3[C_J1A + 2] = p[2];
    if (b1) { r3[C_J2A] = -p[0]; r3[C_J2A + 1] = -p[1]; r3[C_J2A + 2] = -p[2]; }
    r3[C_RHS] = fps * worldERP * (-k);
    int r = 4;
    Real t[3];
    bool ht = false;
    if (odeb_add_limot(j.limot1, ls.limit1, ls.err1, b0, b1, fps, row + r * ROWLEN, ax1, t, &ht)) r++;
    if (ht) { tq[0] += t[0]; tq[1] += t[1]; tq[2] += t[2]; *has_tq = true; }
    ht = false;
    odeb_add_limot(j.limot2, ls.limit2, ls.err2, b0, b1, fps, row + r * ROWLEN, ax2, t, &ht);
    if (ht) { tq[0] += t[0]; tq[1] += t[1]; tq[2] += t[2]; *has_tq = true; }
}

// surface parameters of the contact policy (dSurfaceParameters, include/ode/contact.h:55-72)
struct DSurface {
    int mode, the_m;
    Real mu, mu2, bounce, bounce_vel, soft_erp, soft_cfm, motion1, motion2, motionN, slip1, slip2;
    Real rho, rho2, rhoN;   // rolling / spinning friction (dContactRolling), clamped to >= 0 like getInfo1 does
    Real fdir1[3];          // dContact.fdir1, used when mode has dContactFDir1 (classic per-object API only)
};

// dxJointContact::getInfo1 row count (contact.cpp:48-122). Note the reference's asymmetry: a rolling coefficient of exactly 0
// still counts a row here, while getInfo2 (:299-343) fills rows only for rho > 0; the counted row then stays at its initial
// values (J = 0, world CFM, lo/hi = -+inf).
__host__ __device__ inline int odeb_contact_rows(int mode, Real mu, Real mu2, Real rho, Real rho2, Real rhoN)
{
    int m = 1;
    if (mode & 0x001) {
        if (mu > 0) m++;
        if (mu2 > 0) m++;
        if (mode & 0x400) { if (!(rho < 0)) m++; if (!(rho2 < 0)) m++; if (!(rhoN < 0)) m++; }
    } else {
        if (mu > 0) m += 2;
        if ((mode & 0x400) && !(rho < 0)) m += 3;
    }
    return m;
}

// dxJointContact::getInfo2 joints/contact.cpp:125-347
// STD3: the caller has checked that the contact has exactly its normal row and both friction rows (mu > 0, mu2 > 0, no rolling friction):
// the row positions are compile-time constants then, and a caller that unrolls its own row loops keeps all three rows in registers
template <bool STD3>
__device__ __forceinline__ void odeb_contact_info2(const DSurface &s, const Real *cpos, const Real *cnormal, Real cdepth, int reverse,
                                   const DBody &b0, const DBody *b1, Real fps, Real worldERP, Real min_depth, Real maxvel,
                                   Real *row, int *findex)
{
    const int mode = s.mode;
    Real erp = (mode & 0x008) ? s.soft_erp : worldERP;
    Real k = fps * erp;
    Real depth = cdepth - min_depth;
    if (depth < 0) depth = 0;
    Real motionN = (mode & 0x080) ? s.motionN : R_(0.0);
    const Real pushout = k * depth + motionN;
    bool apply_bounce = (mode & 0x004) != 0 && s.bounce_vel >= 0;
    Real outgoing = 0;
    Real c = pushout > maxvel ? maxvel : pushout;
    Real c1[3], c2[3] = { 0, 0, 0 }, normal[3];
    if (reverse) { normal[0] = -cnormal[0]; normal[1] = -cnormal[1]; normal[2] = -cnormal[2]; }
    else { normal[0] = cnormal[0]; normal[1] = cnormal[1]; normal[2] = cnormal[2]; }
    Real *J1 = row, *J2 = row + C_J2L;
    if (b1) {
        for (int i = 0; i < 3; i++) c2[i] = cpos[i] - b1->pos[i];
        J2[0] = -normal[0]; J2[1] = -normal[1]; J2[2] = -normal[2];
        cross3(J2 + 3, normal, c2);
        if (apply_bounce) outgoing = dot3(J2 + 3, b1->avel) - dot3(normal, b1->lvel);
    }
    for (int i = 0; i < 3; i++) c1[i] = cpos[i] - b0.pos[i];
    J1[0] = normal[0]; J1[1] = normal[1]; J1[2] = normal[2];
    cross3(J1 + 3, c1, normal);
    if (apply_bounce) {
        outgoing += dot3(J1 + 3, b0.avel) + dot3(normal, b0.lvel);
        Real neg_out = motionN - outgoing;
        if (neg_out > s.bounce_vel) {
            const Real newc = s.bounce * neg_out + motionN;
            if (newc > c) c = newc;
        }
    }
    row[C_RHS] = c;
    if (mode & 0x010) row[C_CFM] = s.soft_cfm;
    row[C_LO] = 0; row[C_HI] = R_INF;
    if (STD3 || s.the_m > 1) {
        Real t1[3], t2[3];
        if (mode & 0x002) { t1[0] = s.fdir1[0]; t1[1] = s.fdir1[1]; t1[2] = s.fdir1[2]; cross3(t2, normal, t1); }   // contact.cpp:221-224
        else plane_space(normal, t1, t2);
        int r = 1;
        if (STD3 || s.mu > 0) {
            Real *q = row + r * ROWLEN;
            q[C_J1L] = t1[0]; q[C_J1L + 1] = t1[1]; q[C_J1L + 2] = t1[2];
            cross3(q + C_J1A, c1, t1);
            if (b1) { q[C_J2L] = -t1[0]; q[C_J2L + 1] = -t1[1]; q[C_J2L + 2] = -t1[2]; cross3(q + C_J2A, t1, c2); }
            if (mode & 0x020) q[C_RHS] = s.motion1;
            if (mode & 0x100) q[C_CFM] = s.slip1;
            q[C_LO] = -s.mu; q[C_HI] = s.mu;
            if (mode & 0x1000) findex[r] = 0;
            r++;
        }
        const Real mu2 = (mode & 0x001) ? s.mu2 : s.mu;
        if (STD3 || mu2 > 0) {
            Real *q = row + r * ROWLEN;
            q[C_J1L] = t2[0]; q[C_J1L + 1] = t2[1]; q[C_J1L + 2] = t2[2];
            cross3(q + C_J1A, c1, t2);
            if (b1) { q[C_J2L] = -t2[0]; q[C_J2L + 1] = -t2[1]; q[C_J2L + 2] = -t2[2]; cross3(q + C_J2A, t2, c2); }
            if (mode & 0x040) q[C_RHS] = s.motion2;
            if (mode & 0x200) q[C_CFM] = s.slip2;
            q[C_LO] = -mu2; q[C_HI] = mu2;
            if (mode & 0x2000) findex[r] = 0;
            r++;
        }
        if (!STD3 && (mode & 0x400)) {   // rolling about t1, t2 and spinning about the normal (contact.cpp:299-343)
            const Real *ax[3] = { t1, t2, normal };
            const int approx_bits[3] = { 0x1000, 0x2000, 0x4000 };
            const Real rho[3] = { s.rho, (mode & 0x001) ? s.rho2 : s.rho, (mode & 0x001) ? s.rhoN : s.rho };
            for (int i = 0; i < 3; i++) {
                if (rho[i] > 0) {
                    Real *q = row + r * ROWLEN;
                    q[C_J1A] = ax[i][0]; q[C_J1A + 1] = ax[i][1]; q[C_J1A + 2] = ax[i][2];
                    if (b1) { q[C_J2A] = -ax[i][0]; q[C_J2A + 1] = -ax[i][1]; q[C_J2A + 2] = -ax[i][2]; }
                    q[C_LO] = -rho[i]; q[C_HI] = rho[i];
                    if (mode & approx_bits[i]) findex[r] = 0;
                    r++;
                }
            }
        }
    }
}
#endif
