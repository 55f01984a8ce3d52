// orc_collide.h -- TEST INFRASTRUCTURE (oracle). Scalar restatement of the reference's AABB
// functions and primitive colliders for sphere / box / capsule / plane, operation order preserved.
// Each function cites the reference code it follows.
#ifndef ORC_COLLIDE_H
#define ORC_COLLIDE_H
#include "orc_math.h"

struct OrcGeom {
    int type;        // ODEB_SPHERE / BOX / CAPSULE / CYLINDER / PLANE / RAY
    int body;        // -1 = none
    Real p[4];       // radius | sides | radius,length | plane a,b,c,d (normalised) | ray length
    unsigned cat, col;
    const Real *pos; // -> body pos (4) or own storage
    const Real *R;   // -> body R (12)
    Real aabb[6];
    int has_ofs;     // dGeomSetOffset*: opos / oR relative to the body, fpos / fR = final pose (computePosr collision_kernel.cpp:455-466)
    Real opos[4], oR[12], fpos[4], fR[12];
};

struct OrcContactGeom { Real pos[3], normal[3], depth; };

#define ORC_NUMC_MASK 0xffff
#define ORC_CONTACTS_UNIMPORTANT 0x80000000

// dxSphere::computeAABB sphere.cpp:59-67; dxBox::computeAABB box.cpp:60-77;
// dxCapsule::computeAABB capsule.cpp:60-74; dxPlane::computeAABB plane.cpp:80-107
static inline void orc_compute_aabb(OrcGeom &g)
{
    const Real *pos = g.pos, *R = g.R;
    Real *a = g.aabb;
    if (g.type == 0) {
        Real r = g.p[0];
        a[0] = pos[0] - r; a[1] = pos[0] + r; a[2] = pos[1] - r; a[3] = pos[1] + r; a[4] = pos[2] - r; a[5] = pos[2] + r;
    } else if (g.type == 1) {
        const Real *s = g.p;
        Real xr = R_(0.5) * (RFABS(R[0] * s[0]) + RFABS(R[1] * s[1]) + RFABS(R[2] * s[2]));
        Real yr = R_(0.5) * (RFABS(R[4] * s[0]) + RFABS(R[5] * s[1]) + RFABS(R[6] * s[2]));
        Real zr = R_(0.5) * (RFABS(R[8] * s[0]) + RFABS(R[9] * s[1]) + RFABS(R[10] * s[2]));
        a[0] = pos[0] - xr; a[1] = pos[0] + xr; a[2] = pos[1] - yr; a[3] = pos[1] + yr; a[4] = pos[2] - zr; a[5] = pos[2] + zr;
    } else if (g.type == 2) {
        Real radius = g.p[0], lz = g.p[1];
        Real xr = RFABS(R[2] * lz) * R_(0.5) + radius;
        Real yr = RFABS(R[6] * lz) * R_(0.5) + radius;
        Real zr = RFABS(R[10] * lz) * R_(0.5) + radius;
        a[0] = pos[0] - xr; a[1] = pos[0] + xr; a[2] = pos[1] - yr; a[3] = pos[1] + yr; a[4] = pos[2] - zr; a[5] = pos[2] + zr;
    } else if (g.type == 3) {     // dxCylinder::computeAABB cylinder.cpp:63-80
        Real radius = g.p[0], lz = g.p[1];
        Real m0 = (Real)(R_(1.0) - R[2] * R[2]), m1 = (Real)(R_(1.0) - R[6] * R[6]), m2 = (Real)(R_(1.0) - R[10] * R[10]);
        Real xr = RFABS(R[2] * lz * R_(0.5)) + radius * RSQRT(m0 > R_(0.0) ? m0 : R_(0.0));
        Real yr = RFABS(R[6] * lz * R_(0.5)) + radius * RSQRT(m1 > R_(0.0) ? m1 : R_(0.0));
        Real zr = RFABS(R[10] * lz * R_(0.5)) + radius * RSQRT(m2 > R_(0.0) ? m2 : R_(0.0));
        a[0] = pos[0] - xr; a[1] = pos[0] + xr; a[2] = pos[1] - yr; a[3] = pos[1] + yr; a[4] = pos[2] - zr; a[5] = pos[2] + zr;
    } else if (g.type == 5) {     // dxRay::computeAABB ray.cpp:58-93
        const Real len = g.p[0];
        for (int k = 0; k < 3; k++) {
            Real e = pos[k] + R[4 * k + 2] * len;
            if (pos[k] < e) { a[2 * k] = pos[k]; a[2 * k + 1] = e; } else { a[2 * k] = e; a[2 * k + 1] = pos[k]; }
        }
    } else {
        const Real *p = g.p;
        a[0] = -R_INF; a[1] = R_INF; a[2] = -R_INF; a[3] = R_INF; a[4] = -R_INF; a[5] = R_INF;
        if (p[1] == 0.0f && p[2] == 0.0f) { a[0] = (p[0] > 0) ? -R_INF : -p[3]; a[1] = (p[0] > 0) ? p[3] : R_INF; }
        else if (p[0] == 0.0f && p[2] == 0.0f) { a[2] = (p[1] > 0) ? -R_INF : -p[3]; a[3] = (p[1] > 0) ? p[3] : R_INF; }
        else if (p[0] == 0.0f && p[1] == 0.0f) { a[4] = (p[2] > 0) ? -R_INF : -p[3]; a[5] = (p[2] > 0) ? p[3] : R_INF; }
    }
}

// dCollideSpheres collision_util.cpp:38-67
static inline int orc_collide_spheres(const Real *p1, Real r1, const Real *p2, Real r2, OrcContactGeom *c)
{
    Real t[3] = { p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2] };
    Real d = RSQRT(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
    if (d > (r1 + r2)) return 0;
    if (d <= 0) {
        c->pos[0] = p1[0]; c->pos[1] = p1[1]; c->pos[2] = p1[2];
        c->normal[0] = 1; c->normal[1] = 0; c->normal[2] = 0;
        c->depth = r1 + r2;
    } else {
        Real d1 = rrecip(d);
        c->normal[0] = (p1[0] - p2[0]) * d1; c->normal[1] = (p1[1] - p2[1]) * d1; c->normal[2] = (p1[2] - p2[2]) * d1;
        Real k = R_(0.5) * (r2 - r1 - d);
        c->pos[0] = p1[0] + c->normal[0] * k; c->pos[1] = p1[1] + c->normal[1] * k; c->pos[2] = p1[2] + c->normal[2] * k;
        c->depth = r1 + r2 - d;
    }
    return 1;
}

// dCollideSphereBox sphere.cpp:133-218
static inline int orc_sphere_box(const OrcGeom &o1, const OrcGeom &o2, OrcContactGeom *c)
{
    Real l[3], t[3], p[3], q[3], r[3];
    int onborder = 0;
    const Real *R = o2.R;
    p[0] = o1.pos[0] - o2.pos[0]; p[1] = o1.pos[1] - o2.pos[1]; p[2] = o1.pos[2] - o2.pos[2];
    for (int i = 0; i < 3; i++) {
        l[i] = o2.p[i] * R_(0.5);
        t[i] = dot3s(p, 1, R + i, 4);
        if (t[i] < -l[i]) { t[i] = -l[i]; onborder = 1; }
        if (t[i] > l[i]) { t[i] = l[i]; onborder = 1; }
    }
    if (!onborder) {
        Real mind = l[0] - RFABS(t[0]);
        int mini = 0;
        for (int i = 1; i < 3; i++) { Real fd = l[i] - RFABS(t[i]); if (fd < mind) { mind = fd; mini = i; } }
        c->pos[0] = o1.pos[0]; c->pos[1] = o1.pos[1]; c->pos[2] = o1.pos[2];
        Real tmp[3] = { 0, 0, 0 };
        tmp[mini] = (t[mini] > 0) ? R_(1.0) : R_(-1.0);
        mul0_331(c->normal, R, tmp);
        c->depth = mind + o1.p[0];
        return 1;
    }
    mul0_331(q, R, t);
    r[0] = p[0] - q[0]; r[1] = p[1] - q[1]; r[2] = p[2] - q[2];
    Real depth = o1.p[0] - RSQRT(dot3(r, r));
    if (depth < 0) return 0;
    c->pos[0] = q[0] + o2.pos[0]; c->pos[1] = q[1] + o2.pos[1]; c->pos[2] = q[2] + o2.pos[2];
    c->normal[0] = r[0]; c->normal[1] = r[1]; c->normal[2] = r[2];
    normalize3(c->normal);
    c->depth = depth;
    return 1;
}

// dCollideSpherePlane sphere.cpp:221-251
static inline int orc_sphere_plane(const OrcGeom &o1, const OrcGeom &o2, OrcContactGeom *c)
{
    const Real *pl = o2.p;
    Real k = dot3(o1.pos, pl);
    Real depth = pl[3] - k + o1.p[0];
    if (depth >= 0) {
        c->normal[0] = pl[0]; c->normal[1] = pl[1]; c->normal[2] = pl[2];
        c->pos[0] = o1.pos[0] - pl[0] * o1.p[0]; c->pos[1] = o1.pos[1] - pl[1] * o1.p[0]; c->pos[2] = o1.pos[2] - pl[2] * o1.p[0];
        c->depth = depth;
        return 1;
    }
    return 0;
}

// dLineClosestApproach collision_util.cpp:70-92
static inline void orc_line_closest_approach(const Real *pa, const Real *ua, const Real *pb, const Real *ub, Real *alpha, Real *beta)
{
    Real p[3] = { pb[0] - pa[0], pb[1] - pa[1], pb[2] - pa[2] };
    Real uaub = dot3(ua, ub), q1 = dot3(ua, p), q2 = -dot3(ub, p);
    Real d = 1 - uaub * uaub;
    if (d <= R_(0.0001)) { *alpha = 0; *beta = 0; }
    else { d = rrecip(d); *alpha = (q1 + uaub * q2) * d; *beta = (uaub * q1 + q2) * d; }
}

// intersectRectQuad box.cpp:212-263
static inline int orc_intersect_rect_quad(const Real h[2], Real p[8], Real ret[16])
{
    int nq = 4, nr = 0;
    Real buffer[16];
    Real *q = p, *r = ret;
    for (int dir = 0; dir <= 1; dir++) {
        for (int sign = -1; sign <= 1; sign += 2) {
            Real *pq = q, *pr = r;
            nr = 0;
            for (int i = nq; i > 0; i--) {
                if (sign * pq[dir] < h[dir]) {
                    pr[0] = pq[0]; pr[1] = pq[1]; pr += 2; nr++;
                    if (nr & 8) { q = r; goto done; }
                }
                Real *nextq = (i > 1) ? pq + 2 : q;
                if ((sign * pq[dir] < h[dir]) ^ (sign * nextq[dir] < h[dir])) {
                    pr[1 - dir] = pq[1 - dir] + (nextq[1 - dir] - pq[1 - dir]) / (nextq[dir] - pq[dir]) * (sign * h[dir] - pq[dir]);
                    pr[dir] = sign * h[dir];
                    pr += 2; nr++;
                    if (nr & 8) { q = r; goto done; }
                }
                pq += 2;
            }
            q = r;
            r = (q == ret) ? buffer : ret;
            nq = nr;
        }
    }
done:
    if (q != ret) memcpy(ret, q, nr * 2 * sizeof(Real));
    return nr;
}

// cullPoints box.cpp:274-336 (note the double-precision M_PI expressions in the single build)
static inline void orc_cull_points(int n, const Real p[], int m, int i0, int iret[])
{
    int i, j;
    Real a, cx, cy, q;
    if (n == 1) { cx = p[0]; cy = p[1]; }
    else if (n == 2) { cx = R_(0.5) * (p[0] + p[2]); cy = R_(0.5) * (p[1] + p[3]); }
    else {
        a = 0; cx = 0; cy = 0;
        for (i = 0; i < (n - 1); i++) {
            q = p[i * 2] * p[i * 2 + 3] - p[i * 2 + 2] * p[i * 2 + 1];
            a += q;
            cx += q * (p[i * 2] + p[i * 2 + 2]);
            cy += q * (p[i * 2 + 1] + p[i * 2 + 3]);
        }
        q = p[n * 2 - 2] * p[1] - p[0] * p[n * 2 - 1];
        a = rrecip(R_(3.0) * (a + q));
        cx = a * (cx + q * (p[n * 2 - 2] + p[0]));
        cy = a * (cy + q * (p[n * 2 - 1] + p[1]));
    }
    Real A[8];
    for (i = 0; i < n; i++) A[i] = RATAN2(p[i * 2 + 1] - cy, p[i * 2] - cx);
    int avail[8];
    for (i = 0; i < n; i++) avail[i] = 1;
    avail[i0] = 0;
    iret[0] = i0;
    iret++;
    for (j = 1; j < m; j++) {
        a = (Real)((Real)j * (2 * M_PI / m) + A[i0]);
        if (a > M_PI) a -= (Real)(2 * M_PI);
        Real maxdiff = 1e9, diff;
        *iret = i0;
        for (i = 0; i < n; i++) {
            if (avail[i]) {
                diff = RFABS(A[i] - a);
                if (diff > M_PI) diff = (Real)(2 * M_PI - diff);
                if (diff < maxdiff) { maxdiff = diff; *iret = i; }
            }
        }
        avail[*iret] = 0;
        iret++;
    }
}

// dBoxBox box.cpp:356-737. Returns contact count; normal/depth/code as the reference.
static inline int orc_box_box(const Real *p1, const Real *R1, const Real *side1, const Real *p2, const Real *R2, const Real *side2,
                              Real *normal, Real *depth, int *return_code, int flags, OrcContactGeom *contact)
{
    const Real fudge = R_(1.05);
    Real p[3], pp[3], normalC[3] = { 0, 0, 0 };
    const Real *normalR = 0;
    Real A[3], B[3], Rm[3][3], Q[3][3], s, s2, l, e1;
    int i, j, invert_normal, code;
    p[0] = p2[0] - p1[0]; p[1] = p2[1] - p1[1]; p[2] = p2[2] - p1[2];
    mul1_331(pp, R1, p);
    for (i = 0; i < 3; i++) { A[i] = side1[i] * R_(0.5); B[i] = side2[i] * R_(0.5); }
    for (i = 0; i < 3; i++) for (j = 0; j < 3; j++) { Rm[i][j] = dot3s(R1 + i, 4, R2 + j, 4); Q[i][j] = RFABS(Rm[i][j]); }
    s = -R_INF; invert_normal = 0; code = 0;
    const bool unimportant = (flags & ORC_CONTACTS_UNIMPORTANT) != 0;
    do {
#define ORC_TST1(expr1, expr2, norm, cc) \
        e1 = (expr1); s2 = RFABS(e1) - (expr2); if (s2 > 0) return 0; \
        if (s2 > s) { s = s2; normalR = norm; invert_normal = (e1 < 0); code = (cc); if (unimportant) break; }
        ORC_TST1(pp[0], (A[0] + B[0] * Q[0][0] + B[1] * Q[0][1] + B[2] * Q[0][2]), R1 + 0, 1);
        ORC_TST1(pp[1], (A[1] + B[0] * Q[1][0] + B[1] * Q[1][1] + B[2] * Q[1][2]), R1 + 1, 2);
        ORC_TST1(pp[2], (A[2] + B[0] * Q[2][0] + B[1] * Q[2][1] + B[2] * Q[2][2]), R1 + 2, 3);
        ORC_TST1(dot3s(R2 + 0, 4, p, 1), (A[0] * Q[0][0] + A[1] * Q[1][0] + A[2] * Q[2][0] + B[0]), R2 + 0, 4);
        ORC_TST1(dot3s(R2 + 1, 4, p, 1), (A[0] * Q[0][1] + A[1] * Q[1][1] + A[2] * Q[2][1] + B[1]), R2 + 1, 5);
        ORC_TST1(dot3s(R2 + 2, 4, p, 1), (A[0] * Q[0][2] + A[1] * Q[1][2] + A[2] * Q[2][2] + B[2]), R2 + 2, 6);
#undef ORC_TST1
#define ORC_TST2(expr1, expr2, n1, n2, n3, cc) \
        e1 = (expr1); s2 = RFABS(e1) - (expr2); if (s2 > 0) return 0; \
        l = RSQRT((n1) * (n1) + (n2) * (n2) + (n3) * (n3)); \
        if (l > 0) { s2 /= l; if (s2 * fudge > s) { s = s2; normalR = 0; \
            normalC[0] = (n1) / l; normalC[1] = (n2) / l; normalC[2] = (n3) / l; \
            invert_normal = (e1 < 0); code = (cc); if (unimportant) break; } }
        ORC_TST2(pp[2] * Rm[1][0] - pp[1] * Rm[2][0], (A[1] * Q[2][0] + A[2] * Q[1][0] + B[1] * Q[0][2] + B[2] * Q[0][1]), 0, -Rm[2][0], Rm[1][0], 7);
        ORC_TST2(pp[2] * Rm[1][1] - pp[1] * Rm[2][1], (A[1] * Q[2][1] + A[2] * Q[1][1] + B[0] * Q[0][2] + B[2] * Q[0][0]), 0, -Rm[2][1], Rm[1][1], 8);
        ORC_TST2(pp[2] * Rm[1][2] - pp[1] * Rm[2][2], (A[1] * Q[2][2] + A[2] * Q[1][2] + B[0] * Q[0][1] + B[1] * Q[0][0]), 0, -Rm[2][2], Rm[1][2], 9);
        ORC_TST2(pp[0] * Rm[2][0] - pp[2] * Rm[0][0], (A[0] * Q[2][0] + A[2] * Q[0][0] + B[1] * Q[1][2] + B[2] * Q[1][1]), Rm[2][0], 0, -Rm[0][0], 10);
        ORC_TST2(pp[0] * Rm[2][1] - pp[2] * Rm[0][1], (A[0] * Q[2][1] + A[2] * Q[0][1] + B[0] * Q[1][2] + B[2] * Q[1][0]), Rm[2][1], 0, -Rm[0][1], 11);
        ORC_TST2(pp[0] * Rm[2][2] - pp[2] * Rm[0][2], (A[0] * Q[2][2] + A[2] * Q[0][2] + B[0] * Q[1][1] + B[1] * Q[1][0]), Rm[2][2], 0, -Rm[0][2], 12);
        ORC_TST2(pp[1] * Rm[0][0] - pp[0] * Rm[1][0], (A[0] * Q[1][0] + A[1] * Q[0][0] + B[1] * Q[2][2] + B[2] * Q[2][1]), -Rm[1][0], Rm[0][0], 0, 13);
        ORC_TST2(pp[1] * Rm[0][1] - pp[0] * Rm[1][1], (A[0] * Q[1][1] + A[1] * Q[0][1] + B[0] * Q[2][2] + B[2] * Q[2][0]), -Rm[1][1], Rm[0][1], 0, 14);
        ORC_TST2(pp[1] * Rm[0][2] - pp[0] * Rm[1][2], (A[0] * Q[1][2] + A[1] * Q[0][2] + B[0] * Q[2][1] + B[1] * Q[2][0]), -Rm[1][2], Rm[0][2], 0, 15);
#undef ORC_TST2
    } while (0);
    if (!code) return 0;
    if (normalR) { normal[0] = normalR[0]; normal[1] = normalR[4]; normal[2] = normalR[8]; }
    else mul0_331(normal, R1, normalC);
    if (invert_normal) { normal[0] = -normal[0]; normal[1] = -normal[1]; normal[2] = -normal[2]; }
    *depth = -s;

    if (code > 6) {
        Real pa[3], pb[3], sign;
        for (i = 0; i < 3; i++) pa[i] = p1[i];
        for (j = 0; j < 3; j++) {
            sign = (dot3s(normal, 1, R1 + j, 4) > 0) ? R_(1.0) : R_(-1.0);
            for (i = 0; i < 3; i++) pa[i] += sign * A[j] * R1[i * 4 + j];
        }
        for (i = 0; i < 3; i++) pb[i] = p2[i];
        for (j = 0; j < 3; j++) {
            sign = (dot3s(normal, 1, R2 + j, 4) > 0) ? R_(-1.0) : R_(1.0);
            for (i = 0; i < 3; i++) pb[i] += sign * B[j] * R2[i * 4 + j];
        }
        Real alpha, beta, ua[3], ub[3];
        for (i = 0; i < 3; i++) ua[i] = R1[(code - 7) / 3 + i * 4];
        for (i = 0; i < 3; i++) ub[i] = R2[(code - 7) % 3 + i * 4];
        orc_line_closest_approach(pa, ua, pb, ub, &alpha, &beta);
        for (i = 0; i < 3; i++) pa[i] += ua[i] * alpha;
        for (i = 0; i < 3; i++) pb[i] += ub[i] * beta;
        for (i = 0; i < 3; i++) contact[0].pos[i] = R_(0.5) * (pa[i] + pb[i]);
        contact[0].depth = *depth;
        *return_code = code;
        return 1;
    }

    const Real *Ra, *Rb, *pa, *pb, *Sa, *Sb;
    if (code <= 3) { Ra = R1; Rb = R2; pa = p1; pb = p2; Sa = A; Sb = B; }
    else { Ra = R2; Rb = R1; pa = p2; pb = p1; Sa = B; Sb = A; }
    Real normal2[3], nr[3], anr[3];
    if (code <= 3) { normal2[0] = normal[0]; normal2[1] = normal[1]; normal2[2] = normal[2]; }
    else { normal2[0] = -normal[0]; normal2[1] = -normal[1]; normal2[2] = -normal[2]; }
    mul1_331(nr, Rb, normal2);
    anr[0] = RFABS(nr[0]); anr[1] = RFABS(nr[1]); anr[2] = RFABS(nr[2]);
    int lanr, a1, a2;
    if (anr[1] > anr[0]) {
        if (anr[1] > anr[2]) { a1 = 0; lanr = 1; a2 = 2; } else { a1 = 0; a2 = 1; lanr = 2; }
    } else {
        if (anr[0] > anr[2]) { lanr = 0; a1 = 1; a2 = 2; } else { a1 = 0; a2 = 1; lanr = 2; }
    }
    Real center[3];
    if (nr[lanr] < 0) { for (i = 0; i < 3; i++) center[i] = pb[i] - pa[i] + Sb[lanr] * Rb[i * 4 + lanr]; }
    else { for (i = 0; i < 3; i++) center[i] = pb[i] - pa[i] - Sb[lanr] * Rb[i * 4 + lanr]; }
    int codeN, code1, code2;
    codeN = (code <= 3) ? code - 1 : code - 4;
    if (codeN == 0) { code1 = 1; code2 = 2; } else if (codeN == 1) { code1 = 0; code2 = 2; } else { code1 = 0; code2 = 1; }
    Real quad[8], c1, c2, m11, m12, m21, m22;
    c1 = dot3s(center, 1, Ra + code1, 4);
    c2 = dot3s(center, 1, Ra + code2, 4);
    m11 = dot3s(Ra + code1, 4, Rb + a1, 4); m12 = dot3s(Ra + code1, 4, Rb + a2, 4);
    m21 = dot3s(Ra + code2, 4, Rb + a1, 4); m22 = dot3s(Ra + code2, 4, Rb + a2, 4);
    {
        Real k1 = m11 * Sb[a1], k2 = m21 * Sb[a1], k3 = m12 * Sb[a2], k4 = m22 * Sb[a2];
        quad[0] = c1 - k1 - k3; quad[1] = c2 - k2 - k4; quad[2] = c1 - k1 + k3; quad[3] = c2 - k2 + k4;
        quad[4] = c1 + k1 + k3; quad[5] = c2 + k2 + k4; quad[6] = c1 + k1 - k3; quad[7] = c2 + k2 - k4;
    }
    Real rect[2] = { Sa[code1], Sa[code2] };
    Real ret[16];
    int n = orc_intersect_rect_quad(rect, quad, ret);
    if (n < 1) return 0;
    Real point[3 * 8], dep[8];
    Real det1 = rrecip(m11 * m22 - m12 * m21);
    m11 *= det1; m12 *= det1; m21 *= det1; m22 *= det1;
    int cnum = 0;
    for (j = 0; j < n; j++) {
        Real k1 = m22 * (ret[j * 2] - c1) - m12 * (ret[j * 2 + 1] - c2);
        Real k2 = -m21 * (ret[j * 2] - c1) + m11 * (ret[j * 2 + 1] - c2);
        for (i = 0; i < 3; i++) point[cnum * 3 + i] = center[i] + k1 * Rb[i * 4 + a1] + k2 * Rb[i * 4 + a2];
        dep[cnum] = Sa[codeN] - dot3(normal2, point + cnum * 3);
        if (dep[cnum] >= 0) {
            ret[cnum * 2] = ret[j * 2]; ret[cnum * 2 + 1] = ret[j * 2 + 1];
            cnum++;
            if ((unsigned)(cnum | ORC_CONTACTS_UNIMPORTANT) == ((unsigned)flags & (ORC_NUMC_MASK | ORC_CONTACTS_UNIMPORTANT))) break;
        }
    }
    if (cnum < 1) return 0;
    int maxc = flags & ORC_NUMC_MASK;
    if (maxc > cnum) maxc = cnum;
    if (maxc < 1) maxc = 1;
    if (cnum <= maxc) {
        for (j = 0; j < cnum; j++) {
            for (i = 0; i < 3; i++) contact[j].pos[i] = point[j * 3 + i] + pa[i];
            contact[j].depth = dep[j];
        }
    } else {
        int i1 = 0;
        Real maxdepth = dep[0];
        for (i = 1; i < cnum; i++) if (dep[i] > maxdepth) { maxdepth = dep[i]; i1 = i; }
        int iret[8];
        orc_cull_points(cnum, ret, maxc, i1, iret);
        for (j = 0; j < maxc; j++) {
            for (i = 0; i < 3; i++) contact[j].pos[i] = point[iret[j] * 3 + i] + pa[i];
            contact[j].depth = dep[iret[j]];
        }
        cnum = maxc;
    }
    *return_code = code;
    return cnum;
}

// dCollideBoxBox box.cpp:741-767
static inline int orc_collide_box_box(const OrcGeom &o1, const OrcGeom &o2, int flags, OrcContactGeom *c)
{
    Real normal[3], depth; int code;
    int num = orc_box_box(o1.pos, o1.R, o1.p, o2.pos, o2.R, o2.p, normal, &depth, &code, flags, c);
    for (int i = 0; i < num; i++) { c[i].normal[0] = -normal[0]; c[i].normal[1] = -normal[1]; c[i].normal[2] = -normal[2]; }
    return num;
}

// dCollideBoxPlane box.cpp:770-903
static inline int orc_box_plane(const OrcGeom &o1, const OrcGeom &o2, int flags, OrcContactGeom *c)
{
    const Real *R = o1.R, *n = o2.p, *side = o1.p;
    int ret = 0;
    Real Q1 = dot3s(n, 1, R + 0, 4), Q2 = dot3s(n, 1, R + 1, 4), Q3 = dot3s(n, 1, R + 2, 4);
    Real Av[3] = { side[0] * Q1, side[1] * Q2, side[2] * Q3 };
    Real Bv[3] = { RFABS(Av[0]), RFABS(Av[1]), RFABS(Av[2]) };
    Real depth = n[3] + R_(0.5) * (Bv[0] + Bv[1] + Bv[2]) - dot3(n, o1.pos);
    if (depth < 0) return 0;
    int maxc = flags & ORC_NUMC_MASK;
    if (maxc > 4) maxc = 4;
    Real p[3] = { o1.pos[0], o1.pos[1], o1.pos[2] };
    for (int i = 0; i < 3; i++) {
        if (Av[i] > 0) { p[0] -= R_(0.5) * side[i] * R[0 + i]; p[1] -= R_(0.5) * side[i] * R[4 + i]; p[2] -= R_(0.5) * side[i] * R[8 + i]; }
        else { p[0] += R_(0.5) * side[i] * R[0 + i]; p[1] += R_(0.5) * side[i] * R[4 + i]; p[2] += R_(0.5) * side[i] * R[8 + i]; }
    }
    c[0].pos[0] = p[0]; c[0].pos[1] = p[1]; c[0].pos[2] = p[2]; c[0].depth = depth;
    ret = 1;
    if (maxc != 1) {
        // choose the two sides with the smallest projected length, in the reference's order
        int first, second;
        if (Bv[0] < Bv[1]) {
            if (Bv[2] < Bv[0]) { first = 2; second = (Bv[0] < Bv[1]) ? 0 : 1; }
            else { first = 0; second = (Bv[1] < Bv[2]) ? 1 : 2; }
        } else {
            if (Bv[2] < Bv[1]) { first = 2; second = (Bv[0] < Bv[1]) ? 0 : 1; }
            else { first = 1; second = (Bv[0] < Bv[2]) ? 0 : 2; }
        }
        const int order[2] = { first, second };
        for (int k = 0; k < 2; k++) {
            int sd = order[k], ci = k + 1;
            if (depth - Bv[sd] < 0) break;
            if (Av[sd] > 0) { c[ci].pos[0] = p[0] + side[sd] * R[0 + sd]; c[ci].pos[1] = p[1] + side[sd] * R[4 + sd]; c[ci].pos[2] = p[2] + side[sd] * R[8 + sd]; }
            else { c[ci].pos[0] = p[0] - side[sd] * R[0 + sd]; c[ci].pos[1] = p[1] - side[sd] * R[4 + sd]; c[ci].pos[2] = p[2] - side[sd] * R[8 + sd]; }
            c[ci].depth = depth - Bv[sd];
            ret++;
            if (k == 0 && maxc == 2) break;
        }
    }
    if (maxc == 4 && ret == 3) {
        Real d4 = c[1].depth + c[2].depth - depth;
        if (d4 > 0) {
            c[3].pos[0] = c[1].pos[0] + c[2].pos[0] - p[0];
            c[3].pos[1] = c[1].pos[1] + c[2].pos[1] - p[1];
            c[3].pos[2] = c[1].pos[2] + c[2].pos[2] - p[2];
            c[3].depth = d4;
            ret++;
        }
    }
    for (int i = 0; i < ret; i++) { c[i].normal[0] = n[0]; c[i].normal[1] = n[1]; c[i].normal[2] = n[2]; }
    return ret;
}

#include "orc_capsule.h"
#include "orc_ray_cyl.h"

// dCollide collision_kernel.cpp:292-338 with the collider table of dInitColliders (:166-268):
// direct entries (sphere,sphere) (sphere,box) (sphere,plane) (box,box) (box,plane) (capsule,sphere)
// (capsule,box) (capsule,capsule) (capsule,plane); the transposed pairs call the same function with
// swapped arguments and negate the normals.
static inline int orc_collide_direct(const OrcGeom &a, const OrcGeom &b, int flags, OrcContactGeom *c, int *handled)
{
    *handled = 1;
    if (a.type == 0 && b.type == 0) return orc_collide_spheres(a.pos, a.p[0], b.pos, b.p[0], c);
    if (a.type == 0 && b.type == 1) return orc_sphere_box(a, b, c);
    if (a.type == 0 && b.type == 4) return orc_sphere_plane(a, b, c);
    if (a.type == 1 && b.type == 1) return orc_collide_box_box(a, b, flags, c);
    if (a.type == 1 && b.type == 4) return orc_box_plane(a, b, flags, c);
    if (a.type == 2 && b.type == 0) return orc_capsule_sphere(a, b, c);
    if (a.type == 2 && b.type == 1) return orc_capsule_box(a, b, flags, c);
    if (a.type == 2 && b.type == 2) return orc_capsule_capsule(a, b, flags, c);
    if (a.type == 2 && b.type == 4) return orc_capsule_plane(a, b, flags, c);
    if (a.type == 5 && b.type == 0) return orc_ray_sphere(a, b, c);          // collision_kernel.cpp:191-194
    if (a.type == 5 && b.type == 1) return orc_ray_box(a, b, c);
    if (a.type == 5 && b.type == 2) return orc_ray_capsule(a, b, c);
    if (a.type == 5 && b.type == 4) return orc_ray_plane(a, b, c);
    if (a.type == 5 && b.type == 3) return orc_ray_cylinder(a, b, c);        // :195
    if (a.type == 3 && b.type == 1) return orc_cylinder_box(a, b, flags, c);   // :210
    if (a.type == 3 && b.type == 0) return orc_cylinder_sphere(a, b, c);     // :212-213 (cylinder-capsule and cylinder-cylinder have no
    if (a.type == 3 && b.type == 4) return orc_cylinder_plane(a, b, flags, c); //  collider without libccd)
    *handled = 0;
    return 0;
}

static inline int orc_collide(const OrcGeom &o1, const OrcGeom &o2, int flags, OrcContactGeom *c)
{
    if ((flags & ORC_NUMC_MASK) == 0) return 0;
    if (&o1 == &o2) return 0;
    if (o1.body == o2.body && o1.body >= 0) return 0;
    int handled;
    int n = orc_collide_direct(o1, o2, flags, c, &handled);
    if (handled) return n;
    n = orc_collide_direct(o2, o1, flags, c, &handled);
    if (!handled) return 0;
    for (int i = 0; i < n; i++) { c[i].normal[0] = -c[i].normal[0]; c[i].normal[1] = -c[i].normal[1]; c[i].normal[2] = -c[i].normal[2]; }
    return n;
}
#endif
