// orc_capsule.h -- TEST INFRASTRUCTURE (oracle). Capsule colliders and the closest-point helpers they
// use, restated from ode/src/capsule.cpp:130-415 and ode/src/collision_util.cpp:109-384.
// Included from orc_collide.h (needs OrcGeom, OrcContactGeom, orc_collide_spheres, orc_box_box).
#ifndef ORC_CAPSULE_H
#define ORC_CAPSULE_H

// dClosestLineSegmentPoints collision_util.cpp:109-223
static inline void orc_closest_segment_points(const Real *a1, const Real *a2, const Real *b1, const Real *b2, Real *cp1, Real *cp2)
{
    Real a1a2[3], b1b2[3], a1b1[3], a1b2[3], a2b1[3], a2b2[3], n[3];
    Real la, lb, k, da1, da2, da3, da4, db1, db2, db3, db4, det;
    for (int i = 0; i < 3; i++) { a1a2[i] = a2[i] - a1[i]; b1b2[i] = b2[i] - b1[i]; a1b1[i] = b1[i] - a1[i]; }
    da1 = dot3(a1a2, a1b1); db1 = dot3(b1b2, a1b1);
    if (da1 <= 0 && db1 >= 0) { for (int i = 0; i < 3; i++) { cp1[i] = a1[i]; cp2[i] = b1[i]; } return; }
    for (int i = 0; i < 3; i++) a1b2[i] = b2[i] - a1[i];
    da2 = dot3(a1a2, a1b2); db2 = dot3(b1b2, a1b2);
    if (da2 <= 0 && db2 <= 0) { for (int i = 0; i < 3; i++) { cp1[i] = a1[i]; cp2[i] = b2[i]; } return; }
    for (int i = 0; i < 3; i++) a2b1[i] = b1[i] - a2[i];
    da3 = dot3(a1a2, a2b1); db3 = dot3(b1b2, a2b1);
    if (da3 >= 0 && db3 >= 0) { for (int i = 0; i < 3; i++) { cp1[i] = a2[i]; cp2[i] = b1[i]; } return; }
    for (int i = 0; i < 3; i++) a2b2[i] = b2[i] - a2[i];
    da4 = dot3(a1a2, a2b2); db4 = dot3(b1b2, a2b2);
    if (da4 >= 0 && db4 <= 0) { for (int i = 0; i < 3; i++) { cp1[i] = a2[i]; cp2[i] = b2[i]; } return; }
    la = dot3(a1a2, a1a2);
    if (da1 >= 0 && da3 <= 0) {
        k = da1 / la;
        for (int i = 0; i < 3; i++) n[i] = a1b1[i] - k * a1a2[i];
        if (dot3(b1b2, n) >= 0) { for (int i = 0; i < 3; i++) { cp1[i] = a1[i] + k * a1a2[i]; cp2[i] = b1[i]; } return; }
    }
    if (da2 >= 0 && da4 <= 0) {
        k = da2 / la;
        for (int i = 0; i < 3; i++) n[i] = a1b2[i] - k * a1a2[i];
        if (dot3(b1b2, n) <= 0) { for (int i = 0; i < 3; i++) { cp1[i] = a1[i] + k * a1a2[i]; cp2[i] = b2[i]; } return; }
    }
    lb = dot3(b1b2, b1b2);
    if (db1 <= 0 && db2 >= 0) {
        k = -db1 / lb;
        for (int i = 0; i < 3; i++) n[i] = -a1b1[i] - k * b1b2[i];
        if (dot3(a1a2, n) >= 0) { for (int i = 0; i < 3; i++) { cp1[i] = a1[i]; cp2[i] = b1[i] + k * b1b2[i]; } return; }
    }
    if (db3 <= 0 && db4 >= 0) {
        k = -db3 / lb;
        for (int i = 0; i < 3; i++) n[i] = -a2b1[i] - k * b1b2[i];
        if (dot3(a1a2, n) <= 0) { for (int i = 0; i < 3; i++) { cp1[i] = a2[i]; cp2[i] = b1[i] + k * b1b2[i]; } return; }
    }
    k = dot3(a1a2, b1b2);
    det = la * lb - k * k;
    if (det <= 0) { for (int i = 0; i < 3; i++) { cp1[i] = a1[i]; cp2[i] = b1[i]; } return; }
    det = rrecip(det);
    Real alpha = (lb * da1 - k * db1) * det;
    Real beta = (k * da1 - la * db1) * det;
    for (int i = 0; i < 3; i++) { cp1[i] = a1[i] + alpha * a1a2[i]; cp2[i] = b1[i] + beta * b1b2[i]; }
}

// dClosestLineBoxPoints collision_util.cpp:247-384
static inline void orc_closest_line_box_points(const Real *p1, const Real *p2, const Real *c, const Real *R, const Real *side, Real *lret, Real *bret)
{
    int i;
    Real tmp[3], s[3], v[3], sign[3], v2[3], h[3], tanchor[3];
    int region[3];
    tmp[0] = p1[0] - c[0]; tmp[1] = p1[1] - c[1]; tmp[2] = p1[2] - c[2];
    mul1_331(s, R, tmp);
    tmp[0] = p2[0] - p1[0]; tmp[1] = p2[1] - p1[1]; tmp[2] = p2[2] - p1[2];
    mul1_331(v, R, tmp);
    for (i = 0; i < 3; i++) { if (v[i] < 0) { s[i] = -s[i]; v[i] = -v[i]; sign[i] = -1; } else sign[i] = 1; }
    for (i = 0; i < 3; i++) { v2[i] = v[i] * v[i]; h[i] = R_(0.5) * side[i]; }
#if defined(ODEB_DOUBLE)
    const Real eps = R_(1e-307);
#else
    const Real eps = R_(1e-19);
#endif
    for (i = 0; i < 3; i++) {
        if (v[i] > eps) {
            if (s[i] < -h[i]) { region[i] = -1; tanchor[i] = (-h[i] - s[i]) / v[i]; }
            else { region[i] = (s[i] > h[i]); tanchor[i] = (h[i] - s[i]) / v[i]; }
        } else { region[i] = 0; tanchor[i] = 2; }
    }
    Real t = 0, dd2dt = 0;
    for (i = 0; i < 3; i++) dd2dt -= (region[i] ? v2[i] : 0) * tanchor[i];
    if (!(dd2dt >= 0)) {
        bool answered = false;
        do {
            Real next_t = 1;
            for (i = 0; i < 3; i++) if (tanchor[i] > t && tanchor[i] < 1 && tanchor[i] < next_t) next_t = tanchor[i];
            Real next_dd2dt = 0;
            for (i = 0; i < 3; i++) next_dd2dt += (region[i] ? v2[i] : 0) * (next_t - tanchor[i]);
            if (next_dd2dt >= 0) {
                Real m = (next_dd2dt - dd2dt) / (next_t - t);
                t -= dd2dt / m;
                answered = true;
                break;
            }
            for (i = 0; i < 3; i++) if (tanchor[i] == next_t) { tanchor[i] = (h[i] - s[i]) / v[i]; region[i]++; }
            t = next_t;
            dd2dt = next_dd2dt;
        } while (t < 1);
        if (!answered) t = 1;
    }
    for (i = 0; i < 3; i++) lret[i] = p1[i] + t * tmp[i];
    for (i = 0; i < 3; i++) {
        tmp[i] = sign[i] * (s[i] + t * v[i]);
        if (tmp[i] < -h[i]) tmp[i] = -h[i]; else if (tmp[i] > h[i]) tmp[i] = h[i];
    }
    mul0_331(s, R, tmp);
    for (i = 0; i < 3; i++) bret[i] = s[i] + c[i];
}

// dCollideCapsuleSphere capsule.cpp:130-163
static inline int orc_capsule_sphere(const OrcGeom &o1, const OrcGeom &o2, OrcContactGeom *c)
{
    const Real *R = o1.R;
    Real alpha = R[2] * (o2.pos[0] - o1.pos[0]) + R[6] * (o2.pos[1] - o1.pos[1]) + R[10] * (o2.pos[2] - o1.pos[2]);
    Real lz2 = o1.p[1] * R_(0.5);
    if (alpha > lz2) alpha = lz2;
    if (alpha < -lz2) alpha = -lz2;
    Real p[3] = { o1.pos[0] + alpha * R[2], o1.pos[1] + alpha * R[6], o1.pos[2] + alpha * R[10] };
    return orc_collide_spheres(p, o1.p[0], o2.pos, o2.p[0], c);
}

static inline int orc_box_box(const Real *p1, const Real *R1, const Real *side1, const Real *p2, const Real *R2, const Real *side2,
                              Real *normal, Real *depth, int *return_code, int flags, OrcContactGeom *contact);

// dCollideCapsuleBox capsule.cpp:166-237
static inline int orc_capsule_box(const OrcGeom &o1, const OrcGeom &o2, int flags, OrcContactGeom *contact)
{
    const Real *R1 = o1.R;
    Real p1[3], p2[3];
    Real clen = o1.p[1] * R_(0.5);
    p1[0] = o1.pos[0] + clen * R1[2]; p1[1] = o1.pos[1] + clen * R1[6]; p1[2] = o1.pos[2] + clen * R1[10];
    p2[0] = o1.pos[0] - clen * R1[2]; p2[1] = o1.pos[1] - clen * R1[6]; p2[2] = o1.pos[2] - clen * R1[10];
    Real radius = o1.p[0];
    Real pl[3], pb[3];
    orc_closest_line_box_points(p1, p2, o2.pos, o2.R, o2.p, pl, pb);
#if defined(ODEB_DOUBLE)
    Real mindist = R_(1e-15);
#else
    Real mindist = R_(1e-6);
#endif
    Real d[3] = { pl[0] - pb[0], pl[1] - pb[1], pl[2] - pb[2] };
    if (RSQRT(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) < mindist) {
        Real normal[3], depth; int code;
        Real r2 = radius * R_(2.0);
        const Real capboxside[3] = { r2, r2, o1.p[1] + r2 };
        int num = orc_box_box(o2.pos, o2.R, o2.p, o1.pos, o1.R, capboxside, normal, &depth, &code, flags, contact);
        for (int i = 0; i < num; i++) { contact[i].normal[0] = normal[0]; contact[i].normal[1] = normal[1]; contact[i].normal[2] = normal[2]; }
        return num;
    }
    return orc_collide_spheres(pl, radius, pb, 0, contact);
}

// dCollideCapsuleCapsule capsule.cpp:240-353
static inline int orc_capsule_capsule(const OrcGeom &o1, const OrcGeom &o2, int flags, OrcContactGeom *contact)
{
    int i;
    const Real tolerance = R_(1e-5);
    Real lz1 = o1.p[1] * R_(0.5), lz2 = o2.p[1] * R_(0.5);
    const Real *pos1 = o1.pos, *pos2 = o2.pos;
    Real axis1[3] = { o1.R[2], o1.R[6], o1.R[10] }, axis2[3] = { o2.R[2], o2.R[6], o2.R[10] };
    Real sphere1[3], sphere2[3];
    Real a1a2 = dot3(axis1, axis2);
    Real det = R_(1.0) - a1a2 * a1a2;
    if (det < tolerance) {
        if (a1a2 < 0) { axis2[0] = -axis2[0]; axis2[1] = -axis2[1]; axis2[2] = -axis2[2]; }
        Real q[3];
        for (i = 0; i < 3; i++) q[i] = pos1[i] - pos2[i];
        Real k = dot3(axis1, q);
        Real a1lo = -lz1, a1hi = lz1, a2lo = -lz2 - k, a2hi = lz2 - k;
        Real lo = (a1lo > a2lo) ? a1lo : a2lo;
        Real hi = (a1hi < a2hi) ? a1hi : a2hi;
        if (lo <= hi) {
            int num_contacts = flags & ORC_NUMC_MASK;
            if (num_contacts >= 2 && lo < hi) {
                for (i = 0; i < 3; i++) sphere1[i] = pos1[i] + lo * axis1[i];
                for (i = 0; i < 3; i++) sphere2[i] = pos2[i] + (lo + k) * axis2[i];
                int n1 = orc_collide_spheres(sphere1, o1.p[0], sphere2, o2.p[0], contact);
                if (n1) {
                    for (i = 0; i < 3; i++) sphere1[i] = pos1[i] + hi * axis1[i];
                    for (i = 0; i < 3; i++) sphere2[i] = pos2[i] + (hi + k) * axis2[i];
                    int n2 = orc_collide_spheres(sphere1, o1.p[0], sphere2, o2.p[0], contact + 1);
                    if (n2) return 2;
                }
            }
            Real alpha1 = (lo + hi) * R_(0.5);
            Real alpha2 = alpha1 + k;
            for (i = 0; i < 3; i++) sphere1[i] = pos1[i] + alpha1 * axis1[i];
            for (i = 0; i < 3; i++) sphere2[i] = pos2[i] + alpha2 * axis2[i];
            return orc_collide_spheres(sphere1, o1.p[0], sphere2, o2.p[0], contact);
        }
    }
    Real a1[3], a2[3], b1[3], b2[3];
    for (i = 0; i < 3; i++) {
        a1[i] = pos1[i] + axis1[i] * lz1; a2[i] = pos1[i] - axis1[i] * lz1;
        b1[i] = pos2[i] + axis2[i] * lz2; b2[i] = pos2[i] - axis2[i] * lz2;
    }
    orc_closest_segment_points(a1, a2, b1, b2, sphere1, sphere2);
    return orc_collide_spheres(sphere1, o1.p[0], sphere2, o2.p[0], contact);
}

// dCollideCapsulePlane capsule.cpp:356-415
static inline int orc_capsule_plane(const OrcGeom &o1, const OrcGeom &o2, int flags, OrcContactGeom *contact)
{
    const Real *R = o1.R, *pl = o2.p;
    Real radius = o1.p[0], lz = o1.p[1];
    Real sign = (dot3s(pl, 1, R + 2, 4) > 0) ? R_(-1.0) : R_(1.0);
    Real p[3];
    p[0] = o1.pos[0] + R[2] * lz * R_(0.5) * sign;
    p[1] = o1.pos[1] + R[6] * lz * R_(0.5) * sign;
    p[2] = o1.pos[2] + R[10] * lz * R_(0.5) * sign;
    Real k = dot3(p, pl);
    Real depth = pl[3] - k + radius;
    if (depth < 0) return 0;
    contact[0].normal[0] = pl[0]; contact[0].normal[1] = pl[1]; contact[0].normal[2] = pl[2];
    contact[0].pos[0] = p[0] - pl[0] * radius; contact[0].pos[1] = p[1] - pl[1] * radius; contact[0].pos[2] = p[2] - pl[2] * radius;
    contact[0].depth = depth;
    int ncontacts = 1;
    if ((flags & ORC_NUMC_MASK) >= 2) {
        p[0] = o1.pos[0] - R[2] * lz * R_(0.5) * sign;
        p[1] = o1.pos[1] - R[6] * lz * R_(0.5) * sign;
        p[2] = o1.pos[2] - R[10] * lz * R_(0.5) * sign;
        k = dot3(p, pl);
        depth = pl[3] - k + radius;
        if (depth >= 0) {
            contact[1].normal[0] = pl[0]; contact[1].normal[1] = pl[1]; contact[1].normal[2] = pl[2];
            contact[1].pos[0] = p[0] - pl[0] * radius; contact[1].pos[1] = p[1] - pl[1] * radius; contact[1].pos[2] = p[2] - pl[2] * radius;
            contact[1].depth = depth;
            ncontacts = 2;
        }
    }
    return ncontacts;
}
#endif
