// odeb_ray_cyl.cuh -- device-side ray colliders (ray vs sphere / box / capsule / plane / cylinder, ode/src/ray.cpp) and the flat
// cylinder's plane, sphere and box colliders (collision_cylinder_plane.cpp, collision_cylinder_sphere.cpp, collision_cylinder_box.cpp):
// one thread per geom pair, the reference's operation order (bit-identical contacts under -fmad=false; none of these routines calls a
// libm transcendental at run time: the cap octagon's normals are constants).
#ifndef ODEB_RAY_CYL_CUH
#define ODEB_RAY_CYL_CUH

// ray_sphere_helper ray.cpp:207-247 (mode 1: the exit point of the sphere)
__device__ int odeb_ray_sphere_helper(const DGeom &ray, const Real *sphere_pos, Real radius, DContactGeom *c, int mode)
{
    Real q[3] = { ray.pos[0] - sphere_pos[0], ray.pos[1] - sphere_pos[1], ray.pos[2] - sphere_pos[2] };
    Real B = dot3s(q, 1, ray.R + 2, 4);
    Real C = dot3(q, q) - radius * radius;
    Real k = B * B - C;
    if (k < 0) return 0;
    k = RSQRT(k);
    Real alpha;
    if (mode && C >= 0) {
        alpha = -B + k;
        if (alpha < 0) return 0;
    } else {
        alpha = -B - k;
        if (alpha < 0) {
            alpha = -B + k;
            if (alpha < 0) return 0;
        }
    }
    if (alpha > ray.p[0]) return 0;
    c->pos[0] = ray.pos[0] + alpha * ray.R[2]; c->pos[1] = ray.pos[1] + alpha * ray.R[6]; c->pos[2] = ray.pos[2] + alpha * ray.R[10];
    Real nsign = (C < 0 || mode) ? R_(-1.0) : R_(1.0);
    c->normal[0] = nsign * (c->pos[0] - sphere_pos[0]); c->normal[1] = nsign * (c->pos[1] - sphere_pos[1]); c->normal[2] = nsign * (c->pos[2] - sphere_pos[2]);
    normalize3(c->normal);
    c->depth = alpha;
    return 1;
}

// dCollideRaySphere ray.cpp:250-266
__device__ int odeb_ray_sphere(const DGeom &ray, const DGeom &s, DContactGeom *c) { return odeb_ray_sphere_helper(ray, s.pos, s.p[0], c, 0); }

// dCollideRayBox ray.cpp:269-383
__device__ int odeb_ray_box(const DGeom &ray, const DGeom &box, DContactGeom *c)
{
    Real tmp[3], s[3], v[3], sign[3];
    tmp[0] = ray.pos[0] - box.pos[0]; tmp[1] = ray.pos[1] - box.pos[1]; tmp[2] = ray.pos[2] - box.pos[2];
    mul1_331(s, box.R, tmp);
    tmp[0] = ray.R[2]; tmp[1] = ray.R[6]; tmp[2] = ray.R[10];
    mul1_331(v, box.R, tmp);
    for (int i = 0; i < 3; i++) {
        if (v[i] < 0) { s[i] = -s[i]; v[i] = -v[i]; sign[i] = 1; }
        else sign[i] = -1;
    }
    Real h[3] = { R_(0.5) * box.p[0], R_(0.5) * box.p[1], R_(0.5) * box.p[2] };
    if ((s[0] < -h[0] && v[0] <= 0) || s[0] > h[0] || (s[1] < -h[1] && v[1] <= 0) || s[1] > h[1] || (s[2] < -h[2] && v[2] <= 0) || s[2] > h[2]
        || (v[0] == 0 && v[1] == 0 && v[2] == 0)) return 0;
    Real lo = -R_INF, hi = R_INF;
    int nlo = 0, nhi = 0;
    for (int i = 0; i < 3; i++) {
        if (v[i] != 0) {
            Real k = (-h[i] - s[i]) / v[i];
            if (k > lo) { lo = k; nlo = i; }
            k = (h[i] - s[i]) / v[i];
            if (k < hi) { hi = k; nhi = i; }
        }
    }
    if (lo > hi) return 0;
    Real alpha; int n;
    if (lo >= 0) { alpha = lo; n = nlo; } else { alpha = hi; n = nhi; }
    if (alpha < 0 || alpha > ray.p[0]) return 0;
    c->pos[0] = ray.pos[0] + alpha * ray.R[2]; c->pos[1] = ray.pos[1] + alpha * ray.R[6]; c->pos[2] = ray.pos[2] + alpha * ray.R[10];
    c->normal[0] = box.R[0 + n] * sign[n]; c->normal[1] = box.R[4 + n] * sign[n]; c->normal[2] = box.R[8 + n] * sign[n];
    c->depth = alpha;
    return 1;
}

// dCollideRayCapsule ray.cpp:386-507
__device__ int odeb_ray_capsule(const DGeom &ray, const DGeom &cc, DContactGeom *c)
{
    const Real radius = cc.p[0], lz2 = cc.p[1] * R_(0.5);
    Real cs[3], q[3], r[3], C, k;
    cs[0] = ray.pos[0] - cc.pos[0]; cs[1] = ray.pos[1] - cc.pos[1]; cs[2] = ray.pos[2] - cc.pos[2];
    k = dot3s(cc.R + 2, 4, cs, 1);
    q[0] = k * cc.R[2] - cs[0]; q[1] = k * cc.R[6] - cs[1]; q[2] = k * cc.R[10] - cs[2];
    C = dot3(q, q) - radius * radius;
    int inside = 0;
    if (C < 0) {
        if (k < -lz2) k = -lz2; else if (k > lz2) k = lz2;
        r[0] = cc.pos[0] + k * cc.R[2]; r[1] = cc.pos[1] + k * cc.R[6]; r[2] = cc.pos[2] + k * cc.R[10];
        if ((ray.pos[0] - r[0]) * (ray.pos[0] - r[0]) + (ray.pos[1] - r[1]) * (ray.pos[1] - r[1]) + (ray.pos[2] - r[2]) * (ray.pos[2] - r[2]) < radius * radius) inside = 1;
    }
    if (!inside && C < 0) {
        if (k < 0) k = -lz2; else k = lz2;
    } else {
        Real uv = dot3s(cc.R + 2, 4, ray.R + 2, 4);
        r[0] = uv * cc.R[2] - ray.R[2]; r[1] = uv * cc.R[6] - ray.R[6]; r[2] = uv * cc.R[10] - ray.R[10];
        Real A = dot3(r, r);
        if (A == 0) {
            if (uv < 0) k = -lz2; else k = lz2;
        } else {
            Real B = 2 * dot3(q, r);
            k = B * B - 4 * A * C;
            if (k < 0) {
                if (!inside) return 0;
                if (uv < 0) k = -lz2; else k = lz2;
            } else {
                k = RSQRT(k);
                A = rrecip(2 * A);
                Real alpha = (-B - k) * A;
                if (alpha < 0) {
                    alpha = (-B + k) * A;
                    if (alpha < 0) return 0;
                }
                if (alpha > ray.p[0]) return 0;
                c->pos[0] = ray.pos[0] + alpha * ray.R[2]; c->pos[1] = ray.pos[1] + alpha * ray.R[6]; c->pos[2] = ray.pos[2] + alpha * ray.R[10];
                q[0] = c->pos[0] - cc.pos[0]; q[1] = c->pos[1] - cc.pos[1]; q[2] = c->pos[2] - cc.pos[2];
                k = dot3s(q, 1, cc.R + 2, 4);
                Real nsign = inside ? R_(-1.0) : R_(1.0);
                if (k >= -lz2 && k <= lz2) {
                    c->normal[0] = nsign * (c->pos[0] - (cc.pos[0] + k * cc.R[2]));
                    c->normal[1] = nsign * (c->pos[1] - (cc.pos[1] + k * cc.R[6]));
                    c->normal[2] = nsign * (c->pos[2] - (cc.pos[2] + k * cc.R[10]));
                    normalize3(c->normal);
                    c->depth = alpha;
                    return 1;
                }
                if (k < 0) k = -lz2; else k = lz2;
            }
        }
    }
    q[0] = cc.pos[0] + k * cc.R[2]; q[1] = cc.pos[1] + k * cc.R[6]; q[2] = cc.pos[2] + k * cc.R[10];
    return odeb_ray_sphere_helper(ray, q, radius, c, inside);
}

// dCollideRayPlane ray.cpp:510-543
__device__ int odeb_ray_plane(const DGeom &ray, const DGeom &pl, DContactGeom *c)
{
    Real alpha = pl.p[3] - dot3(pl.p, ray.pos);
    Real nsign = (alpha > 0) ? R_(-1.0) : R_(1.0);
    Real k = dot3s(pl.p, 1, ray.R + 2, 4);
    if (k == 0) return 0;
    alpha /= k;
    if (alpha < 0 || alpha > ray.p[0]) return 0;
    c->pos[0] = ray.pos[0] + alpha * ray.R[2]; c->pos[1] = ray.pos[1] + alpha * ray.R[6]; c->pos[2] = ray.pos[2] + alpha * ray.R[10];
    c->normal[0] = nsign * pl.p[0]; c->normal[1] = nsign * pl.p[1]; c->normal[2] = nsign * pl.p[2];
    c->depth = alpha;
    return 1;
}

// dCollideRayCylinder ray.cpp:538-735 (Joseph Cooper's case analysis: caps first, then the lateral surface, in the cylinder's frame)
__device__ int odeb_ray_cylinder(const DGeom &ray, const DGeom &cyl, DContactGeom *c)
{
    const Real half_length = cyl.p[1] * R_(0.5), radius = cyl.p[0];
    Real tmp[3], pos[3], dir[3];
    tmp[0] = ray.pos[0] - cyl.pos[0]; tmp[1] = ray.pos[1] - cyl.pos[1]; tmp[2] = ray.pos[2] - cyl.pos[2];
    mul1_331(pos, cyl.R, tmp);
    tmp[0] = ray.R[2]; tmp[1] = ray.R[6]; tmp[2] = ray.R[10];
    mul1_331(dir, cyl.R, tmp);
    const Real r2 = radius * radius;
    const Real C = pos[0] * pos[0] + pos[1] * pos[1] - r2;
    const int parallel = (dir[0] == 0 && dir[1] == 0), perpendicular = (dir[2] == 0);
    const int inRadius = (C <= 0), inCaps = (RFABS(pos[2]) <= half_length);
    const int checkCaps = (!perpendicular && (!inCaps || inRadius));
    int checkCyl = (!parallel && (!inRadius || inCaps));
    const int flipNormals = (inCaps && inRadius);
    Real tt = -R_INF, nrm[3] = { 0, 0, 0 };
    if (checkCaps) {
        int flipDir = 0;
        if ((dir[2] < 0 && flipNormals) || (dir[2] > 0 && !flipNormals)) { flipDir = 1; dir[2] = -dir[2]; pos[2] = -pos[2]; }
        tt = (half_length - pos[2]) / dir[2];
        if (tt >= 0 && tt <= ray.p[0]) {
            tmp[0] = pos[0] + tt * dir[0];
            tmp[1] = pos[1] + tt * dir[1];
            if (tmp[0] * tmp[0] + tmp[1] * tmp[1] <= r2) {
                tmp[2] = flipDir ? -half_length : half_length;
                nrm[0] = 0; nrm[1] = 0; nrm[2] = (flipDir != flipNormals) ? -R_(1.0) : R_(1.0);
                checkCyl = 0;
            } else tt = -R_INF;
        } else tt = -R_INF;
        if (flipDir) { dir[2] = -dir[2]; pos[2] = -pos[2]; }
    }
    if (checkCyl) {
        Real A = dir[0] * dir[0] + dir[1] * dir[1];
        Real B = 2 * (pos[0] * dir[0] + pos[1] * dir[1]);
        Real k = B * B - 4 * A * C;
        if (k >= 0 && (B < 0 || B * B <= k)) {
            k = RSQRT(k);
            A = rrecip(2 * A);
            if (RFABS(B) <= k) tt = (-B + k) * A; else tt = (-B - k) * A;
            if (tt <= ray.p[0]) {
                tmp[2] = pos[2] + tt * dir[2];
                if (RFABS(tmp[2]) <= half_length) {
                    tmp[0] = pos[0] + tt * dir[0];
                    tmp[1] = pos[1] + tt * dir[1];
                    nrm[0] = tmp[0] / radius; nrm[1] = tmp[1] / radius; nrm[2] = 0;
                    if (flipNormals) { nrm[0] = -nrm[0]; nrm[1] = -nrm[1]; }
                } else tt = -R_INF;
            } else tt = -R_INF;
        }
    }
    if (tt > 0) {
        c->depth = tt;
        mul0_331(c->normal, cyl.R, nrm);
        mul0_331(c->pos, cyl.R, tmp);
        c->pos[0] += cyl.pos[0]; c->pos[1] += cyl.pos[1]; c->pos[2] += cyl.pos[2];
        return 1;
    }
    return 0;
}

// dCollideCylinderPlane collision_cylinder_plane.cpp:43-264.  Like the reference, a candidate point is written to the next contact slot
// before its depth is tested, so c[] needs one slot more than maxc.
__device__ int odeb_cylinder_plane(const DGeom &cy, const DGeom &pl, int flags, DContactGeom *c)
{
    const int maxc = flags & ODEB_NUMC_MASK;
    int n = 0;
#ifdef ODEB_DOUBLE
    const Real toleranz = R_(0.0000001);
#else
    const Real toleranz = R_(0.0001);
#endif
    const Real radius = cy.p[0], length = cy.p[1];
    const Real *pv = pl.p;
    Real P1[3], P2[3], d[3] = { cy.R[2], cy.R[6], cy.R[10] };
    Real s = length * R_(0.5);
    P2[0] = d[0] * s + cy.pos[0]; P2[1] = d[1] * s + cy.pos[1]; P2[2] = d[2] * s + cy.pos[2];
    P1[0] = d[0] * -s + cy.pos[0]; P1[1] = d[1] * -s + cy.pos[1]; P1[2] = d[2] * -s + cy.pos[2];
    s = d[0] * pv[0] + d[1] * pv[1] + d[2] * pv[2];
    if (s < 0) s += R_(1.0); else s -= R_(1.0);
    if (s < toleranz && s > (-toleranz)) {
        Real P[3], t;
        s = pv[3] - dot3(pv, P1);
        t = pv[3] - dot3(pv, P2);
        if (s >= t) { if (s >= 0) { P[0] = P1[0]; P[1] = P1[1]; P[2] = P1[2]; } else return n; }
        else { if (t >= 0) { P[0] = P2[0]; P[1] = P2[1]; P[2] = P2[2]; } else return n; }
        Real V1[3], V2[3];
        if (d[0] < toleranz && d[0] > (-toleranz)) { V1[0] = d[0] + R_(1.0); V1[1] = d[1]; V1[2] = d[2]; }
        else { V1[0] = d[0]; V1[1] = d[1] + R_(1.0); V1[2] = d[2]; }
        cross3(V2, V1, d);
        t = RSQRT(V2[0] * V2[0] + V2[1] * V2[1] + V2[2] * V2[2]);
        t = radius / t;
        V2[0] *= t; V2[1] *= t; V2[2] *= t;
        cross3(V1, V2, d);
        for (int k = 0; k < 4; k++) {
            const Real *V = (k < 2) ? V1 : V2;
            if ((k & 1) == 0) { c[n].pos[0] = P[0] + V[0]; c[n].pos[1] = P[1] + V[1]; c[n].pos[2] = P[2] + V[2]; }
            else { c[n].pos[0] = P[0] - V[0]; c[n].pos[1] = P[1] - V[1]; c[n].pos[2] = P[2] - V[2]; }
            c[n].depth = pv[3] - dot3(pv, c[n].pos);
            if (c[n].depth > 0) {
                c[n].normal[0] = pv[0]; c[n].normal[1] = pv[1]; c[n].normal[2] = pv[2];
                n++;
                if (n >= maxc) return n;
            }
        }
    } else {
        Real C[3], t = dot3(pv, d);
        C[0] = d[0] * t - pv[0]; C[1] = d[1] * t - pv[1]; C[2] = d[2] * t - pv[2];
        s = RSQRT(C[0] * C[0] + C[1] * C[1] + C[2] * C[2]);
        s = radius / s;
        C[0] *= s; C[1] *= s; C[2] *= s;
        c[n].pos[0] = C[0] + P1[0]; c[n].pos[1] = C[1] + P1[1]; c[n].pos[2] = C[2] + P1[2];
        c[n].depth = pv[3] - dot3(pv, c[n].pos);
        if (c[n].depth >= 0) {
            c[n].normal[0] = pv[0]; c[n].normal[1] = pv[1]; c[n].normal[2] = pv[2];
            n++;
            if (n >= maxc) return n;
        }
        c[n].pos[0] = C[0] + P2[0]; c[n].pos[1] = C[1] + P2[1]; c[n].pos[2] = C[2] + P2[2];
        c[n].depth = pv[3] - pv[0] * c[n].pos[0] - pv[1] * c[n].pos[1] - pv[2] * c[n].pos[2];
        if (c[n].depth >= 0) {
            c[n].normal[0] = pv[0]; c[n].normal[1] = pv[1]; c[n].normal[2] = pv[2];
            n++;
            if (n >= maxc) return n;
        }
    }
    return n;
}

// dCollideCylinderSphere collision_cylinder_sphere.cpp:53-275
__device__ int odeb_cylinder_sphere(const DGeom &cy, const DGeom &sp, DContactGeom *c)
{
#ifdef ODEB_DOUBLE
    const Real toleranz = R_(0.0000001);
#else
    const Real toleranz = R_(0.0001);
#endif
    const Real radius = cy.p[0], length = cy.p[1], radius2 = sp.p[0];
    const Real *S = sp.pos;
    Real P1[3], P2[3], d[3] = { cy.R[2], cy.R[6], cy.R[10] }, C[3], t;
    Real s = length * R_(0.5);
    P2[0] = d[0] * s + cy.pos[0]; P2[1] = d[1] * s + cy.pos[1]; P2[2] = d[2] * s + cy.pos[2];
    P1[0] = d[0] * -s + cy.pos[0]; P1[1] = d[1] * -s + cy.pos[1]; P1[2] = d[2] * -s + cy.pos[2];
    s = (S[0] - P1[0]) * d[0] - (P1[1] - S[1]) * d[1] - (P1[2] - S[2]) * d[2];
    if (s < (-radius2) || s > (length + radius2)) return 0;
    C[0] = s * d[0] + P1[0] - S[0]; C[1] = s * d[1] + P1[1] - S[1]; C[2] = s * d[2] + P1[2] - S[2];
    t = RSQRT(C[0] * C[0] + C[1] * C[1] + C[2] * C[2]);
    if (t > (radius + radius2)) return 0;
    if (t > radius && (s < 0 || s > length)) {
        const Real *Pe = (s <= 0) ? P1 : P2;
        if (s <= 0) c->depth = radius2 - RSQRT((s) * (s) + (t - radius) * (t - radius));
        else c->depth = radius2 - RSQRT((s - length) * (s - length) + (t - radius) * (t - radius));
        if (c->depth < 0) return 0;
        c->pos[0] = C[0] / t * -radius + Pe[0]; c->pos[1] = C[1] / t * -radius + Pe[1]; c->pos[2] = C[2] / t * -radius + Pe[2];
        c->normal[0] = (c->pos[0] - S[0]) / (radius2 - c->depth);
        c->normal[1] = (c->pos[1] - S[1]) / (radius2 - c->depth);
        c->normal[2] = (c->pos[2] - S[2]) / (radius2 - c->depth);
        return 1;
    } else if ((radius - t) <= s && (radius - t) <= (length - s)) {
        c->depth = (radius2 + radius) - t;
        if (c->depth < 0) return 0;
        if (t > (radius2 + toleranz)) {
            C[0] /= t; C[1] /= t; C[2] /= t;
            c->pos[0] = C[0] * radius2 + S[0]; c->pos[1] = C[1] * radius2 + S[1]; c->pos[2] = C[2] * radius2 + S[2];
            c->normal[0] = C[0]; c->normal[1] = C[1]; c->normal[2] = C[2];
        } else {
            c->pos[0] = C[0] + S[0]; c->pos[1] = C[1] + S[1]; c->pos[2] = C[2] + S[2];
            c->normal[0] = C[0] / t; c->normal[1] = C[1] / t; c->normal[2] = C[2] / t;
        }
        return 1;
    } else {
        if (s <= (length * R_(0.5))) {
            c->depth = s + radius2;
            if (c->depth < 0) return 0;
            c->pos[0] = radius2 * d[0] + S[0]; c->pos[1] = radius2 * d[1] + S[1]; c->pos[2] = radius2 * d[2] + S[2];
            c->normal[0] = d[0]; c->normal[1] = d[1]; c->normal[2] = d[2];
        } else {
            c->depth = (radius2 + length - s);
            if (c->depth < 0) return 0;
            c->pos[0] = radius2 * -d[0] + S[0]; c->pos[1] = radius2 * -d[1] + S[1]; c->pos[2] = radius2 * -d[2] + S[2];
            c->normal[0] = -d[0]; c->normal[1] = -d[1]; c->normal[2] = -d[2];
        }
        return 1;
    }
}
// ---- dCollideCylinderBox collision_cylinder_box.cpp:1026-1038 (Croteam's collider: 40 candidate separating axes, then either the
// cylinder's side line clipped to the box or the nearest box face clipped to an octagon inscribed in the cylinder's cap)
struct DCylBox {
    Real cylR[12], cylPos[3], cylAxis[3], radius, size;
    Real boxR[12], boxPos[3], half[3], vert[8][3];
    Real diff[3], normal[3], bestDepth, bestrb, bestrc;
    int bestAxis;
};
// the cap octagon's outward normals: -cos / -sin of pi/8 + k*pi/4 accumulated in dReal (collision_cylinder_box.cpp:183-197), as evaluated
// by glibc (values generated with the reference's own expressions; tests/test_oracle.py pins them against the compiled reference)
#ifdef ODEB_DOUBLE
#define ODEB_CYLN { { -0x1.d906bcf328d46p-1, -0x1.87de2a6aea963p-2 }, { -0x1.87de2a6aea964p-2, -0x1.d906bcf328d46p-1 }, { 0x1.87de2a6aea962p-2, -0x1.d906bcf328d46p-1 }, \
    { 0x1.d906bcf328d46p-1, -0x1.87de2a6aea965p-2 }, { 0x1.d906bcf328d47p-1, 0x1.87de2a6aea961p-2 }, { 0x1.87de2a6aea96dp-2, 0x1.d906bcf328d44p-1 }, \
    { -0x1.87de2a6aea958p-2, 0x1.d906bcf328d48p-1 }, { -0x1.d906bcf328d44p-1, 0x1.87de2a6aea96ep-2 } }
#else
#define ODEB_CYLN { { -0x1.d906bcp-1f, -0x1.87de2cp-2f }, { -0x1.87de2ap-2f, -0x1.d906bcp-1f }, { 0x1.87de3p-2f, -0x1.d906bcp-1f }, { 0x1.d906cp-1f, -0x1.87de2p-2f }, \
    { 0x1.d906bap-1f, 0x1.87de3ap-2f }, { 0x1.87de16p-2f, 0x1.d906c2p-1f }, { -0x1.87de36p-2f, 0x1.d906bap-1f }, { -0x1.d906bep-1f, 0x1.87de2ap-2f } }
#endif

// _cldTestAxis :206-286
__device__ int odeb_cb_test_axis(DCylBox &d, Real *n, int iAxis)
{
    Real fL = RSQRT(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    if (fL < R_(1e-5)) return 1;
    normalize3(n);
    Real fdot1 = dot3(d.cylAxis, n), frc;
    if (fdot1 > R_(1.0)) frc = d.size * R_(0.5);
    else if (fdot1 < R_(-1.0)) frc = d.size * R_(0.5);
    else frc = RFABS(fdot1 * (d.size * R_(0.5))) + d.radius * RSQRT(R_(1.0) - (fdot1 * fdot1));
    Real t[3];
    t[0] = d.boxR[0]; t[1] = d.boxR[4]; t[2] = d.boxR[8];
    Real frb = RFABS(dot3(t, n)) * d.half[0];
    t[0] = d.boxR[1]; t[1] = d.boxR[5]; t[2] = d.boxR[9];
    frb += RFABS(dot3(t, n)) * d.half[1];
    t[0] = d.boxR[2]; t[1] = d.boxR[6]; t[2] = d.boxR[10];
    frb += RFABS(dot3(t, n)) * d.half[2];
    Real fd = dot3(d.diff, n);
    Real fDepth = frc + frb;
    if (RFABS(fd) > fDepth) return 0;
    fDepth -= RFABS(fd);
    if (fDepth < d.bestDepth) {
        d.bestDepth = fDepth;
        d.normal[0] = n[0]; d.normal[1] = n[1]; d.normal[2] = n[2];
        d.bestAxis = iAxis; d.bestrb = frb; d.bestrc = frc;
        if (fd > 0) { d.normal[0] = -d.normal[0]; d.normal[1] = -d.normal[1]; d.normal[2] = -d.normal[2]; }
    }
    return 1;
}
// _cldTestEdgeCircleAxis :289-333
__device__ int odeb_cb_test_edge_circle(DCylBox &d, const Real *cc, const Real *v0, const Real *v1, int iAxis)
{
    Real e[3] = { v1[0] - v0[0], v1[1] - v0[1], v1[2] - v0[2] };
    normalize3(e);
    Real fdot2 = dot3(e, d.cylAxis);
    if (RFABS(fdot2) < R_(1e-5)) return 1;
    Real t1[3] = { cc[0] - v0[0], cc[1] - v0[1], cc[2] - v0[2] };
    Real fdot1 = dot3(t1, d.cylAxis);
    Real pnt[3] = { v0[0] + e[0] * (fdot1 / fdot2), v0[1] + e[1] * (fdot1 / fdot2), v0[2] + e[2] * (fdot1 / fdot2) };
    Real tangent[3], axis[3];
    t1[0] = cc[0] - pnt[0]; t1[1] = cc[1] - pnt[1]; t1[2] = cc[2] - pnt[2];
    cross3(tangent, t1, d.cylAxis);
    cross3(axis, tangent, e);
    return odeb_cb_test_axis(d, axis, iAxis);
}
// dClipEdgeToPlane collision_util.cpp:471-509
__device__ int odeb_clip_edge_to_plane(Real *p0, Real *p1, const Real *pl)
{
    Real d0 = pl[0] * p0[0] + pl[1] * p0[1] + pl[2] * p0[2] + pl[3], d1 = pl[0] * p1[0] + pl[1] * p1[1] + pl[2] * p1[2] + pl[3];
    if (d0 < 0 && d1 < 0) return 0;
    else if (d0 > 0 && d1 > 0) return 1;
    else if ((d0 > 0 && d1 < 0) || (d0 < 0 && d1 > 0)) {
        Real ip[3];
        ip[0] = p0[0] - (p0[0] - p1[0]) * d0 / (d0 - d1);
        ip[1] = p0[1] - (p0[1] - p1[1]) * d0 / (d0 - d1);
        ip[2] = p0[2] - (p0[2] - p1[2]) * d0 / (d0 - d1);
        if (d0 < 0) { p0[0] = ip[0]; p0[1] = ip[1]; p0[2] = ip[2]; } else { p1[0] = ip[0]; p1[1] = ip[1]; p1[2] = ip[2]; }
        return 1;
    }
    return 1;
}
// dClipPolyToPlane collision_util.cpp:512-557
__device__ void odeb_clip_poly_to_plane(const Real (*in)[3], int ctIn, Real (*out)[3], int *ctOut, const Real *pl)
{
    int n = 0, i0 = ctIn - 1;
    for (int i1 = 0; i1 < ctIn; i0 = i1, i1++) {
        Real d0 = pl[0] * in[i0][0] + pl[1] * in[i0][1] + pl[2] * in[i0][2] + pl[3];
        Real d1 = pl[0] * in[i1][0] + pl[1] * in[i1][1] + pl[2] * in[i1][2] + pl[3];
        if (d0 >= 0) { out[n][0] = in[i0][0]; out[n][1] = in[i0][1]; out[n][2] = in[i0][2]; n++; }
        if ((d0 > 0 && d1 < 0) || (d0 < 0 && d1 > 0)) {
            out[n][0] = in[i0][0] - (in[i0][0] - in[i1][0]) * d0 / (d0 - d1);
            out[n][1] = in[i0][1] - (in[i0][1] - in[i1][1]) * d0 / (d0 - d1);
            out[n][2] = in[i0][2] - (in[i0][2] - in[i1][2]) * d0 / (d0 - d1);
            n++;
        }
    }
    *ctOut = n;
}
// dMatrix3Inv collision_util.h:199-235 (the adjugate is scaled by a double reciprocal of the dReal determinant)
__device__ void odeb_matrix3_inv(const Real *ma, Real *dst)
{
    Real det = ma[0] * (ma[5] * ma[10] - ma[9] * ma[6]) - ma[1] * (ma[4] * ma[10] - ma[8] * ma[6]) + ma[2] * (ma[4] * ma[9] - ma[8] * ma[5]);
    if (RFABS(det) < R_(0.0005)) { for (int k = 0; k < 12; k++) dst[k] = 0; dst[0] = dst[5] = dst[10] = 1; return; }
    const double dr = (double)(R_(1.0) / det);
    dst[0] = (Real)((ma[5] * ma[10] - ma[6] * ma[9]) * dr); dst[1] = (Real)((ma[9] * ma[2] - ma[1] * ma[10]) * dr); dst[2] = (Real)((ma[1] * ma[6] - ma[5] * ma[2]) * dr);
    dst[4] = (Real)((ma[6] * ma[8] - ma[4] * ma[10]) * dr); dst[5] = (Real)((ma[0] * ma[10] - ma[8] * ma[2]) * dr); dst[6] = (Real)((ma[4] * ma[2] - ma[0] * ma[6]) * dr);
    dst[8] = (Real)((ma[4] * ma[9] - ma[8] * ma[5]) * dr); dst[9] = (Real)((ma[8] * ma[1] - ma[0] * ma[9]) * dr); dst[10] = (Real)((ma[0] * ma[5] - ma[1] * ma[4]) * dr);
    dst[3] = dst[7] = dst[11] = 0;
}

__device__ __noinline__ int odeb_cylinder_box(const DGeom &cy, const DGeom &bx, int flags, DContactGeom *c)
{
    const int maxc = flags & ODEB_NUMC_MASK;
    DCylBox d;
    // _cldInitCylinderBox :100-203
    for (int k = 0; k < 12; k++) { d.cylR[k] = cy.R[k]; d.boxR[k] = bx.R[k]; }
    for (int k = 0; k < 3; k++) { d.cylPos[k] = cy.pos[k]; d.boxPos[k] = bx.pos[k]; d.half[k] = bx.p[k]; }
    d.cylAxis[0] = d.cylR[2]; d.cylAxis[1] = d.cylR[6]; d.cylAxis[2] = d.cylR[10];
    d.radius = cy.p[0]; d.size = cy.p[1];
    d.half[0] *= R_(0.5); d.half[1] *= R_(0.5); d.half[2] *= R_(0.5);
    {
        const int sx[8] = { -1, 1, -1, 1, 1, 1, -1, -1 }, sy[8] = { 1, 1, -1, -1, 1, -1, -1, 1 }, sz[8] = { -1, -1, -1, -1, 1, 1, 1, 1 };
        for (int i = 0; i < 8; i++) {
            Real v[3] = { sx[i] < 0 ? -d.half[0] : d.half[0], sy[i] < 0 ? -d.half[1] : d.half[1], sz[i] < 0 ? -d.half[2] : d.half[2] }, t[3];
            mul0_331(t, d.boxR, v);
            d.vert[i][0] = t[0] + d.boxPos[0]; d.vert[i][1] = t[1] + d.boxPos[1]; d.vert[i][2] = t[2] + d.boxPos[2];
        }
    }
    d.diff[0] = d.cylPos[0] - d.boxPos[0]; d.diff[1] = d.cylPos[1] - d.boxPos[1]; d.diff[2] = d.cylPos[2] - d.boxPos[2];
    d.bestDepth = R_INF; d.normal[0] = d.normal[1] = d.normal[2] = 0; d.bestrb = 0; d.bestrc = 0; d.bestAxis = 0;
    // _cldTestSeparatingAxes :336-573
    {
        Real ax[3], t1[3], t2[3], col[3];
        const Real eps = R_(1e-6);
        for (int k = 0; k < 3; k++) { ax[0] = d.boxR[k]; ax[1] = d.boxR[4 + k]; ax[2] = d.boxR[8 + k]; if (!odeb_cb_test_axis(d, ax, 1 + k)) return 0; }
        ax[0] = d.cylAxis[0]; ax[1] = d.cylAxis[1]; ax[2] = d.cylAxis[2];
        if (!odeb_cb_test_axis(d, ax, 4)) return 0;
        for (int k = 0; k < 3; k++) {
            col[0] = d.boxR[k]; col[1] = d.boxR[4 + k]; col[2] = d.boxR[8 + k];
            cross3(ax, d.cylAxis, col);
            if (ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2] > eps) if (!odeb_cb_test_axis(d, ax, 5 + k)) return 0;
        }
        for (int i = 0; i < 8; i++) {
            t1[0] = d.vert[i][0] - d.cylPos[0]; t1[1] = d.vert[i][1] - d.cylPos[1]; t1[2] = d.vert[i][2] - d.cylPos[2];
            cross3(t2, d.cylAxis, t1);
            cross3(ax, d.cylAxis, t2);
            if (ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2] > eps) if (!odeb_cb_test_axis(d, ax, 8 + i)) return 0;
        }
        const int e0[12] = { 1, 1, 2, 2, 4, 4, 0, 5, 5, 2, 4, 6 }, e1[12] = { 0, 3, 3, 0, 1, 7, 7, 3, 6, 6, 5, 7 };
        Real cc[3];
        for (int cap = 0; cap < 2; cap++) {
            const Real hs = d.size * R_(0.5);
            if (cap == 0) { cc[0] = d.cylPos[0] + d.cylAxis[0] * hs; cc[1] = d.cylPos[1] + d.cylAxis[1] * hs; cc[2] = d.cylPos[2] + d.cylAxis[2] * hs; }
            else { cc[0] = d.cylPos[0] - d.cylAxis[0] * hs; cc[1] = d.cylPos[1] - d.cylAxis[1] * hs; cc[2] = d.cylPos[2] - d.cylAxis[2] * hs; }
            for (int k = 0; k < 12; k++) if (!odeb_cb_test_edge_circle(d, cc, d.vert[e0[k]], d.vert[e1[k]], 16 + 12 * cap + k)) return 0;
        }
    }
    if (d.bestAxis == 0) return 0;
    int nc = 0;
    const Real fdot = dot3(d.normal, d.cylAxis);
    if (RFABS(fdot) < R_(0.9)) {
        // _cldClipCylinderToBox :576-717
        Real vN[3], ep0[3], ep1[3], ct[3], pl[4], t[3];
        const Real f1 = dot3(d.cylAxis, d.normal), hs = d.size * R_(0.5);
        vN[0] = d.normal[0] - d.cylAxis[0] * f1; vN[1] = d.normal[1] - d.cylAxis[1] * f1; vN[2] = d.normal[2] - d.cylAxis[2] * f1;
        normalize3(vN);
        ct[0] = d.cylPos[0] + vN[0] * d.radius; ct[1] = d.cylPos[1] + vN[1] * d.radius; ct[2] = d.cylPos[2] + vN[2] * d.radius;
        for (int k = 0; k < 3; k++) { ep0[k] = ct[k] + d.cylAxis[k] * hs; ep1[k] = ct[k] - d.cylAxis[k] * hs; }
        for (int k = 0; k < 3; k++) { ep0[k] -= d.boxPos[k]; ep1[k] -= d.boxPos[k]; }
        for (int k = 0; k < 6; k++) {
            const int a = k % 3;
            t[0] = d.boxR[a]; t[1] = d.boxR[4 + a]; t[2] = d.boxR[8 + a];
            if (k >= 3) { t[0] = -t[0]; t[1] = -t[1]; t[2] = -t[2]; }
            pl[0] = t[0]; pl[1] = t[1]; pl[2] = t[2]; pl[3] = d.half[a];
            if (!odeb_clip_edge_to_plane(ep0, ep1, pl)) return 0;
        }
        Real dep0 = d.bestrb + dot3(ep0, d.normal), dep1 = d.bestrb + dot3(ep1, d.normal);
        if (dep0 < 0) dep0 = R_(0.0);
        if (dep1 < 0) dep1 = R_(0.0);
        for (int k = 0; k < 3; k++) { ep0[k] += d.boxPos[k]; ep1[k] += d.boxPos[k]; }
        c[nc].depth = dep0;
        for (int k = 0; k < 3; k++) { c[nc].normal[k] = -d.normal[k]; c[nc].pos[k] = ep0[k]; }
        nc++;
        if (nc != maxc) {
            c[nc].depth = dep1;
            for (int k = 0; k < 3; k++) { c[nc].normal[k] = -d.normal[k]; c[nc].pos[k] = ep1[k]; }
            nc++;
        }
        return nc;
    }
    // _cldClipBoxToCylinder :720-982
    {
        Real ccp[3], cn[3] = { 0, 0, 0 };
        const Real hs = d.size * R_(0.5);
        if (dot3(d.cylAxis, d.normal) > R_(0.0)) { for (int k = 0; k < 3; k++) ccp[k] = d.cylPos[k] + d.cylAxis[k] * hs; cn[2] = R_(-1.0); }
        else { for (int k = 0; k < 3; k++) ccp[k] = d.cylPos[k] - d.cylAxis[k] * hs; cn[2] = R_(1.0); }
        Real vNr[3], inv[12], an[3];
        odeb_matrix3_inv(d.boxR, inv);
        mul0_331(vNr, inv, d.normal);
        an[0] = RFABS(vNr[0]); an[1] = RFABS(vNr[1]); an[2] = RFABS(vNr[2]);
        int iB0, iB1, iB2;
        if (an[1] > an[0]) {
            if (an[0] > an[2]) { iB0 = 1; iB1 = 0; iB2 = 2; }
            else if (an[1] > an[2]) { iB0 = 1; iB1 = 2; iB2 = 0; }
            else { iB0 = 2; iB1 = 1; iB2 = 0; }
        } else {
            if (an[1] > an[2]) { iB0 = 0; iB1 = 1; iB2 = 2; }
            else if (an[0] > an[2]) { iB0 = 0; iB1 = 2; iB2 = 1; }
            else { iB0 = 2; iB1 = 0; iB2 = 1; }
        }
        Real ctr[3], t[3], a1[3], a2[3];
        t[0] = d.boxR[iB0]; t[1] = d.boxR[4 + iB0]; t[2] = d.boxR[8 + iB0];
        if (vNr[iB0] > 0) { for (int k = 0; k < 3; k++) ctr[k] = d.boxPos[k] - d.half[iB0] * t[k]; }
        else { for (int k = 0; k < 3; k++) ctr[k] = d.boxPos[k] + d.half[iB0] * t[k]; }
        Real pts[4][3], A1[16][3], A2[16][3];
        for (int i = 0; i < 16; i++) for (int k = 0; k < 3; k++) { A1[i][k] = R_(0.0); A2[i][k] = R_(0.0); }
        a1[0] = d.boxR[iB1]; a1[1] = d.boxR[4 + iB1]; a1[2] = d.boxR[8 + iB1];
        a2[0] = d.boxR[iB2]; a2[1] = d.boxR[4 + iB2]; a2[2] = d.boxR[8 + iB2];
        for (int k = 0; k < 3; k++) {
            pts[0][k] = ctr[k] + d.half[iB1] * a1[k] - d.half[iB2] * a2[k];
            pts[1][k] = ctr[k] - d.half[iB1] * a1[k] - d.half[iB2] * a2[k];
            pts[2][k] = ctr[k] - d.half[iB1] * a1[k] + d.half[iB2] * a2[k];
            pts[3][k] = ctr[k] + d.half[iB1] * a1[k] + d.half[iB2] * a2[k];
        }
        Real cinv[12];
        odeb_matrix3_inv(d.cylR, cinv);
        for (int i = 0; i < 4; i++) {
            t[0] = pts[i][0] - ccp[0]; t[1] = pts[i][1] - ccp[1]; t[2] = pts[i][2] - ccp[2];
            mul0_331(pts[i], cinv, t);
        }
        int n1 = 0, n2 = 0;
        Real pl[4] = { cn[0], cn[1], cn[2], R_(0.0) };
        odeb_clip_poly_to_plane(pts, 4, A1, &n1, pl);
        const Real cyln[8][2] = ODEB_CYLN;
        int seg;
        for (seg = 0; seg < 8; seg++) {
            pl[0] = cyln[seg][0]; pl[1] = cyln[seg][1]; pl[2] = 0; pl[3] = d.radius;
            if ((seg % 2) == 0) odeb_clip_poly_to_plane(A1, n1, A2, &n2, pl);
            else odeb_clip_poly_to_plane(A2, n2, A1, &n1, pl);
        }
        const Real (*fin)[3] = (seg % 2) ? A2 : A1;
        const int nf = (seg % 2) ? n2 : n1;
        for (int i = 0; i < nf; i++) {
            Real pt[3];
            mul0_331(pt, d.cylR, fin[i]);
            pt[0] += ccp[0]; pt[1] += ccp[1]; pt[2] += ccp[2];
            t[0] = pt[0] - d.cylPos[0]; t[1] = pt[1] - d.cylPos[1]; t[2] = pt[2] - d.cylPos[2];
            const Real depth = d.bestrc - dot3(t, d.normal);
            if (depth > R_(0.0)) {
                c[nc].depth = depth;
                for (int k = 0; k < 3; k++) { c[nc].normal[k] = -d.normal[k]; c[nc].pos[k] = pt[k]; }
                nc++;
                if (nc == maxc) break;
            }
        }
    }
    return nc;
}
#endif
