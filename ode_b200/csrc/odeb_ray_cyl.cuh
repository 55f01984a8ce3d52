// odeb_ray_cyl.cuh -- device-side ray colliders (ray vs sphere / box / capsule / plane / cylinder, ode/src/ray.cpp) and the flat
// cylinder's plane and sphere colliders (collision_cylinder_plane.cpp, collision_cylinder_sphere.cpp): one thread per geom pair, the
// reference's operation order (bit-identical contacts under -fmad=false; none of these routines calls a libm transcendental).
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
#endif
