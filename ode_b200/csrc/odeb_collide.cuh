// odeb_collide.cuh -- device-side AABBs and primitive colliders (sphere / box / capsule / plane) of the
// B200 step path: one thread evaluates one geom pair. Operation order follows the reference functions
// cited at each routine (ode/src/box.cpp, sphere.cpp, capsule.cpp, plane.cpp, collision_util.cpp,
// collision_kernel.cpp) so contacts are bit-identical under -fmad=false; the one libm call on the path, atan2 in the
// contact culling, is the host libm's algorithm in single precision (odeb_math.cuh) and CUDA's libm (last-ulp differences) in double.
#ifndef ODEB_COLLIDE_CUH
#define ODEB_COLLIDE_CUH
#include "odeb_math.cuh"


struct DGeom {
    int type;        // ODEB_SPHERE / BOX / CAPSULE / CYLINDER / PLANE / RAY
    int body;        // -1 = none
    Real p[4];       // radius | sides | radius,length | plane a,b,c,d (normalised)
    Real pos[3];
    Real R[12];
};

struct DContactGeom { Real pos[3], normal[3], depth; };

#define ODEB_NUMC_MASK 0xffff
#define ODEB_CONTACTS_UNIMPORTANT 0x80000000

// dxSphere::computeAABB sphere.cpp:59-67; dxBox::computeAABB box.cpp:60-77;
// dxCapsule::computeAABB capsule.cpp:60-74; dxPlane::computeAABB plane.cpp:80-107
__device__ __forceinline__ void odeb_compute_aabb(const DGeom &g, Real *a)
{
    const Real *pos = g.pos, *R = g.R;
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

// dxHashSpace::collide collision_space.cpp:421-614 beyond the AABB test: does the cell walk bring two overlapping AABBs together?
// level = frexp exponent of the largest extent (findLevel :329-349) clamped up to minlevel; cell bounds = floor(aabb / 2^level) (:448-462);
// levels above maxlevel (infinite AABBs: MAXINT) sit in the big list and are tested against everything (:590-607).  The lower-level AABB
// probes the higher level with bounds >>= 1 (:582).  The hash index of a cell column starts at (level*1000UL + x*100UL + y*10UL + zbegin) % sz
// (:358-361, :499, :533; level/x/y through unsigned int, zbegin a plain int) and is incremented per z: a negative zbegin larger than the
// base wraps the sum around 2^64, which shifts that column's slots by 2^64 % sz (sz is an odd prime), so the two sides only see each
// other in columns they address with the same wrap state.
__device__ __forceinline__ int odeb_hash_level(const Real *a, int minlevel)
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
__device__ __forceinline__ bool odeb_hash_column_wraps(int level, int x, int y, int zbegin)
{
    unsigned long long base = (unsigned long long)(unsigned)level * 1000ULL + (unsigned long long)(unsigned)x * 100ULL + (unsigned long long)(unsigned)y * 10ULL;
    return zbegin < 0 && base < (unsigned long long)(-(long long)zbegin);
}
__device__ __noinline__ bool odeb_hash_space_meets(const Real *a, const Real *b, int minlevel, int maxlevel)
{
    int la = odeb_hash_level(a, minlevel), lb = odeb_hash_level(b, minlevel);
    if (la > maxlevel || lb > maxlevel) return true;
    if (la > lb) { const Real *t = a; a = b; b = t; int tl = la; la = lb; lb = tl; }
    int da[6], db[6];
    const Real ra = ldexp((Real)1, -la), rb = ldexp((Real)1, -lb);
    for (int i = 0; i < 6; i++) { da[i] = (int)floor(a[i] * ra) >> (lb - la); db[i] = (int)floor(b[i] * rb); }
    if (da[4] >= 0 && db[4] >= 0) return true;                  // nothing wraps: overlapping AABBs always share a cell
    const int x0 = max(da[0], db[0]), x1 = min(da[1], db[1]), y0 = max(da[2], db[2]), y1 = min(da[3], db[3]);
    if (max(da[4], db[4]) > min(da[5], db[5])) return false;
    for (int x = x0; x <= x1; x++) for (int y = y0; y <= y1; y++)
        if (odeb_hash_column_wraps(lb, x, y, da[4]) == odeb_hash_column_wraps(lb, x, y, db[4])) return true;
    return false;
}

// dCollideSpheres collision_util.cpp:38-67
__device__ int odeb_collide_spheres(const Real *p1, Real r1, const Real *p2, Real r2, DContactGeom *c)
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
__device__ int odeb_sphere_box(const DGeom &o1, const DGeom &o2, DContactGeom *c)
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
__device__ int odeb_sphere_plane(const DGeom &o1, const DGeom &o2, DContactGeom *c)
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

// box-box: odeb_boxbox.cuh (separating-axis loop, edge / face manifold builders, rectangle clipper, contact spreading).  Kept out of
// line: k_narrow and the capsule-box collider both call it, and inlining it twice costs k_narrow ~40 registers.
#include "odeb_boxbox.cuh"
__device__ __noinline__ int odeb_box_box(const Real *p1, const Real *R1, const Real *side1, const Real *p2, const Real *R2, const Real *side2,
                                         Real *normal, Real *depth, int *return_code, int flags, DContactGeom *contact)
{
    return odeb_bb_collide(p1, R1, side1, p2, R2, side2, normal, depth, return_code, flags, contact);
}

// dCollideBoxBox box.cpp:741-767
__device__ int odeb_collide_box_box(const DGeom &o1, const DGeom &o2, int flags, DContactGeom *c)
{
    Real normal[3], depth; int code;
    int num = odeb_box_box(o1.pos, o1.R, o1.p, o2.pos, o2.R, o2.p, normal, &depth, &code, flags, c);
    for (int i = 0; i < num; i++) { c[i].normal[0] = -normal[0]; c[i].normal[1] = -normal[1]; c[i].normal[2] = -normal[2]; }
    return num;
}

// dCollideBoxPlane box.cpp:770-903
__device__ int odeb_box_plane(const DGeom &o1, const DGeom &o2, int flags, DContactGeom *c)
{
    const Real *R = o1.R, *n = o2.p, *side = o1.p;
    int ret = 0;
    Real Q1 = dot3s(n, 1, R + 0, 4), Q2 = dot3s(n, 1, R + 1, 4), Q3 = dot3s(n, 1, R + 2, 4);
    Real Av[3] = { side[0] * Q1, side[1] * Q2, side[2] * Q3 };
    Real Bv[3] = { RFABS(Av[0]), RFABS(Av[1]), RFABS(Av[2]) };
    Real depth = n[3] + R_(0.5) * (Bv[0] + Bv[1] + Bv[2]) - dot3(n, o1.pos);
    if (depth < 0) return 0;
    int maxc = flags & ODEB_NUMC_MASK;
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



// dClosestLineSegmentPoints collision_util.cpp:109-223
__device__ __noinline__ void odeb_closest_segment_points(const Real *a1, const Real *a2, const Real *b1, const Real *b2, Real *cp1, Real *cp2)
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
__device__ __noinline__ void odeb_closest_line_box_points(const Real *p1, const Real *p2, const Real *c, const Real *R, const Real *side, Real *lret, Real *bret)
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
__device__ int odeb_capsule_sphere(const DGeom &o1, const DGeom &o2, DContactGeom *c)
{
    const Real *R = o1.R;
    Real alpha = R[2] * (o2.pos[0] - o1.pos[0]) + R[6] * (o2.pos[1] - o1.pos[1]) + R[10] * (o2.pos[2] - o1.pos[2]);
    Real lz2 = o1.p[1] * R_(0.5);
    if (alpha > lz2) alpha = lz2;
    if (alpha < -lz2) alpha = -lz2;
    Real p[3] = { o1.pos[0] + alpha * R[2], o1.pos[1] + alpha * R[6], o1.pos[2] + alpha * R[10] };
    return odeb_collide_spheres(p, o1.p[0], o2.pos, o2.p[0], c);
}

// dCollideCapsuleBox capsule.cpp:166-237
__device__ int odeb_capsule_box(const DGeom &o1, const DGeom &o2, int flags, DContactGeom *contact)
{
    const Real *R1 = o1.R;
    Real p1[3], p2[3];
    Real clen = o1.p[1] * R_(0.5);
    p1[0] = o1.pos[0] + clen * R1[2]; p1[1] = o1.pos[1] + clen * R1[6]; p1[2] = o1.pos[2] + clen * R1[10];
    p2[0] = o1.pos[0] - clen * R1[2]; p2[1] = o1.pos[1] - clen * R1[6]; p2[2] = o1.pos[2] - clen * R1[10];
    Real radius = o1.p[0];
    Real pl[3], pb[3];
    odeb_closest_line_box_points(p1, p2, o2.pos, o2.R, o2.p, pl, pb);
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
        int num = odeb_box_box(o2.pos, o2.R, o2.p, o1.pos, o1.R, capboxside, normal, &depth, &code, flags, contact);
        for (int i = 0; i < num; i++) { contact[i].normal[0] = normal[0]; contact[i].normal[1] = normal[1]; contact[i].normal[2] = normal[2]; }
        return num;
    }
    return odeb_collide_spheres(pl, radius, pb, 0, contact);
}

// dCollideCapsuleCapsule capsule.cpp:240-353
__device__ int odeb_capsule_capsule(const DGeom &o1, const DGeom &o2, int flags, DContactGeom *contact)
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
            int num_contacts = flags & ODEB_NUMC_MASK;
            if (num_contacts >= 2 && lo < hi) {
                for (i = 0; i < 3; i++) sphere1[i] = pos1[i] + lo * axis1[i];
                for (i = 0; i < 3; i++) sphere2[i] = pos2[i] + (lo + k) * axis2[i];
                int n1 = odeb_collide_spheres(sphere1, o1.p[0], sphere2, o2.p[0], contact);
                if (n1) {
                    for (i = 0; i < 3; i++) sphere1[i] = pos1[i] + hi * axis1[i];
                    for (i = 0; i < 3; i++) sphere2[i] = pos2[i] + (hi + k) * axis2[i];
                    int n2 = odeb_collide_spheres(sphere1, o1.p[0], sphere2, o2.p[0], contact + 1);
                    if (n2) return 2;
                }
            }
            Real alpha1 = (lo + hi) * R_(0.5);
            Real alpha2 = alpha1 + k;
            for (i = 0; i < 3; i++) sphere1[i] = pos1[i] + alpha1 * axis1[i];
            for (i = 0; i < 3; i++) sphere2[i] = pos2[i] + alpha2 * axis2[i];
            return odeb_collide_spheres(sphere1, o1.p[0], sphere2, o2.p[0], contact);
        }
    }
    Real a1[3], a2[3], b1[3], b2[3];
    for (i = 0; i < 3; i++) {
        a1[i] = pos1[i] + axis1[i] * lz1; a2[i] = pos1[i] - axis1[i] * lz1;
        b1[i] = pos2[i] + axis2[i] * lz2; b2[i] = pos2[i] - axis2[i] * lz2;
    }
    odeb_closest_segment_points(a1, a2, b1, b2, sphere1, sphere2);
    return odeb_collide_spheres(sphere1, o1.p[0], sphere2, o2.p[0], contact);
}

// dCollideCapsulePlane capsule.cpp:356-415
__device__ int odeb_capsule_plane(const DGeom &o1, const DGeom &o2, int flags, DContactGeom *contact)
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
    if ((flags & ODEB_NUMC_MASK) >= 2) {
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


#include "odeb_ray_cyl.cuh"

// dCollide collision_kernel.cpp:292-338 with the collider table of dInitColliders (:166-268):
// direct entries (sphere,sphere) (sphere,box) (sphere,plane) (box,box) (box,plane) (capsule,sphere)
// (capsule,box) (capsule,capsule) (capsule,plane); the transposed pairs call the same function with
// swapped arguments and negate the normals.
__device__ int odeb_collide_direct(const DGeom &a, const DGeom &b, int flags, DContactGeom *c, int *handled)
{
    *handled = 1;
    if (a.type == 0 && b.type == 0) return odeb_collide_spheres(a.pos, a.p[0], b.pos, b.p[0], c);
    if (a.type == 0 && b.type == 1) return odeb_sphere_box(a, b, c);
    if (a.type == 0 && b.type == 4) return odeb_sphere_plane(a, b, c);
    if (a.type == 1 && b.type == 1) return odeb_collide_box_box(a, b, flags, c);
    if (a.type == 1 && b.type == 4) return odeb_box_plane(a, b, flags, c);
    if (a.type == 2 && b.type == 0) return odeb_capsule_sphere(a, b, c);
    if (a.type == 2 && b.type == 1) return odeb_capsule_box(a, b, flags, c);
    if (a.type == 2 && b.type == 2) return odeb_capsule_capsule(a, b, flags, c);
    if (a.type == 2 && b.type == 4) return odeb_capsule_plane(a, b, flags, c);
    if (a.type == 5 && b.type == 0) return odeb_ray_sphere(a, b, c);          // collision_kernel.cpp:191-195
    if (a.type == 5 && b.type == 1) return odeb_ray_box(a, b, c);
    if (a.type == 5 && b.type == 2) return odeb_ray_capsule(a, b, c);
    if (a.type == 5 && b.type == 4) return odeb_ray_plane(a, b, c);
    if (a.type == 5 && b.type == 3) return odeb_ray_cylinder(a, b, c);
    if (a.type == 3 && b.type == 1) return odeb_cylinder_box(a, b, flags, c);   // :210
    if (a.type == 3 && b.type == 0) return odeb_cylinder_sphere(a, b, c);     // :212-213; no cylinder-capsule / cylinder-cylinder collider
    if (a.type == 3 && b.type == 4) return odeb_cylinder_plane(a, b, flags, c); // exists without libccd
    *handled = 0;
    return 0;
}

__device__ int odeb_collide(const DGeom &o1, const DGeom &o2, int flags, DContactGeom *c)
{
    if ((flags & ODEB_NUMC_MASK) == 0) return 0;
    if (o1.body == o2.body && o1.body >= 0) return 0;
    int handled;
    int n = odeb_collide_direct(o1, o2, flags, c, &handled);
    if (handled) return n;
    n = odeb_collide_direct(o2, o1, flags, c, &handled);
    if (!handled) return 0;
    for (int i = 0; i < n; i++) { c[i].normal[0] = -c[i].normal[0]; c[i].normal[1] = -c[i].normal[1]; c[i].normal[2] = -c[i].normal[2]; }
    return n;
}
#endif
