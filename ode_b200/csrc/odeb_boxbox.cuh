// odeb_boxbox.cuh -- box-box narrowphase of the B200 step path.
//
// What it computes is what the reference's dBoxBox computes (ode/src/box.cpp:356-737: separating-axis search over the 15 candidate
// axes with the 1.05 preference for face axes, then either the closest points of two edges or the incident face clipped against the
// reference face and culled to the requested number of contacts, box.cpp:212-336), and it has to do so with the reference's rounding:
// every sum below keeps the reference's operand order (a build with -fmad=false is then bit-identical, tests/golden/collide_*.npz,
// tools/boxbox_host_check.py).  How it is organised is this library's own:
//   * the 15 axes are ONE loop over a descriptor computed from the axis number (3 face axes of A, 3 of B, 9 edge pairs through the
//     index tables LO/HI), not 15 macro expansions; the running best axis is a small struct;
//   * the manifold builders are separate functions (edge-edge support points; face clipping) over a BoxFrame view of a box, so the
//     "which box is the reference" case is a swap of two views;
//   * the rectangle clipper is a four-pass half-plane clipper over a fixed-capacity polygon with an explicit "full" result instead
//     of pointer juggling between two buffers;
//   * the culling step works on polar angles about the area centroid and hands back indices.
// Everything is __host__ __device__: the same source is compiled into a host harness and checked against the compiled reference on
// tens of thousands of random and degenerate pairs without a GPU.
#ifndef ODEB_BOXBOX_CUH
#define ODEB_BOXBOX_CUH
#include "odeb_math.cuh"

#define ODEB_BB_HD __host__ __device__

// a box as the manifold builders see it: centre, rotation (3 x 4 row-major, columns = box axes in world space), half extents
struct BoxFrame {
    const Real *c, *R; Real h[3];
    ODEB_BB_HD Real axis(int k, int comp) const { return R[4 * comp + k]; }      // component `comp` of box axis k
};

struct BestAxis {
    Real gap;            // largest (least negative) signed separation found so far; penetration depth = -gap
    int id;              // 0 none, 1..3 face of A, 4..6 face of B, 7..15 edge i of A x edge j of B (7 + 3 i + j)
    bool flip;           // the normal points from B to A along the axis: negate
    Real n_in_a[3];      // edge axes: unit normal in A's frame
};

// the two indices other than k, ascending
#define ODEB_BB_LO(k) ((k) == 0 ? 1 : 0)
#define ODEB_BB_HI(k) ((k) == 2 ? 1 : 2)

// ---------------------------------------------------------------------------------------------------------------------------
// rectangle clipper: the quadrilateral q (4 vertices, 2-D) against |x| < hx, |y| < hy, one half-plane at a time.  At most 8 output
// vertices are ever kept: when a pass produces its 8th vertex the clipper stops there (box.cpp:212-263 does the same).
struct ClipPoly { Real v[8][2]; int n; };

ODEB_BB_HD inline int odeb_bb_clip_quad(const Real half[2], const Real quad[4][2], ClipPoly &out)
{
    ClipPoly cur, nxt;
    cur.n = 4;
    for (int k = 0; k < 4; k++) { cur.v[k][0] = quad[k][0]; cur.v[k][1] = quad[k][1]; }
    for (int pass = 0; pass < 4; pass++) {
        const int ax = pass >> 1, other = 1 - ax;
        const Real sgn = (pass & 1) ? R_(1.0) : R_(-1.0);
        const Real lim = half[ax];
        nxt.n = 0;
        bool full = false;
        for (int k = 0; k < cur.n && !full; k++) {
            const Real *a = cur.v[k], *b = cur.v[k + 1 < cur.n ? k + 1 : 0];
            const bool a_in = sgn * a[ax] < lim, b_in = sgn * b[ax] < lim;
            if (a_in) {
                nxt.v[nxt.n][0] = a[0]; nxt.v[nxt.n][1] = a[1];
                full = ++nxt.n == 8;
            }
            if (!full && a_in != b_in) {      // the edge a -> b crosses the boundary sgn * x[ax] = lim
                nxt.v[nxt.n][other] = a[other] + (b[other] - a[other]) / (b[ax] - a[ax]) * (sgn * lim - a[ax]);
                nxt.v[nxt.n][ax] = sgn * lim;
                full = ++nxt.n == 8;
            }
        }
        cur = nxt;
        if (full) break;
    }
    out = cur;
    return cur.n;
}

// ---------------------------------------------------------------------------------------------------------------------------
// keep `want` of the n coplanar points (2-D coordinates xy), spread around their area centroid: `first` is always kept, the others
// are the points whose polar angle is nearest to first + j * 2 pi / want (box.cpp:274-336; the pi arithmetic is double in both builds)
ODEB_BB_HD inline void odeb_bb_spread(int n, const Real xy[][2], int want, int first, int keep[])
{
    Real gx, gy;
    if (n == 1) { gx = xy[0][0]; gy = xy[0][1]; }
    else if (n == 2) { gx = R_(0.5) * (xy[0][0] + xy[1][0]); gy = R_(0.5) * (xy[0][1] + xy[1][1]); }
    else {
        Real area = 0, sx = 0, sy = 0, w;
        for (int k = 0; k + 1 < n; k++) {
            w = xy[k][0] * xy[k + 1][1] - xy[k + 1][0] * xy[k][1];
            area += w;
            sx += w * (xy[k][0] + xy[k + 1][0]);
            sy += w * (xy[k][1] + xy[k + 1][1]);
        }
        w = xy[n - 1][0] * xy[0][1] - xy[0][0] * xy[n - 1][1];
        const Real inv = rrecip(R_(3.0) * (area + w));
        gx = inv * (sx + w * (xy[n - 1][0] + xy[0][0]));
        gy = inv * (sy + w * (xy[n - 1][1] + xy[0][1]));
    }
    Real ang[8];
    bool taken[8];
    for (int k = 0; k < n; k++) { ang[k] = RATAN2(xy[k][1] - gy, xy[k][0] - gx); taken[k] = false; }
    taken[first] = true;
    keep[0] = first;
    for (int j = 1; j < want; j++) {
        Real target = (Real)((Real)j * (2 * M_PI / want) + ang[first]);
        if (target > M_PI) target -= (Real)(2 * M_PI);
        Real best = 1e9;
        int pick = first;
        for (int k = 0; k < n; k++) {
            if (taken[k]) continue;
            Real d = RFABS(ang[k] - target);
            if (d > M_PI) d = (Real)(2 * M_PI - d);
            if (d < best) { best = d; pick = k; }
        }
        taken[pick] = true;
        keep[j] = pick;
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// closest points of the lines pa + s ua and pb + t ub (collision_util.cpp:70-92)
ODEB_BB_HD inline void odeb_bb_lines_nearest(const Real *pa, const Real *ua, const Real *pb, const Real *ub, Real &s, Real &t)
{
    const Real d[3] = { pb[0] - pa[0], pb[1] - pa[1], pb[2] - pa[2] };
    const Real c = dot3(ua, ub), qa = dot3(ua, d), qb = -dot3(ub, d);
    Real det = 1 - c * c;
    if (det <= R_(0.0001)) { s = 0; t = 0; return; }
    det = rrecip(det);
    s = (qa + c * qb) * det;
    t = (c * qa + qb) * det;
}

// the corner of box F furthest along +n (towards = +1) or along -n (towards = -1), written into out
ODEB_BB_HD inline void odeb_bb_extreme_corner(const BoxFrame &F, const Real n[3], Real towards, Real out[3])
{
    out[0] = F.c[0]; out[1] = F.c[1]; out[2] = F.c[2];
    for (int k = 0; k < 3; k++) {
        const Real along = n[0] * F.axis(k, 0) + n[1] * F.axis(k, 1) + n[2] * F.axis(k, 2);
        const Real sg = along > 0 ? towards : -towards;
        for (int comp = 0; comp < 3; comp++) out[comp] += sg * F.h[k] * F.axis(k, comp);
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// face manifold: Fr owns the face whose outward normal is n (axis nk of Fr); the face of Fi that looks at it most directly is
// projected into Fr's face plane, clipped against Fr's face and every clipped vertex that lies below the face becomes a contact
// candidate.  Returns the number of contacts written (positions in world space, depths).
ODEB_BB_HD inline int odeb_bb_face_manifold(const BoxFrame &Fr, const BoxFrame &Fi, const Real n[3], int nk, int flags, DContactGeom *contact)
{
    // the incident face: the axis of Fi with the largest |n . axis| (ties go to the higher index, as the reference's comparisons do)
    Real ni[3], mag[3];
    mul1_331(ni, Fi.R, n);
    for (int k = 0; k < 3; k++) mag[k] = RFABS(ni[k]);
    int big = mag[1] > mag[0] ? 1 : 0;
    if (!(mag[big] > mag[2])) big = 2;
    const int u = big == 0 ? 1 : 0, v = big == 2 ? 1 : 2;          // the incident face's own two axes
    // centre of the incident face relative to Fr's centre
    Real fc[3];
    if (ni[big] < 0) { for (int comp = 0; comp < 3; comp++) fc[comp] = Fi.c[comp] - Fr.c[comp] + Fi.h[big] * Fi.axis(big, comp); }
    else { for (int comp = 0; comp < 3; comp++) fc[comp] = Fi.c[comp] - Fr.c[comp] - Fi.h[big] * Fi.axis(big, comp); }
    // Fr's face plane: in-plane axes s, t (the two axes other than nk, ascending)
    const int s = ODEB_BB_LO(nk), t = ODEB_BB_HI(nk);
    const Real cs = dot3s(fc, 1, Fr.R + s, 4), ct = dot3s(fc, 1, Fr.R + t, 4);
    Real msu = dot3s(Fr.R + s, 4, Fi.R + u, 4), msv = dot3s(Fr.R + s, 4, Fi.R + v, 4);
    Real mtu = dot3s(Fr.R + t, 4, Fi.R + u, 4), mtv = dot3s(Fr.R + t, 4, Fi.R + v, 4);
    Real quad[4][2];
    {
        const Real su = msu * Fi.h[u], tu = mtu * Fi.h[u], sv = msv * Fi.h[v], tv = mtv * Fi.h[v];
        quad[0][0] = cs - su - sv; quad[0][1] = ct - tu - tv;
        quad[1][0] = cs - su + sv; quad[1][1] = ct - tu + tv;
        quad[2][0] = cs + su + sv; quad[2][1] = ct + tu + tv;
        quad[3][0] = cs + su - sv; quad[3][1] = ct + tu - tv;
    }
    const Real half[2] = { Fr.h[s], Fr.h[t] };
    ClipPoly poly;
    if (odeb_bb_clip_quad(half, quad, poly) < 1) return 0;
    // back to 3-D through the inverse of the 2 x 2 projection, keep what penetrates
    const Real inv = rrecip(msu * mtv - msv * mtu);
    msu *= inv; msv *= inv; mtu *= inv; mtv *= inv;
    Real p3[8][3], pen[8], xy[8][2];
    int kept = 0;
    const unsigned stop_at = (unsigned)flags & (ODEB_NUMC_MASK | ODEB_CONTACTS_UNIMPORTANT);
    for (int k = 0; k < poly.n; k++) {
        const Real a = mtv * (poly.v[k][0] - cs) - msv * (poly.v[k][1] - ct);
        const Real b = -mtu * (poly.v[k][0] - cs) + msu * (poly.v[k][1] - ct);
        for (int comp = 0; comp < 3; comp++) p3[kept][comp] = fc[comp] + a * Fi.axis(u, comp) + b * Fi.axis(v, comp);
        pen[kept] = Fr.h[nk] - dot3(n, p3[kept]);
        if (pen[kept] >= 0) {
            xy[kept][0] = poly.v[k][0]; xy[kept][1] = poly.v[k][1];
            kept++;
            if (((unsigned)kept | ODEB_CONTACTS_UNIMPORTANT) == stop_at) break;
        }
    }
    if (kept < 1) return 0;
    int want = flags & ODEB_NUMC_MASK;
    if (want > kept) want = kept;
    if (want < 1) want = 1;
    if (kept <= want) {
        for (int k = 0; k < kept; k++) {
            for (int comp = 0; comp < 3; comp++) contact[k].pos[comp] = p3[k][comp] + Fr.c[comp];
            contact[k].depth = pen[k];
        }
        return kept;
    }
    int deepest = 0;
    for (int k = 1; k < kept; k++) if (pen[k] > pen[deepest]) deepest = k;
    int keep[8];
    odeb_bb_spread(kept, xy, want, deepest, keep);
    for (int k = 0; k < want; k++) {
        for (int comp = 0; comp < 3; comp++) contact[k].pos[comp] = p3[keep[k]][comp] + Fr.c[comp];
        contact[k].depth = pen[keep[k]];
    }
    return want;
}

// ---------------------------------------------------------------------------------------------------------------------------
// Returns the contact count; `normal` (from box 2 towards box 1 is NOT applied here: see odeb_collide_box_box), depth and the axis
// code (1..15) as the reference reports them.
ODEB_BB_HD inline int odeb_bb_collide(const Real *p1, const Real *R1, const Real *side1, const Real *p2, const Real *R2, const Real *side2,
                                   Real *normal, Real *depth, int *return_code, int flags, DContactGeom *contact)
{
    BoxFrame A, B;
    A.c = p1; A.R = R1; B.c = p2; B.R = R2;
    for (int k = 0; k < 3; k++) { A.h[k] = side1[k] * R_(0.5); B.h[k] = side2[k] * R_(0.5); }
    const Real d[3] = { p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2] };
    Real da[3];                                   // the centre offset in A's frame
    mul1_331(da, R1, d);
    Real C[3][3], aC[3][3];                       // C = R1^T R2 and its absolute values
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { C[i][j] = dot3s(R1 + i, 4, R2 + j, 4); aC[i][j] = RFABS(C[i][j]); }

    BestAxis best;
    best.gap = -R_INF; best.id = 0; best.flip = false;
    best.n_in_a[0] = best.n_in_a[1] = best.n_in_a[2] = 0;
    const bool first_hit_wins = (flags & ODEB_CONTACTS_UNIMPORTANT) != 0;
    const Real face_preference = R_(1.05);
    for (int ax = 1; ax <= 15; ax++) {
        Real proj, reach;
        if (ax <= 3) {                            // face k of A
            const int k = ax - 1;
            proj = da[k];
            reach = A.h[k] + B.h[0] * aC[k][0] + B.h[1] * aC[k][1] + B.h[2] * aC[k][2];
        } else if (ax <= 6) {                     // face k of B
            const int k = ax - 4;
            proj = dot3s(R2 + k, 4, d, 1);
            reach = A.h[0] * aC[0][k] + A.h[1] * aC[1][k] + A.h[2] * aC[2][k] + B.h[k];
        } else {                                  // edge i of A x edge j of B
            const int i = (ax - 7) / 3, j = (ax - 7) % 3, i1 = (i + 1) % 3, i2 = (i + 2) % 3;
            proj = da[i2] * C[i1][j] - da[i1] * C[i2][j];
            reach = A.h[ODEB_BB_LO(i)] * aC[ODEB_BB_HI(i)][j] + A.h[ODEB_BB_HI(i)] * aC[ODEB_BB_LO(i)][j]
                  + B.h[ODEB_BB_LO(j)] * aC[i][ODEB_BB_HI(j)] + B.h[ODEB_BB_HI(j)] * aC[i][ODEB_BB_LO(j)];
        }
        Real gap = RFABS(proj) - reach;
        if (gap > 0) return 0;                    // a separating axis
        if (ax <= 6) {
            if (gap > best.gap) { best.gap = gap; best.id = ax; best.flip = proj < 0; if (first_hit_wins) break; }
        } else {
            const int i = (ax - 7) / 3, j = (ax - 7) % 3, i1 = (i + 1) % 3, i2 = (i + 2) % 3;
            Real n[3];
            n[i] = 0; n[i1] = -C[i2][j]; n[i2] = C[i1][j];
            const Real len = RSQRT(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
            if (len > 0) {                        // parallel edges give no axis
                gap /= len;
                if (gap * face_preference > best.gap) {
                    best.gap = gap; best.id = ax; best.flip = proj < 0;
                    best.n_in_a[0] = n[0] / len; best.n_in_a[1] = n[1] / len; best.n_in_a[2] = n[2] / len;
                    if (first_hit_wins) break;
                }
            }
        }
    }
    if (!best.id) return 0;
    if (best.id <= 3) { const int k = best.id - 1; normal[0] = R1[k]; normal[1] = R1[4 + k]; normal[2] = R1[8 + k]; }
    else if (best.id <= 6) { const int k = best.id - 4; normal[0] = R2[k]; normal[1] = R2[4 + k]; normal[2] = R2[8 + k]; }
    else mul0_331(normal, R1, best.n_in_a);
    if (best.flip) { normal[0] = -normal[0]; normal[1] = -normal[1]; normal[2] = -normal[2]; }
    *depth = -best.gap;
    *return_code = best.id;

    if (best.id > 6) {
        // two edges: a point on each (the corners that face each other along the normal), then the nearest points of the two lines
        Real pa[3], pb[3], ua[3], ub[3], s, t;
        odeb_bb_extreme_corner(A, normal, R_(1.0), pa);
        odeb_bb_extreme_corner(B, normal, R_(-1.0), pb);
        const int i = (best.id - 7) / 3, j = (best.id - 7) % 3;
        for (int comp = 0; comp < 3; comp++) { ua[comp] = A.axis(i, comp); ub[comp] = B.axis(j, comp); }
        odeb_bb_lines_nearest(pa, ua, pb, ub, s, t);
        for (int comp = 0; comp < 3; comp++) pa[comp] += ua[comp] * s;
        for (int comp = 0; comp < 3; comp++) pb[comp] += ub[comp] * t;
        for (int comp = 0; comp < 3; comp++) contact[0].pos[comp] = R_(0.5) * (pa[comp] + pb[comp]);
        contact[0].depth = *depth;
        return 1;
    }
    if (best.id <= 3) return odeb_bb_face_manifold(A, B, normal, best.id - 1, flags, contact);
    const Real back[3] = { -normal[0], -normal[1], -normal[2] };
    return odeb_bb_face_manifold(B, A, back, best.id - 4, flags, contact);
}
#endif
