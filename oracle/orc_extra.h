// orc_extra.h -- TEST INFRASTRUCTURE (oracle). Single-function entry points used by the golden-vector
// tests (same signatures as the ref_* helpers in oracle/ref_driver.cpp).
extern "C" {

static void orc_make_geom(OrcGeom &g, int type, const Real *p, const Real *pos, const Real *R)
{
    g.type = type; g.body = -1; g.cat = g.col = ~0u;
    for (int k = 0; k < 4; k++) g.p[k] = (type == ODEB_PLANE || k < 3) ? p[k] : 0;
    if (type == ODEB_PLANE) {
        Real l = g.p[0] * g.p[0] + g.p[1] * g.p[1] + g.p[2] * g.p[2];
        if (l > 0) { l = rrecipsqrt(l); g.p[0] *= l; g.p[1] *= l; g.p[2] *= l; g.p[3] *= l; }
        else { g.p[0] = 1; g.p[1] = 0; g.p[2] = 0; g.p[3] = 0; }
    }
    g.pos = pos; g.R = R;
}

int orc_collide_pair(int type1, const Real *p1, const Real *pos1, const Real *R1,
                     int type2, const Real *p2, const Real *pos2, const Real *R2, int flags, Real *geom7, int cap)
{
    OrcGeom a, b;
    orc_make_geom(a, type1, p1, pos1, R1); orc_make_geom(b, type2, p2, pos2, R2);
    OrcContactGeom c[16];
    int n = orc_collide(a, b, flags, c);
    for (int i = 0; i < n && i < cap; i++) {
        for (int k = 0; k < 3; k++) { geom7[7 * i + k] = c[i].pos[k]; geom7[7 * i + 3 + k] = c[i].normal[k]; }
        geom7[7 * i + 6] = c[i].depth;
    }
    return n;
}

void orc_geom_aabb(int type, const Real *p, const Real *pos, const Real *R, Real *aabb6)
{
    OrcGeom g; orc_make_geom(g, type, p, pos, R);
    orc_compute_aabb(g);
    for (int k = 0; k < 6; k++) aabb6[k] = g.aabb[k];
}

/* dxJointContact::getInfo1 + getInfo2 on an explicit contact between two free bodies at pos1/pos2 (identity
 * rotation, zero velocity), 16-wide row layout -- the call pattern of the reference's tests/friction.cpp:69-173. */
int orc_contact_rows(int mode, Real mu, Real mu2, const Real *cpos, const Real *cnormal, Real depth, const Real *fdir1,
                     const Real *pos1, const Real *pos2, Real fps, Real erp, Real *rows48, int *findex3)
{
    Batch B;
    memset(&B.wp, 0, sizeof(B.wp));
    B.wp.surf_mode = mode & ~ODEB_CONTACT_FDIR1; B.wp.mu = mu; B.wp.mu2 = mu2;
    B.max_vel = R_INF; B.min_depth = 0; B.cfm = 0;
    World W;
    W.bodies.resize(2);
    for (int i = 0; i < 2; i++) {
        Body &b = W.bodies[i];
        memset(b.pos, 0, sizeof(Real) * (4 + 12 + 4 + 4 + 4 + 4 + 4));
        const Real *p = i ? pos2 : pos1;
        b.pos[0] = p[0]; b.pos[1] = p[1]; b.pos[2] = p[2];
        b.R[0] = b.R[5] = b.R[10] = 1; b.q[0] = 1;
    }
    Joint j;
    memset(&j, 0, sizeof(j));
    j.type = ODEB_JOINT_CONTACT; j.b0 = 0; j.b1 = 1;
    for (int k = 0; k < 3; k++) { j.cg.pos[k] = cpos[k]; j.cg.normal[k] = cnormal[k]; }
    j.cg.depth = depth;
    contact_info1(B, j);
    for (int k = 0; k < 48; k++) rows48[k] = 0;
    for (int k = 0; k < 3; k++) findex3[k] = -1;
    contact_info2(B, W, j, fps, erp, rows48, findex3, (mode & ODEB_CONTACT_FDIR1) ? fdir1 : 0);
    return j.m;
}

unsigned long orc_rand_next(uint32_t *seed) { return orc_rand(seed); }
int orc_rand_int_(uint32_t *seed, int n) { return orc_rand_int(seed, n); }
int orc_sizeof_real(void) { return (int)sizeof(Real); }

}
