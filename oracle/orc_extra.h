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

unsigned long orc_rand_next(uint32_t *seed) { return orc_rand(seed); }
int orc_rand_int_(uint32_t *seed, int n) { return orc_rand_int(seed, n); }
int orc_sizeof_real(void) { return (int)sizeof(Real); }

}
