// Host harness of the box-box collider (odeb_boxbox.cuh compiled for the CPU): lets tools/boxbox_host_check.py compare it with the
// compiled reference on random / degenerate pairs without a GPU.  Build: nvcc -x cu -O2 -fmad=false [-DODEB_DOUBLE] -shared -Xcompiler -fPIC
#include <cuda_runtime.h>
#include "../ode_b200/csrc/odeb_math.cuh"
struct DContactGeom { Real pos[3], normal[3], depth; };
#define ODEB_NUMC_MASK 0xffff
#define ODEB_CONTACTS_UNIMPORTANT 0x80000000
#include "../ode_b200/csrc/odeb_boxbox.cuh"
extern "C" int bbh_box_box(const Real *side1, const Real *pos1, const Real *R1, const Real *side2, const Real *pos2, const Real *R2, int flags, Real *geom7, int *code)
{
    DContactGeom c[8];
    Real normal[3], depth;
    int n = odeb_bb_collide(pos1, R1, side1, pos2, R2, side2, normal, &depth, code, flags, c);
    for (int i = 0; i < n; i++) {   // dCollideBoxBox box.cpp:741-767: the contact normal is the negated dBoxBox normal
        geom7[7 * i] = c[i].pos[0]; geom7[7 * i + 1] = c[i].pos[1]; geom7[7 * i + 2] = c[i].pos[2];
        geom7[7 * i + 3] = -normal[0]; geom7[7 * i + 4] = -normal[1]; geom7[7 * i + 5] = -normal[2]; geom7[7 * i + 6] = c[i].depth;
    }
    return n;
}
