/* Two user-defined device operators, written the way an application would write them against
 * dc-lib_b200/csrc/dc_device_operator.cuh (test of the device-operator hook):
 *   user_mass_op       P1 mass matrix, scalar, symmetric:   M[a][b] = vol * (a == b ? 0.1 : 0.05)
 *   user_advection_op  P1 advection, scalar, NOT symmetric: C[a][b] = vol * 0.25 * (g_b . w),
 *                      w = the 3 operator parameters -- needs the general (non-symmetric) schedule.
 * Built by __graft_entry__.build() into tests/custom_op/libcustom_ops.so. */
#include "dc_device_operator.cuh"

struct UserMassOp : dcdev::KeOperator<UserMassOp, 1, true> {
    static constexpr int NPARAM = 0;
    __device__ static void ke(const double (&x)[4][3], const double *, double *k)
    {
        double g[12], vol;
        dcdev::tet_gradients(x, g, vol);
        const double off = __dmul_rn(vol, 0.05), diag = __dmul_rn(vol, 0.1);
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int b = 0; b < 4; b++) k[4 * a + b] = (a == b) ? diag : off;
    }
};

struct UserAdvectionOp : dcdev::KeOperator<UserAdvectionOp, 1, false> {
    static constexpr int NPARAM = 3;
    __device__ static void ke(const double (&x)[4][3], const double *w, double *k)
    {
        double g[12], vol;
        dcdev::tet_gradients(x, g, vol);
        const double q = __dmul_rn(vol, 0.25);
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const double gw = __fma_rn(g[3 * b + 2], w[2], __fma_rn(g[3 * b + 1], w[1], __dmul_rn(g[3 * b], w[0])));
            const double v = __dmul_rn(q, gw);
#pragma unroll
            for (int a = 0; a < 4; a++) k[4 * a + b] = v;
        }
    }
};

DC_DEVICE_OPERATOR(user_mass_op, UserMassOp)
DC_DEVICE_OPERATOR(user_advection_op, UserAdvectionOp)
