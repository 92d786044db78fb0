/* dc_operators.cuh -- device-side element operators for the assembly kernels.
 *
 * DC-lib itself contains no element kernel: the reference hands each leaf to a
 * HOST callback (src/tree_traversal.cc:51,56,61) that computes the element matrix
 * and `+=`s it into nodeToNodeValue.  Host function pointers cannot run on the
 * GPU, so the callbacks are replaced by device operators with this contract:
 *
 *   setup(x[4][3], params)      once per element: whatever the operator wants to
 *                               keep for its 16 blocks (here the 4 basis-function
 *                               gradients and 1-2 scalars), NS doubles in total;
 *   block(slot, a, b, out[D*D]) the D x D block K_e[a][b] from that record.
 *
 * Every product / fma / sum is written with an explicit rounding intrinsic so
 * that nvcc cannot contract or reassociate: the sequence is the operator's
 * numeric definition and deterministic-mode results are compared bit for bit
 * against a CPU evaluation of the same sequence.
 */
#ifndef DC_OPERATORS_CUH
#define DC_OPERATORS_CUH

namespace dcdev {

/* record layout: g[a][k] at 3*a+k (a = 0..3), scalars from index 12 */
#define DC_SLOT_G 12

__device__ __forceinline__ double cross_comp(double p, double q, double r, double s)
{
    return __fma_rn(p, q, -__dmul_rn(r, s));
}

/* Gradients of the P1 basis functions and |det|/6.  a,b,c = edges from node 0,
 * g1 = (b x c)/det, g2 = (c x a)/det, g3 = (a x b)/det, g0 = -((g1+g2)+g3). */
__device__ __forceinline__ void tet_gradients(const double (&x)[4][3], double *g, double &vol)
{
    double a[3], b[3], c[3], n1[3], n2[3], n3[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        a[k] = __dsub_rn(x[1][k], x[0][k]);
        b[k] = __dsub_rn(x[2][k], x[0][k]);
        c[k] = __dsub_rn(x[3][k], x[0][k]);
    }
    n1[0] = cross_comp(b[1], c[2], b[2], c[1]);
    n1[1] = cross_comp(b[2], c[0], b[0], c[2]);
    n1[2] = cross_comp(b[0], c[1], b[1], c[0]);
    n2[0] = cross_comp(c[1], a[2], c[2], a[1]);
    n2[1] = cross_comp(c[2], a[0], c[0], a[2]);
    n2[2] = cross_comp(c[0], a[1], c[1], a[0]);
    n3[0] = cross_comp(a[1], b[2], a[2], b[1]);
    n3[1] = cross_comp(a[2], b[0], a[0], b[2]);
    n3[2] = cross_comp(a[0], b[1], a[1], b[0]);
    double det = __fma_rn(a[2], n1[2], __fma_rn(a[1], n1[1], __dmul_rn(a[0], n1[0])));
    double inv = __ddiv_rn(1.0, det);
#pragma unroll
    for (int k = 0; k < 3; k++) {
        double g1 = __dmul_rn(n1[k], inv), g2 = __dmul_rn(n2[k], inv), g3 = __dmul_rn(n3[k], inv);
        g[3 + k] = g1;
        g[6 + k] = g2;
        g[9 + k] = g3;
        g[k] = -__dadd_rn(__dadd_rn(g1, g2), g3);
    }
    vol = __dmul_rn(fabs(det), 0.16666666666666666);
}

/* P1 Laplacian: K[a][b] = vol * (ga . gb), dot = fma(z,z, fma(y,y, x*x)). */
struct OpLaplacian {
    static constexpr int D = 1;
    static constexpr int NS = DC_SLOT_G + 1;      /* gradients + vol */
    static constexpr int NPARAM = 0;
    __device__ __forceinline__ static void setup(const double (&x)[4][3], const double *, double *rec)
    {
        double vol;
        tet_gradients(x, rec, vol);
        rec[DC_SLOT_G] = vol;
    }
    template <typename R>
    __device__ __forceinline__ static void block(R rec, int a, int b, double (&out)[1])
    {
        double ax = rec[3 * a], ay = rec[3 * a + 1], az = rec[3 * a + 2];
        double bx = rec[3 * b], by = rec[3 * b + 1], bz = rec[3 * b + 2];
        double d = __fma_rn(az, bz, __fma_rn(ay, by, __dmul_rn(ax, bx)));
        out[0] = __dmul_rn(rec[DC_SLOT_G], d);
    }
};

/* Isotropic linear elasticity, 3x3 blocks:
 * K[a][b][i][j] = lV*(ga_i gb_j) + mV*(ga_j gb_i) + delta_ij * mV*(ga . gb),
 * p[i][j] = ga_i*gb_j, dot = (p00+p11)+p22, lV = lambda*vol, mV = mu*vol,
 * entry = fma(mV, p[j][i], lV*p[i][j]), diagonal: fma(mV, dot, entry).
 * params = {lambda, mu}. */
struct OpElasticity {
    static constexpr int D = 3;
    static constexpr int NS = DC_SLOT_G + 2;      /* gradients + lV + mV */
    static constexpr int NPARAM = 2;
    __device__ __forceinline__ static void setup(const double (&x)[4][3], const double *params, double *rec)
    {
        double vol;
        tet_gradients(x, rec, vol);
        rec[DC_SLOT_G] = __dmul_rn(params[0], vol);
        rec[DC_SLOT_G + 1] = __dmul_rn(params[1], vol);
    }
    template <typename R>
    __device__ __forceinline__ static void block(R rec, int a, int b, double (&out)[9])
    {
        double ga[3] = {rec[3 * a], rec[3 * a + 1], rec[3 * a + 2]};
        double gb[3] = {rec[3 * b], rec[3 * b + 1], rec[3 * b + 2]};
        double lV = rec[DC_SLOT_G], mV = rec[DC_SLOT_G + 1];
        double p[3][3];
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) p[i][j] = __dmul_rn(ga[i], gb[j]);
        double dot = __dadd_rn(__dadd_rn(p[0][0], p[1][1]), p[2][2]);
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) {
                double v = __fma_rn(mV, p[j][i], __dmul_rn(lV, p[i][j]));
                if (i == j) v = __fma_rn(mV, dot, v);
                out[i * 3 + j] = v;
            }
    }
};

}  // namespace dcdev
#endif
