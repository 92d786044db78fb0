/* dc_operators.cuh -- device-side element operators for the assembly kernels.
 *
 * DC-lib itself contains no element kernel: the reference hands each leaf to a
 * HOST callback (src/tree_traversal.cc:51,56,61) that computes the element matrix
 * and `+=`s it into nodeToNodeValue.  Host function pointers cannot run on the
 * GPU, so the callbacks are replaced by device operators with this contract:
 *
 *   setup(x[4][3], params, rec)         once per element: whatever the operator wants
 *                                       to keep for its 16 blocks, NS doubles;
 *   accumulate(rec, params, 4*a+b, acc) acc[D*D] += K_e[a][b], the `+=` of the host
 *                                       callback, in the operator's own fixed sequence;
 *   SYMMETRIC                           K_e[b][a] == transpose(K_e[a][b]) bit for bit and
 *                                       accumulate keeps that, so the symmetric schedule
 *                                       may compute (a,b) once and store the transpose.
 *
 * Every product / fma / sum is written with an explicit rounding intrinsic so
 * that nvcc cannot contract or reassociate: the sequence is the operator's
 * numeric definition and deterministic-mode results are compared bit for bit
 * against a CPU evaluation of the same sequence.
 */
#ifndef DC_OPERATORS_CUH
#define DC_OPERATORS_CUH

namespace dcdev {

__device__ __forceinline__ double cross_comp(double p, double q, double r, double s)
{
    return __fma_rn(p, q, -__dmul_rn(r, s));
}

/* Gradients of the P1 basis functions (g[3*a+k]) and |det|/6.  a,b,c = edges from node 0,
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
    double inv = __drcp_rn(det);          /* correctly rounded 1/det */
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

/* P1 Laplacian: K[a][b] = vol * (ga . gb), dot = fma(z,z, fma(y,y, x*x)).
 * The element matrix is symmetric bit for bit, so the record is its 10 distinct entries. */
struct OpLaplacian {
    static constexpr int D = 1;
    static constexpr int NS = 10;
    static constexpr int NPARAM = 0;
    __device__ __forceinline__ static void setup(const double (&x)[4][3], const double *, double *rec)
    {
        double g[12], vol;
        tet_gradients(x, g, vol);
        int q = 0;
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int b = a; b < 4; b++) {
                double d = __fma_rn(g[3 * a + 2], g[3 * b + 2],
                                    __fma_rn(g[3 * a + 1], g[3 * b + 1], __dmul_rn(g[3 * a], g[3 * b])));
                rec[q++] = __dmul_rn(vol, d);
            }
    }
    static constexpr bool SYMMETRIC = true;
    /* ab = 4*a+b */
    template <typename R>
    __device__ __forceinline__ static void accumulate(R rec, const double *, uint32_t ab, double (&acc)[1])
    {
        const uint32_t q = (uint32_t)(0x9863875265413210ULL >> (4 * ab)) & 15u;
        acc[0] = __dadd_rn(acc[0], rec[q]);
    }
    /* diagonal block (a == b): same update */
    template <typename R>
    __device__ __forceinline__ static void accumulate_diag(R rec, const double *params, uint32_t a, double (&acc)[1])
    {
        accumulate(rec, params, 5u * a, acc);
    }
    __device__ __forceinline__ static void mirror(double (&)[1]) {}
    /* the same updates split in two for software pipelining: fetch = the shared-memory reads,
     * apply = the arithmetic */
    static constexpr int NOPS = 1, NOPS_DIAG = 1;
    template <typename R>
    __device__ __forceinline__ static void fetch(R rec, uint32_t ab, double (&ops)[1])
    {
        ops[0] = rec[(uint32_t)(0x9863875265413210ULL >> (4 * ab)) & 15u];
    }
    template <typename R>
    __device__ __forceinline__ static void fetch_diag(R rec, uint32_t a, double (&ops)[1]) { fetch(rec, 5u * a, ops); }
    __device__ __forceinline__ static void apply(const double (&ops)[1], const double *, double (&acc)[1])
    {
        acc[0] = __dadd_rn(acc[0], ops[0]);
    }
    __device__ __forceinline__ static void apply_diag(const double (&ops)[1], const double *p, double (&acc)[1])
    {
        apply(ops, p, acc);
    }
};

/* Isotropic linear elasticity, 3x3 blocks, volume folded symmetrically into the gradients:
 * h_a = g_a * sqrt(vol); p[i][j] = ha_i*hb_j; dot = (p00+p11)+p22; the scatter-add of block
 * (a,b) into its CSR entry v is, per component,
 *     v[i][j] = fma(lambda, p[i][j], v[i][j]);  v[i][j] = fma(mu, p[j][i], v[i][j]);
 *     and for i == j:  v[i][i] = fma(mu, dot, v[i][i]);
 * (= v += vol*(lambda ga_i gb_j + mu ga_j gb_i + delta_ij mu ga.gb), three fused updates).
 * Swapping a and b transposes p bit for bit, so block (b,a) receives exactly the transposed
 * updates.  params = {lambda, mu}.  The record is the 12 scaled gradient components. */
struct OpElasticity {
    static constexpr int D = 3;
    static constexpr int NS = 12;
    static constexpr int NPARAM = 2;
    __device__ __forceinline__ static void setup(const double (&x)[4][3], const double *, double *rec)
    {
        double vol;
        tet_gradients(x, rec, vol);
        const double s = __dsqrt_rn(vol);
#pragma unroll
        for (int i = 0; i < 12; i++) rec[i] = __dmul_rn(rec[i], s);
    }
    static constexpr bool SYMMETRIC = true;
    template <typename R>
    __device__ __forceinline__ static void accumulate(R rec, const double *params, uint32_t ab, double (&acc)[9])
    {
        const uint32_t a = ab >> 2, b = ab & 3u;
        const double ga[3] = {rec[3 * a], rec[3 * a + 1], rec[3 * a + 2]};
        const double gb[3] = {rec[3 * b], rec[3 * b + 1], rec[3 * b + 2]};
        const double lam = params[0], mu = params[1];
        double p[3][3];
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) p[i][j] = __dmul_rn(ga[i], gb[j]);
        const double dot = __dadd_rn(__dadd_rn(p[0][0], p[1][1]), p[2][2]);
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) {
                double v = __fma_rn(lam, p[i][j], acc[i * 3 + j]);
                v = __fma_rn(mu, p[j][i], v);
                if (i == j) v = __fma_rn(mu, dot, v);
                acc[i * 3 + j] = v;
            }
    }
    /* diagonal block (a == b): p is symmetric bit for bit (p[i][j] = h_i*h_j), so the lower
     * triangle repeats the upper one -- 6 products, 6 update chains, one gradient read.
     * acc must be (and stays) symmetric; only its upper triangle is updated here and
     * mirror() completes it before the block is stored. */
    template <typename R>
    __device__ __forceinline__ static void accumulate_diag(R rec, const double *params, uint32_t a, double (&acc)[9])
    {
        const double h[3] = {rec[3 * a], rec[3 * a + 1], rec[3 * a + 2]};
        const double lam = params[0], mu = params[1];
        double p[3][3];
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = i; j < 3; j++) p[i][j] = __dmul_rn(h[i], h[j]);
        const double dot = __dadd_rn(__dadd_rn(p[0][0], p[1][1]), p[2][2]);
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = i; j < 3; j++) {
                double v = __fma_rn(lam, p[i][j], acc[i * 3 + j]);
                v = __fma_rn(mu, p[i][j], v);
                if (i == j) v = __fma_rn(mu, dot, v);
                acc[i * 3 + j] = v;
            }
    }
    __device__ __forceinline__ static void mirror(double (&acc)[9])
    {
        acc[3] = acc[1];
        acc[6] = acc[2];
        acc[7] = acc[5];
    }
    /* the same updates split in two for software pipelining: fetch = the shared-memory reads,
     * apply = the arithmetic */
    static constexpr int NOPS = 6, NOPS_DIAG = 3;
    template <typename R>
    __device__ __forceinline__ static void fetch(R rec, uint32_t ab, double (&ops)[6])
    {
        const uint32_t a = ab >> 2, b = ab & 3u;
        ops[0] = rec[3 * a]; ops[1] = rec[3 * a + 1]; ops[2] = rec[3 * a + 2];
        ops[3] = rec[3 * b]; ops[4] = rec[3 * b + 1]; ops[5] = rec[3 * b + 2];
    }
    template <typename R>
    __device__ __forceinline__ static void fetch_diag(R rec, uint32_t a, double (&ops)[3])
    {
        ops[0] = rec[3 * a]; ops[1] = rec[3 * a + 1]; ops[2] = rec[3 * a + 2];
    }
    __device__ __forceinline__ static void apply(const double (&ops)[6], const double *params, double (&acc)[9])
    {
        const double lam = params[0], mu = params[1];
        double p[3][3];
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) p[i][j] = __dmul_rn(ops[i], ops[3 + j]);
        const double dot = __dadd_rn(__dadd_rn(p[0][0], p[1][1]), p[2][2]);
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) {
                double v = __fma_rn(lam, p[i][j], acc[i * 3 + j]);
                v = __fma_rn(mu, p[j][i], v);
                if (i == j) v = __fma_rn(mu, dot, v);
                acc[i * 3 + j] = v;
            }
    }
    __device__ __forceinline__ static void apply_diag(const double (&h)[3], const double *params, double (&acc)[9])
    {
        const double lam = params[0], mu = params[1];
        double p[3][3];
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = i; j < 3; j++) p[i][j] = __dmul_rn(h[i], h[j]);
        const double dot = __dadd_rn(__dadd_rn(p[0][0], p[1][1]), p[2][2]);
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = i; j < 3; j++) {
                double v = __fma_rn(lam, p[i][j], acc[i * 3 + j]);
                v = __fma_rn(mu, p[i][j], v);
                if (i == j) v = __fma_rn(mu, dot, v);
                acc[i * 3 + j] = v;
            }
    }
};

}  // namespace dcdev
#endif
