/* dc_oracle.c -- CPU ORACLE for the DC-lib assembly path.  TEST INFRASTRUCTURE ONLY
 * (never shipped, never on the product path; see dc_oracle.h).
 *
 * What is restated here, in plain C, and the reference lines it follows:
 *   - the leaf call shape and (left || right) -> sep order of the traversal
 *     (src/tree_traversal.cc:27-103), serially;
 *   - the user-callback contract of the assembly (call sites tree_traversal.cc:51,
 *     56,61; README.md:120-140): zero the leaf's edge interval unless isSep, then
 *     for every element gather 4 nodes, build the element matrix and `+=` its 16
 *     blocks into the CSR row of the first node at the column of the second;
 *   - the edge intervals of DC_finalize_tree (src/tree_creation.cc:176-239);
 *   - the tree file format (src/IO.cc:24-158);
 *   - the permutation family, literally (src/permutations.cc:22-180).
 * The element kernels themselves (P1 Laplacian, isotropic elasticity) are NOT in
 * the reference -- they are user code there.  They are defined here with an
 * explicit operation order (every product, fma and sum spelled out, compiled with
 * -ffp-contract=off) and the device operators use the same sequence, so
 * "deterministic mode" can be compared bit for bit.
 *
 * Pinning: the reference has no tests, golden vectors or fixtures.
 * The oracle is pinned against the reference ITSELF, compiled unmodified into
 * oracle/_ref/ (oracle/Makefile): tests/test_oracle_vs_reference.py hands these
 * callbacks to the reference's dc_tree_traversal_ and checks that the restated
 * traversal, the serial loop and the reference's OpenMP run agree bitwise, and
 * that the permutation / finalize / tree-file restatements reproduce the
 * reference's outputs; tests/golden/ holds vectors generated that way
 * (oracle/gen_golden.py) for the GPU box, where /root/reference does not exist.
 * Tree creation, partitioning and colouring are not restated: they are checked
 * directly against oracle/_ref and the golden tree files.
 */
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include "dc_oracle.h"

/* ------------------------------------------------------------------ kernels */

/* Gradients of the four P1 basis functions and the volume of one tetrahedron.
 * a,b,c = edges from node 0; g1 = (b x c)/det, g2 = (c x a)/det, g3 = (a x b)/det,
 * g0 = -((g1+g2)+g3); det = a.(b x c); vol = |det| * (1/6). */
static inline double cross_comp(double p, double q, double r, double s)
{
    double t = r * s;                 /* p*q - r*s with one rounding on r*s and one fma */
    return fma(p, q, -t);
}

void dco_tet_gradients(const double x[4][3], double g[4][3], double *vol)
{
    double a[3], b[3], c[3], n1[3], n2[3], n3[3];
    for (int k = 0; k < 3; k++) {
        a[k] = x[1][k] - x[0][k];
        b[k] = x[2][k] - x[0][k];
        c[k] = x[3][k] - x[0][k];
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
    double det = fma(a[2], n1[2], fma(a[1], n1[1], a[0] * n1[0]));
    double inv = 1.0 / det;
    for (int k = 0; k < 3; k++) {
        g[1][k] = n1[k] * inv;
        g[2][k] = n2[k] * inv;
        g[3][k] = n3[k] * inv;
        g[0][k] = -((g[1][k] + g[2][k]) + g[3][k]);
    }
    *vol = fabs(det) * 0.16666666666666666;   /* 0x3FC5555555555555 */
}

/* ke[a*4+b] = vol * (ga . gb), dot = fma(z,z, fma(y,y, x*x)) */
void dco_laplacian_ke(const double x[4][3], double ke[16])
{
    double g[4][3], vol;
    dco_tet_gradients(x, g, &vol);
    for (int a = 0; a < 4; a++)
        for (int b = 0; b < 4; b++) {
            double d = fma(g[a][2], g[b][2], fma(g[a][1], g[b][1], g[a][0] * g[b][0]));
            ke[a * 4 + b] = vol * d;
        }
}

/* Two more scalar operators, used to test the device-operator hook with user-written
 * operators (tests/custom_op/custom_ops.cu): a symmetric mass matrix and a non-symmetric
 * advection matrix.  ke[a*4+b]. */
void dco_mass_ke(const double x[4][3], double ke[16])
{
    double g[4][3], vol;
    dco_tet_gradients(x, g, &vol);
    double off = vol * 0.05, diag = vol * 0.1;
    for (int a = 0; a < 4; a++)
        for (int b = 0; b < 4; b++) ke[a * 4 + b] = (a == b) ? diag : off;
}

void dco_advection_ke(const double x[4][3], const double w[3], double ke[16])
{
    double g[4][3], vol;
    dco_tet_gradients(x, g, &vol);
    double q = vol * 0.25;
    for (int b = 0; b < 4; b++) {
        double gw = fma(g[b][2], w[2], fma(g[b][1], w[1], g[b][0] * w[0]));
        double v = q * gw;
        for (int a = 0; a < 4; a++) ke[a * 4 + b] = v;
    }
}

/* Isotropic elasticity with the volume folded symmetrically into the gradients:
 * h_a = g_a * sqrt(vol);  p[i][j] = ha_i*hb_j;  dot = (p00+p11)+p22.  The scatter-add of
 * block (a,b) into its CSR entry v is three fused updates per component,
 *     v[i][j] = fma(lambda, p[i][j], v[i][j]);  v[i][j] = fma(mu, p[j][i], v[i][j]);
 *     i == j:  v[i][i] = fma(mu, dot, v[i][i]);
 * i.e. v += vol*(lambda ga_i gb_j + mu ga_j gb_i + delta_ij mu ga.gb).  dco_elasticity_h
 * computes the h_a; dco_elasticity_add performs the update of one block. */
void dco_elasticity_h(const double x[4][3], double h[4][3])
{
    double g[4][3], vol;
    dco_tet_gradients(x, g, &vol);
    double s = sqrt(vol);
    for (int a = 0; a < 4; a++)
        for (int k = 0; k < 3; k++) h[a][k] = g[a][k] * s;
}

void dco_elasticity_add(const double ha[3], const double hb[3], double lambda, double mu, double v[9])
{
    double p[3][3];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) p[i][j] = ha[i] * hb[j];
    double dot = (p[0][0] + p[1][1]) + p[2][2];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double t = fma(lambda, p[i][j], v[i * 3 + j]);
            t = fma(mu, p[j][i], t);
            if (i == j) t = fma(mu, dot, t);
            v[i * 3 + j] = t;
        }
}

/* the element matrix itself (what the updates add up to), for reference */
void dco_elasticity_ke(const double x[4][3], double lambda, double mu, double ke[16][9])
{
    double h[4][3];
    dco_elasticity_h(x, h);
    for (int a = 0; a < 4; a++)
        for (int b = 0; b < 4; b++) {
            for (int k = 0; k < 9; k++) ke[a * 4 + b][k] = 0.0;
            dco_elasticity_add(h[a], h[b], lambda, mu, ke[a * 4 + b]);
        }
}

/* ---------------------------------------------------------------- callbacks */

static inline int64_t find_edge(const dco_user_t *u, int rowNode, int colNode)
{
    int target = colNode + u->colBase;
    for (int k = u->row[rowNode]; k < u->row[rowNode + 1]; k++)
        if (u->col[k] == target) return k;
    abort();   /* the CSR pattern must contain every element edge */
}

void dco_assemble_range(const dco_user_t *u, int firstElem, int lastElem)
{
    const int d2 = u->opDim * u->opDim;
    for (int e = firstElem; e <= lastElem; e++) {
        int n[4];
        double x[4][3];
        for (int a = 0; a < 4; a++) {
            n[a] = u->elemToNode[(int64_t)e * 4 + a] - 1;
            for (int k = 0; k < 3; k++) x[a][k] = u->coord[(int64_t)n[a] * 3 + k];
        }
        if (u->opDim == 1) {
            double ke[16];
            if (u->opKind == 1) dco_mass_ke(x, ke);
            else if (u->opKind == 2) dco_advection_ke(x, u->w, ke);
            else dco_laplacian_ke(x, ke);
            for (int a = 0; a < 4; a++)
                for (int b = 0; b < 4; b++)
                    u->values[find_edge(u, n[a], n[b])] += ke[a * 4 + b];
        } else {
            double h[4][3];
            dco_elasticity_h(x, h);
            for (int a = 0; a < 4; a++)
                for (int b = 0; b < 4; b++)
                    dco_elasticity_add(h[a], h[b], u->lambda, u->mu, u->values + find_edge(u, n[a], n[b]) * d2);
        }
    }
}

static void reset_leaf(const dco_user_t *u, const dco_args_t *a)
{
    if (a->isSep || a->lastEdge < a->firstEdge) return;
    const int64_t d2 = (int64_t)u->opDim * u->opDim;
    memset(u->values + (int64_t)a->firstEdge * d2, 0,
           sizeof(double) * (size_t)((int64_t)(a->lastEdge - a->firstEdge + 1) * d2));
}

void dco_seq_callback(void *user, void *dcargs)
{
    const dco_user_t *u = (const dco_user_t *)user;
    const dco_args_t *a = (const dco_args_t *)dcargs;
    if (u->resetInSeq) reset_leaf(u, a);
    dco_assemble_range(u, a->firstElem, a->lastElem);
}

/* README-generation callback shape fn(userArgs, firstElem, lastElem) (README.md:122-124): the
 * library resets the leaf itself, the callback only assembles the range */
void dco_range_callback(void *user, int firstElem, int lastElem)
{
    dco_assemble_range((const dco_user_t *)user, firstElem, lastElem);
}

/* In DC_VEC builds the vec callback runs first on a leaf (tree_traversal.cc:47-56),
 * so the reset lives here and the seq callback must not repeat it. */
void dco_vec_callback(void *user, void *dcargs)
{
    const dco_user_t *u = (const dco_user_t *)user;
    const dco_args_t *a = (const dco_args_t *)dcargs;
    reset_leaf(u, a);
    dco_assemble_range(u, a->firstElem, a->lastElem);
}

/* ---------------------------------------------------------------- traversal */

void dco_traverse(const dco_node_t *nodes, int root, const dco_user_t *u, int split)
{
    const dco_node_t *t = &nodes[root];
    if (t->isLeaf) {
        dco_args_t a;
        a.firstNode = t->firstNode; a.lastNode = t->lastNode;
        a.firstEdge = t->firstEdge; a.lastEdge = t->lastEdge;
        a.isSep = t->isSep;
        dco_user_t v = *u;
        if (split) {
            v.resetInSeq = 0;
            a.firstElem = t->firstElem;     a.lastElem = t->vecOffset;
            dco_vec_callback(&v, &a);
            a.firstElem = t->vecOffset + 1; a.lastElem = t->lastElem;
            dco_seq_callback(&v, &a);
        } else {
            v.resetInSeq = 1;
            a.firstElem = t->firstElem;     a.lastElem = t->lastElem;
            dco_seq_callback(&v, &a);
        }
        return;
    }
    dco_traverse(nodes, t->right, u, split);   /* spawned in the reference */
    dco_traverse(nodes, t->left, u, split);
    /* sync */
    if (t->sep >= 0) dco_traverse(nodes, t->sep, u, split);
}

void dco_assemble_serial(const dco_user_t *u, int nbElem, int64_t nnz)
{
    memset(u->values, 0, sizeof(double) * (size_t)(nnz * u->opDim * u->opDim));
    dco_assemble_range(u, 0, nbElem - 1);
}

/* ---------------------------------------------------------------- tree file */

typedef struct { const unsigned char *p, *end; int bad; } rd_t;
static void rd(rd_t *r, void *dst, size_t n)
{
    if ((size_t)(r->end - r->p) < n) { r->bad = 1; memset(dst, 0, n); return; }
    memcpy(dst, r->p, n);
    r->p += n;
}

static int64_t parse_rec(rd_t *r, dco_node_t *nodes, int64_t *count, int64_t maxNodes)
{
    int64_t me = (*count)++;
    int rec[8];
    unsigned char isSep, isLeaf;
    rd(r, rec, sizeof rec);
    rd(r, &isSep, 1);
    rd(r, &isLeaf, 1);
    if (r->bad) return -1;
    dco_node_t t;
    t.firstElem = rec[0]; t.lastElem = rec[1]; t.lastSep = rec[2];
    t.firstNode = rec[3]; t.lastNode = rec[4];
    t.firstEdge = rec[5]; t.lastEdge = rec[6]; t.vecOffset = rec[7];
    t.isSep = isSep; t.isLeaf = isLeaf; t.left = t.right = t.sep = -1;
    if (isLeaf) {
        int owned, intf;
        rd(r, &owned, 4);
        rd(r, &intf, 4);
        if (owned > 0 || intf > 0) r->bad = 1;   /* MULTITHREADED_COMM payloads: out of scope */
    } else {
        t.left = (int)parse_rec(r, nodes, count, maxNodes);
        t.right = (int)parse_rec(r, nodes, count, maxNodes);
        if (t.lastSep > t.lastElem) t.sep = (int)parse_rec(r, nodes, count, maxNodes);
    }
    if (nodes && me < maxNodes) nodes[me] = t;
    return me;
}

int64_t dco_parse_tree(const unsigned char *buf, int64_t len, int nbElem, int nbNodes,
                       int *nbMaxComm, int *elemPermOut, int *nodePermOut,
                       dco_node_t *nodes, int64_t maxNodes)
{
    rd_t r = {buf, buf + len, 0};
    int comm;
    rd(&r, &comm, 4);
    if (nbMaxComm) *nbMaxComm = comm;
    if (elemPermOut) rd(&r, elemPermOut, 4 * (size_t)nbElem); else r.p += 4 * (size_t)nbElem;
    if (nodePermOut) rd(&r, nodePermOut, 4 * (size_t)nbNodes); else r.p += 4 * (size_t)nbNodes;
    int64_t count = 0;
    parse_rec(&r, nodes, &count, maxNodes);
    if (r.bad || r.p != r.end) return -1;
    return count;
}

void dco_finalize(dco_node_t *nodes, int64_t nbTreeNodes, const int *row)
{
    for (int64_t i = 0; i < nbTreeNodes; i++)
        if (nodes[i].isLeaf) {
            nodes[i].firstEdge = row[nodes[i].firstNode];
            nodes[i].lastEdge = row[nodes[i].lastNode + 1] - 1;
        }
}

/* ------------------------------------------------------------- permutations */

void dco_permute_double_2d(double *tab, const int *perm, int nbItem, int dimItem)
{
    char *done = (char *)calloc((size_t)(nbItem > 0 ? nbItem : 1), 1);
    double *carry = (double *)malloc(sizeof(double) * (size_t)dimItem);
    for (int i = 0; i < nbItem; i++) {
        if (done[i]) continue;
        int src = i;
        memcpy(carry, tab + (int64_t)i * dimItem, sizeof(double) * (size_t)dimItem);
        do {
            int dst = perm[src];
            for (int j = 0; j < dimItem; j++) {
                double keep = tab[(int64_t)dst * dimItem + j];
                tab[(int64_t)dst * dimItem + j] = carry[j];
                carry[j] = keep;
            }
            src = dst;
            done[src] = 1;
        } while (src != i);
    }
    free(carry);
    free(done);
}

void dco_permute_int_2d(int *tab, const int *perm, int nbItem, int dimItem, int offset)
{
    char *done = (char *)calloc((size_t)(nbItem > 0 ? nbItem : 1), 1);
    int *carry = (int *)malloc(sizeof(int) * (size_t)dimItem);
    int *base = tab + (int64_t)offset * dimItem;
    for (int i = 0; i < nbItem; i++) {
        if (done[i]) continue;
        int src = i;
        memcpy(carry, base + (int64_t)i * dimItem, sizeof(int) * (size_t)dimItem);
        do {
            int dst = perm[src];
            for (int j = 0; j < dimItem; j++) {
                int keep = base[(int64_t)dst * dimItem + j];
                base[(int64_t)dst * dimItem + j] = carry[j];
                carry[j] = keep;
            }
            src = dst;
            done[src] = 1;
        } while (src != i);
    }
    free(carry);
    free(done);
}

void dco_permute_int_1d(int *tab, const int *perm, int size)
{
    dco_permute_int_2d(tab, perm, size, 1, 0);
}

void dco_renumber(int *tab, const int *perm, int size, int isFortran)
{
    int base = isFortran ? 1 : 0;
    for (int i = 0; i < size; i++) tab[i] = perm[tab[i] - base] + base;
}

void dco_create_permutation(int *perm, const int *part, int size, int nbPart)
{
    int next = 0;
    for (int p = 0; p < nbPart; p++)
        for (int j = 0; j < size; j++)
            if (part[j] == p) perm[j] = next++;
}
