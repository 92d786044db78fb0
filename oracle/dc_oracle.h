/* dc_oracle.h -- CPU ORACLE for the DC-lib assembly path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library; nothing under dc-lib_b200/ includes,
 * links or calls it.  See dc_oracle.c for what is restated and how it is pinned.
 */
#ifndef DC_ORACLE_H
#define DC_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Same layout as the reference's DCargs_t in BULK_SYNCHRONOUS builds
 * (include/DC.h:39-48) so the callbacks below can be handed to the reference's
 * own DC_tree_traversal (tree_traversal.cc:106-117). */
typedef struct {
    int firstElem, lastElem, firstNode, lastNode, firstEdge, lastEdge, isSep;
} dco_args_t;

/* "userArgs" of the assembly callbacks: the user-side data the reference never
 * sees (README.md:120-140). */
typedef struct {
    const int    *elemToNode;   /* [nbElem][4], 1-based, renumbered, tree order */
    const double *coord;        /* [nbNodes][3], permuted */
    const int    *row;          /* [nbNodes+1] 0-based edge offsets */
    const int    *col;          /* [nnz] node of each edge, numbering base colBase */
    double       *values;       /* [nnz][opDim*opDim] */
    int    colBase;
    int    opDim;               /* 1 = P1 Laplacian, 3 = isotropic elasticity */
    double lambda, mu;          /* Lame coefficients (opDim == 3) */
    int    resetInSeq;          /* 1: seq callback zeroes the leaf's edges (pure D&C) */
    int    opKind;              /* opDim == 1 only: 0 Laplacian, 1 mass, 2 advection  */
    double w[3];                /* advection velocity                                 */
} dco_user_t;

/* flat copy of one tree record (IO.cc:102-117), children as indices (-1 = none) */
typedef struct {
    int firstElem, lastElem, lastSep, firstNode, lastNode, firstEdge, lastEdge, vecOffset;
    int isSep, isLeaf, left, right, sep;
} dco_node_t;

/* element kernels (the numeric contract shared with the device operators) */
void dco_tet_gradients(const double x[4][3], double g[4][3], double *vol);
void dco_laplacian_ke(const double x[4][3], double ke[16]);
void dco_mass_ke(const double x[4][3], double ke[16]);
void dco_advection_ke(const double x[4][3], const double w[3], double ke[16]);
void dco_elasticity_h(const double x[4][3], double h[4][3]);
void dco_elasticity_add(const double ha[3], const double hb[3], double lambda, double mu, double v[9]);
void dco_elasticity_ke(const double x[4][3], double lambda, double mu, double ke[16][9]);

/* leaf callbacks with the reference's signature void(*)(void*, DCargs_t*) */
void dco_seq_callback(void *user, void *dcargs);
void dco_vec_callback(void *user, void *dcargs);
void dco_range_callback(void *user, int firstElem, int lastElem);
void dco_assemble_range(const dco_user_t *u, int firstElem, int lastElem);

/* tree_traversal.cc:27-103 restated over a flat tree; split = DC_VEC call shape */
void dco_traverse(const dco_node_t *nodes, int root, const dco_user_t *u, int split);
/* secondary oracle: zero everything, then elements in ascending index */
void dco_assemble_serial(const dco_user_t *u, int nbElem, int64_t nnz);

/* IO.cc:80-99 / 24-77: parse a tree file; returns node count (nodes may be NULL to count) */
int64_t dco_parse_tree(const unsigned char *buf, int64_t len, int nbElem, int nbNodes,
                       int *nbMaxComm, int *elemPermOut, int *nodePermOut,
                       dco_node_t *nodes, int64_t maxNodes);
/* tree_creation.cc:176-239: edge intervals of the leaves */
void dco_finalize(dco_node_t *nodes, int64_t nbTreeNodes, const int *row);

/* permutations.cc restated literally (cycle chasing / nested scans) */
void dco_permute_double_2d(double *tab, const int *perm, int nbItem, int dimItem);
void dco_permute_int_2d(int *tab, const int *perm, int nbItem, int dimItem, int offset);
void dco_permute_int_1d(int *tab, const int *perm, int size);
void dco_renumber(int *tab, const int *perm, int size, int isFortran);
void dco_create_permutation(int *perm, const int *part, int size, int nbPart);

#ifdef __cplusplus
}
#endif
#endif
