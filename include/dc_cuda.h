/* dc_cuda.h -- thin C ABI between the host library (C/C++/Fortran callers) and the
 * CUDA side of dc-b200.  Plain pointers, sizes and int return codes only; no C++
 * or torch types.  0 = success, negative = error (dc_cuda_last_error() explains).
 *
 * What each entry point replaces in the reference (all CPU there):
 *   dc_cuda_assemble          the per-step hot path: DC_tree_traversal walking the
 *                             tree and running the user's element callbacks
 *                             (src/tree_traversal.cc:27-117), including the CSR
 *                             reset over the DC_finalize_tree edge intervals
 *                             (src/tree_creation.cc:185-186, tree_traversal.cc:39-41)
 *   dc_cuda_set_mesh / plan   the user data reached through `userArgs`
 *                             (tree_traversal.cc:51,56,61) and the tree_t schedule
 *                             (include/DC.h:27-36)
 *   dc_cuda_permute_* /       DC_permute_double_2d_array / DC_permute_int_2d_array /
 *   dc_cuda_renumber          DC_permute_int_1d_array / DC_renumber_int_array
 *                             (src/permutations.cc:22-115)
 * Pointers named d_* are DEVICE pointers, h_* are HOST pointers.  `stream` is a
 * cudaStream_t passed as void* (NULL = legacy default stream).
 */
#ifndef DC_CUDA_H
#define DC_CUDA_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct dc_cuda_ctx dc_cuda_ctx;
struct dc_launch_args;
typedef int (*dc_operator_launcher)(struct dc_launch_args *args, const double *h_params, int nbParams,
                                    double *d_values, void *stream);

struct dc_plan_s;   /* dc_plan_t, see dc-lib_b200/csrc/dc_plan.h */

enum { DC_OP_P1_TET_LAPLACIAN = 0, DC_OP_P1_TET_ELASTICITY = 1 };
enum {
    DC_MODE_DETERMINISTIC = 0,  /* work units, every edge summed in ascending element order */
    DC_MODE_ATOMIC        = 1,  /* reset + one thread per element + red.global.add.f64      */
    DC_MODE_LIST          = 2   /* one thread per edge over explicit contribution lists     */
};

const char *dc_cuda_last_error(void);
int  dc_cuda_device_count(void);

int  dc_cuda_create(dc_cuda_ctx **ctx, int device);
void dc_cuda_destroy(dc_cuda_ctx *ctx);

/* Mesh + CSR pattern, copied from host memory.  h_conn: [nbElem][4] 1-based,
 * renumbered, tree order; h_coord: [nbNodes][3]; h_row: [nbNodes+1]; h_col: [nnz]. */
int  dc_cuda_set_mesh(dc_cuda_ctx *ctx, const int *h_conn, const double *h_coord,
                      const int *h_row, const int *h_col, int nbElem, int nbNodes, int colBase);
/* Same, adopting arrays that already live in device memory (not copied, not freed).
 * d_coord must have room for 2 nodes (48 bytes) beyond nbNodes: the kernel copies a unit's
 * own coordinates as 48-byte granules. */
int  dc_cuda_set_mesh_device(dc_cuda_ctx *ctx, const int *d_conn, const double *d_coord,
                             const int *d_row, const int *d_col, int nbElem, int nbNodes,
                             int64_t nnz, int colBase);
/* New node coordinates for the next assembly (host -> device, asynchronous). */
int  dc_cuda_update_coords(dc_cuda_ctx *ctx, const double *h_coord, void *stream);

/* Schedule built by dc_plan_build() (host); copied to the device. */
int  dc_cuda_upload_plan(dc_cuda_ctx *ctx, const struct dc_plan_s *plan);
/* Same, adopting schedule arrays that already live in device memory (e.g. received from
 * another rank over NCCL); not copied, not freed. */
int  dc_cuda_set_plan_device(dc_cuda_ctx *ctx, const void *d_units, const void *d_tasks,
                             const void *d_items, const void *d_slotNodes, const int *d_haloNodes,
                             const int *d_diagRel, const int *d_tgt, int64_t nbUnits, int maxSlots,
                             int maxHalo, int maxNodes, int maxTasks, int maxItems, int maxTgt,
                             int symmetric);
/* ... and its fallback lists for the rows that fit no unit (dc_plan_t.big: edge, ptr, item and, for the
 * symmetric schedule, edgeT), adopted the same way. */
int  dc_cuda_set_plan_big_device(dc_cuda_ctx *ctx, const int64_t *d_bigEdge, const int64_t *d_bigPtr,
                                 const uint32_t *d_bigItem, const int64_t *d_bigEdgeT, int64_t nbBigEdges);
/* Explicit contribution lists for DC_MODE_LIST (all edges). */
int  dc_cuda_upload_lists(dc_cuda_ctx *ctx, const int64_t *h_edge, const int64_t *h_ptr,
                          const uint32_t *h_item, int64_t nbEdges, int64_t nbItems);

/* One assembly: d_values[nnz * dim*dim] is reset and filled.  params: operator
 * coefficients on the host ({} for the Laplacian, {lambda, mu} for elasticity). */
int  dc_cuda_assemble(dc_cuda_ctx *ctx, int op, const double *h_params, int nbParams, int mode,
                      double *d_values, void *stream);
/* Same with a user-defined device operator: `launcher` is the function a user's .cu file
 * defines with DC_DEVICE_OPERATOR(name, Functor) (dc-lib_b200/csrc/dc_device_operator.cuh) --
 * the device-side replacement of the reference's host function pointers. */
int  dc_cuda_assemble_with(dc_cuda_ctx *ctx, dc_operator_launcher launcher, const double *h_params,
                           int nbParams, int mode, double *d_values, void *stream);
/* Convenience: values owned by the context. */
int  dc_cuda_values(dc_cuda_ctx *ctx, int dim, double **d_values);
int  dc_cuda_download_values(dc_cuda_ctx *ctx, int dim, double *h_values, void *stream);

/* Device -> host copy of d_values[first, first+count) (asynchronous; for streaming a matrix
 * larger than the host's pinned buffer). */
int  dc_cuda_download_range(const double *d_values, int64_t first, int64_t count, double *h_buf,
                            void *stream);

/* Permutation family on device arrays: d_out[perm[i]] = d_in[i] (rows of dimItem). */
int  dc_cuda_permute_double_2d(const double *d_in, double *d_out, const int *d_perm,
                               int64_t nbItem, int dimItem, void *stream);
int  dc_cuda_permute_int_2d(const int *d_in, int *d_out, const int *d_perm, int64_t nbItem,
                            int dimItem, void *stream);
/* The same permutation told by destination: d_out[d] = d_in[inv[d]] with inv the inverse of perm
 * (new position -> old position).  Whole sectors are written, only the reads are scattered: the fast
 * form on B200, where a partial sector write costs a DRAM fetch (profiles/README.md). */
int  dc_cuda_gather_double_2d(const double *d_in, double *d_out, const int *d_inv, int64_t nbItem,
                              int dimItem, void *stream);
int  dc_cuda_gather_int_2d(const int *d_in, int *d_out, const int *d_inv, int64_t nbItem, int dimItem,
                           void *stream);
int  dc_cuda_renumber(int *d_tab, const int *d_perm, int64_t size, int isFortran, void *stream);
/* Zero d_values[first*perEdge, (last+1)*perEdge) -- the reset of one edge interval. */
int  dc_cuda_reset_edges(double *d_values, int64_t firstEdge, int64_t lastEdge, int perEdge,
                         void *stream);

/* Interface exchange between the GPUs of one node.  The reference leaves the halo exchange to the
 * user's MPI/GASPI code (userCommFct, src/tree_traversal.cc:65-73) and prepares, per neighbour domain,
 * the interface nodes and their place in the neighbour's receive buffer (intfIndex / intfNodes /
 * intfDstOffsets, include/DC.h:51-53,141-143; src/tree_creation.cc:48-169).  Here the same lists exist at
 * CSR-edge granularity and the transport is peer memory over NVLink (csrc/dc_cuda_exchange.cu): per
 * neighbour, the local CSR edges both ranks hold, in an order both sides agree on (dc_interface_edges
 * below derives them from the reference's node lists).
 *   create   edges of neighbour i = edges[edgePtr[i] .. edgePtr[i+1]), neighbourRank ascending;
 *            _device: the concatenated list already lives on the device (not copied, not freed)
 *   export   128-byte handle of MY receive buffer for neighbour i -- hand it to that rank by any
 *            means (MPI, torch.distributed, a file) ...
 *   import   ... which imports it as the buffer it pushes into (CUDA IPC between processes, the
 *            plain pointer + peer access inside one process)
 *   pack     gather the partial sums of the shared edges into the neighbours' buffers and raise
 *            their flags; unpack: wait for the own flags and add the buffers (neighbours in the
 *            order of creation: a fixed summation order); run = pack + unpack.
 * All calls are asynchronous on `stream`; every rank calls pack/unpack once per step. */
#define DC_EXCHANGE_HANDLE_BYTES 128
typedef struct dc_cuda_exchange dc_cuda_exchange;
int  dc_cuda_exchange_create(dc_cuda_exchange **x, int device, int perEdge, int nbNeighbours,
                             const int *neighbourRank, const int64_t *edgePtr, const int64_t *h_edges);
int  dc_cuda_exchange_create_device(dc_cuda_exchange **x, int device, int perEdge, int nbNeighbours,
                                    const int *neighbourRank, const int64_t *edgePtr, const int64_t *d_edges);
int  dc_cuda_exchange_neighbours(const dc_cuda_exchange *x);
int  dc_cuda_exchange_export(dc_cuda_exchange *x, int nbIndex, void *handle);
int  dc_cuda_exchange_import(dc_cuda_exchange *x, int nbIndex, const void *handle);
int  dc_cuda_exchange_pack(dc_cuda_exchange *x, const double *d_values, void *stream);
int  dc_cuda_exchange_unpack(dc_cuda_exchange *x, double *d_values, void *stream);
int  dc_cuda_exchange_run(dc_cuda_exchange *x, double *d_values, void *stream);
int64_t dc_cuda_exchange_bytes(const dc_cuda_exchange *x);       /* sent + received per step */
int  dc_cuda_exchange_set_blocks(dc_cuda_exchange *x, int maxBlocks);   /* CTAs per pack / unpack kernel (default 64) */
void dc_cuda_exchange_destroy(dc_cuda_exchange *x);
const char *dc_cuda_exchange_last_error(void);

/* One assembly step of a rank of a multi-GPU run: the units that own interface rows first
 * (dc_cuda_set_front_units), pack on a side stream as soon as they are done, the other units with a few
 * CTA slots left free so that the pack / unpack kernels find room beside the persistent grid
 * ("exchange_reserve" option, default 4), unpack, join: the exchange hides behind the interior units
 * (the reference's MULTITHREADED_COMM idea, src/tree_creation.cc:48-169, at unit granularity). */
int  dc_cuda_set_front_units(dc_cuda_ctx *ctx, int64_t nbFrontUnits);
/* the pieces of that step, for callers that order them themselves: units [unitFirst, +unitCount) of the
 * schedule, reserveCtas CTA slots of the persistent grid left free, fallback rows skipped on request */
int  dc_cuda_assemble_units(dc_cuda_ctx *ctx, int op, const double *h_params, int nbParams, double *d_values,
                            void *stream, int64_t unitFirst, int64_t unitCount, int reserveCtas, int skipBigRows);
int  dc_cuda_assemble_exchange(dc_cuda_ctx *ctx, dc_cuda_exchange *x, int op, const double *h_params,
                               int nbParams, double *d_values, void *stream);
int  dc_cuda_assemble_exchange_with(dc_cuda_ctx *ctx, dc_cuda_exchange *x, dc_operator_launcher launcher,
                                    const double *h_params, int nbParams, double *d_values, void *stream);

/* Lower-level pieces (kept for callers that bring their own transport, e.g. NCCL or MPI buffers):
 * pack gathers the partial sums of the listed CSR edges into a contiguous send buffer,
 * unpack_add adds a received buffer (local, or a peer GPU's buffer mapped over NVLink). */
int  dc_cuda_pack_edges(const double *d_values, const int64_t *d_edge, int64_t nbEdges, int perEdge,
                        double *d_buf, void *stream);
int  dc_cuda_unpack_add_edges(double *d_values, const int64_t *d_edge, int64_t nbEdges, int perEdge,
                              const double *d_buf, void *stream);

/* The CSR edges two domains share, from the reference's interface layout: `nodes` = the interface
 * nodes shared with ONE neighbour (1-based local ids, in the order both sides list them -- the
 * reference's intfNodes slice intfIndex[i] .. intfIndex[i+1]).  Writes, ordered by (position of the
 * row node, position of the column node) in that list, the local edge of every pair of listed nodes
 * that is an entry of the local pattern, and its key position(row) * nbListed + position(col); returns
 * the count (edges / keys may be NULL to size the arrays).  Both sides must exchange only the keys
 * they BOTH hold: on conforming interfaces the lists are equal; otherwise intersect the key lists. */
int64_t dc_interface_edges(const int *h_row, const int *h_col, int colBase, const int *nodes, int nbListed,
                           int64_t *edges, int64_t *keys);

/* Consumer of the assembled matrix (what the reference's user, Mini-FEM, does next;
 * README.md:32): y = A x for the D x D-block CSR values the assembly left on the device, with the
 * context's CSR pattern; x and y hold D doubles per node (dim = 1 or 3).  Fixed summation order
 * (ascending edge, then column component, one fma per entry). */
int  dc_cuda_spmv(dc_cuda_ctx *ctx, int dim, const double *d_values, const double *d_x, double *d_y, void *stream);
/* ... and its block-Jacobi preconditioner: d_invDiag[nbNodes][dim*dim] = inverse of the diagonal block of
 * every row (adjugate over determinant, fixed operation order).  *nbBad (host pointer, may be NULL: then
 * the call stays asynchronous) = rows without a diagonal entry or with a singular block (they get zeros). */
int  dc_cuda_block_jacobi(dc_cuda_ctx *ctx, int dim, const double *d_values, double *d_invDiag, int *nbBad,
                          void *stream);

/* CSR pattern of the renumbered mesh, built on the device.  In the reference this is the user's
 * step between DC_renumber_int_array and DC_finalize_tree (the library only ships the node ->
 * element helper DC_create_nodeToElem, src/coloring.cc:79-112); at 10^8 elements the host build
 * dominates the set-up.  d_conn: [nbElem][4] 1-based.  begin() fills d_row[nbNodes+1] (edge (i,j)
 * iff an element holds both nodes, diagonal included) and returns nnz; columns() then writes the
 * nnz column indices, ascending inside a row, shifted by colBase; node_to_elem() copies the
 * node -> element lists (d_ptr[nbNodes+1], d_val[4*nbElem] ascending per node; either may be NULL).
 * The arrays equal the host builder's bit for bit.  end() frees the work space. */
typedef struct dc_cuda_csr dc_cuda_csr;
int  dc_cuda_csr_begin(const int *d_conn, int nbElem, int nbNodes, int *d_row, int64_t *nnz,
                       dc_cuda_csr **handle, void *stream);
int  dc_cuda_csr_columns(dc_cuda_csr *handle, int *d_col, int colBase, void *stream);
int  dc_cuda_csr_node_to_elem(dc_cuda_csr *handle, int64_t *d_ptr, int *d_val, void *stream);
void dc_cuda_csr_end(dc_cuda_csr *handle);

/* The level-wise element split of DC_create_tree on the device ("GPU-assisted
 * create_elem_part / 3-way stable partition / permute per level; METIS stays host").  At every node of
 * the tree the reference classes the node's elements by the partition of their nodes -- all left of the
 * pivot part, all right, straddling = separator (create_elem_part, src/tree_creation.cc:352-375) --
 * orders them stably by class (DC_create_permutation, src/permutations.cc:169-180) and moves the
 * connectivity rows (DC_permute_int_2d_array, :51-80; merge into elemPerm, :118-130).  Here one LEVEL
 * of the tree proper is one pass over all elements on the device; the caller keeps the recursion's
 * bookkeeping (tree nodes, part and node ranges) and builds the separator subtrees on the host.
 *   begin     h_conn [nbElem][dimElem] 0-BASED node ids in the current element order, h_nodePart[nbNodes]
 *             the node partition (METIS or coordinate bisection); every element starts in segment 0
 *   classify  segments of the level = disjoint position ranges [h_segFirst[s], h_segEnd[s]) with pivot part
 *             h_pivot[s]; returns h_totals[3 s + {0, 1, 2}] = elements left / right / straddling
 *   scatter   moves rows (stable inside a class: left, right, separator) and numbers the elements for the
 *             next level: h_child[3 s + k] = next segment of class k of segment s, -1 = not split further
 *   end       copies rows and the position -> original element map back (either may be NULL), frees
 * Rows and map equal the host builder's byte for byte (tests/test_gpu_tree_device.py). */
typedef struct dc_cuda_tree dc_cuda_tree;
int  dc_cuda_tree_begin(dc_cuda_tree **handle, int device, const int *h_conn, int nbElem, int dimElem,
                        const int *h_nodePart, int nbNodes);
int  dc_cuda_tree_classify(dc_cuda_tree *handle, int nbSeg, const int *h_segFirst, const int *h_segEnd,
                           const int *h_pivot, int *h_totals);
int  dc_cuda_tree_scatter(dc_cuda_tree *handle, const int *h_child);
int  dc_cuda_tree_end(dc_cuda_tree *handle, int *h_conn, int *h_posToOrig, int64_t *launches);
const char *dc_cuda_tree_last_error(void);

/* Tuning knobs: "unit_threads" = CTA size of the unit kernel (128, 256 or 512; default 256);
 * "persistent" = 1 (default): SMs x resident CTAs stride over the units and prefetch the next
 * unit's inputs, 0: one CTA per unit; "exchange_reserve" = CTA slots dc_cuda_assemble_exchange leaves
 * free beside the interior units for the pack / unpack kernels (default -1: about 8 slots per byte sent per interior element,
 * between 4 -- long interiors, the exchange is hidden completely -- and 48 -- one mesh cut over many
 * GPUs); "exchange_pack_early" = 1: the pack runs on the whole GPU between the interface
 * units and the interior launch instead of beside it (default 0; measured slower). */
int  dc_cuda_set_option(dc_cuda_ctx *ctx, const char *name, int value);

/* Number of kernels this context has launched (for honest launch accounting). */
int64_t dc_cuda_launch_count(const dc_cuda_ctx *ctx);
int  dc_cuda_sync(void *stream);

#ifdef __cplusplus
}
#endif
#endif
