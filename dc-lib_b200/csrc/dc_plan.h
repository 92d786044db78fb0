/* dc_plan.h -- device schedule ("plan") of one assembly, shared by the host-side
 * plan builder (host/dc_plan.cc) and the CUDA side (csrc/dc_cuda_api.cu).
 *
 * The reference schedules the assembly by walking tree_t at run time and calling
 * a host callback per leaf (src/tree_traversal.cc:27-103).  On the GPU the same
 * tree is flattened once into WORK UNITS: a unit is a non-separator subtree of
 * the D&C tree (or a row sub-range of one leaf) small enough for one CTA.  It
 * owns the subtree's node interval, hence the CSR edge interval
 * [row[firstNode], row[lastNode+1]) that DC_finalize_tree computes for the reset
 * (src/tree_creation.cc:185-186), and writes every value of that interval
 * exactly once -- the reset is fused into that single write.  The elements a unit
 * needs are its own contiguous range [firstElem, lastSep] plus the "halo":
 * elements of ancestor separators that touch one of its nodes.  Each element gets a
 * SLOT (numbered so that the record reads of a half-warp spread over the banks); slotNodes gives its 4 nodes as indices into the unit's node table (the own
 * node interval, whose coordinates are one contiguous block of the coordinate array,
 * followed by the halo nodes), so the kernel stages all coordinates in shared memory
 * once instead of gathering them per element.
 *
 * Inside a unit every CSR edge is summed in ascending element index, which is the
 * order the reference's traversal produces (left < right < separator in element
 * index, executed (left || right) -> separator), so values are bit-identical.
 */
#ifndef DC_PLAN_H
#define DC_PLAN_H
#include <stdint.h>

#define DC_SLOT_LIMIT  4094         /* slots are 12-bit: item = slot<<4 | a<<2 | b; item rows are  */
                                    /* padded with (nSlots << 4): the kernel keeps an all-zero     */
                                    /* record in slot nSlots, so padding adds exact zeros          */
#define DC_TASK_EDGES  0            /* 32 consecutive edges of the unit (general schedule)   */
#define DC_TASK_DIAG   1            /* the diagonal edges of 32 consecutive nodes            */
#define DC_TASK_UPPER  2            /* symmetric schedule: 32 owned edges (diagonal or col > row) */
                                    /* listed in tgt; an upper-edge lane also writes the        */
                                    /* transposed block to (col,row)                            */
#define DC_TASK_OWNDIAG 3           /* symmetric schedule: 32 owned DIAGONAL edges (a == b in    */
                                    /* every item, no transposed target)                        */

typedef struct {
    int64_t edgeFirst;      /* first CSR edge of the unit                          */
    int64_t itemFirst;      /* first uint16 item of the unit                       */
    int64_t tgtFirst;       /* symmetric schedule: first (edge, transposed edge) pair */
    int64_t slotFirst;      /* first entry of slotNodes (4 uint16 per slot)        */
    int64_t hnodeFirst;     /* first entry of haloNodes                            */
    int64_t reserved_;
    int32_t edgeCount;
    int32_t nodeFirst, nodeCount;
    int32_t ownElemFirst, ownElemCount;   /* own elements [ownElemFirst, +ownElemCount)          */
    int32_t haloCount;                    /* halo elements; slots = ownElemCount + haloCount,   */
                                          /* numbered for the shared-memory banks (slotElems)    */
    int32_t taskFirst, taskCount;
    int32_t itemCount;                    /* uint16 items of the unit (multiple of 32) */
    int32_t tgtCount;                     /* owned edges of the unit (symmetric schedule) */
    int32_t hnodeCount;                   /* nodes used by the unit outside its own interval */
    int32_t pad_;
} dc_unit_t;                /* 96 bytes */

typedef struct {
    uint32_t itemRel;       /* items of this task start at unit.itemFirst+itemRel  */
    uint16_t iters;         /* item rows ([iter][lane], 32 lanes each)             */
    uint16_t kind;          /* DC_TASK_EDGES / DC_TASK_DIAG                        */
    uint32_t mask;          /* EDGES: lanes that are diagonal edges (skipped);     */
                            /* DIAG : lanes that hold a node                       */
    uint32_t first;         /* EDGES: edge offset in unit; DIAG: node offset;      */
                            /* UPPER: index of the first upper edge of the unit    */
} dc_task_t;                /* 16 bytes */

/* Fallback list for rows too large for a unit (and for generic operators):
 * explicit per-edge contribution lists, item = elem<<4 | a<<2 | b. */
typedef struct {
    int64_t  nbEdges;
    int64_t  nbItems;
    int64_t *edge;          /* [nbEdges]   CSR edge index                          */
    int64_t *ptr;           /* [nbEdges+1] offsets into item                       */
    uint32_t *item;         /* [nbItems]                                           */
    int64_t *edgeT;         /* [nbEdges] symmetric lists: edge that receives the   */
                            /* transposed block, -1 for none; NULL otherwise       */
} dc_contrib_list_t;

typedef struct dc_plan_s {
    int32_t  maxSlots;      /* largest slot count of any unit                      */
    int32_t  maxHalo;       /* largest halo-node list (entries, padded to 4)       */
    int32_t  maxNodes;      /* most own nodes in one unit                          */
    int32_t  maxTasks;      /* most tasks in one unit                              */
    int32_t  maxItems;      /* most uint16 items in one unit                       */
    int32_t  maxTgt;        /* most upper edges in one unit (symmetric schedule)   */
    int32_t  symmetric;     /* 1: UPPER tasks + tgt pairs, 0: EDGES tasks          */
    int64_t  nbTgt;         /* pairs in tgt                                        */
    int32_t *tgt;           /* [nbTgt][2] (edge of (row,col), edge of (col,row) or -1) */
    int64_t  nbUnits, nbTasks, nbItems, nbHalo;
    dc_unit_t *units;
    dc_task_t *tasks;
    uint16_t  *items;
    int32_t   *slotElems;   /* [nbSlotNodes/4] element of every slot, -1 for padding (host-side bookkeeping) */
    int64_t    nbSlotNodes; /* uint16 entries in slotNodes (4 per slot, units padded to 2 slots) */
    uint16_t  *slotNodes;   /* per slot: the element's 4 nodes as indices into the unit's node table: */
                            /* own nodes first (from nodeFirst & ~1), then haloNodes                */
    int64_t    nbHaloNodes;
    int32_t   *haloNodes;   /* node ids outside the unit's own interval            */
    int32_t   *diagRel;     /* [nbNodes] edge offset (in its unit) of node's diagonal, -1 if none */
    dc_contrib_list_t big;  /* rows that did not fit a unit                        */
    /* statistics */
    int64_t  nbSlotsTotal;  /* sum of slots over units (element setups performed)  */
    int64_t  nbItemsUseful; /* items that are not padding                          */
} dc_plan_t;

#ifdef __cplusplus
extern "C" {
#endif
/* Build the plan on the host.  conn: [nbElem][4] 1-based renumbered tree-order
 * connectivity; row/col: CSR pattern (col numbering base colBase, any order
 * within a row); leaves: nbLeaves x 8 ints {firstElem,lastElem,lastSep,firstNode,
 * lastNode,isSep,isLeaf,parent-free flag} in PRE-ORDER of the tree -- see
 * dc_plan_build_from_tree for the pointer-tree front end.  Returns 0 on success; -5 when rows
 * fall back to the per-edge lists (or dc_contrib_build is called) on a mesh of 2^28 or more
 * elements: list items hold the element index in 28 bits. */
int  dc_plan_build(dc_plan_t *plan, const int *conn, int nbElem, int nbNodes, const int *row,
                   const int *col, int colBase, const int *treeRec, int64_t nbTreeRec,
                   int maxSlots, int nbThreads, int symmetric);
int  dc_contrib_build(dc_contrib_list_t *out, const int *conn, int nbElem, int nbNodes,
                      const int *row, const int *col, int colBase, const int *rows,
                      int64_t nbRows, int symmetric);
/* Record layout the slot numbering of the following dc_plan_build calls is optimised for (shared-memory
 * banks of the accumulation reads): 0 = four 3-word node gradients in 13-word records (elasticity, default),
 * 1 = the 10 entries of a symmetric 4x4 element matrix in 11-word records (P1 Laplacian).  Any schedule is
 * valid for any operator; the model only decides how many bank conflicts the record reads have. */
void dc_plan_set_bank_model(int model);
/* Multi-GPU overlap: move the units that own a flagged row (nodeFlag[node] != 0: the interface
 * nodes of this rank) to the front of plan->units, keeping the order inside both groups; returns
 * their number.  Units are independent of one another, so any order is a valid schedule; with the
 * interface rows first their partial sums can travel while the other units are assembled. */
int64_t dc_plan_order_units(dc_plan_t *plan, const unsigned char *nodeFlag);
void dc_plan_free(dc_plan_t *plan);
void dc_contrib_free(dc_contrib_list_t *c);
#ifdef __cplusplus
}
#endif
#endif
