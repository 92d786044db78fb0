/* dc_traversal.cc -- host-callback traversal of the D&C tree.
 *
 * Compatibility entry points for programs that hand DC-lib HOST function
 * pointers: DC_tree_traversal (reference: tree_traversal.cc:27-117) and the
 * README-generation DC_assembly (README.md:120-140).  Scheduling contract, as in
 * the reference: a leaf runs its callback(s); an inner node runs its left and
 * right subtrees concurrently, waits, then runs its separator subtree.  Host
 * callbacks cannot execute on the GPU -- the accelerated path is
 * DC_assembly_device() (dc_device.cc), which never comes through here.
 */
#include <cstdio>
#include <cstring>
#include <omp.h>
#include "dc_internal.h"

namespace {

struct LeafVisitor {
    void (*seq)(void *, DCargs_t *);
    void (*vec)(void *, DCargs_t *);
    void (*comm)(void *, DCcommArgs_t *);
    void *userArgs, *userCommArgs;
    bool split;   /* colouring variant: vec part then seq part */
};

void visit(const LeafVisitor &v, tree_t *t)
{
    if (t->left == nullptr && t->right == nullptr) {
        DCargs_t a;
        memset(&a, 0, sizeof a);
        a.firstNode = t->firstNode;  a.lastNode = t->lastNode;
        a.firstEdge = t->firstEdge;  a.lastEdge = t->lastEdge;
        a.isSep = t->isSep;
        if (v.split) {
            a.firstElem = t->firstElem;      a.lastElem = t->vecOffset;
            v.vec(v.userArgs, &a);
            a.firstElem = t->vecOffset + 1;  a.lastElem = t->lastElem;
            v.seq(v.userArgs, &a);
        } else {
            a.firstElem = t->firstElem;      a.lastElem = t->lastElem;
            v.seq(v.userArgs, &a);
        }
        return;
    }
    #pragma omp task default(shared) if (t->lastElem - t->firstElem > 256)
    visit(v, t->right);
    #pragma omp task default(shared) if (t->lastElem - t->firstElem > 256)
    visit(v, t->left);
    #pragma omp taskwait
    if (t->sep) visit(v, t->sep);
}

void run(const LeafVisitor &v)
{
    if (!treeHead) dc::fatal("Error: D&C tree traversal called without a D&C tree.");
    int nt = dc::options().nbThreads > 0 ? dc::options().nbThreads : omp_get_max_threads();
    #pragma omp parallel num_threads(nt)
    #pragma omp single nowait
    visit(v, treeHead);
}

/* adapter for the README-generation callbacks fn(userArgs, firstElem, lastElem) */
struct AssemblyCtx {
    void (*seq)(void *, int, int);
    void (*vec)(void *, int, int);
    void *userArgs;
    double *values;
    int64_t perEdge;
    bool split;
};

void assembly_reset(AssemblyCtx *c, DCargs_t *a)
{
    if (a->isSep || !c->values || a->lastEdge < a->firstEdge) return;
    memset(c->values + (int64_t)a->firstEdge * c->perEdge, 0,
           sizeof(double) * (size_t)((int64_t)(a->lastEdge - a->firstEdge + 1) * c->perEdge));
}
void assembly_vec(void *p, DCargs_t *a)
{
    AssemblyCtx *c = (AssemblyCtx *)p;
    assembly_reset(c, a);                      /* the vec part runs first on a leaf */
    if (a->lastElem >= a->firstElem) c->vec(c->userArgs, a->firstElem, a->lastElem);
}
void assembly_seq(void *p, DCargs_t *a)
{
    AssemblyCtx *c = (AssemblyCtx *)p;
    if (!c->split) assembly_reset(c, a);
    if (a->lastElem >= a->firstElem) c->seq(c->userArgs, a->firstElem, a->lastElem);
}

}  // namespace

void DC_tree_traversal (void (*userSeqFct)(void *, DCargs_t *), void (*userVecFct)(void *, DCargs_t *),
                        void (*userCommFct)(void *, DCcommArgs_t *), void *userArgs,
                        void *userCommArgs)
{
    /* userCommFct is only called by the reference's MULTITHREADED_COMM variant (tree_traversal.cc:65-73);
     * like its BULK_SYNCHRONOUS variant this traversal never calls it -- said once, not silently */
    if (userCommFct) {
        static bool told = false;
        if (!told) {
            told = true;
            fprintf(stderr, "DC_tree_traversal: userCommFct is not called (bulk-synchronous library): exchange the interface "
                            "after the traversal.\n");
        }
    }
    LeafVisitor v{userSeqFct, userVecFct, userCommFct, userArgs, userCommArgs,
                  dc::options().vecSize > 0 && userVecFct != nullptr};
    run(v);
}

void DC_assembly (void (*userSeqFct)(void *, int, int), void (*userVecFct)(void *, int, int),
                  void *userArgs, double *nodeToNodeValue, int operatorDim)
{
    AssemblyCtx c{userSeqFct, userVecFct, userArgs, nodeToNodeValue,
                  (int64_t)operatorDim * operatorDim,
                  dc::options().vecSize > 0 && userVecFct != nullptr};
    LeafVisitor v{assembly_seq, assembly_vec, nullptr, &c, nullptr, c.split};
    run(v);
}
