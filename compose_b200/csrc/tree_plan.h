// Host-side flattening of the caller's tree (reference: cedr_tree_caller.hpp:12-24,
// consumed by tree::analyze, cedr_tree.cpp:215-231) into the device-resident plan
// the sweep kernels run on.
//
// The reference schedules the whole tree level by level (level = height,
// cedr_tree.cpp:55-70) and launches one kernel per level. Here the tree is cut
// into BLOCKS: a block is a subtree with at most `max_block_leaves` leaves, small
// enough that one CTA sweeps it entirely in shared memory. Because the reference
// numbers leaves in DFS order (cedr_tree.cpp:85, :148-180), every subtree owns a
// contiguous range of local cell indices, so a block's leaf data is one contiguous
// segment per field. The roots of the blocks of one TIER are the leaves of the
// next tier, and so on until a tier consists of a single block holding the root.
// Node arithmetic is order-independent data flow (each node = f(its two kids)),
// so this schedule produces bit-identical results to the reference's.
#ifndef CEDR_B200_TREE_PLAN_H
#define CEDR_B200_TREE_PLAN_H

#include <cstdint>
#include <string>
#include <vector>

namespace cedr_b200 {

// Block-local topology, shared by all blocks of identical shape.
struct Shape {
  int nl = 0;    // leaves
  int ni = 0;    // internal nodes (nl - 1)
  int nlev = 0;  // number of internal levels (block-root height); 0 if nl == 1
  // Internal nodes are ordered by (height, DFS post-order). lvlptr has nlev+1
  // entries indexing that order; level l (1-based height) is
  // [lvlptr[l-1], lvlptr[l]).
  std::vector<int> lvlptr;
  // Kids as block-local node ids: 0..nl-1 are leaves (DFS order), nl+j is
  // internal node j.
  std::vector<int> kid0, kid1;
  // True if the shape is the recursive bisection n -> (n/2, n - n/2) of
  // cedr_tree.cpp:391-413 all the way down (enables the register fast path).
  bool bisection = false;
  // Offsets into the packed device arrays (filled by pack()).
  int dev_lvlptr_off = 0, dev_kid_off = 0;

  // Fast-path tables (fast_kernels.cuh), present iff `fast`: the shape is a perfect
  // binary tree down to depth 9 whose 512 depth-9 nodes are single leaves or pairs
  // of leaves (true for bisection shapes with 512 < nl <= 1024).
  bool fast = false;
  std::vector<unsigned short> dtab;  // [512] leaf offset | (is pair << 15)
  std::vector<unsigned short> ptab;  // depth-9 positions of the pairs, increasing
  std::vector<unsigned short> fpos;  // [ni] internal node j -> fast const position
  // Down-sweep work assignment chosen to avoid shared-memory bank conflicts (see
  // build_down_tables in tree_plan.cpp): perm[thread] = depth-7 node of a leaf thread;
  // pent = [start of warp 0..3's pair entries, end | entries], an entry being
  // leaf offset | d9x slot << 11 | pair's const rank << 20.
  std::vector<unsigned short> perm;
  std::vector<unsigned> pent;
  // Up-sweep: [0, 128) thread -> depth-7 node, a permutation WITHIN each warp (the levels
  // above go through warp shuffles); [128, 256) node -> thread.
  std::vector<unsigned short> perm_up;
  int dev_dtab_off = 0, dev_ptab_off = 0, dev_fpos_off = 0, dev_perm_off = 0,
    dev_pent_off = 0, dev_perm_up_off = 0;
};

struct Block {
  int leaf0 = 0;   // first leaf, as an index into this tier's leaf array
  int nl = 0;
  int shape = 0;   // index into Plan::shapes
  int ibase = 0;   // first internal node of this block in the plan-global
                   // internal numbering (node constants live there)
  int owner = 0;   // owning rank, or -1 if the block spans ranks (replicated)
};

struct Tier {
  int nleaves = 0;            // leaves of this tier (tier 0: cells)
  std::vector<Block> blocks;  // in DFS order; block b's root is leaf b of tier+1
  int max_nl = 0;
};

struct Plan {
  int ncells = 0;
  int max_block_leaves = 0;
  std::vector<Tier> tiers;       // tiers.back() has exactly one block
  std::vector<Shape> shapes;
  int ninternal = 0;             // total internal nodes == ncells - 1
  std::vector<int64_t> lci2gci;  // local cell index -> caller's cellidx
  std::vector<int> leaf_rank;    // per lci: owning rank of the leaf
  int nlevels_ref = 0;           // the reference's level count (tree height + 1)

  // Packed topology for the device.
  std::vector<int> dev_lvlptr, dev_kid0, dev_kid1;
  std::vector<unsigned short> dev_dtab, dev_ptab, dev_fpos, dev_perm;
  std::vector<unsigned> dev_pent;
  // True if every tier-0 block has a fast shape (then the fast kernels can run it).
  bool tier0_fast = false;

  // Build from a flat tree: kids[2*i], kids[2*i+1] (-1,-1 for a leaf),
  // cellidx[i] for leaves, rank[i] for leaves (may be null: all rank 0).
  // Throws std::logic_error on malformed input.
  void build(int ncells, int nnodes, int root, const int* kids,
             const int64_t* cellidx, const int* rank, int max_block_leaves);
};

// The recursive-bisection tree of cedr_tree.cpp:391-413
// (make_tree_over_1d_mesh) as flat arrays, nodes in pre-order (root = 0).
// rank_of_cell may be null (all rank 0).
void make_bisection_tree(int ncells, bool imbalanced, std::vector<int>& kids,
                         std::vector<int64_t>& cellidx);

// Union of the ranks' PARTIAL trees (tree::Node::level, cedr_tree_caller.hpp:20-22,
// consumed at cedr_tree.cpp:71-76): each rank may hand over the global tree with every
// subtree that holds none of its cells cut down to a stub (a node without kids whose
// rank is not this rank's, cedr_tree.cpp:96-108). The reference never forms the whole
// tree -- its level lists only need a rank's own nodes and their neighbours; the block
// plan does need it (the tier sweeps above the block roots run on every rank), so the
// parts are all-gathered once at setup and merged here. A node is identified by its
// path from the root (kid slots are kept in the parts); a position is internal if any
// part expands it, and a leaf's cell and rank are read from its owner's part. Part p is
// `nnodes[p]` nodes at offset sum(nnodes[0..p)) of the concatenated arrays, kids
// part-local, root = part-local node `root[p]`. Output in pre-order, root = 0.
void merge_partial_trees(int nparts, const int* nnodes, const int* root, const int* kids,
                         const int64_t* cellidx, const int* rank,
                         std::vector<int>& out_kids, std::vector<int64_t>& out_cellidx,
                         std::vector<int>& out_rank);

} // namespace cedr_b200

#endif
