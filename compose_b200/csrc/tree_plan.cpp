#include "tree_plan.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <map>
#include <mutex>
#include <sstream>
#include <stdexcept>

namespace cedr_b200 {

namespace {

[[noreturn]] void fail (const std::string& msg) {
  throw std::logic_error("cedr_b200 tree plan: " + msg);
}

void bisect (int cs, int ce, bool imbalanced, std::vector<int>& kids,
             std::vector<int64_t>& cellidx) {
  // Iterative pre-order emission; a frame remembers where to patch kid links.
  struct Frame { int cs, ce, parent, which; };
  std::vector<Frame> st;
  st.push_back({cs, ce, -1, 0});
  while (!st.empty()) {
    const Frame f = st.back();
    st.pop_back();
    const int me = static_cast<int>(cellidx.size());
    kids.push_back(-1);
    kids.push_back(-1);
    cellidx.push_back(-1);
    if (f.parent >= 0) kids[2*f.parent + f.which] = me;
    const int cn = f.ce - f.cs;
    if (cn == 1) {
      cellidx[me] = f.cs;
      continue;
    }
    // cedr_tree.cpp:393-397
    const int cn0 = (imbalanced && cn > 2) ? cn/3 : cn/2;
    // Push right first so the left subtree is emitted first (pre-order).
    st.push_back({f.cs + cn0, f.ce, me, 1});
    st.push_back({f.cs, f.cs + cn0, me, 0});
  }
}

} // namespace

// Which leaf thread of the down-sweep works on which depth-7 node, and which lane solves
// which depth-9 pair. A thread reads its micro-subtree's leaves at offsets that advance by
// the irregular leaf counts of the nodes (5 or 6 for 675 leaves), so with thread = node
// the 16 lanes of a half-warp -- the unit of a 64-bit shared-memory access, 16 banks of
// 8 bytes -- hit the same bank up to 5 times (measured 8.75 wavefronts per load against
// an ideal of 2 at ne120). The leaf threads need no particular order among themselves in
// the down-sweep (no shuffles), so nodes are dealt to the 8 half-warps such that the leaf
// offsets within each are as distinct mod 16 as possible: simulated annealing over
// swaps, on the exact wavefront count of the kernel's 8 leaf accesses (first / second
// leaf of each of a thread's 4 depth-9 nodes). Deterministic (fixed seed).
static void build_down_tables_uncached (Shape& sh);

// The search depends on the shape's depth-9 table only; QLT and CAAS objects over the same
// tree (and every rank's plan) share its result within a process.
static void build_down_tables (Shape& sh) {
  struct Tables { std::vector<unsigned short> perm, perm_up; std::vector<unsigned> pent; };
  static std::mutex mtx;
  static std::map<std::pair<std::vector<unsigned short>, std::vector<unsigned short> >,
                  Tables> cache;
  const auto key = std::make_pair(sh.dtab, sh.ptab);
  {
    std::lock_guard<std::mutex> lock(mtx);
    const auto it = cache.find(key);
    if (it != cache.end()) {
      sh.perm = it->second.perm;
      sh.perm_up = it->second.perm_up;
      sh.pent = it->second.pent;
      return;
    }
  }
  build_down_tables_uncached(sh);
  std::lock_guard<std::mutex> lock(mtx);
  if (cache.size() < 64) cache[key] = Tables{sh.perm, sh.perm_up, sh.pent};
}

static void build_down_tables_uncached (Shape& sh) {
  const std::vector<unsigned short>& dtab = sh.dtab;
  auto off9 = [&] (int p) { return dtab[p] & 0x7fff; };
  auto pair9 = [&] (int p) { return (dtab[p] >> 15) != 0; };
  // Wavefronts of the 8 accesses for a half-warp of 16 nodes.
  auto cost = [&] (const int* nodes) {
    int tot = 0;
    for (int q = 0; q < 4; ++q)
      for (int j = 0; j < 2; ++j) {
        int cnt[16] = {0}, mx = 0;
        for (int i = 0; i < 16; ++i) {
          const int p = 4*nodes[i] + q;
          if (j == 1 && ! pair9(p)) continue;
          mx = std::max(mx, ++cnt[(off9(p) + j) & 15]);
        }
        tot += mx;
      }
    return tot;
  };
  // Pairs are solved 32 at a time by the warp that owns their nodes: the number of such
  // rounds, sum over warps of ceil(pairs / 32), comes first (nodes with few pairs dealt
  // together; a round is a whole node problem for the warp, a conflict one wavefront).
  auto npairs_of = [&] (int n) {
    return static_cast<int>(pair9(4*n)) + pair9(4*n + 1) + pair9(4*n + 2) + pair9(4*n + 3);
  };
  const int kRound = 200;
  auto rounds = [&] (const int* h0, const int* h1) {
    int np = 0;
    for (int i = 0; i < 16; ++i) np += npairs_of(h0[i]) + npairs_of(h1[i]);
    return (np + 31)/32;
  };
  int g[8][16];
  {
    std::vector<int> order(128);
    for (int i = 0; i < 128; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(),
                     [&] (int x, int y) { return npairs_of(x) < npairs_of(y); });
    for (int i = 0; i < 128; ++i) g[i >> 4][i & 15] = order[i];
  }
  // gc[k]: wavefronts of half-warp k; wc[w]: pair rounds of warp w (half-warps 2w, 2w+1).
  int gc[8], wc[4], cur = 0;
  for (int k = 0; k < 8; ++k) { gc[k] = cost(g[k]); cur += gc[k]; }
  for (int w = 0; w < 4; ++w) { wc[w] = rounds(g[2*w], g[2*w + 1]); cur += kRound*wc[w]; }
  int best = cur, bestg[8][16];
  std::memcpy(bestg, g, sizeof(g));
  unsigned long long rng = 0x9E3779B97F4A7C15ull;
  auto next = [&] () { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return rng; };
  double T = 2.0;
  for (int it = 0; it < 150000; ++it) {
    const int g1 = static_cast<int>(next() % 8), g2 = static_cast<int>((g1 + 1 + next() % 7) % 8);
    const int i1 = static_cast<int>(next() % 16), i2 = static_cast<int>(next() % 16);
    std::swap(g[g1][i1], g[g2][i2]);
    const int c1 = cost(g[g1]), c2 = cost(g[g2]);
    const int w1 = g1 >> 1, w2 = g2 >> 1;
    const int r1 = rounds(g[2*w1], g[2*w1 + 1]), r2 = rounds(g[2*w2], g[2*w2 + 1]);
    const int d = c1 + c2 - gc[g1] - gc[g2] +
      (w1 == w2 ? 0 : kRound*(r1 + r2 - wc[w1] - wc[w2]));
    const double u = (next() >> 11)*(1.0/9007199254740992.0);
    if (d <= 0 || u < std::exp(-d/T)) {
      gc[g1] = c1; gc[g2] = c2; wc[w1] = r1; wc[w2] = r2; cur += d;
      if (cur < best) { best = cur; std::memcpy(bestg, g, sizeof(g)); }
    } else {
      std::swap(g[g1][i1], g[g2][i2]);
    }
    T = std::max(0.05, T*0.99995);
  }
  sh.perm.resize(128);
  std::vector<int> thread_of(128);
  for (int i = 0; i < 128; ++i) {
    sh.perm[i] = static_cast<unsigned short>(bestg[i >> 4][i & 15]);
    thread_of[sh.perm[i]] = i;
  }
  // Pairs: solved by the warp that owns their depth-7 node, dealt to its lanes 32 at a
  // time; within a warp ordered so that consecutive entries have distinct leaf offsets
  // mod 16 (k-th pair of each residue class first).
  const int np = static_cast<int>(sh.ptab.size());
  std::vector<std::vector<unsigned> > per_warp(4);
  std::vector<std::vector<std::pair<int,unsigned> > > keyed(4);
  int seen[4][16];
  std::memset(seen, 0, sizeof(seen));
  for (int r = 0; r < np; ++r) {
    const int p = sh.ptab[r], owner = thread_of[p >> 2], w = owner >> 5;
    const unsigned o = off9(p), slot = (p & 3)*128 + owner;
    const unsigned e = o | (slot << 11) | (static_cast<unsigned>(r) << 20);
    keyed[w].push_back(std::make_pair(seen[w][o & 15]++*16 + static_cast<int>(o & 15), e));
  }
  sh.pent.assign(5, 0);
  for (int w = 0; w < 4; ++w) {
    std::stable_sort(keyed[w].begin(), keyed[w].end(),
                     [] (const std::pair<int,unsigned>& a, const std::pair<int,unsigned>& b) {
                       return a.first < b.first; });
    sh.pent[w] = static_cast<unsigned>(sh.pent.size());
    for (size_t i = 0; i < keyed[w].size(); ++i) sh.pent.push_back(keyed[w][i].second);
  }
  sh.pent[4] = static_cast<unsigned>(sh.pent.size());

  // Up-sweep: the same wavefront count, nodes moved only between the two half-warps of
  // their own warp (the kernel puts the depth-7 records back into lane order with one
  // shuffle per field before its shuffle tree).
  for (int i = 0; i < 128; ++i) g[i >> 4][i & 15] = i;
  for (int k = 0; k < 8; ++k) gc[k] = cost(g[k]);
  T = 2.0;
  for (int it = 0; it < 60000; ++it) {
    const int g1 = static_cast<int>(next() % 8), g2 = g1 ^ 1;
    const int i1 = static_cast<int>(next() % 16), i2 = static_cast<int>(next() % 16);
    std::swap(g[g1][i1], g[g2][i2]);
    const int c1 = cost(g[g1]), c2 = cost(g[g2]);
    const int d = c1 + c2 - gc[g1] - gc[g2];
    const double u = (next() >> 11)*(1.0/9007199254740992.0);
    if (d <= 0 || u < std::exp(-d/T)) { gc[g1] = c1; gc[g2] = c2; }
    else std::swap(g[g1][i1], g[g2][i2]);
    T = std::max(0.02, T*0.9999);
  }
  sh.perm_up.resize(256);
  for (int i = 0; i < 128; ++i) {
    sh.perm_up[i] = static_cast<unsigned short>(g[i >> 4][i & 15]);
    sh.perm_up[128 + g[i >> 4][i & 15]] = static_cast<unsigned short>(i);
  }
}

// Try to describe `sh` as "perfect to depth 9, leaves or pairs below" (see
// fast_kernels.cuh). Leaves sh.fast false if it is not of that form.
static void build_fast_tables (Shape& sh) {
  sh.fast = false;
  if (sh.nl < 513 || sh.nl > 1024 || sh.ni != sh.nl - 1) return;
  const int nl = sh.nl, ni = sh.ni;
  std::vector<int> depth(ni, -1), pos(ni, -1);
  std::vector<unsigned short> dtab(512, 0xffff), fpos(ni, 0xffff);
  std::vector<char> is_pair(512, 0);
  depth[ni-1] = 0;
  pos[ni-1] = 0;
  // Internal nodes are ordered kids-before-parents, so walk from the back.
  for (int j = ni - 1; j >= 0; --j) {
    if (depth[j] < 0) return;
    const int d = depth[j], p = pos[j];
    if (d > 9) return;
    const int kid[2] = {sh.kid0[j], sh.kid1[j]};
    if (d == 9) {
      // Must be a pair of adjacent leaves.
      if (kid[0] >= nl || kid[1] != kid[0] + 1) return;
      dtab[p] = static_cast<unsigned short>(kid[0] | 0x8000);
      is_pair[p] = 1;
      continue;
    }
    fpos[j] = static_cast<unsigned short>((1 << d) - 1 + p);
    for (int k = 0; k < 2; ++k) {
      const int cp = 2*p + k;
      if (kid[k] >= nl) {
        depth[kid[k] - nl] = d + 1;
        pos[kid[k] - nl] = cp;
      } else {
        if (d + 1 != 9) return;  // a leaf shallower than depth 9
        dtab[cp] = static_cast<unsigned short>(kid[k]);
      }
    }
  }
  std::vector<unsigned short> ptab;
  for (int p = 0; p < 512; ++p) {
    if (dtab[p] == 0xffff) return;
    if (is_pair[p]) ptab.push_back(static_cast<unsigned short>(p));
  }
  if (static_cast<int>(ptab.size()) != nl - 512) return;
  // Pairs take the const positions after the 511 heap nodes, by increasing p.
  for (int j = 0; j < ni; ++j)
    if (depth[j] == 9) {
      const int p = pos[j];
      const int r = static_cast<int>(std::lower_bound(ptab.begin(), ptab.end(), p) -
                                     ptab.begin());
      fpos[j] = static_cast<unsigned short>(511 + r);
    }
  // Leaf offsets must increase with p by 1 or 2 (DFS order), starting at 0.
  int expect = 0;
  for (int p = 0; p < 512; ++p) {
    if ((dtab[p] & 0x7fff) != expect) return;
    expect += is_pair[p] ? 2 : 1;
  }
  sh.dtab.swap(dtab);
  sh.ptab.swap(ptab);
  sh.fpos.swap(fpos);
  sh.fast = true;
  build_down_tables(sh);
}

void make_bisection_tree (int ncells, bool imbalanced, std::vector<int>& kids,
                          std::vector<int64_t>& cellidx) {
  if (ncells < 1) fail("ncells must be >= 1");
  kids.clear();
  cellidx.clear();
  kids.reserve(2*(2*static_cast<size_t>(ncells) - 1));
  cellidx.reserve(2*static_cast<size_t>(ncells) - 1);
  bisect(0, ncells, imbalanced, kids, cellidx);
}

void merge_partial_trees (int nparts, const int* nnodes, const int* root, const int* kids,
                          const int64_t* cellidx, const int* rank,
                          std::vector<int>& out_kids, std::vector<int64_t>& out_cellidx,
                          std::vector<int>& out_rank) {
  if (nparts < 1) fail("merge_partial_trees: need at least one part");
  if (nparts > 1 && ! rank) fail("merge_partial_trees: node ranks are required");
  std::vector<size_t> off(nparts + 1, 0);
  for (int p = 0; p < nparts; ++p) {
    if (nnodes[p] < 1) fail("merge_partial_trees: empty part");
    if (root[p] < 0 || root[p] >= nnodes[p]) fail("merge_partial_trees: root out of range");
    off[p+1] = off[p] + static_cast<size_t>(nnodes[p]);
  }
  std::vector<char> seen(off[nparts], 0);
  out_kids.clear();
  out_cellidx.clear();
  out_rank.clear();

  // One position of the global tree = the nodes the parts have there (global node ids
  // into the concatenated arrays), plus the output slot waiting for this position's id.
  struct Item { std::vector<size_t> views; int patch; };
  std::vector<Item> stack(1);
  stack[0].patch = -1;
  for (int p = 0; p < nparts; ++p) stack[0].views.push_back(off[p] + root[p]);
  std::vector<size_t> views, kid[2];
  while ( ! stack.empty()) {
    views.swap(stack.back().views);
    const int patch = stack.back().patch;
    stack.pop_back();
    if (out_cellidx.size() >= static_cast<size_t>(std::numeric_limits<int>::max()/2))
      fail("merge_partial_trees: tree too large");
    const int me = static_cast<int>(out_cellidx.size());
    if (patch >= 0) out_kids[patch] = me;
    out_kids.push_back(-1);
    out_kids.push_back(-1);
    out_cellidx.push_back(-1);
    out_rank.push_back(0);
    kid[0].clear();
    kid[1].clear();
    long owner_view = -1;
    for (size_t v : views) {
      if (seen[v]) fail("merge_partial_trees: a part is not a tree (node reached twice)");
      seen[v] = 1;
      const int p = static_cast<int>(std::upper_bound(off.begin(), off.end(), v) -
                                     off.begin()) - 1;
      const int k0 = kids[2*v], k1 = kids[2*v+1];
      if (k0 < 0 && k1 < 0) {
        // A leaf, or the stub of a subtree this part has cut off.
        if ( ! rank || rank[v] == p) owner_view = static_cast<long>(v);
        continue;
      }
      if (k0 < 0 || k1 < 0)
        fail("merge_partial_trees: an internal node must keep both kid slots "
             "(cut a subtree down to a stub instead of dropping it)");
      if (k0 >= nnodes[p] || k1 >= nnodes[p] || k0 == k1)
        fail("merge_partial_trees: kid index out of range");
      kid[0].push_back(off[p] + k0);
      kid[1].push_back(off[p] + k1);
    }
    if (kid[0].empty()) {
      if (owner_view < 0)
        fail("merge_partial_trees: no part holds this leaf as its own (every rank's "
             "partial tree must reach the cells it owns)");
      out_cellidx[me] = cellidx[owner_view];
      out_rank[me] = rank ? rank[owner_view] : 0;
      if (out_cellidx[me] < 0) fail("merge_partial_trees: leaf without a cell index");
    } else {
      for (int k = 1; k >= 0; --k) {
        stack.push_back(Item());
        stack.back().views = kid[k];
        stack.back().patch = 2*me + k;
      }
    }
  }
}

void Plan::build (int ncells_, int nnodes, int root, const int* kids,
                  const int64_t* cellidx, const int* rank, int max_block_leaves_) {
  if (ncells_ < 1) fail("ncells must be >= 1");
  if (nnodes != 2*ncells_ - 1) fail("a binary tree over ncells leaves has 2*ncells-1 nodes");
  if (root < 0 || root >= nnodes) fail("root out of range");
  if (max_block_leaves_ < 2) fail("max_block_leaves must be >= 2");
  ncells = ncells_;
  max_block_leaves = max_block_leaves_;
  tiers.clear();
  shapes.clear();
  dev_lvlptr.clear();
  dev_kid0.clear();
  dev_kid1.clear();
  dev_dtab.clear();
  dev_ptab.clear();
  dev_fpos.clear();
  dev_perm.clear();
  dev_pent.clear();
  tier0_fast = false;

  // ---- DFS: leaf order (= the reference's lci), post-order, heights.
  std::vector<int> post;
  post.reserve(nnodes);
  std::vector<int> height(nnodes, 0), seen(nnodes, 0);
  lci2gci.assign(ncells, -1);
  leaf_rank.assign(ncells, 0);
  std::vector<int> node_lci(nnodes, -1);
  {
    int nleaf = 0;
    std::vector<std::pair<int,int> > st; // (node, state)
    st.push_back(std::make_pair(root, 0));
    while (!st.empty()) {
      const int n = st.back().first;
      int& state = st.back().second;
      const int k0 = kids[2*n], k1 = kids[2*n+1];
      if (state == 0) {
        if (seen[n]++) fail("node reachable twice (not a tree)");
        if ((k0 < 0) != (k1 < 0))
          fail("every internal node must have exactly 2 kids (cedr_qlt.cpp:355)");
        if (k0 < 0) {
          if (nleaf >= ncells) fail("more leaves than ncells");
          node_lci[n] = nleaf;
          lci2gci[nleaf] = cellidx[n];
          leaf_rank[nleaf] = rank ? rank[n] : 0;
          ++nleaf;
          post.push_back(n);
          st.pop_back();
          continue;
        }
        if (k0 >= nnodes || k1 >= nnodes) fail("kid index out of range");
        state = 1;
        st.push_back(std::make_pair(k0, 0));
      } else if (state == 1) {
        state = 2;
        st.push_back(std::make_pair(k1, 0));
      } else {
        height[n] = 1 + std::max(height[k0], height[k1]);
        post.push_back(n);
        st.pop_back();
      }
    }
    if (nleaf != ncells) fail("tree does not have ncells leaves");
    if (static_cast<int>(post.size()) != nnodes) fail("tree does not reach every node");
  }
  nlevels_ref = height[root] + 1;

  // ---- Tiers.
  // vleaf[n] >= 0: node n is leaf number vleaf[n] of the current tier.
  std::vector<int> vleaf = node_lci;
  std::vector<int> vrank(nnodes, 0);
  for (int n = 0; n < nnodes; ++n)
    if (node_lci[n] >= 0) vrank[n] = leaf_rank[node_lci[n]];
  std::vector<int> cnt(nnodes), first(nnodes), local_id(nnodes, -1);
  std::map<std::vector<int>, int> shape_index;
  ninternal = 0;
  int nvleaves = ncells;

  for (int tier_idx = 0; ; ++tier_idx) {
    if (tier_idx > 256) fail("tree is too deep/unbalanced for the block plan");
    // cnt/first over the current virtual tree, bottom-up along the post-order.
    for (size_t i = 0; i < post.size(); ++i) {
      const int n = post[i];
      if (vleaf[n] >= 0) { cnt[n] = 1; first[n] = vleaf[n]; }
      else if (kids[2*n] >= 0 && cnt[kids[2*n]] >= 0 && vleaf[n] != -2) {
        cnt[n] = cnt[kids[2*n]] + cnt[kids[2*n+1]];
        first[n] = first[kids[2*n]];
      }
    }
    Tier tier;
    tier.nleaves = nvleaves;
    // Cut: maximal subtrees with <= max_block_leaves current-tier leaves, DFS.
    std::vector<int> block_roots;
    {
      std::vector<int> st;
      st.push_back(root);
      while (!st.empty()) {
        const int n = st.back();
        st.pop_back();
        if (cnt[n] <= max_block_leaves) { block_roots.push_back(n); continue; }
        st.push_back(kids[2*n+1]);
        st.push_back(kids[2*n]);
      }
    }
    for (size_t b = 0; b < block_roots.size(); ++b) {
      const int broot = block_roots[b];
      Block blk;
      blk.leaf0 = first[broot];
      blk.nl = cnt[broot];
      // Local topology: DFS inside the block down to the tier's leaves.
      std::vector<int> internal_post;   // internal nodes, post-order
      std::vector<int> lheight;         // local heights, parallel to internal_post
      std::map<int,int> hmap;
      int nleaf_local = 0;
      int owner = -2;
      bool bis = true;
      {
        std::vector<std::pair<int,int> > st;
        st.push_back(std::make_pair(broot, 0));
        while (!st.empty()) {
          const int n = st.back().first;
          int& state = st.back().second;
          if (vleaf[n] >= 0) {
            local_id[n] = nleaf_local++;
            hmap[n] = 0;
            owner = (owner == -2) ? vrank[n] : (owner == vrank[n] ? owner : -1);
            st.pop_back();
            continue;
          }
          if (state == 0) { state = 1; st.push_back(std::make_pair(kids[2*n], 0)); }
          else if (state == 1) { state = 2; st.push_back(std::make_pair(kids[2*n+1], 0)); }
          else {
            const int h = 1 + std::max(hmap[kids[2*n]], hmap[kids[2*n+1]]);
            hmap[n] = h;
            internal_post.push_back(n);
            lheight.push_back(h);
            if (cnt[kids[2*n]] != cnt[n]/2) bis = false;
            st.pop_back();
          }
        }
      }
      blk.owner = owner;
      const int ni = static_cast<int>(internal_post.size());
      // Order internal nodes by (height, post-order position): stable sort.
      std::vector<int> ord(ni);
      for (int i = 0; i < ni; ++i) ord[i] = i;
      std::stable_sort(ord.begin(), ord.end(),
                       [&] (int a, int c) { return lheight[a] < lheight[c]; });
      for (int j = 0; j < ni; ++j) local_id[internal_post[ord[j]]] = blk.nl + j;
      Shape sh;
      sh.nl = blk.nl;
      sh.ni = ni;
      sh.nlev = ni ? lheight[ord[ni-1]] : 0;
      sh.bisection = bis;
      sh.lvlptr.assign(sh.nlev + 1, 0);
      sh.kid0.resize(ni);
      sh.kid1.resize(ni);
      for (int j = 0; j < ni; ++j) {
        const int n = internal_post[ord[j]];
        sh.kid0[j] = local_id[kids[2*n]];
        sh.kid1[j] = local_id[kids[2*n+1]];
        ++sh.lvlptr[lheight[ord[j]]];
      }
      for (int l = 0; l < sh.nlev; ++l) sh.lvlptr[l+1] += sh.lvlptr[l];
      // Dedupe.
      std::vector<int> key;
      key.reserve(2*ni + 1);
      key.push_back(sh.nl);
      key.insert(key.end(), sh.kid0.begin(), sh.kid0.end());
      key.insert(key.end(), sh.kid1.begin(), sh.kid1.end());
      std::map<std::vector<int>, int>::iterator it = shape_index.find(key);
      if (it == shape_index.end()) {
        build_fast_tables(sh);
        if (sh.fast) {
          sh.dev_dtab_off = static_cast<int>(dev_dtab.size());
          sh.dev_ptab_off = static_cast<int>(dev_ptab.size());
          sh.dev_fpos_off = static_cast<int>(dev_fpos.size());
          dev_dtab.insert(dev_dtab.end(), sh.dtab.begin(), sh.dtab.end());
          dev_ptab.insert(dev_ptab.end(), sh.ptab.begin(), sh.ptab.end());
          // keep ptab/dtab offsets even so ushort4 loads stay aligned
          while (dev_ptab.size() % 4) dev_ptab.push_back(0);
          dev_fpos.insert(dev_fpos.end(), sh.fpos.begin(), sh.fpos.end());
          sh.dev_perm_off = static_cast<int>(dev_perm.size());
          dev_perm.insert(dev_perm.end(), sh.perm.begin(), sh.perm.end());
          sh.dev_perm_up_off = static_cast<int>(dev_perm.size());
          dev_perm.insert(dev_perm.end(), sh.perm_up.begin(), sh.perm_up.end());
          sh.dev_pent_off = static_cast<int>(dev_pent.size());
          dev_pent.insert(dev_pent.end(), sh.pent.begin(), sh.pent.end());
        }
        sh.dev_lvlptr_off = static_cast<int>(dev_lvlptr.size());
        sh.dev_kid_off = static_cast<int>(dev_kid0.size());
        dev_lvlptr.insert(dev_lvlptr.end(), sh.lvlptr.begin(), sh.lvlptr.end());
        dev_kid0.insert(dev_kid0.end(), sh.kid0.begin(), sh.kid0.end());
        dev_kid1.insert(dev_kid1.end(), sh.kid1.begin(), sh.kid1.end());
        blk.shape = static_cast<int>(shapes.size());
        shape_index[key] = blk.shape;
        shapes.push_back(sh);
      } else {
        blk.shape = it->second;
      }
      blk.ibase = ninternal;
      ninternal += ni;
      tier.max_nl = std::max(tier.max_nl, blk.nl);
      tier.blocks.push_back(blk);
    }
    const bool last = block_roots.size() == 1 && block_roots[0] == root;
    // The block roots become the next tier's leaves. Nodes strictly inside a
    // block are retired (marked -2 so cnt is no longer recomputed for them).
    for (size_t b = 0; b < block_roots.size(); ++b) {
      std::vector<int> st;
      st.push_back(block_roots[b]);
      while (!st.empty()) {
        const int n = st.back();
        st.pop_back();
        if (vleaf[n] < 0 && kids[2*n] >= 0) {
          st.push_back(kids[2*n]);
          st.push_back(kids[2*n+1]);
        }
        vleaf[n] = -2;
        cnt[n] = -1;
      }
    }
    for (size_t b = 0; b < block_roots.size(); ++b) {
      vleaf[block_roots[b]] = static_cast<int>(b);
      vrank[block_roots[b]] = tier.blocks[b].owner;
    }
    nvleaves = static_cast<int>(block_roots.size());
    tiers.push_back(tier);
    if (last) break;
  }
  tier0_fast = tiers.size() > 1;
  for (size_t b = 0; b < tiers[0].blocks.size(); ++b)
    if ( ! shapes[tiers[0].blocks[b].shape].fast) tier0_fast = false;
  if (ninternal != ncells - 1) {
    std::stringstream ss;
    ss << "internal error: counted " << ninternal << " internal nodes, expected "
       << ncells - 1;
    fail(ss.str());
  }
}

} // namespace cedr_b200
