// phybin_host.cpp -- host side of the RF path (see include/phybin_host.h).
//
// What a Haskell integration keeps in Haskell is re-stated here in C++ (no GHC in this image):
// Newick parsing with PhyBin's label numbering, the HashRF leaf-count filter, and bipartition
// extraction straight into the bit-packed layout the CUDA library consumes.  Trees are held in
// flat pre-order arrays (a child always has a larger index than its parent), so a clade bit set is
// one backwards sweep of ORs per tree -- no per-node allocation, unlike the reference's IntSet
// unions (RFDistance.hs:216-221).  No CUDA, no oracle.

#include "phybin_host.h"

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>
#include <limits>
#include <unordered_map>
#include <vector>

namespace {

constexpr uint32_t NONE = 0xFFFFFFFFu;

struct Forest {
  std::vector<std::string> label_names;
  std::unordered_map<std::string, uint32_t> name_to_label;
  // All trees back to back, each in pre-order; the root of tree t is node tree_begin[t].
  std::vector<int32_t> label;  // >= 0: leaf with that label; -1: interior
  std::vector<uint32_t> parent, first_child, next_sib;
  std::vector<uint64_t> tree_begin;  // T+1
  std::vector<uint32_t> num_leaves;  // numLeaves (CoreTypes.hs:289): duplicates count twice
  std::vector<uint32_t> file_trees;  // trees parsed from every input text (phb_parse only; not kept by filters)
  std::string err;

  uint32_t T() const { return (uint32_t)num_leaves.size(); }
  uint32_t new_node(int32_t lab, uint32_t par) {
    uint32_t v = (uint32_t)label.size();
    label.push_back(lab);
    parent.push_back(par);
    first_child.push_back(NONE);
    next_sib.push_back(NONE);
    return v;
  }
};

// ---------------------------------------------------------------------------------------------
// Newick (grammar of Parser.hs:127-202; whitespace pre-filtered, Parser.hs:35)
// ---------------------------------------------------------------------------------------------
struct NewickReader {
  Forest& f;
  const std::string& s;
  size_t p = 0;
  int name_prefix;
  std::vector<std::string> leaf_names;  // parallel to nodes created by this reader (by node id - base)
  uint64_t base;

  NewickReader(Forest& fo, const std::string& str, int np)
      : f(fo), s(str), name_prefix(np), base(fo.label.size()) {}

  [[noreturn]] void fail(const char* what) const {
    throw std::runtime_error("parse error at offset " + std::to_string(p) + ": " + what);
  }
  bool at(char c) const { return p < s.size() && s[p] == c; }
  static bool reserved(char c) { return std::strchr("()[]:;,", c) || std::isspace((unsigned char)c); }
  bool digits() {
    size_t b = p;
    while (p < s.size() && std::isdigit((unsigned char)s[p])) ++p;
    return p > b;
  }
  void skip_name() {
    while (p < s.size() && !reserved(s[p])) ++p;
  }
  void metadata() {  // ":len[boot]", len = d+(.d+)?e-?d+ | -?d+(.d+)?
    if (!at(':')) return;
    ++p;
    size_t save = p;
    bool sci = false;
    if (digits()) {
      if (at('.')) { size_t q = p; ++p; if (!digits()) p = q; }
      if (at('e')) {
        ++p;
        if (at('-')) ++p;
        sci = digits();
      }
    }
    if (!sci) {
      p = save;
      if (at('-')) ++p;
      if (!digits()) fail("expected a branch length");
      if (at('.')) { size_t q = p; ++p; if (!digits()) p = q; }
    }
    if (at('[')) {
      ++p;
      if (!digits()) fail("expected bootstrap digits");
      if (!at(']')) fail("expected ]");
      ++p;
    }
  }
  uint32_t subtree(uint32_t par) {
    if (at('(')) {
      ++p;
      uint32_t v = f.new_node(-1, par);
      leaf_names.emplace_back();
      uint32_t last = NONE;
      for (;;) {
        uint32_t c = subtree(v);
        metadata();
        if (last == NONE) f.first_child[v] = c; else f.next_sib[last] = c;
        last = c;
        if (at(',')) { ++p; continue; }
        break;
      }
      if (!at(')')) fail("expected )");
      ++p;
      skip_name();  // interior names are ignored (Parser.hs:152)
      return v;
    }
    size_t b = p;
    skip_name();
    std::string nm = s.substr(b, p - b);
    if (name_prefix >= 0 && (int)nm.size() > name_prefix) nm.resize(name_prefix);
    uint32_t v = f.new_node(0, par);
    leaf_names.push_back(std::move(nm));
    return v;
  }
  void tree() {
    f.tree_begin.push_back(f.label.size());
    subtree(NONE);
    metadata();
    if (!at(';')) fail("expected ;");
    ++p;
  }
};

// Label numbering of extractLabelTable (Parser.hs:85-110): next label = table size at first
// sight; an interior node's children are visited right-to-left.  `name_of(v)` names a leaf.
template <class NameOf>
void assign_labels(Forest& f, uint64_t root, NameOf&& name_of) {
  std::vector<uint32_t> stack{(uint32_t)root}, kids;
  while (!stack.empty()) {
    uint32_t v = stack.back();
    stack.pop_back();
    if (f.first_child[v] == NONE && f.label[v] >= 0) {
      const std::string& nm = name_of(v);
      auto it = f.name_to_label.find(nm);
      if (it == f.name_to_label.end()) {
        uint32_t lab = (uint32_t)f.label_names.size();
        f.name_to_label.emplace(nm, lab);
        f.label_names.push_back(nm);
        f.label[v] = (int32_t)lab;
      } else {
        f.label[v] = (int32_t)it->second;
      }
      continue;
    }
    // push children left-to-right so the rightmost is popped (visited) first
    for (uint32_t c = f.first_child[v]; c != NONE; c = f.next_sib[c]) stack.push_back(c);
  }
}

void finish_counts(Forest& f) {
  uint32_t T = (uint32_t)f.tree_begin.size();
  f.tree_begin.push_back(f.label.size());
  f.num_leaves.assign(T, 0);
  for (uint32_t t = 0; t < T; ++t) {
    uint32_t n = 0;
    for (uint64_t v = f.tree_begin[t]; v < f.tree_begin[t + 1]; ++v) n += f.label[v] >= 0;
    f.num_leaves[t] = n;
  }
}

// Keeps the trees flagged in `keep`, compacting the node arrays.
void compact(Forest& f, const std::vector<char>& keep) {
  Forest g;
  uint32_t T = f.T();
  for (uint32_t t = 0; t < T; ++t) {
    if (!keep[t]) continue;
    uint64_t b = f.tree_begin[t], e = f.tree_begin[t + 1];
    int64_t shift = (int64_t)g.label.size() - (int64_t)b;
    g.tree_begin.push_back(g.label.size());
    for (uint64_t v = b; v < e; ++v) {
      g.label.push_back(f.label[v]);
      auto mv = [&](uint32_t x) { return x == NONE ? NONE : (uint32_t)((int64_t)x + shift); };
      g.parent.push_back(mv(f.parent[v]));
      g.first_child.push_back(mv(f.first_child[v]));
      g.next_sib.push_back(mv(f.next_sib[v]));
    }
    g.num_leaves.push_back(f.num_leaves[t]);
  }
  g.tree_begin.push_back(g.label.size());
  f.label.swap(g.label);
  f.parent.swap(g.parent);
  f.first_child.swap(g.first_child);
  f.next_sib.swap(g.next_sib);
  f.tree_begin.swap(g.tree_begin);
  f.num_leaves.swap(g.num_leaves);
}

// ---------------------------------------------------------------------------------------------
// Synthetic trees
// ---------------------------------------------------------------------------------------------
struct SplitMix64 {
  uint64_t s;
  explicit SplitMix64(uint64_t seed) : s(seed) {}
  uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  uint32_t below(uint32_t m) { return (uint32_t)(((unsigned __int128)next() * m) >> 64); }
  double unit() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};

// Unrooted binary tree kept as a rooted tree whose root has three children.
struct BinTree {
  uint32_t n = 0;
  std::vector<int32_t> taxon;       // >= 0 for leaves
  std::vector<uint32_t> par;
  std::vector<uint32_t> kid;        // 3 slots per node
  std::vector<uint8_t> nk;
  uint32_t root = 0;

  uint32_t add(int32_t tx, uint32_t p) {
    uint32_t v = (uint32_t)taxon.size();
    taxon.push_back(tx);
    par.push_back(p);
    kid.insert(kid.end(), {NONE, NONE, NONE});
    nk.push_back(0);
    return v;
  }
  void push_kid(uint32_t p, uint32_t c) { kid[3 * p + nk[p]++] = c; }
  void replace_kid(uint32_t p, uint32_t from, uint32_t to) {
    for (int i = 0; i < nk[p]; ++i)
      if (kid[3 * p + i] == from) { kid[3 * p + i] = to; return; }
  }
  // random-edge insertion of taxa 0..n-1 starting from the 3-star on 0,1,2
  void random_topology(uint32_t n_, SplitMix64& rng) {
    n = n_;
    taxon.clear(); par.clear(); kid.clear(); nk.clear();
    root = add(-1, NONE);
    for (uint32_t k = 0; k < std::min(n, 3u); ++k) push_kid(root, add((int32_t)k, root));
    for (uint32_t k = 3; k < n; ++k) {
      uint32_t v = 1 + rng.below((uint32_t)taxon.size() - 1);  // uniform non-root node = edge
      uint32_t p = par[v];
      uint32_t u = add(-1, p);
      replace_kid(p, v, u);
      par[v] = u;
      push_kid(u, v);
      push_kid(u, add((int32_t)k, u));
    }
  }
  // one random nearest-neighbour interchange across a random internal edge
  void random_nni(SplitMix64& rng) {
    std::vector<uint32_t>& cand = scratch;
    cand.clear();
    for (uint32_t v = 0; v < taxon.size(); ++v)
      if (v != root && taxon[v] < 0) cand.push_back(v);
    if (cand.empty()) return;
    uint32_t v = cand[rng.below((uint32_t)cand.size())];
    uint32_t p = par[v];
    uint32_t c = kid[3 * v + rng.below(nk[v])];
    uint32_t sibs[3], ns = 0;
    for (int i = 0; i < nk[p]; ++i)
      if (kid[3 * p + i] != v) sibs[ns++] = kid[3 * p + i];
    uint32_t s = sibs[rng.below(ns)];
    replace_kid(v, c, s);
    replace_kid(p, s, c);
    par[s] = v;
    par[c] = p;
  }
  std::vector<uint32_t> scratch;
};

// Appends `bt` restricted to the kept taxa to the forest (pre-order), splicing out interiors left
// with a single child -- the shape a Newick file of the smaller taxon set would have.
void emit_tree(Forest& f, const BinTree& bt, const std::vector<char>& kept) {
  size_t N = bt.taxon.size();
  std::vector<uint32_t> alive(N, 0);
  // NNI can break the "child id > parent id" ordering, so accumulate over an explicit pre-order.
  std::vector<uint32_t> order;
  order.reserve(N);
  std::vector<uint32_t> st{bt.root};
  while (!st.empty()) {
    uint32_t v = st.back();
    st.pop_back();
    order.push_back(v);
    for (int i = 0; i < bt.nk[v]; ++i) st.push_back(bt.kid[3 * v + i]);
  }
  for (size_t i = order.size(); i-- > 0;) {
    uint32_t v = order[i];
    if (bt.taxon[v] >= 0) alive[v] = kept[bt.taxon[v]] ? 1 : 0;
    if (bt.par[v] != NONE) alive[bt.par[v]] += alive[v];
  }
  f.tree_begin.push_back(f.label.size());
  struct Item { uint32_t v, par; };
  std::vector<Item> work{{bt.root, NONE}};
  std::vector<uint32_t> last_child;  // per forest node created here: last child id, for sibling links
  uint64_t base = f.label.size();
  while (!work.empty()) {
    Item it = work.back();
    work.pop_back();
    uint32_t v = it.v;
    for (;;) {  // descend through single-alive-child chains
      if (bt.taxon[v] >= 0) break;
      uint32_t cnt = 0, only = NONE;
      for (int i = 0; i < bt.nk[v]; ++i) {
        uint32_t c = bt.kid[3 * v + i];
        if (alive[c]) { ++cnt; only = c; }
      }
      if (cnt != 1) break;
      v = only;
    }
    uint32_t me = f.new_node(bt.taxon[v] >= 0 ? bt.taxon[v] : -1, it.par);
    last_child.push_back(NONE);
    if (it.par != NONE) {
      uint32_t& lc = last_child[it.par - base];
      if (lc == NONE) f.first_child[it.par] = me; else f.next_sib[lc] = me;
      lc = me;
    }
    if (bt.taxon[v] < 0)  // push children in reverse so they are created left-to-right
      for (int i = bt.nk[v]; i-- > 0;) {
        uint32_t c = bt.kid[3 * v + i];
        if (alive[c]) work.push_back({c, me});
      }
  }
}

// ---------------------------------------------------------------------------------------------
// Bit-set helpers
// ---------------------------------------------------------------------------------------------
inline uint32_t popcount_w(const uint64_t* a, uint32_t W) {
  uint32_t c = 0;
  for (uint32_t w = 0; w < W; ++w) c += (uint32_t)__builtin_popcountll(a[w]);
  return c;
}
inline int lowest_bit(const uint64_t* a, uint32_t W) {  // -1 when empty
  for (uint32_t w = 0; w < W; ++w)
    if (a[w]) return (int)(w * 64 + __builtin_ctzll(a[w]));
  return -1;
}

// Clade bit sets of every node of tree t into `bits` (node-major, W words each).
void clade_bits(const Forest& f, uint32_t t, uint32_t W, std::vector<uint64_t>& bits) {
  uint64_t b = f.tree_begin[t], e = f.tree_begin[t + 1];
  bits.assign((e - b) * W, 0);
  for (uint64_t v = e; v-- > b;) {
    uint64_t* mine = &bits[(v - b) * W];
    if (f.label[v] >= 0) mine[f.label[v] >> 6] |= 1ull << (f.label[v] & 63);
    if (f.parent[v] != NONE) {
      uint64_t* up = &bits[(f.parent[v] - b) * W];
      for (uint32_t w = 0; w < W; ++w) up[w] |= mine[w];
    }
  }
}

// normBip (RFDistance.hs:229-239) on bit sets.  `universe` = {0..size-1}.  Writes the canonical
// side to `out` and returns its cardinality.
uint32_t norm_bip(const uint64_t* b, const uint64_t* universe, uint32_t size, uint32_t W,
                  uint64_t* out) {
  uint32_t cnt = popcount_w(b, W);
  uint32_t half = size / 2;
  if (cnt < half) {
    std::memcpy(out, b, W * 8);
    return cnt;
  }
  uint64_t flipped[64];
  for (uint32_t w = 0; w < W; ++w) flipped[w] = universe[w] & ~b[w];  // invertDense (:127-131)
  bool take_flipped = true;
  if (cnt == half) {
    // min on IntSet's lexicographic order.  b and flipped are disjoint, so the first elements
    // decide; the empty set is below everything.
    int lf = lowest_bit(flipped, W), lb = lowest_bit(b, W);
    take_flipped = (lf < 0) || (lb >= 0 && lf < lb);
  }
  const uint64_t* src = take_flipped ? flipped : b;
  std::memcpy(out, src, W * 8);
  return popcount_w(src, W);
}

// Removes duplicates from the W-word keys in `keys` (a tree's bipartitions are a Set), keeping the first
// occurrence of every key in input order -- the order prf_allbips produces on the device.
void unique_in_order(std::vector<uint64_t>& keys, uint32_t W, std::vector<uint64_t>& out) {
  const size_t k = keys.size() / W;
  size_t cap = 16;
  while (cap < 2 * k) cap <<= 1;
  std::vector<uint32_t> tab(cap, NONE);
  for (size_t i = 0; i < k; ++i) {
    const uint64_t* key = &keys[i * W];
    uint64_t h = 0x9E3779B97F4A7C15ull;
    for (uint32_t w = 0; w < W; ++w) {
      h ^= key[w] + 0x632BE59BD9B4E019ull * (w + 1);
      h = (h ^ (h >> 30)) * 0xBF58476D1CE4E5B9ull;
      h ^= h >> 27;
    }
    size_t s = h & (cap - 1);
    bool fresh = true;
    while (tab[s] != NONE) {
      if (!std::memcmp(&keys[(size_t)tab[s] * W], key, W * 8)) { fresh = false; break; }
      s = (s + 1) & (cap - 1);
    }
    if (!fresh) continue;
    tab[s] = (uint32_t)i;
    out.insert(out.end(), key, key + W);
  }
}

template <class PerTree>
int parallel_trees(const Forest& f, int threads, uint32_t W, PerTree per_tree, uint64_t** data,
                   uint64_t** tree_off, uint64_t* nnz) {
  uint32_t T = f.T();
  unsigned hw = std::thread::hardware_concurrency();
  unsigned nt = threads > 0 ? (unsigned)threads : (hw ? hw : 1);
  nt = std::max(1u, std::min(nt, std::max(1u, T / 64)));
  std::vector<std::vector<uint64_t>> part(nt);
  std::vector<uint64_t> counts(T, 0);
  auto run = [&](unsigned k) {
    uint32_t t0 = (uint32_t)((uint64_t)T * k / nt), t1 = (uint32_t)((uint64_t)T * (k + 1) / nt);
    std::vector<uint64_t> bits, keys;
    for (uint32_t t = t0; t < t1; ++t) {
      size_t before = part[k].size();
      per_tree(t, bits, keys, part[k]);
      counts[t] = (part[k].size() - before) / W;
    }
  };
  std::vector<std::thread> th;
  for (unsigned k = 1; k < nt; ++k) th.emplace_back(run, k);
  run(0);
  for (auto& x : th) x.join();
  uint64_t total = 0;
  uint64_t* off = (uint64_t*)std::malloc((T + 1) * sizeof(uint64_t));
  if (!off) return -1;
  for (uint32_t t = 0; t < T; ++t) { off[t] = total; total += counts[t]; }
  off[T] = total;
  uint64_t* out = (uint64_t*)std::malloc(std::max<uint64_t>(total * W, 1) * sizeof(uint64_t));
  if (!out) { std::free(off); return -1; }
  uint64_t pos = 0;
  for (unsigned k = 0; k < nt; ++k) {
    if (!part[k].empty()) std::memcpy(out + pos, part[k].data(), part[k].size() * 8);
    pos += part[k].size();
    std::vector<uint64_t>().swap(part[k]);
  }
  *data = out;
  *tree_off = off;
  *nnz = total;
  return 0;
}

}  // namespace

struct phb_forest {
  Forest f;
};

extern "C" {

phb_forest* phb_parse(int nfiles, const char* const* texts, int name_prefix) {
  auto* h = new phb_forest;
  Forest& f = h->f;
  try {
    // parseNewicks folds the files from the right (Parser.hs:71): the last file is labelled first,
    // yet the trees keep command-line order.  Parse right-to-left into separate forests, then
    // stitch them left-to-right.
    std::vector<Forest> per(nfiles);
    for (int fi = nfiles; fi-- > 0;) {
      std::string filtered;
      for (const char* c = texts[fi]; *c; ++c)
        if (!std::isspace((unsigned char)*c)) filtered.push_back(*c);
      Forest& g = per[fi];
      g.label_names.swap(f.label_names);
      g.name_to_label.swap(f.name_to_label);
      NewickReader rd(g, filtered, name_prefix);
      do rd.tree(); while (rd.p < filtered.size());
      for (size_t t = 0; t < g.tree_begin.size(); ++t)
        assign_labels(g, g.tree_begin[t], [&](uint32_t v) -> const std::string& { return rd.leaf_names[v]; });
      finish_counts(g);
      f.label_names.swap(g.label_names);
      f.name_to_label.swap(g.name_to_label);
    }
    for (int fi = 0; fi < nfiles; ++fi) {
      Forest& g = per[fi];
      uint64_t shift = f.label.size();
      auto mv = [&](uint32_t x) { return x == NONE ? NONE : (uint32_t)(x + shift); };
      for (size_t v = 0; v < g.label.size(); ++v) {
        f.label.push_back(g.label[v]);
        f.parent.push_back(mv(g.parent[v]));
        f.first_child.push_back(mv(g.first_child[v]));
        f.next_sib.push_back(mv(g.next_sib[v]));
      }
      for (uint32_t t = 0; t < g.T(); ++t) {
        f.tree_begin.push_back(g.tree_begin[t] + shift);
        f.num_leaves.push_back(g.num_leaves[t]);
      }
      f.file_trees.push_back(g.T());
    }
    f.tree_begin.push_back(f.label.size());
  } catch (const std::exception& e) {
    f = Forest{};
    f.tree_begin.push_back(0);
    f.err = e.what();
  }
  return h;
}

phb_forest* phb_synth(uint32_t T, uint32_t n, uint64_t seed, int family, uint32_t nni_max,
                      double p_missing) {
  auto* h = new phb_forest;
  Forest& f = h->f;
  if (n < 3 || (uint64_t)T * (2ull * n) > 0xFFFFFFF0ull) {
    f.tree_begin.push_back(0);
    f.err = "phb_synth: need n >= 3 and T*2n < 2^32 nodes";
    return h;
  }
  BinTree base, cur;
  if (family == PHB_NNI_FAMILY) {
    SplitMix64 r0(seed ^ 0xBA5EBA5EBA5EBA5Eull);
    base.random_topology(n, r0);
  }
  std::vector<char> kept(n, 1);
  for (uint32_t t = 0; t < T; ++t) {
    SplitMix64 rng(seed + 0x632BE59BD9B4E019ull * (t + 1));  // one stream per tree
    if (family == PHB_NNI_FAMILY) {
      cur = base;
      uint32_t r = rng.below(nni_max + 1);
      for (uint32_t k = 0; k < r; ++k) cur.random_nni(rng);
    } else {
      cur.random_topology(n, rng);
    }
    std::fill(kept.begin(), kept.end(), 1);
    if (p_missing > 0) {
      uint32_t alive = n;
      for (uint32_t x = 0; x < n; ++x)
        if (rng.unit() < p_missing && alive > 4) { kept[x] = 0; --alive; }
    }
    emit_tree(f, cur, kept);
  }
  // name taxa t<k>, then number them the way parseNewicks numbers the emitted text
  std::vector<std::string> names(n);
  for (uint32_t k = 0; k < n; ++k) names[k] = "t" + std::to_string(k);
  std::vector<int32_t> taxon_of(f.label.begin(), f.label.end());
  for (size_t t = 0; t < f.tree_begin.size(); ++t)
    assign_labels(f, f.tree_begin[t], [&](uint32_t v) -> const std::string& { return names[taxon_of[v]]; });
  finish_counts(f);
  return h;
}

void phb_free(phb_forest* f) { delete f; }
const char* phb_error(const phb_forest* f) { return f->f.err.empty() ? nullptr : f->f.err.c_str(); }
uint32_t phb_num_trees(const phb_forest* f) { return f->f.T(); }
uint32_t phb_num_labels(const phb_forest* f) { return (uint32_t)f->f.label_names.size(); }
const char* phb_label_name(const phb_forest* f, uint32_t l) {
  return l < f->f.label_names.size() ? f->f.label_names[l].c_str() : "";
}
uint32_t phb_file_num_trees(const phb_forest* f, uint32_t i) { return i < f->f.file_trees.size() ? f->f.file_trees[i] : 0; }
uint32_t phb_tree_num_leaves(const phb_forest* f, uint32_t t) { return t < f->f.T() ? f->f.num_leaves[t] : 0; }

uint32_t phb_filter_num_leaves(phb_forest* h, uint32_t expected) {
  Forest& f = h->f;
  std::vector<char> keep(f.T());
  for (uint32_t t = 0; t < f.T(); ++t) keep[t] = f.num_leaves[t] == expected;
  compact(f, keep);
  return f.T();
}
uint32_t phb_slice(phb_forest* h, uint32_t t0, uint32_t t1) {
  Forest& f = h->f;
  std::vector<char> keep(f.T());
  for (uint32_t t = 0; t < f.T(); ++t) keep[t] = t >= t0 && t < t1;
  compact(f, keep);
  return f.T();
}

int phb_has_duplicate_leaves(const phb_forest* h) {
  const Forest& f = h->f;
  std::vector<uint32_t> stamp(f.label_names.size(), NONE);
  for (uint32_t t = 0; t < f.T(); ++t)
    for (uint64_t v = f.tree_begin[t]; v < f.tree_begin[t + 1]; ++v)
      if (f.label[v] >= 0) {
        if (stamp[f.label[v]] == t) return 1;
        stamp[f.label[v]] = t;
      }
  return 0;
}

char* phb_to_newick(const phb_forest* h, uint32_t t0, uint32_t t1) {
  const Forest& f = h->f;
  std::string o;
  t1 = std::min(t1, f.T());
  for (uint32_t t = t0; t < t1; ++t) {
    // iterative pre-order with explicit close markers
    struct It { uint32_t v; bool close; };
    std::vector<It> st{{(uint32_t)f.tree_begin[t], false}};
    while (!st.empty()) {
      It it = st.back();
      st.pop_back();
      if (it.close) {
        o += ')';
      } else if (f.label[it.v] >= 0) {
        o += f.label_names[f.label[it.v]];
      } else {
        o += '(';
        st.push_back({it.v, true});
        std::vector<uint32_t> kids;
        for (uint32_t c = f.first_child[it.v]; c != NONE; c = f.next_sib[c]) kids.push_back(c);
        for (size_t i = kids.size(); i-- > 0;) st.push_back({kids[i], false});
        continue;
      }
      // separator: a comma unless the next thing closes a group or the tree ended
      if (!st.empty() && !st.back().close) o += ',';
    }
    o += ";\n";
  }
  char* r = (char*)std::malloc(o.size() + 1);
  if (r) std::memcpy(r, o.c_str(), o.size() + 1);
  return r;
}

uint32_t phb_words(const phb_forest* h) {
  const Forest& f = h->f;
  uint32_t m = (uint32_t)f.label_names.size();
  for (uint32_t x : f.num_leaves) m = std::max(m, x);
  return std::max(1u, (m + 63) / 64);
}

int phb_allbips(const phb_forest* h, uint32_t W, int threads, uint64_t** bips, uint64_t** tree_off,
                uint64_t* nnz) {
  const Forest& f = h->f;
  if (!bips || !tree_off || !nnz || W < phb_words(h) || W > 64) return -1;
  auto per_tree = [&f, W](uint32_t t, std::vector<uint64_t>& bits, std::vector<uint64_t>& keys,
                          std::vector<uint64_t>& out) {
    clade_bits(f, t, W, bits);
    uint64_t b = f.tree_begin[t], e = f.tree_begin[t + 1];
    uint32_t size = f.num_leaves[t];  // labelBips: size = numLeaves tr (RFDistance.hs:215)
    uint64_t universe[64];
    for (uint32_t w = 0; w < W; ++w) {
      int64_t rem = (int64_t)size - 64ll * w;
      universe[w] = rem >= 64 ? ~0ull : (rem <= 0 ? 0ull : ((1ull << rem) - 1));
    }
    keys.clear();
    uint64_t tmp[64];
    // every node except an interior root contributes its leaf set (labelBips :217-221); a leaf's
    // singleton can only survive the `> 1` filter when size < 4 (it stays a singleton otherwise).
    for (uint64_t v = b; v < e; ++v) {
      bool is_leaf = f.label[v] >= 0;
      if (v == b && !is_leaf) continue;
      if (is_leaf && size >= 4) continue;
      if (norm_bip(&bits[(v - b) * W], universe, size, W, tmp) > 1)  // allBips filter (:248)
        keys.insert(keys.end(), tmp, tmp + W);
    }
    unique_in_order(keys, W, out);
  };
  return parallel_trees(f, threads, W, per_tree, bips, tree_off, nnz);
}

int phb_raw_clades(const phb_forest* h, uint32_t W, int threads, uint64_t** clades,
                   uint64_t** tree_off, uint64_t** leafsets, uint64_t* nnz) {
  const Forest& f = h->f;
  if (!clades || !tree_off || !leafsets || !nnz || W < phb_words(h) || W > 64) return -1;
  uint32_t T = f.T();
  uint64_t* ls = (uint64_t*)std::calloc(std::max<uint64_t>((uint64_t)T * W, 1), 8);
  if (!ls) return -1;
  auto per_tree = [&f, W, ls](uint32_t t, std::vector<uint64_t>& bits, std::vector<uint64_t>& keys,
                              std::vector<uint64_t>& out) {
    clade_bits(f, t, W, bits);
    uint64_t b = f.tree_begin[t], e = f.tree_begin[t + 1];
    std::memcpy(ls + (uint64_t)t * W, &bits[0], W * 8);  // the root's set = treeLabels
    keys.clear();
    for (uint64_t v = b + 1; v < e; ++v)
      if (f.label[v] < 0) keys.insert(keys.end(), &bits[(v - b) * W], &bits[(v - b) * W] + W);
    unique_in_order(keys, W, out);
  };
  int rc = parallel_trees(f, threads, W, per_tree, clades, tree_off, nnz);
  if (rc) { std::free(ls); return rc; }
  *leafsets = ls;
  return 0;
}

int phb_raw_clades_ex(const phb_forest* h, uint32_t W, uint64_t** clades, uint64_t** outsets, uint64_t** tree_off,
                      uint64_t** leafsets, uint64_t** dup_off, uint32_t** dup_label, uint32_t** dup_extra, uint64_t* nnz) {
  const Forest& f = h->f;
  if (!clades || !outsets || !tree_off || !leafsets || !dup_off || !dup_label || !dup_extra || !nnz || W < phb_words(h) || W > 64)
    return -1;
  const uint32_t T = f.T();
  std::vector<uint64_t> cl, ou, off(1, 0), ls((uint64_t)T * W, 0), doff(1, 0), bits, out, acc(W);
  std::vector<uint32_t> dlab, dext, count(f.label_names.size(), 0), kids;
  for (uint32_t t = 0; t < T; ++t) {
    clade_bits(f, t, W, bits);
    const uint64_t b = f.tree_begin[t], e = f.tree_begin[t + 1];
    std::memcpy(&ls[(uint64_t)t * W], &bits[0], W * 8);
    // out(root) = {}; out(child c of v) = out(v) | clades of c's siblings (label SETS: with repeated labels
    // out(c) may intersect clade(c))
    out.assign((e - b) * W, 0);
    for (uint64_t v = b; v < e; ++v) {   // pre-order: parents first
      kids.clear();
      for (uint32_t c = f.first_child[v]; c != NONE; c = f.next_sib[c]) kids.push_back(c);
      if (kids.empty()) continue;
      std::vector<uint64_t> suffix((kids.size() + 1) * W, 0);
      for (size_t i = kids.size(); i-- > 0;)
        for (uint32_t w = 0; w < W; ++w) suffix[i * W + w] = suffix[(i + 1) * W + w] | bits[(kids[i] - b) * W + w];
      for (uint32_t w = 0; w < W; ++w) acc[w] = out[(v - b) * W + w];
      for (size_t i = 0; i < kids.size(); ++i) {
        for (uint32_t w = 0; w < W; ++w) {
          out[(kids[i] - b) * W + w] = acc[w] | suffix[(i + 1) * W + w];
          acc[w] |= bits[(kids[i] - b) * W + w];
        }
      }
    }
    for (uint64_t v = b + 1; v < e; ++v)   // every non-root interior node, NOT made unique: equal clades can differ in out
      if (f.label[v] < 0) {
        cl.insert(cl.end(), &bits[(v - b) * W], &bits[(v - b) * W] + W);
        ou.insert(ou.end(), &out[(v - b) * W], &out[(v - b) * W] + W);
      }
    off.push_back(cl.size() / W);
    for (uint64_t v = b; v < e; ++v)
      if (f.label[v] >= 0) ++count[f.label[v]];
    for (uint64_t v = b; v < e; ++v)
      if (f.label[v] >= 0 && count[f.label[v]]) {
        if (count[f.label[v]] > 1) { dlab.push_back((uint32_t)f.label[v]); dext.push_back(count[f.label[v]] - 1); }
        count[f.label[v]] = 0;
      }
    doff.push_back(dlab.size());
  }
  auto give = [](const auto& v, auto** dst) -> bool {
    using E = typename std::remove_reference<decltype(v)>::type::value_type;
    E* p = (E*)std::malloc(std::max<size_t>(v.size(), 1) * sizeof(E));
    if (!p) return false;
    if (!v.empty()) std::memcpy(p, v.data(), v.size() * sizeof(E));
    *dst = p;
    return true;
  };
  if (!give(cl, clades) || !give(ou, outsets) || !give(off, tree_off) || !give(ls, leafsets) || !give(doff, dup_off) ||
      !give(dlab, dup_label) || !give(dext, dup_extra))
    return -1;
  *nnz = cl.size() / W;
  return 0;
}

int phb_tree_arrays(const phb_forest* h, const int32_t** node_label, const uint32_t** node_parent,
                    const uint64_t** tree_begin, const uint32_t** num_leaves, uint64_t* n_nodes) {
  if (!h) return -1;
  const Forest& f = h->f;
  if (f.tree_begin.size() != (size_t)f.T() + 1) return -1;
  if (node_label) *node_label = f.label.data();
  if (node_parent) *node_parent = f.parent.data();
  if (tree_begin) *tree_begin = f.tree_begin.data();
  if (num_leaves) *num_leaves = f.num_leaves.data();
  if (n_nodes) *n_nodes = f.label.size();
  return 0;
}

// Ord IntSet = lexicographic on the ascending element lists: at the lowest label where the two sets differ,
// the set that has it is the smaller one unless the other set has nothing above it (a proper prefix is smaller).
static bool intset_less(const uint64_t* a, const uint64_t* b, uint32_t W) {
  for (uint32_t w = 0; w < W; ++w) {
    uint64_t x = a[w] ^ b[w];
    if (!x) continue;
    int d = __builtin_ctzll(x);
    bool a_has = (a[w] >> d) & 1;
    const uint64_t* other = a_has ? b : a;
    bool other_has_more = d < 63 && (other[w] >> (d + 1)) != 0;
    for (uint32_t v = w + 1; v < W && !other_has_more; ++v) other_has_more = other[v] != 0;
    return a_has ? other_has_more : !other_has_more;
  }
  return false;
}

char* phb_consensus_newick(const phb_forest* h, const uint64_t* keys, uint64_t k, uint32_t W, uint32_t num_taxa) {
  if (!h || (k && !keys) || W == 0 || W > 64 || num_taxa > 64 * W) return nullptr;
  const Forest& f = h->f;
  if (num_taxa > f.label_names.size()) return nullptr;
  // sorted = L.sortBy (compare `on` bipSize) (S.toList origbip): stable by size over Set order
  std::vector<const uint64_t*> bip(k);
  for (uint64_t i = 0; i < k; ++i) bip[i] = keys + i * W;
  std::sort(bip.begin(), bip.end(), [W](const uint64_t* a, const uint64_t* b) {
    uint32_t ca = popcount_w(a, W), cb = popcount_w(b, W);
    return ca != cb ? ca < cb : intset_less(a, b, W);
  });
  struct Node { int32_t leaf; std::vector<uint32_t> kids; };
  std::vector<Node> nodes;
  struct Sub { std::vector<uint64_t> set; uint32_t node; };
  std::vector<Sub> subs;  // lvl0: the singleton leaves 0 .. num_taxa-1 (:416-417)
  for (uint32_t ix = 0; ix < num_taxa; ++ix) {
    Sub s{std::vector<uint64_t>(W, 0), (uint32_t)nodes.size()};
    s.set[ix >> 6] = 1ull << (ix & 63);
    nodes.push_back(Node{(int32_t)ix, {}});
    subs.push_back(std::move(s));
  }
  for (const uint64_t* b : bip) {
    std::vector<Sub> in, out;
    for (Sub& s : subs) {
      bool inside = true;  // isMatch bip x = denseIsSubset x bip (:424)
      for (uint32_t w = 0; w < W; ++w) inside &= (s.set[w] & ~b[w]) == 0;
      (inside ? in : out).push_back(std::move(s));
    }
    if (in.empty()) return nullptr;  // "bipsToTree: Internal error!  No match for bip" (:436)
    Sub merged{std::vector<uint64_t>(W, 0), (uint32_t)nodes.size()};
    Node nd{-1, {}};
    for (const Sub& s : in) {
      for (uint32_t w = 0; w < W; ++w) merged.set[w] |= s.set[w];
      nd.kids.push_back(s.node);
    }
    nodes.push_back(std::move(nd));
    subs.clear();
    subs.push_back(std::move(merged));  // the merged subtree goes to the front (:443-444)
    for (Sub& s : out) subs.push_back(std::move(s));
  }
  uint32_t root;
  if (subs.empty()) return nullptr;
  if (subs.size() == 1) {
    root = subs[0].node;
  } else {
    Node nd{-1, {}};
    for (const Sub& s : subs) nd.kids.push_back(s.node);
    root = (uint32_t)nodes.size();
    nodes.push_back(std::move(nd));
  }
  std::string o;
  struct It { uint32_t v; size_t next; };
  std::vector<It> st{{root, 0}};
  while (!st.empty()) {
    It& it = st.back();
    const Node& nd = nodes[it.v];
    if (nd.leaf >= 0) { o += f.label_names[nd.leaf]; st.pop_back(); continue; }
    if (it.next == 0) o += '(';
    if (it.next == nd.kids.size()) { o += ')'; st.pop_back(); continue; }
    if (it.next) o += ", ";
    uint32_t child = nd.kids[it.next++];
    st.push_back({child, 0});
  }
  o += ';';
  char* r = (char*)std::malloc(o.size() + 1);
  if (r) std::memcpy(r, o.c_str(), o.size() + 1);
  return r;
}

namespace {

// Agglomerative clustering of m items cut at `thresh`; see phybin_host.h.  d: packed upper triangle, modified.
int agglomerate(std::vector<double>& d, uint32_t m, int linkage, double thresh, uint32_t* labels, uint32_t* merge_a = nullptr,
                uint32_t* merge_b = nullptr, double* merge_d = nullptr, uint32_t* n_merges = nullptr) {
  auto at = [&](uint32_t i, uint32_t j) -> double& {  // i < j
    return d[(uint64_t)i * (2ull * m - i - 1) / 2 + (j - i - 1)];
  };
  auto get = [&](uint32_t i, uint32_t j) -> double& { return i < j ? at(i, j) : at(j, i); };
  const double inf = std::numeric_limits<double>::infinity();
  std::vector<char> active(m, 1);
  std::vector<uint32_t> size(m, 1), parent(m), nn(m, NONE);
  std::vector<double> nnd(m, inf);
  for (uint32_t i = 0; i < m; ++i) parent[i] = i;
  auto rescan = [&](uint32_t i) {  // nearest active j > i, the first one on ties
    nn[i] = NONE;
    nnd[i] = inf;
    for (uint32_t j = i + 1; j < m; ++j)
      if (active[j] && at(i, j) < nnd[i]) { nnd[i] = at(i, j); nn[i] = j; }
  };
  for (uint32_t i = 0; i < m; ++i) rescan(i);
  uint32_t step = 0;
  for (uint32_t left = m; left > 1; --left) {
    uint32_t a = NONE;
    double best = inf;
    for (uint32_t i = 0; i < m; ++i)
      if (active[i] && nn[i] != NONE && nnd[i] < best) { best = nnd[i]; a = i; }
    if (a == NONE || best > thresh) break;  // sliceDendro: dist > dstThresh splits (PhyBin.hs:418-419)
    const uint32_t b = nn[a];
    if (merge_a) merge_a[step] = a;
    if (merge_b) merge_b[step] = b;
    if (merge_d) merge_d[step] = best;
    ++step;
    const double na = size[a], nb = size[b];
    for (uint32_t x = 0; x < m; ++x) {
      if (!active[x] || x == a || x == b) continue;
      double &da = get(a, x), db = get(b, x);
      da = linkage == 0 ? std::min(da, db) : linkage == 1 ? std::max(da, db) : (na * da + nb * db) / (na + nb);
    }
    size[a] += size[b];
    active[b] = 0;
    parent[b] = a;
    for (uint32_t i = 0; i < b; ++i) {
      if (!active[i]) continue;
      if (i == a || nn[i] == a || nn[i] == b) { rescan(i); continue; }
      if (i < a) {
        const double v = at(i, a);
        if (v < nnd[i] || (v == nnd[i] && a < nn[i])) { nnd[i] = v; nn[i] = a; }
      }
    }
  }
  if (n_merges) *n_merges = step;
  for (uint32_t i = 0; labels && i < m; ++i) {
    uint32_t r = i;
    while (parent[r] != r) r = parent[r];
    labels[i] = r;  // merges keep the lower key, so the root is the cluster's smallest index
  }
  return 0;
}

// `show` of a Haskell Double (Numeric.showFloat via floatToDigits): the shortest digits that read back to the
// same double; plain decimal for 0.1 <= x < 10^7, d.ddde<n> otherwise.
void show_double(std::string& o, double x) {
  if (x != x) { o += "NaN"; return; }
  if (x < 0) { o += '-'; x = -x; }
  if (x == std::numeric_limits<double>::infinity()) { o += "Infinity"; return; }
  if (x == 0) { o += "0.0"; return; }
  char buf[40];
  int prec = 0;
  for (prec = 0; prec < 17; ++prec) {
    std::snprintf(buf, sizeof buf, "%.*e", prec, x);
    if (std::strtod(buf, nullptr) == x) break;
  }
  std::string digits;
  const char* e = std::strchr(buf, 'e');
  for (const char* c = buf; c < e; ++c)
    if (*c != '.') digits += *c;
  const int exp10 = std::atoi(e + 1);  // x = d.ddd * 10^exp10
  if (x >= 0.1 && x < 1e7) {
    if (exp10 < 0) {  // 0.1 <= x < 1
      o += "0.";
      o += digits;
    } else {
      for (int i = 0; i <= exp10; ++i) o += i < (int)digits.size() ? digits[i] : '0';
      o += '.';
      if ((int)digits.size() > exp10 + 1) o.append(digits, exp10 + 1, std::string::npos);
      else o += '0';
    }
  } else {
    o += digits[0];
    o += '.';
    if (digits.size() > 1) o.append(digits, 1, std::string::npos);
    else o += '0';
    o += 'e';
    o += std::to_string(exp10);
  }
}

// `show` of a Haskell String.
void show_string(std::string& o, const char* s) {
  o += '"';
  bool num_escape = false;  // the previous character was a numeric escape: a following digit needs "\&"
  for (const unsigned char* c = (const unsigned char*)s; *c; ++c) {
    if (num_escape && *c >= '0' && *c <= '9') o += "\\&";
    num_escape = false;
    switch (*c) {
      case '"': o += "\\\""; break;
      case '\\': o += "\\\\"; break;
      case '\n': o += "\\n"; break;
      case '\t': o += "\\t"; break;
      case '\r': o += "\\r"; break;
      default:
        if (*c < 32 || *c >= 127) { o += '\\'; o += std::to_string((unsigned)*c); num_escape = true; }
        else o += (char)*c;
    }
  }
  o += '"';
}

}  // namespace

int phb_cluster_component_f64(const double* sub, uint32_t m, int linkage, double thresh, uint32_t* labels) {
  if ((m > 1 && !sub) || (m && !labels) || linkage < 0 || linkage > 2) return -1;
  if (m > 46000) return -2;  // 8.5 GB of doubles
  std::vector<double> d(sub, sub + (uint64_t)m * (m ? m - 1 : 0) / 2);
  return agglomerate(d, m, linkage, thresh, labels);
}

int phb_cluster_component(const uint16_t* sub, uint32_t m, int linkage, double thresh, uint32_t* labels) {
  if ((m > 1 && !sub) || (m && !labels) || linkage < 0 || linkage > 2) return -1;
  if (m > 46000) return -2;
  std::vector<double> d(sub, sub + (uint64_t)m * (m ? m - 1 : 0) / 2);  // fromIntegral (PhyBin.hs:355-356)
  return agglomerate(d, m, linkage, thresh, labels);
}

int phb_agglomerate(const uint16_t* sub, uint32_t m, int linkage, double thresh, uint32_t* labels, uint32_t* merge_a,
                    uint32_t* merge_b, double* merge_d, uint32_t* n_merges) {
  if ((m > 1 && !sub) || linkage < 0 || linkage > 2) return -1;
  if (m > 46000) return -2;
  std::vector<double> d(sub, sub + (uint64_t)m * (m ? m - 1 : 0) / 2);
  return agglomerate(d, m, linkage, thresh, labels, merge_a, merge_b, merge_d, n_merges);
}

char* phb_dendrogram_text(const char* const* names, uint32_t m, const uint32_t* merge_a, const uint32_t* merge_b,
                          const double* merge_d, uint32_t n_merges) {
  if (!m || !names || n_merges + 1 != m || (n_merges && (!merge_a || !merge_b || !merge_d))) return nullptr;
  // node k < m: Leaf k; node m + k: the Branch made by merge k.  cur[key] = node currently standing for that cluster.
  std::vector<uint32_t> cur(m), left(n_merges), right(n_merges);
  for (uint32_t i = 0; i < m; ++i) cur[i] = i;
  for (uint32_t k = 0; k < n_merges; ++k) {
    const uint32_t a = merge_a[k], b = merge_b[k];
    if (a >= b || b >= m || cur[a] == NONE || cur[b] == NONE) return nullptr;
    left[k] = cur[a];
    right[k] = cur[b];
    cur[a] = m + k;
    cur[b] = NONE;
  }
  // derived Show, showsPrec 0: "Branch d (l) (r)" / "Leaf \"name\""; iterative (a chain dendrogram is m deep)
  std::string o;
  struct It { uint32_t node; int state; };
  std::vector<It> st;
  st.push_back({n_merges ? m + n_merges - 1 : 0u, 0});
  while (!st.empty()) {
    It& it = st.back();
    if (it.node < m) {
      o += "Leaf ";
      show_string(o, names[it.node] ? names[it.node] : "");
      st.pop_back();
      continue;
    }
    const uint32_t k = it.node - m;
    if (it.state == 0) {
      o += "Branch ";
      show_double(o, merge_d[k]);
      o += " (";
      it.state = 1;
      st.push_back({left[k], 0});
    } else if (it.state == 1) {
      o += ") (";
      it.state = 2;
      st.push_back({right[k], 0});
    } else {
      o += ')';
      st.pop_back();
    }
  }
  char* r = (char*)std::malloc(o.size() + 1);
  if (r) std::memcpy(r, o.c_str(), o.size() + 1);
  return r;
}

static void append_uint(std::string& o, unsigned x) {
  char buf[12];
  int n = 0;
  do { buf[n++] = (char)('0' + x % 10); x /= 10; } while (x);
  while (n) o += buf[--n];
}

// printDistMat (RFDistance.hs:352-361): row i lists d(i,0..i-1), each followed by one space, then
// the implicit diagonal "0".
static void dist_mat_row(std::string& o, const uint16_t* upper, uint32_t T, uint32_t i) {
  for (uint32_t j = 0; j < i; ++j) {
    uint64_t p = (uint64_t)j * T - (uint64_t)j * (j + 1) / 2 + (i - j - 1);
    append_uint(o, upper[p]);
    o += ' ';
  }
  o += "0\n";
}
static const char* const kHeader = "Robinson-Foulds distance (matrix format):\n";
static const char* const kRule = "-----------------------------------------\n";

char* phb_print_dist_mat(const uint16_t* upper, uint32_t T) {
  std::string o = kHeader;
  o += kRule;
  for (uint32_t i = 0; i < T; ++i) dist_mat_row(o, upper, T, i);
  o += kRule;
  char* r = (char*)std::malloc(o.size() + 1);
  if (r) std::memcpy(r, o.c_str(), o.size() + 1);
  return r;
}

int phb_write_dist_mat(const char* path, const uint16_t* upper, uint32_t T) {
  FILE* fp = std::fopen(path, "w");
  if (!fp) return -1;
  std::fputs(kHeader, fp);
  std::fputs(kRule, fp);
  std::string o;
  for (uint32_t i = 0; i < T; ++i) {
    o.clear();
    dist_mat_row(o, upper, T, i);
    std::fwrite(o.data(), 1, o.size(), fp);
  }
  std::fputs(kRule, fp);
  return std::fclose(fp) == 0 ? 0 : -1;
}

namespace {
struct DistBinHeader {
  char magic[8];
  uint32_t version, T;
  uint64_t P;
  uint32_t bits, layout;
};
static_assert(sizeof(DistBinHeader) == 32, "header layout");
}  // namespace

int phb_dist_bin_begin(const char* path, uint32_t T) {
  FILE* fp = path ? std::fopen(path, "wb") : nullptr;
  if (!fp) return -1;
  DistBinHeader h{{'P', 'H', 'Y', 'B', 'I', 'N', 'R', 'F'}, 1, T, (uint64_t)T * (T ? T - 1 : 0) / 2, 16, 0};
  const bool ok = std::fwrite(&h, sizeof h, 1, fp) == 1;
  return (std::fclose(fp) == 0 && ok) ? 0 : -1;
}

int phb_dist_bin_append(const char* path, const uint16_t* entries, uint64_t n) {
  if (!n) return 0;
  FILE* fp = (path && entries) ? std::fopen(path, "ab") : nullptr;
  if (!fp) return -1;
  const bool ok = std::fwrite(entries, 2, n, fp) == n;
  return (std::fclose(fp) == 0 && ok) ? 0 : -1;
}

int phb_dist_bin_read(const char* path, uint32_t* T, uint16_t** upper) {
  FILE* fp = (path && T && upper) ? std::fopen(path, "rb") : nullptr;
  if (!fp) return -1;
  DistBinHeader h;
  int rc = -1;
  uint16_t* buf = nullptr;
  if (std::fread(&h, sizeof h, 1, fp) == 1 && !std::memcmp(h.magic, "PHYBINRF", 8) && h.version == 1 && h.bits == 16 &&
      h.layout == 0 && h.P == (uint64_t)h.T * (h.T ? h.T - 1 : 0) / 2) {
    buf = (uint16_t*)std::malloc(std::max<uint64_t>(h.P, 1) * 2);
    if (buf && std::fread(buf, 2, h.P, fp) == h.P && std::fgetc(fp) == EOF) {
      *T = h.T;
      *upper = buf;
      rc = 0;
    }
  }
  if (rc) std::free(buf);
  std::fclose(fp);
  return rc;
}

void phb_free_buf(void* p) { std::free(p); }

}  // extern "C"
