// phybin_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A CPU restatement of the Robinson-Foulds distance-matrix path of rrnewton/PhyBin
// (Bio.Phylogeny.PhyBin.RFDistance: hashRF / naiveDistMatrix and the functions they call).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library; the product path (phybin_b200/) never does.
//
// The Haskell reference cannot be compiled in this image (no ghc/cabal/stack), so this file
// restates it function by function with the same data structures (sets of integer label sets
// ordered lexicographically on their ascending element list, a Map from bipartition to tree-id
// set, a ragged lower-triangular matrix).  Each function cites the reference lines it follows.
//
// Parity pin: the one known-answer test the reference asserts for this path
// (TestAll.hs:22-32, naiveDistMatrix on the three t30 trees == [[],[0],[1,0]]) plus the
// fixture-derived values of SURVEY.md section 4; see tests/test_oracle_golden.py.
//
// Build: make -C oracle   (g++ -O2 -shared -fPIC; no dependencies)

#ifdef _OPENMP
#include <omp.h>
#endif
#include <algorithm>
#include <cctype>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iterator>
#include <map>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace orc {

// ---------------------------------------------------------------------------------------------
// DenseLabelSet = Data.IntSet (RFDistance.hs:116-134).  An IntSet is represented by its ascending
// element list; std::vector's operator< is lexicographic, which is exactly the Ord instance of
// IntSet (compare on toAscList) that normBip's `min` relies on (RFDistance.hs:239).
// ---------------------------------------------------------------------------------------------
using IntSet = std::vector<int>;

static IntSet markLabel(int lab, const IntSet& s) {  // RFDistance.hs:117  SI.insert
  IntSet r = s;
  auto it = std::lower_bound(r.begin(), r.end(), lab);
  if (it == r.end() || *it != lab) r.insert(it, lab);
  return r;
}
static IntSet denseUnions(const std::vector<IntSet>& ls) {  // RFDistance.hs:120  SI.unions
  IntSet r;
  for (const IntSet& s : ls) r.insert(r.end(), s.begin(), s.end());
  std::sort(r.begin(), r.end());
  r.erase(std::unique(r.begin(), r.end()), r.end());
  return r;
}
static bool member(int x, const IntSet& s) { return std::binary_search(s.begin(), s.end(), x); }

// RFDistance.hs:127-131: { ix | ix <- [0..size-1], ix not in bip }.  NB the universe is [0,size),
// not the tree's own label set.
static IntSet invertDense(int size, const IntSet& bip) {
  IntSet r;
  for (int ix = 0; ix < size; ++ix)
    if (!member(ix, bip)) r.push_back(ix);
  return r;
}

// ---------------------------------------------------------------------------------------------
// NewickTree (CoreTypes.hs:62-65): rose tree; leaves carry an Int label, interior nodes a child
// list.  Decorations (bootstrap, branch length) are parsed but play no role on this path.
// ---------------------------------------------------------------------------------------------
struct Tree {
  bool leaf = false;
  int label = 0;       // leaves only
  std::string name;    // leaves only, before label assignment
  std::vector<Tree> kids;
};

static int numLeaves(const Tree& t) {  // CoreTypes.hs:289-291 (counts duplicate labels twice)
  if (t.leaf) return 1;
  int s = 0;
  for (const Tree& k : t.kids) s += numLeaves(k);
  return s;
}

// ---------------------------------------------------------------------------------------------
// Newick grammar (Parser.hs:127-202).  Whitespace is filtered from the whole input first
// (Parser.hs:35); interior names are parsed and ignored (:152); metadata is ":len[boot]" with
// len = sciNotation | number (:168-195); names are many1 (noneOf "()[]:;, \t\n\r\f\v") or empty.
// ---------------------------------------------------------------------------------------------
struct Parser {
  const std::string& s;
  size_t p = 0;
  int name_prefix;  // NameHack: `take n` (Main.hs:290); <0 = id
  explicit Parser(const std::string& str, int np) : s(str), name_prefix(np) {}

  [[noreturn]] void fail(const char* what) const {
    throw std::runtime_error(std::string("parse error at offset ") + std::to_string(p) + ": " + what);
  }
  bool at(char c) const { return p < s.size() && s[p] == c; }
  static bool reserved(char c) { return std::strchr("()[]:;,", c) != nullptr || std::isspace((unsigned char)c); }

  std::string name() {  // Parser.hs:201-202
    size_t b = p;
    while (p < s.size() && !reserved(s[p])) ++p;
    return s.substr(b, p - b);
  }
  bool digits() {
    size_t b = p;
    while (p < s.size() && std::isdigit((unsigned char)s[p])) ++p;
    return p > b;
  }
  bool sciNotation() {  // Parser.hs:187-195 (wrapped in `try`, so it backtracks on failure)
    size_t save = p;
    if (!digits()) { p = save; return false; }
    if (at('.')) {
      size_t q = p; ++p;
      if (!digits()) p = q;  // `option "0" $ try ...`
    }
    if (!at('e')) { p = save; return false; }
    ++p;
    if (at('-')) ++p;
    if (!digits()) { p = save; return false; }
    return true;
  }
  void number() {  // Parser.hs:179-184
    if (at('-')) ++p;
    if (!digits()) fail("expected a number");
    if (at('.')) {
      size_t q = p; ++p;
      if (!digits()) p = q;
    }
  }
  void branchMetadat() {  // Parser.hs:166-177
    if (!at(':')) return;
    ++p;
    if (!sciNotation()) number();
    if (at('[')) {
      ++p;
      if (!digits()) fail("expected bootstrap digits");
      if (!at(']')) fail("expected ]");
      ++p;
    }
  }
  Tree subtree() {  // Parser.hs:135-136
    if (at('(')) {  // internal (:146-152)
      ++p;
      Tree t;
      for (;;) {  // branchset (:154-158): at least one branch, comma separated
        Tree b = subtree();
        branchMetadat();
        t.kids.push_back(std::move(b));
        if (at(',')) { ++p; continue; }
        break;
      }
      if (!at(')')) fail("expected )");
      ++p;
      (void)name();  // IGNORED, internal names
      return t;
    }
    Tree l;  // leaf (:141-144); the name may be empty
    l.leaf = true;
    l.name = name();
    if (name_prefix >= 0 && (int)l.name.size() > name_prefix) l.name.resize(name_prefix);
    return l;
  }
  Tree newick() {  // newick_parser (:127-133)
    Tree t = subtree();
    branchMetadat();
    if (!at(';')) fail("expected ;");
    ++p;
    return t;
  }
};

struct LabelTable {
  std::map<std::string, int> seen;  // name -> label
  std::vector<std::string> names;   // label -> name
};

// extractLabelTable / loop2 (Parser.hs:85-110): next label = current table size; the children of
// an interior node are visited RIGHT-TO-LEFT (P.foldr), trees of one file left-to-right (loop1).
static void assignLabels(LabelTable& tb, Tree& t) {
  if (t.leaf) {
    auto it = tb.seen.find(t.name);
    if (it != tb.seen.end()) {
      t.label = it->second;
    } else {
      t.label = (int)tb.names.size();
      tb.seen.emplace(t.name, t.label);
      tb.names.push_back(t.name);
    }
    return;
  }
  for (size_t i = t.kids.size(); i-- > 0;) assignLabels(tb, t.kids[i]);
}

struct Forest {
  LabelTable labels;
  std::vector<Tree> trees;
  std::string error;
};

// parseNewicks (Parser.hs:57-75): P.foldr over the (file, contents) pairs, i.e. the LAST file is
// parsed (and labelled) first, but the resulting tree list keeps command-line order.
static std::unique_ptr<Forest> parseNewicks(const std::vector<std::string>& texts, int name_prefix) {
  auto f = std::make_unique<Forest>();
  std::vector<std::vector<Tree>> per_file(texts.size());
  for (size_t fi = texts.size(); fi-- > 0;) {
    std::string filtered;
    for (char c : texts[fi])
      if (!std::isspace((unsigned char)c)) filtered.push_back(c);  // Parser.hs:35
    Parser ps(filtered, name_prefix);
    std::vector<Tree> trs;
    do {  // many1 newick_parser
      trs.push_back(ps.newick());
    } while (ps.p < filtered.size());
    for (Tree& t : trs) assignLabels(f->labels, t);
    per_file[fi] = std::move(trs);
  }
  for (auto& v : per_file)
    for (Tree& t : v) f->trees.push_back(std::move(t));
  return f;
}

// ---------------------------------------------------------------------------------------------
// Bipartitions
// ---------------------------------------------------------------------------------------------

// normBip (RFDistance.hs:229-239)
static IntSet normBip(int totsize, const IntSet& bip) {
  int halfSize = totsize / 2;  // `quot`, operands are non-negative
  IntSet flipped = invertDense(totsize, bip);
  int sz = (int)bip.size();
  if (sz < halfSize) return bip;
  if (sz > halfSize) return flipped;
  return std::min(bip, flipped);  // the painful case: tie-break on the IntSet order
}

// labelBips + foldBips (RFDistance.hs:207-225, 242-244): every node contributes a list of sets --
// a leaf the singleton of its label, an interior node the leaf set of each child -- and every set
// is normalised with size = numLeaves of the whole tree.  `emit` sees each normalised set.
static IntSet leafSetOf(const Tree& t) {
  if (t.leaf) return IntSet{t.label};
  std::vector<IntSet> parts;
  for (const Tree& k : t.kids) parts.push_back(leafSetOf(k));
  return denseUnions(parts);
}
template <class F>
static void foldBipsRec(const Tree& t, int size, F& emit) {
  if (t.leaf) {
    emit(normBip(size, markLabel(t.label, IntSet{})));
    return;
  }
  for (const Tree& k : t.kids) emit(normBip(size, leafSetOf(k)));
  for (const Tree& k : t.kids) foldBipsRec(k, size, emit);
}

using BipSet = std::set<IntSet>;

// allBips (RFDistance.hs:247-248): the Set of normalised sets, filtered to bipSize > 1 AFTER
// normalisation.
static BipSet allBips(const Tree& t) {
  BipSet acc;
  int size = numLeaves(t);
  auto emit = [&](IntSet b) { acc.insert(std::move(b)); };
  foldBipsRec(t, size, emit);
  for (auto it = acc.begin(); it != acc.end();)
    it = (it->size() > 1) ? std::next(it) : acc.erase(it);
  return acc;
}

// pruneTreeLeaves (PreProcessor.hs:21-32): keep leaves in `set`; drop interiors that lose all
// children; splice out interiors left with one child.  Returns false for Nothing.
static bool pruneTreeLeaves(const std::set<int>& set, const Tree& t, Tree& out) {
  if (t.leaf) {
    if (!set.count(t.label)) return false;
    out = t;
    return true;
  }
  std::vector<Tree> kept;
  for (const Tree& k : t.kids) {
    Tree o;
    if (pruneTreeLeaves(set, k, o)) kept.push_back(std::move(o));
  }
  if (kept.empty()) return false;
  if (kept.size() == 1) { out = std::move(kept[0]); return true; }
  out = Tree{};
  out.kids = std::move(kept);
  return true;
}

static void treeLabels(const Tree& t, std::set<int>& acc) {  // RFDistance.hs:200-203
  if (t.leaf) { acc.insert(t.label); return; }
  for (const Tree& k : t.kids) treeLabels(k, acc);
}

static size_t setDifferenceSize(const BipSet& a, const BipSet& b) {
  size_t n = 0;
  for (const IntSet& x : a)
    if (!b.count(x)) ++n;
  return n;
}

// DistanceMatrix (RFDistance.hs:156): row i holds d(i,0..i-1).  Stored flat: (i,j) at i(i-1)/2+j.
using Matrix = std::vector<int64_t>;
static inline size_t lowerIx(size_t i, size_t j) { return i * (i - 1) / 2 + j; }

// naiveDistMatrix (RFDistance.hs:168-198)
static Matrix naiveDistMatrix(const std::vector<Tree>& trees) {
  size_t sz = trees.size();
  std::vector<std::set<int>> labelSets(sz);
  std::vector<BipSet> eachbips(sz);
  for (size_t i = 0; i < sz; ++i) {
    treeLabels(trees[i], labelSets[i]);
    eachbips[i] = allBips(trees[i]);
  }
  Matrix mat(sz * (sz - (sz ? 1 : 0)) / 2, 0);
  for (size_t i = 0; i < sz; ++i)
    for (size_t j = 0; j < i; ++j) {
      const std::set<int>& inI = labelSets[i];
      const std::set<int>& inJ = labelSets[j];
      std::set<int> inBoth;
      std::set_intersection(inI.begin(), inI.end(), inJ.begin(), inJ.end(),
                            std::inserter(inBoth, inBoth.begin()));
      if (inBoth.empty()) { mat[lowerIx(i, j)] = 0; continue; }  // :195-196
      BipSet bi, bj;
      const BipSet* bipsI = &eachbips[i];
      const BipSet* bipsJ = &eachbips[j];
      if (inBoth.size() != inI.size()) {  // memoisation test, :186-188
        Tree pr;
        pruneTreeLeaves(inBoth, trees[i], pr);
        bi = allBips(pr);
        bipsI = &bi;
      }
      if (inBoth.size() != inJ.size()) {  // :189-191
        Tree pr;
        pruneTreeLeaves(inBoth, trees[j], pr);
        bj = allBips(pr);
        bipsJ = &bj;
      }
      mat[lowerIx(i, j)] =
          (int64_t)(setDifferenceSize(*bipsI, *bipsJ) + setDifferenceSize(*bipsJ, *bipsI));  // :193-197
    }
  return mat;
}

// hashRF (RFDistance.hs:284-341).  `build`: Map bip -> IntSet of tree ids (:289-297).  `ingest`:
// for every bip, bump (max,min) for each pair in haveIt x dontHave (:316-338).  num_taxa is only
// handed to mkSingleDense, where the IntSet variant ignores it (:295, :119).
static Matrix hashRF(int /*num_taxa*/, const std::vector<Tree>& trees, size_t* n_unique = nullptr,
                     double* build_s = nullptr, double* ingest_s = nullptr) {
  using clk = std::chrono::steady_clock;
  auto t0 = clk::now();
  int num_trees = (int)trees.size();
  std::map<IntSet, IntSet> bipTable;
  for (int ix = 0; ix < num_trees; ++ix) {
    BipSet bips = allBips(trees[ix]);
    for (const IntSet& bip : bips) bipTable[bip].push_back(ix);  // ascending ids: already an IntSet
  }
  auto t1 = clk::now();
  size_t T = (size_t)num_trees;
  Matrix matr(T * (T - (T ? 1 : 0)) / 2, 0);
  std::vector<char> have(T);
  for (const auto& kv : bipTable) {
    const IntSet& haveIt = kv.second;
    std::fill(have.begin(), have.end(), 0);
    for (int t : haveIt) have[t] = 1;
    for (int i : haveIt)
      for (int j = 0; j < num_trees; ++j) {  // dontHave = invertDense num_trees haveIt
        if (have[j]) continue;
        if (j < i) ++matr[lowerIx(i, j)];   // bumpMatr (:316-317)
        else       ++matr[lowerIx(j, i)];
      }
  }
  auto t2 = clk::now();
  if (n_unique) *n_unique = bipTable.size();
  if (build_s) *build_s = std::chrono::duration<double>(t1 - t0).count();
  if (ingest_s) *ingest_s = std::chrono::duration<double>(t2 - t1).count();
  return matr;
}

// The same algorithm on ALL HOST CORES (variant ii): the reference's ingest loop is a sequential
// runST (RFDistance.hs:300-341, "Not concurrency safe yet" :319); this is what a straightforward parallelisation of
// it gives -- bipartition extraction parallel over the trees, the bump loop parallel over the table's bipartitions
// with atomic increments.  Same result as hashRF above (tests/test_oracle_golden.py); used only as a CPU baseline.
static Matrix hashRFAllCores(const std::vector<Tree>& trees, size_t* n_unique = nullptr, double* build_s = nullptr,
                             double* ingest_s = nullptr) {
  const int num_trees = (int)trees.size();
  std::vector<BipSet> per_tree((size_t)num_trees);
  const bool trace = getenv("ORC_TRACE") != nullptr;
  auto t0 = std::chrono::steady_clock::now();
  double secs[3] = {0, 0, 0};
  int phase = 0;
  auto lap = [&](const char* what) {
    auto t1 = std::chrono::steady_clock::now();
    secs[phase++] = std::chrono::duration<double>(t1 - t0).count();
    if (trace) fprintf(stderr, "[orc all-cores] %s %.3f s\n", what, secs[phase - 1]);
    t0 = t1;
  };
#pragma omp parallel for schedule(dynamic, 16)
  for (int ix = 0; ix < num_trees; ++ix) per_tree[ix] = allBips(trees[ix]);
  lap("allBips");
  // the table, sharded by a hash of the bipartition: every thread owns the shards s = thread (mod threads) and
  // inserts, in tree order, the bipartitions that fall into them (same lists as the sequential insertWith)
  int shards = 1;
#ifdef _OPENMP
  shards = omp_get_max_threads();
#endif
  std::vector<std::map<IntSet, IntSet>> tables((size_t)shards);
#pragma omp parallel for schedule(static, 1)
  for (int s = 0; s < shards; ++s)
    for (int ix = 0; ix < num_trees; ++ix)
      for (const IntSet& bip : per_tree[ix]) {
        uint64_t h = 0x9E3779B97F4A7C15ull;
        for (int l : bip) h = (h ^ (uint64_t)l) * 0x100000001B3ull;
        if ((int)((h >> 20) % (uint64_t)shards) == s) tables[(size_t)s][bip].push_back(ix);
      }
  std::vector<const IntSet*> lists;
  size_t n_keys = 0;
  for (const auto& tb : tables) n_keys += tb.size();
  lists.reserve(n_keys);
  for (const auto& tb : tables)
    for (const auto& kv : tb) lists.push_back(&kv.second);
  lap("table");
  const size_t T = (size_t)num_trees;
  Matrix matr(T * (T - (T ? 1 : 0)) / 2, 0);
  int threads = 1;
#ifdef _OPENMP
  threads = omp_get_max_threads();
#endif
  // every thread bumps a private matrix (no atomics), summed afterwards; atomics on the shared one if they would not fit
  const bool priv = matr.size() * sizeof(uint32_t) * (size_t)threads <= (size_t)8 << 30;
  std::vector<std::vector<uint32_t>> mine(priv ? (size_t)threads : 0);
#pragma omp parallel num_threads(threads)
  {
    int me = 0;
#ifdef _OPENMP
    me = omp_get_thread_num();
#endif
    uint32_t* m32 = nullptr;
    if (priv) {
      mine[(size_t)me].assign(matr.size(), 0);
      m32 = mine[(size_t)me].data();
    }
    std::vector<char> have(T);
#pragma omp for schedule(dynamic, 64)
    for (long long b = 0; b < (long long)lists.size(); ++b) {
      const IntSet& haveIt = *lists[(size_t)b];
      std::fill(have.begin(), have.end(), 0);
      for (int t : haveIt) have[t] = 1;
      for (int i : haveIt)
        for (int j = 0; j < num_trees; ++j) {
          if (have[j]) continue;
          const size_t at = j < i ? lowerIx(i, j) : lowerIx(j, i);
          if (priv) {
            ++m32[at];
          } else {
            int64_t* cell = &matr[at];
#pragma omp atomic
            ++*cell;
          }
        }
    }
    if (priv) {
#pragma omp for schedule(static)
      for (long long c = 0; c < (long long)matr.size(); ++c) {
        int64_t v = 0;
        for (int t = 0; t < threads; ++t) v += mine[(size_t)t][(size_t)c];
        matr[(size_t)c] = v;
      }
    }
  }
  lap("matrix");
  if (n_unique) *n_unique = n_keys;
  if (build_s) *build_s = secs[0] + secs[1];
  if (ingest_s) *ingest_s = secs[2];
  return matr;
}

// A DERIVED (not literal) evaluation of the same matrix used only to make large parity cases
// finish in seconds: d(i,j) = |B_i| + |B_j| - 2 |B_i n B_j|, the closed form of the bump loop
// (each bip in exactly one of the two trees bumps the pair once).  Checked against the literal
// hashRF above in tests/test_oracle_golden.py before it is trusted anywhere.
static Matrix hashRFClosedForm(const std::vector<Tree>& trees, size_t* n_unique = nullptr) {
  size_t T = trees.size();
  std::map<IntSet, std::vector<int>> bipTable;
  std::vector<int64_t> nb(T);
  for (size_t ix = 0; ix < T; ++ix) {
    BipSet bips = allBips(trees[ix]);
    nb[ix] = (int64_t)bips.size();
    for (const IntSet& bip : bips) bipTable[bip].push_back((int)ix);
  }
  Matrix m(T * (T - (T ? 1 : 0)) / 2);
  for (size_t i = 1; i < T; ++i)
    for (size_t j = 0; j < i; ++j) m[lowerIx(i, j)] = nb[i] + nb[j];
  for (const auto& kv : bipTable) {
    const std::vector<int>& mem = kv.second;
    for (size_t a = 0; a < mem.size(); ++a)
      for (size_t b = 0; b < a; ++b) m[lowerIx(mem[a], mem[b])] -= 2;
  }
  if (n_unique) *n_unique = bipTable.size();
  return m;
}

// printDistMat (RFDistance.hs:352-361)
static std::string printDistMat(const Matrix& mat, size_t T) {
  std::string o = "Robinson-Foulds distance (matrix format):\n";
  o += "-----------------------------------------\n";
  for (size_t i = 0; i < T; ++i) {
    for (size_t j = 0; j < i; ++j) {
      o += std::to_string(mat[lowerIx(i, j)]);
      o += " ";
    }
    o += "0\n";
  }
  o += "-----------------------------------------\n";
  return o;
}

}  // namespace orc

// ---------------------------------------------------------------------------------------------
// C interface for ctypes (tests/, smoke(), bench.py cpu_baseline only).
// ---------------------------------------------------------------------------------------------
extern "C" {

struct orc_forest {
  std::unique_ptr<orc::Forest> f;
  std::string err;
};

orc_forest* orc_parse(int nfiles, const char* const* texts, int name_prefix) {
  auto* h = new orc_forest;
  try {
    std::vector<std::string> v;
    for (int i = 0; i < nfiles; ++i) v.emplace_back(texts[i]);
    h->f = orc::parseNewicks(v, name_prefix);
  } catch (const std::exception& e) {
    h->err = e.what();
    h->f.reset();
  }
  return h;
}
void orc_free(orc_forest* h) { delete h; }
const char* orc_error(const orc_forest* h) { return h->err.empty() ? nullptr : h->err.c_str(); }
int orc_num_trees(const orc_forest* h) { return h->f ? (int)h->f->trees.size() : -1; }
int orc_num_labels(const orc_forest* h) { return h->f ? (int)h->f->labels.names.size() : -1; }
const char* orc_label_name(const orc_forest* h, int lab) { return h->f->labels.names[lab].c_str(); }
int orc_tree_num_leaves(const orc_forest* h, int t) { return orc::numLeaves(h->f->trees[t]); }

// Keep only trees whose numLeaves == expected (the HashRF input filter, PhyBin.hs:187-194).
int orc_filter_num_leaves(orc_forest* h, int expected) {
  std::vector<orc::Tree> keep;
  for (auto& t : h->f->trees)
    if (orc::numLeaves(t) == expected) keep.push_back(std::move(t));
  h->f->trees = std::move(keep);
  return (int)h->f->trees.size();
}

// allBips of one tree, flattened: returns the number of bips; sizes[k] = |bip k|, labels holds
// the ascending elements back to back (in Set order).  Pass nulls to query counts only.
int orc_all_bips(const orc_forest* h, int t, int* sizes, int* labels, int cap_labels) {
  orc::BipSet b = orc::allBips(h->f->trees[t]);
  int k = 0, pos = 0;
  for (const auto& s : b) {
    if (sizes) sizes[k] = (int)s.size();
    for (int x : s) {
      if (labels && pos < cap_labels) labels[pos] = x;
      ++pos;
    }
    ++k;
  }
  return (cap_labels >= pos || !labels) ? k : -pos;
}
int orc_all_bips_total_labels(const orc_forest* h, int t) {
  orc::BipSet b = orc::allBips(h->f->trees[t]);
  int pos = 0;
  for (const auto& s : b) pos += (int)s.size();
  return pos;
}

// threads the all-cores variant uses (1 when the oracle was built without OpenMP)
int orc_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// out: T(T-1)/2 int64, (i,j), j<i at i(i-1)/2+j.  variant 0 = literal hashRF, 1 = closed form, 2 = the literal
// algorithm on all host cores (orc_max_threads() of them).
int orc_hashrf(const orc_forest* h, int num_taxa, int variant, int64_t* out, int64_t* n_unique,
               double* build_s, double* ingest_s) {
  try {
    size_t nu = 0;
    double b = 0, g = 0;
    orc::Matrix m = variant == 0   ? orc::hashRF(num_taxa, h->f->trees, &nu, &b, &g)
                    : variant == 2 ? orc::hashRFAllCores(h->f->trees, &nu, &b, &g)
                                   : orc::hashRFClosedForm(h->f->trees, &nu);
    if (out) std::copy(m.begin(), m.end(), out);
    if (n_unique) *n_unique = (int64_t)nu;
    if (build_s) *build_s = b;
    if (ingest_s) *ingest_s = g;
    return 0;
  } catch (const std::exception&) {
    return -1;
  }
}

int orc_naive(const orc_forest* h, int64_t* out) {
  try {
    orc::Matrix m = orc::naiveDistMatrix(h->f->trees);
    std::copy(m.begin(), m.end(), out);
    return 0;
  } catch (const std::exception&) {
    return -1;
  }
}

// printDistMat into a malloc'd C string (caller frees with orc_free_str).
char* orc_print_dist_mat(const int64_t* lower, int T) {
  size_t n = (size_t)T * (T > 0 ? T - 1 : 0) / 2;
  orc::Matrix m(lower, lower + n);
  std::string s = orc::printDistMat(m, (size_t)T);
  char* r = (char*)std::malloc(s.size() + 1);
  std::memcpy(r, s.c_str(), s.size() + 1);
  return r;
}
void orc_free_str(char* s) { std::free(s); }

}  // extern "C"
