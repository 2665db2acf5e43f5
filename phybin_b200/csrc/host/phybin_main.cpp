// phybin -- command-line host for the B200 RF distance-matrix path.
//
// Flag-compatible with the in-scope subset of the reference CLI (Main.hs:79-176) and mirroring the
// driver's flow for that path (PhyBin.hs:84-243, doCluster :343-364): read Newick files, build the
// global label table, (HashRF) keep trees with numLeaves == #taxa, compute the matrix ON THE GPU
// through libphybin_rf.so, print the reference's log lines, write DIR/distance_matrix.txt in
// printDistMat format (RFDistance.hs:352-361) and, with --rfdist, the same text on stdout.
// Everything else of the reference CLI (binning, graphviz, consensus, --prune, ...) stays with the
// Haskell program and is rejected here with a message.  There is no CPU path: without a B200 the
// program exits with the library's error.
#include <sys/stat.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <cmath>
#include <string>
#include <vector>

#include "phybin_host.h"
#include "phybin_rf.h"

namespace {

struct Options {
  bool tolerant = false, rfdist = false, host_bips = false, binary = false;
  int linkage = 2;  // C.UPGMA, the reference's default (CoreTypes.hs:265); 0 single, 1 complete
  int editdist = -1, name_prefix = -1, variant = PRF_VARIANT_AUTO;
  std::string out_dir = "phybin_out";
  std::vector<std::string> files;
};

[[noreturn]] void die(const std::string& m) {
  std::fprintf(stderr, "phybin: %s\n", m.c_str());
  std::exit(1);
}

bool starts(const char* a, const char* pfx, const char** rest) {
  size_t n = std::strlen(pfx);
  if (std::strncmp(a, pfx, n)) return false;
  *rest = a + n;
  return true;
}

Options parse_args(int argc, char** argv) {
  Options o;
  for (int i = 1; i < argc; ++i) {
    const char* a = argv[i];
    const char* v = nullptr;
    auto need = [&](const char* what) -> const char* {
      if (i + 1 >= argc) die(std::string("option ") + what + " needs an argument");
      return argv[++i];
    };
    if (!std::strcmp(a, "--hashrf")) o.tolerant = false;
    else if (!std::strcmp(a, "--tolerant")) o.tolerant = true;
    else if (!std::strcmp(a, "--rfdist")) o.rfdist = true;
    else if (!std::strcmp(a, "--single")) o.linkage = 0;  // Main.hs:96-98
    else if (!std::strcmp(a, "--complete")) o.linkage = 1;
    else if (!std::strcmp(a, "--UPGMA")) o.linkage = 2;
    else if (starts(a, "--editdist=", &v)) o.editdist = std::atoi(v);
    else if (!std::strcmp(a, "--editdist")) o.editdist = std::atoi(need(a));
    else if (starts(a, "--output=", &v)) o.out_dir = v;
    else if (!std::strcmp(a, "-o") || !std::strcmp(a, "--output")) o.out_dir = need(a);
    else if (starts(a, "--nameprefix=", &v)) o.name_prefix = std::atoi(v);
    else if (!std::strcmp(a, "-p") || !std::strcmp(a, "--nameprefix")) o.name_prefix = std::atoi(need(a));
    else if (!std::strcmp(a, "--sparse")) o.variant = PRF_VARIANT_SPARSE;   // not in the reference: force a kernel variant
    else if (!std::strcmp(a, "--dense")) o.variant = PRF_VARIANT_DENSE;
    else if (!std::strcmp(a, "--binary-matrix")) o.binary = true;  // not in the reference: also write distance_matrix.bin
    else if (!std::strcmp(a, "--host-bips")) o.host_bips = true;  // not in the reference: allBips in the C++ host code
    else if (!std::strcmp(a, "-V") || !std::strcmp(a, "--version")) {
      std::printf("phybin (B200 RF path) -- %s\n", prf_version());
      std::exit(0);
    } else if (a[0] == '-' && a[1])
      die(std::string("option ") + a + " belongs to the part of PhyBin that stays in the Haskell host (see DESIGN.md)");
    else o.files.push_back(a);
  }
  if (o.files.empty()) die("no input files");
  return o;
}

std::string slurp(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) die("cannot read " + path);
  std::ostringstream ss;
  ss << f.rdbuf();
  return ss.str();
}

// prf_allbips on the forest's own arrays.  PRF_E_UNSUPPORTED (a tree too large for one CTA's shared memory)
// is reported with the way out; nothing silently changes path.
int extract_on_device(prf_ctx* ctx, const phb_forest* forest, uint32_t T, uint32_t W, int mode) {
  const int32_t* label = nullptr;
  const uint32_t *parent = nullptr, *leaves = nullptr;
  const uint64_t* begin = nullptr;
  if (phb_tree_arrays(forest, &label, &parent, &begin, &leaves, nullptr)) die("cannot read the parsed trees");
  const int rc = prf_allbips(ctx, label, parent, begin, leaves, T, W, PRF_MEM_HOST, mode, nullptr);
  if (rc == PRF_E_UNSUPPORTED)
    die(std::string(prf_last_error(ctx)) + " -- rerun with --host-bips to extract the bipartitions on the host");
  return rc;
}

}  // namespace

int main(int argc, char** argv) {
  Options opt = parse_args(argc, argv);
  std::vector<std::string> texts;
  for (auto& f : opt.files) texts.push_back(slurp(f));
  std::vector<const char*> ptrs;
  for (auto& t : texts) ptrs.push_back(t.c_str());
  phb_forest* forest = phb_parse((int)ptrs.size(), ptrs.data(), opt.name_prefix);
  if (const char* e = phb_error(forest)) die(e);
  const uint32_t n_taxa = phb_num_labels(forest);
  std::printf("Parsed %u trees, %u taxa.\n", phb_num_trees(forest), n_taxa);
  // treename (Parser.hs:60-68): base name of "<file>[_<index in file>]" plus "_<global index>" when there is
  // more than one tree in all
  std::vector<std::string> names;
  {
    std::vector<std::string> raw;
    for (size_t i = 0; i < opt.files.size(); ++i) {
      const uint32_t k = phb_file_num_trees(forest, (uint32_t)i);
      for (uint32_t j = 0; j < k; ++j) raw.push_back(k > 1 ? opt.files[i] + "_" + std::to_string(j) : opt.files[i]);
    }
    for (size_t g = 0; g < raw.size(); ++g) {
      std::string b = raw[g].substr(raw[g].find_last_of('/') + 1);  // takeBaseName: no directory, no extension
      const size_t dot = b.find_last_of('.');
      if (dot != std::string::npos) b.erase(dot);
      names.push_back(raw.size() > 1 ? b + "_" + std::to_string(g) : b);
    }
  }
  if (!opt.tolerant) {  // HashRF keeps only trees with every taxon (PhyBin.hs:187-194)
    std::vector<std::string> kept_names;
    for (uint32_t t = 0; t < phb_num_trees(forest); ++t)
      if (phb_tree_num_leaves(forest, t) == n_taxa && t < names.size()) kept_names.push_back(names[t]);
    names.swap(kept_names);
    const uint32_t before = phb_num_trees(forest), kept = phb_filter_num_leaves(forest, n_taxa);
    if (kept != before) std::printf("WARNING: dropped %u trees whose leaf count differs from %u\n", before - kept, n_taxa);
  }
  const bool dup_leaves = opt.tolerant && phb_has_duplicate_leaves(forest);  // numLeaves counts them twice (CoreTypes.hs:289-291)
  const uint32_t T = phb_num_trees(forest), W = phb_words(forest);
  const uint64_t P = (uint64_t)T * (T ? T - 1 : 0) / 2;

  prf_ctx* ctx = prf_create(0);
  if (!ctx) die(std::string("cannot use the GPU: ") + prf_last_error(nullptr));
  const bool text_out = T <= 20000;  // the text dump is O(T^2) `show` in the reference too; beyond this only binary
  std::vector<uint16_t> upper(text_out && P ? P : 1);
  std::vector<uint32_t> mask(opt.editdist >= 0 ? (P + 31) / 32 + 1 : 1);
  prf_stats st{};
  const auto t0 = std::chrono::steady_clock::now();
  int rc;
  if (!opt.tolerant) {
    std::printf(" Using HashRF-style algorithm...\n");  // PhyBin.hs:346
    if (opt.host_bips) {
      uint64_t *bips = nullptr, *off = nullptr, nnz = 0;
      if (phb_allbips(forest, W, 0, &bips, &off, &nnz)) die("bipartition extraction failed");
      rc = prf_hashrf_load(ctx, bips, off, T, W, PRF_MEM_HOST, &st);
      phb_free_buf(bips);
      phb_free_buf(off);
    } else {  // allBips on the device from the parsed trees' flat arrays
      const uint64_t *d_keys = nullptr, *d_off = nullptr;
      rc = extract_on_device(ctx, forest, T, W, PRF_EXTRACT_BIPS);
      if (!rc) rc = prf_allbips_result(ctx, &d_keys, &d_off, nullptr, nullptr);
      if (!rc) rc = prf_hashrf_load(ctx, d_keys, d_off, T, W, PRF_MEM_DEVICE, &st);
    }
    if (!rc) rc = prf_hashrf_compute(ctx, 0, P, opt.editdist, opt.variant, nullptr, nullptr, &st);
  } else {
    if (dup_leaves) {
      uint64_t *clades = nullptr, *outs = nullptr, *off = nullptr, *leaf = nullptr, *doff = nullptr, nnz = 0;
      uint32_t *dlab = nullptr, *dext = nullptr;
      if (phb_raw_clades_ex(forest, W, &clades, &outs, &off, &leaf, &doff, &dlab, &dext, &nnz)) die("clade extraction failed");
      rc = prf_tolerant_load_dups(ctx, clades, outs, off, leaf, doff, dlab, dext, T, W, PRF_MEM_HOST, &st);
      for (void* p : {(void*)clades, (void*)outs, (void*)off, (void*)leaf, (void*)doff, (void*)dlab, (void*)dext}) phb_free_buf(p);
    } else if (opt.host_bips) {
      uint64_t *clades = nullptr, *off = nullptr, *leaf = nullptr, nnz = 0;
      if (phb_raw_clades(forest, W, 0, &clades, &off, &leaf, &nnz)) die("clade extraction failed");
      rc = prf_tolerant_load(ctx, clades, off, leaf, T, W, PRF_MEM_HOST, &st);
      phb_free_buf(clades);
      phb_free_buf(off);
      phb_free_buf(leaf);
    } else {
      const uint64_t *d_keys = nullptr, *d_off = nullptr, *d_leaf = nullptr;
      rc = extract_on_device(ctx, forest, T, W, PRF_EXTRACT_CLADES);
      if (!rc) rc = prf_allbips_result(ctx, &d_keys, &d_off, &d_leaf, nullptr);
      if (!rc) rc = prf_tolerant_load(ctx, d_keys, d_off, d_leaf, T, W, PRF_MEM_DEVICE, &st);
    }
    if (!rc) rc = prf_tolerant_compute(ctx, 0, P, opt.editdist, nullptr, nullptr, &st);
  }
  if (!rc && P && text_out) rc = prf_fetch(ctx, 0, P, upper.data());
  if (!rc && P && opt.editdist >= 0) rc = prf_fetch_mask(ctx, 0, P, mask.data());
  if (rc) die(std::string("phybin_rf error: ") + prf_last_error(ctx));
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  std::printf(" Built matrix for dim %u\n", T);  // RFDistance.hs:314
  std::printf("Time to compute distance matrix: %.6fs (%s kernel %.3f ms, %llu unique bipartitions)\n", secs,
              st.variant_used == PRF_VARIANT_DENSE ? "dense tcgen05" : (opt.tolerant ? "tolerant" : "sparse"), st.ms_matrix,
              (unsigned long long)st.n_unique);

  mkdir(opt.out_dir.c_str(), 0777);
  if (opt.binary) {  // streamed from the device in 64 MB pieces: the host never holds the matrix
    const std::string path = opt.out_dir + "/distance_matrix.bin";
    if (phb_dist_bin_begin(path.c_str(), T)) die("cannot write " + path);
    const uint64_t chunk = 1ull << 25;
    std::vector<uint16_t> piece(std::min<uint64_t>(chunk, P ? P : 1));
    for (uint64_t b = 0; b < P; b += chunk) {
      const uint64_t e = std::min(P, b + chunk);
      if (prf_fetch(ctx, b, e, piece.data())) die(std::string("phybin_rf error: ") + prf_last_error(ctx));
      if (phb_dist_bin_append(path.c_str(), piece.data(), e - b)) die("cannot write " + path);
    }
    std::printf("Wrote %s (%.1f MB)\n", path.c_str(), (double)P * 2 / 1e6);
  }
  if (text_out) {
    const std::string path = opt.out_dir + "/distance_matrix.txt";
    if (phb_write_dist_mat(path.c_str(), upper.data(), T)) die("cannot write " + path);
    std::printf("Wrote %s\n", path.c_str());
  } else {
    std::printf("distance_matrix.txt skipped: %u trees would be a %.1f GB text file (use --binary-matrix or the C ABI's prf_fetch_rows)\n", T,
                (double)P * 4 / 1e9);
  }
  if (opt.rfdist && text_out) {
    char* txt = phb_print_dist_mat(upper.data(), T);
    std::fputs(txt, stdout);
    phb_free_buf(txt);
  }
  if (opt.editdist >= 0) {
    // sliceDendro D (C.dendrogram linkage trees dist)  (PhyBin.hs:358,415-422): the device splits the trees into the
    // connected components of { d <= D } (union-find over the fused mask); that is the single-linkage answer, and
    // complete linkage / UPGMA agglomerate every component on its own from its fetched sub-matrix.
    static const char* const kLinkage[] = {"SingleLinkage", "CompleteLinkage", "UPGMA"};
    std::printf("Next, clustering using method %s\n", kLinkage[opt.linkage]);  // PhyBin.hs:363
    std::vector<uint32_t> label(T ? T : 1);
    uint32_t comps = 0;
    if ((rc = prf_mask_components(ctx, nullptr, T, label.data(), &comps)))
      die(std::string("phybin_rf error: ") + prf_last_error(ctx));
    std::vector<std::vector<uint32_t>> members(T);
    for (uint32_t t = 0; t < T; ++t) members[label[t]].push_back(t);
    if (opt.linkage != 0) {
      std::vector<uint16_t> sub;
      std::vector<uint32_t> local;
      for (uint32_t c = 0; c < T; ++c) {
        const std::vector<uint32_t>& ids = members[c];
        const uint32_t m = (uint32_t)ids.size();
        if (m < 3) continue;  // two trees within D of each other are one cluster under every linkage
        local.resize(m);
        if (m >= 64) {  // a large component: the agglomeration itself runs on the device (same rule, same arithmetic)
          if ((rc = prf_agglomerate(ctx, nullptr, T, ids.data(), m, opt.linkage, (double)opt.editdist, local.data(), nullptr, nullptr,
                                    nullptr, nullptr)))
            die(std::string("phybin_rf error: ") + prf_last_error(ctx));
        } else {
          sub.resize((size_t)m * (m - 1) / 2);
          if ((rc = prf_fetch_submatrix(ctx, nullptr, T, ids.data(), m, sub.data()))) die(std::string("phybin_rf error: ") + prf_last_error(ctx));
          if (phb_cluster_component(sub.data(), m, opt.linkage, (double)opt.editdist, local.data()))
            die("component of " + std::to_string(m) + " trees is too large for the host clustering step");
        }
        for (uint32_t i = 0; i < m; ++i) label[ids[i]] = ids[local[i]];
      }
      for (auto& v : members) v.clear();
      for (uint32_t t = 0; t < T; ++t) members[label[t]].push_back(t);
    }
    // dendrogram.txt = show (fmap treename dendro) (PhyBin.hs:238-243): the full agglomeration of every tree, on the device
    if (T >= 1 && T <= 40000) {
      std::vector<uint32_t> ma(T), mb(T);
      std::vector<double> md(T);
      prf_agglo_info ai;
      if ((rc = prf_agglomerate(ctx, nullptr, T, nullptr, T, opt.linkage, INFINITY, nullptr, ma.data(), mb.data(), md.data(), &ai)))
        die(std::string("phybin_rf error: ") + prf_last_error(ctx));
      std::vector<const char*> nptr;
      for (uint32_t t = 0; t < T; ++t) nptr.push_back(t < names.size() ? names[t].c_str() : "");
      char* txt = phb_dendrogram_text(nptr.data(), T, ma.data(), mb.data(), md.data(), ai.n_merges);
      if (!txt) die("could not format the dendrogram");
      std::ofstream(opt.out_dir + "/dendrogram.txt") << txt;
      phb_free_buf(txt);
      std::printf("Time for clustering was: %.3fs\n", ai.ms_kernels / 1e3);
      std::printf(" [finished] Wrote full dendrogram to file dendrogram.txt\n");  // PhyBin.hs:246
    } else {
      std::printf("dendrogram.txt skipped: the full dendrogram of %u trees needs a %u x %u working matrix (flat clusters below are exact)\n", T, T, T);
    }
    std::printf("Combining all clusters at distance less than or equal to %d\n", opt.editdist);  // PhyBin.hs:276
    // sorted0 = reverse (sortBy (compare `on` fst) wlens)  (PhyBin.hs:281): sizes descending.  The order of
    // equal-sized clusters follows the library's dendrogram in the reference; here: by smallest member, reversed.
    std::vector<uint32_t> order;
    for (uint32_t c = 0; c < T; ++c)
      if (!members[c].empty()) order.push_back(c);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return members[a].size() < members[b].size(); });
    std::reverse(order.begin(), order.end());
    std::string sizes = "[";
    for (size_t i = 0; i < order.size(); ++i) sizes += (i ? "," : "") + std::to_string(members[order[i]].size());
    sizes += "]";
    std::printf("After flattening, cluster sizes are: %s\n", sizes.c_str());  // PhyBin.hs:284
    size_t non_single = 0;
    for (uint32_t c : order) non_single += members[c].size() > 1;
    std::printf(" Outcome: %zu clusters found, %zu non-singleton\n", order.size(), non_single);  // reportClusts, PhyBin.hs:375
    uint64_t close = 0;
    for (uint64_t w = 0; w < (P + 31) / 32; ++w) close += (uint64_t)__builtin_popcount(mask[w]);
    std::printf("%llu of %llu pairs within edit distance %d; %u connected components\n", (unsigned long long)close,
                (unsigned long long)P, opt.editdist, comps);

    // strict consensus of every cluster: the intersection runs on the device (prf_consensus), bipsToTree here
    std::vector<uint64_t> ckeys, coff;
    if (!opt.tolerant) {
      prf_consensus_info ci{};
      if (opt.host_bips) {
        uint64_t *bips = nullptr, *off = nullptr, nnz = 0;
        if (phb_allbips(forest, W, 0, &bips, &off, &nnz)) die("bipartition extraction failed");
        rc = prf_consensus(ctx, bips, off, T, W, PRF_MEM_HOST, label.data(), &ci);
        phb_free_buf(bips);
        phb_free_buf(off);
      } else {
        const uint64_t *d_keys = nullptr, *d_off = nullptr;
        rc = prf_allbips_result(ctx, &d_keys, &d_off, nullptr, nullptr);  // still resident from the extraction
        if (!rc) rc = prf_consensus(ctx, d_keys, d_off, T, W, PRF_MEM_DEVICE, label.data(), &ci);
      }
      ckeys.resize(std::max<uint64_t>(ci.n_out * W, 1));
      coff.resize((size_t)T + 1);
      if (!rc) rc = prf_consensus_fetch(ctx, ckeys.data(), coff.data());
      if (rc) die(std::string("phybin_rf error: ") + prf_last_error(ctx));
    }
    // outputClusters (PhyBin.hs:428-441): cluster<i>_<size>.txt, _consensus.tr, _alltrees.tr
    const size_t max_files = 1000;
    for (size_t i = 0; i < order.size() && i < max_files; ++i) {
      const std::vector<uint32_t>& ids = members[order[i]];
      const std::string base = opt.out_dir + "/cluster" + std::to_string(i + 1) + "_" + std::to_string(ids.size());
      std::ofstream txt(base + ".txt"), all(base + "_alltrees.tr");
      for (uint32_t t : ids) {
        txt << (t < names.size() ? names[t] : std::to_string(t)) << '\n';
        char* nw = phb_to_newick(forest, t, t + 1);
        all << nw;
        phb_free_buf(nw);
      }
      if (!opt.tolerant) {
        const uint32_t c = order[i];
        char* cons = phb_consensus_newick(forest, ckeys.data() + coff[c] * W, coff[c + 1] - coff[c], W, n_taxa);
        if (!cons) die("bipsToTree failed for cluster " + std::to_string(i + 1));
        std::ofstream(base + "_consensus.tr") << cons << '\n';
        phb_free_buf(cons);
      }
    }
    std::printf(" [finished] Wrote contents of each cluster to cluster<N>_<size>.txt\n");  // PhyBin.hs:443
    if (!opt.tolerant) std::printf(" [finished] Wrote representative (consensus) trees to cluster<N>_<size>_consensus.tr\n");
    if (order.size() > max_files) std::printf("(only the %zu largest clusters were written out)\n", max_files);
    std::ofstream cf(opt.out_dir + "/components.txt");
    for (uint32_t t = 0; t < T; ++t) cf << t << ' ' << label[t] << '\n';
  }
  prf_destroy(ctx);
  phb_free(forest);
  return 0;
}
