/* phybin_rf.h -- C ABI of libphybin_rf.so, the B200 (sm_100a) implementation of PhyBin's
 * all-to-all Robinson-Foulds distance-matrix path.
 *
 * The reference has no FFI: its boundary for this path is two pure Haskell functions and one type
 * (all in /root/reference/src/Bio/Phylogeny/PhyBin/RFDistance.hs):
 *
 *     hashRF          :: Int -> [NewickTree a] -> DistanceMatrix              -- RFDistance.hs:284
 *     naiveDistMatrix :: [NewickTree DefDecor]
 *                     -> (DistanceMatrix, V.Vector (S.Set DenseLabelSet))     -- RFDistance.hs:168
 *     type DistanceMatrix = V.Vector (U.Vector Int)                           -- RFDistance.hs:156
 *
 * called from exactly one production site, doCluster (PhyBin.hs:343-364).  The entry points below
 * are what a `foreign import ccall` inside drop-in replacements for those two functions binds
 * (INTEGRATION.md shows the Haskell shim).  The host keeps Newick parsing and --prune/--minbootstrap
 * preprocessing; the device receives bit-packed taxon sets (allBips, RFDistance.hs:247 -- extracted by the
 * host, or by prf_allbips below from the parsed trees) and returns the matrix.  The entry points after the
 * matrix ones serve its consumers in PhyBin.hs: flat clusters at --editdist and their strict consensus.
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns PRF_OK (0) or a negative prf_status
 *     and never throws, aborts or prints; prf_last_error() has the message.
 *   - the caller owns every pointer it passes; the library owns all device memory of a context.
 *   - one calling thread per context; no global state; work is issued on the context's stream.
 *   - bit sets: taxon label l (the reference's `Label = Int`, CoreTypes.hs:139) is bit (l % 64) of
 *     64-bit word (l / 64); a set is W consecutive little-endian uint64 words.
 *   - matrix layout: PACKED ROW-MAJOR UPPER TRIANGLE of uint16.  For trees a < b,
 *         d(a,b) lives at  p(a,b) = a*T - a*(a+1)/2 + (b - a - 1)        (64-bit arithmetic)
 *     which is the reference's row b, column a (DistanceMatrix row i holds d(i,0..i-1),
 *     RFDistance.hs:306-311).  The diagonal is implicit 0.  There are P = T*(T-1)/2 entries.
 *   - the optional edit-distance mask has bit (p % 32) of uint32 word (p / 32) set iff
 *     d(p) <= editdist (the --editdist value, Main.hs:100-102); ceil(P/32) words.
 *   - there is no CPU fallback: without a CUDA device prf_create fails.
 */
#ifndef PHYBIN_RF_H
#define PHYBIN_RF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define PRF_API __attribute__((visibility("default")))
#else
#define PRF_API
#endif

typedef struct prf_ctx prf_ctx;

typedef enum prf_status {
  PRF_OK = 0,
  PRF_E_INVALID = -1,     /* bad argument (null pointer, T or W out of range, p range unordered) */
  PRF_E_CUDA = -2,        /* a CUDA runtime call or kernel failed; see prf_last_error */
  PRF_E_NOMEM = -3,       /* host or device allocation failed */
  PRF_E_RANGE = -4,       /* a distance does not fit uint16 (a tree with > 32767 bipartitions) */
  PRF_E_STATE = -5,       /* call order: compute/fetch before a successful load */
  PRF_E_UNSUPPORTED = -6  /* shape outside what the kernels are built for (see prf_limits) */
} prf_status;

typedef enum prf_variant {
  PRF_VARIANT_AUTO = 0,   /* pick by measured split density (stats.hits vs stats.n_shared) */
  PRF_VARIANT_SPARSE = 1, /* inverted-list join in shared memory, HBM-write bound */
  PRF_VARIANT_DENSE = 2   /* int8 tcgen05 Gram product S = A*A^T, tensor-pipe bound */
} prf_variant;

typedef enum prf_memkind {
  PRF_MEM_HOST = 0,   /* pointer is host memory (pageable or pinned) */
  PRF_MEM_DEVICE = 1  /* pointer is device memory on the context's device */
} prf_memkind;

/* Filled by the load / compute calls; all times are CUDA-event milliseconds on the context's
 * stream (0 when the phase did not run). */
typedef struct prf_stats {
  uint64_t n_trees;      /* T */
  uint64_t n_words;      /* W */
  uint64_t nnz;          /* bipartition entries handed in (sum of per-tree counts) */
  uint64_t n_unique;     /* U: distinct bipartitions over all trees (size of hashRF's Map, :289) */
  uint64_t n_shared;     /* distinct bipartitions present in >= 2 trees */
  uint64_t n_shared_nnz; /* entries that belong to those */
  uint64_t n_hits;       /* sum over bipartitions of m*(m-1)/2 = sum over pairs of |B_i n B_j| */
  uint64_t n_pairs;      /* T*(T-1)/2 */
  uint32_t max_bips;     /* max per-tree |B_i| */
  uint32_t min_bips;
  int32_t variant_used;  /* prf_variant actually run by the last compute */
  int32_t gpu_launches;  /* kernels launched by the last call */
  float ms_h2d, ms_dedup, ms_incidence, ms_matrix, ms_d2h, ms_total;
  uint32_t n_flipped;    /* bipartitions held by > T/2 trees, stored as the list of trees LACKING them (n_hits,
                            n_shared_nnz and min/max_bips then describe the flipped sets D_t = B_t xor majority) */
  uint32_t reserved;
} prf_stats;

/* ---- lifetime ------------------------------------------------------------------------------- */

/* Creates a context on CUDA device `device` (>= 0).  Returns NULL when there is no such device or
 * CUDA is unavailable (there is no CPU fallback).  One context drives one GPU; prf_multi_create (below) drives
 * every GPU of the box behind the same kind of call. */
PRF_API prf_ctx* prf_create(int device);
PRF_API void prf_destroy(prf_ctx* ctx);
/* Message of the last failure on this context ("" if none).  ctx may be NULL: returns the message
 * of the last failed prf_create on this thread. */
PRF_API const char* prf_last_error(const prf_ctx* ctx);
PRF_API const char* prf_version(void);
/* The context's CUDA stream (a cudaStream_t) so callers holding device buffers can order work. */
PRF_API void* prf_stream(prf_ctx* ctx);

/* ---- hashRF: one-shot (replaces the body of hashRF, RFDistance.hs:284-341) ------------------- */

/* bips      nnz*W words: for tree t the canonical bipartitions allBips(t) (RFDistance.hs:247) are
 *           entries tree_off[t] .. tree_off[t+1]-1.  A tree's entries are treated as a SET
 *           (duplicates count once), like S.Set DenseLabelSet.  Host memory.
 * tree_off  T+1 ascending offsets, tree_off[0] = 0, tree_off[T] = nnz.  Host memory.
 * editdist  < 0: no mask.  >= 0: also produce the d <= editdist bit mask.
 * out_upper host buffer of T*(T-1)/2 uint16, or NULL to leave the result on the device
 *           (read it later with prf_fetch / prf_fetch_rows).
 * out_mask  host buffer of ceil(P/32) uint32 or NULL.
 * stats     optional. */
PRF_API int prf_hashrf(prf_ctx* ctx, const uint64_t* bips, const uint64_t* tree_off, uint32_t T,
                       uint32_t W, int32_t editdist, uint16_t* out_upper, uint32_t* out_mask,
                       prf_stats* stats);

/* ---- hashRF: staged (large T, device-resident inputs, multi-GPU sharding) -------------------- */

/* Stage 1 -- the `build` phase of hashRF (RFDistance.hs:289-297): global bipartition dedup and
 * the trees x bipartitions incidence, kept on the device.  `kind` says where bips/tree_off live. */
PRF_API int prf_hashrf_load(prf_ctx* ctx, const uint64_t* bips, const uint64_t* tree_off,
                            uint32_t T, uint32_t W, int kind /* prf_memkind */, prf_stats* stats);

/* Stage 2 -- the `ingest` phase (RFDistance.hs:300-341) for packed positions [p_begin, p_end):
 * d = |B_a| + |B_b| - 2|B_a n B_b|.  p_begin must be a multiple of prf_shard_align().
 * d_out   DEVICE buffer of (p_end - p_begin) uint16 (16-byte aligned) or NULL to use a buffer the
 *         context owns (readable with prf_fetch);
 * d_mask  DEVICE buffer of ceil((p_end - p_begin)/32) uint32 or NULL (ctx-owned when editdist>=0).
 * Ranks of a multi-GPU job call this with disjoint ranges; no exchange is needed. */
PRF_API int prf_hashrf_compute(prf_ctx* ctx, uint64_t p_begin, uint64_t p_end, int32_t editdist,
                               int variant /* prf_variant */, uint16_t* d_out, uint32_t* d_mask,
                               prf_stats* stats);

/* Copies packed positions [p_begin, p_end) of the last ctx-owned result to host memory. */
PRF_API int prf_fetch(prf_ctx* ctx, uint64_t p_begin, uint64_t p_end, uint16_t* dst);
PRF_API int prf_fetch_mask(prf_ctx* ctx, uint64_t p_begin, uint64_t p_end, uint32_t* dst);
/* Copies the packed entries of upper-triangle rows [row0, row1) (row a holds d(a, a+1..T-1)). */
PRF_API int prf_fetch_rows(prf_ctx* ctx, uint32_t row0, uint32_t row1, uint16_t* dst);

/* ---- hashRF on all GPUs of one box ------------------------------------------------------------ */
/* The `build` phase (RFDistance.hs:289-297) is hash-partitioned: GPU r deduplicates the bipartitions of hash class r
 * and builds their inverted lists; the classes' structures are exchanged by kernels that store straight into the
 * peers' memory over NVLink (no NCCL call, no host-side size exchange); the `ingest` phase (RFDistance.hs:300-341)
 * then needs no exchange at all: GPU r computes the aligned slice prf_shard_range(T, r, world) of the packed triangle.
 * Results are bit-identical to the single-GPU path.  The tensor-core (dense) variant is single-GPU only.
 *
 * (a) One process drives every GPU -- what a drop-in `hashRF` binds: */
typedef struct prf_multi prf_multi;
/* n_gpus <= 0: every visible GPU.  devices: n_gpus CUDA device indices or NULL for 0 .. n_gpus-1.  NULL on failure. */
PRF_API prf_multi* prf_multi_create(int n_gpus, const int* devices);
PRF_API void prf_multi_destroy(prf_multi* m);
PRF_API const char* prf_multi_last_error(const prf_multi* m);
PRF_API int prf_multi_gpus(const prf_multi* m);
/* The context of rank `rank` (its slice of the last result stays there: prf_fetch / prf_fetch_mask, consumers). */
PRF_API prf_ctx* prf_multi_ctx(prf_multi* m, int rank);
/* Arguments as prf_hashrf (host buffers).  Trees are dealt to the GPUs in contiguous shards; every GPU copies its
 * slice of the matrix (and mask) into out_upper / out_mask over its own PCIe link. */
/* Flat single-linkage clusters after a prf_multi_hashrf with editdist >= 0: every GPU runs the union-find over its
 * resident mask shard, GPU 0 merges the label arrays (prf_mask_components_shard).  labels: host, T entries. */
PRF_API int prf_multi_components(prf_multi* m, uint32_t* labels, uint32_t* n_components);
PRF_API int prf_multi_hashrf(prf_multi* m, const uint64_t* bips, const uint64_t* tree_off, uint32_t T, uint32_t W,
                             int32_t editdist, uint16_t* out_upper, uint32_t* out_mask, prf_stats* stats);

/* (b) One context per GPU, in one process or one process per GPU (MPI / torchrun style):
 *   prf_team_init     every rank, same T / W / nnz_total / max_local_entries / max_tree_entries: allocates the rank's
 *                     exchange arena and returns its handle (prf_team_handle_bytes() bytes, a cudaIpcMemHandle_t);
 *   prf_team_connect  with the handles of all ranks in rank order (exchanged by the caller, e.g. an all-gather);
 *   prf_team_load     collective: rank r passes the canonical bipartitions of ITS trees, trees
 *                     [T r / world, T (r + 1) / world), laid out like prf_hashrf_load's input, in DEVICE memory.
 *                     Afterwards every rank holds the whole structure: prf_hashrf_compute on any range.
 * nnz_total = entries over all ranks; max_local_entries = most entries one rank holds; max_tree_entries = most entries
 * of one tree.  A hash class that outgrows its region (2 x the mean + 64 K entries) is reported as PRF_E_NOMEM. */
PRF_API size_t prf_team_handle_bytes(void);
PRF_API int prf_team_init(prf_ctx* ctx, uint32_t rank, uint32_t world, uint32_t T, uint32_t W, uint64_t nnz_total,
                          uint64_t max_local_entries, uint32_t max_tree_entries, void* handle_out);
PRF_API int prf_team_connect(prf_ctx* ctx, const void* handles);
PRF_API int prf_team_load(prf_ctx* ctx, const uint64_t* d_bips, const uint64_t* d_tree_off, uint32_t T_local,
                          prf_stats* stats);
PRF_API void prf_team_release(prf_ctx* ctx);

/* ---- tolerant mode (replaces fst . naiveDistMatrix, RFDistance.hs:168-198) ------------------- */

/* clades    nnz*W words: for tree t, the UN-normalised leaf set of every non-root interior node
 *           (the sets labelBips attaches to interior nodes, RFDistance.hs:218-221).  Singleton
 *           leaf sets need not be passed (the kernel derives them from `leafsets`).
 * tree_off  T+1 offsets into clades.
 * leafsets  T*W words: the tree's taxon set (treeLabels, RFDistance.hs:200-203).
 * Each tree's leaf labels must be distinct (forests where a tree repeats a label: prf_tolerant_load_dups).
 * Per pair the kernel restricts both trees to the
 * shared taxa, re-normalises (normBip with size = #shared, RFDistance.hs:229-239, including its
 * [0,size) complement universe and odd-size tie rule) and counts the symmetric difference. */
PRF_API int prf_tolerant(prf_ctx* ctx, const uint64_t* clades, const uint64_t* tree_off,
                         const uint64_t* leafsets, uint32_t T, uint32_t W, int32_t editdist,
                         uint16_t* out_upper, uint32_t* out_mask, prf_stats* stats);
PRF_API int prf_tolerant_load(prf_ctx* ctx, const uint64_t* clades, const uint64_t* tree_off,
                              const uint64_t* leafsets, uint32_t T, uint32_t W, int kind,
                              prf_stats* stats);
PRF_API int prf_tolerant_compute(prf_ctx* ctx, uint64_t p_begin, uint64_t p_end, int32_t editdist,
                                 uint16_t* d_out, uint32_t* d_mask, prf_stats* stats);
/* Load for forests in which a tree may REPEAT a leaf label.  The reference's label sets stay sets, but numLeaves
 * -- normBip's size, RFDistance.hs:215 -- counts such leaves twice (CoreTypes.hs:289-291), and whether
 * pruneTreeLeaves (PreProcessor.hs:21-32) leaves a node below the pruned root is no longer visible from its own
 * label set.  Additional inputs (phb_raw_clades_ex produces them):
 *   clades     every non-root interior node of a tree, NOT made unique (equal sets may differ in outsets)
 *   outsets    nnz*W words: per clade the label set of the leaves OUTSIDE that node      (`kind` memory)
 *   dup_off    T+1 offsets into dup_label / dup_extra                                     (HOST memory)
 *   dup_label, dup_extra   per tree its labels that occur more than once, and how many extra times
 * Followed by prf_tolerant_compute as usual (a slower instantiation of the same kernel: per-tree sizes). */
PRF_API int prf_tolerant_load_dups(prf_ctx* ctx, const uint64_t* clades, const uint64_t* outsets,
                                   const uint64_t* tree_off, const uint64_t* leafsets, const uint64_t* dup_off,
                                   const uint32_t* dup_label, const uint32_t* dup_extra, uint32_t T, uint32_t W,
                                   int kind, prf_stats* stats);

/* ---- bipartition extraction on the device (replaces allBips, RFDistance.hs:207-248) ------------ */

typedef enum prf_extract_mode {
  PRF_EXTRACT_BIPS = 0,   /* allBips: normalised bipartitions of > 1 taxa (prf_hashrf_load's input)        */
  PRF_EXTRACT_CLADES = 1  /* un-normalised clades of the non-root interior nodes + the trees' taxon sets  */
                          /* (prf_tolerant_load's input)                                                   */
} prf_extract_mode;

typedef struct prf_extract_info {
  uint64_t nnz;        /* sets produced */
  uint32_t max_nodes;  /* largest tree, in nodes */
  int32_t gpu_launches;
  float ms_h2d, ms_kernels;
} prf_extract_info;

/* Trees as flat arrays, the nodes of tree t being tree_begin[t] .. tree_begin[t+1]-1 in PRE-ORDER (a
 * parent precedes its children; the first node is the root):
 *   node_label   >= 0: a leaf carrying that taxon label (Label, CoreTypes.hs:139); < 0: interior node
 *   node_parent  index of the parent in the same arrays; ignored for a tree's first node
 *   tree_begin   T+1 ascending offsets, tree_begin[0] = 0
 *   num_leaves   numLeaves of every tree (CoreTypes.hs:289-291; duplicate labels count twice) -- the
 *                `size` normBip works with (RFDistance.hs:215,229-239)
 * `kind` says where the four arrays live.  The result stays on the device until the next extraction on
 * this context; it is, per tree, the SET (every set once, in the pre-order of the first node that yields it) of
 *   mode BIPS:   normBip size c for every non-root node's leaf set c, kept when it has > 1 members
 *                (labelBips :207-221, normBip :229-239 with its [0,size) complement universe and odd-size
 *                tie rule, allBips' filter :248);
 *   mode CLADES: the leaf set of every non-root interior node, plus leafsets[t] = the root's set.
 * PRF_E_UNSUPPORTED when one tree's clade sets do not fit a CTA's shared memory
 * (about interior_nodes * (8 W + 4) + 2 * nodes bytes must stay under 200 KB; at most 16383 nodes per tree). */
PRF_API int prf_allbips(prf_ctx* ctx, const int32_t* node_label, const uint32_t* node_parent,
                        const uint64_t* tree_begin, const uint32_t* num_leaves, uint32_t T, uint32_t W,
                        int kind /* prf_memkind */, int mode /* prf_extract_mode */, prf_extract_info* info);
/* Device pointers of the last extraction: keys (nnz*W words), tree_off (T+1), leafsets (T*W words, NULL in
 * mode BIPS) -- pass them to prf_hashrf_load / prf_tolerant_load with PRF_MEM_DEVICE.  Any out pointer may
 * be NULL. */
PRF_API int prf_allbips_result(prf_ctx* ctx, const uint64_t** d_keys, const uint64_t** d_tree_off,
                               const uint64_t** d_leafsets, uint64_t* nnz);
/* Copies the last extraction to host memory (each destination may be NULL). */
PRF_API int prf_allbips_fetch(prf_ctx* ctx, uint64_t* keys, uint64_t* tree_off, uint64_t* leafsets);

/* One-shot from parsed trees: prf_allbips (mode BIPS) + prf_hashrf_load + prf_hashrf_compute (+ fetch) -- what a
 * drop-in hashRF binds when the host does not want to run allBips itself; the tree arrays are host memory, laid
 * out as for prf_allbips.  out_upper / out_mask / stats as for prf_hashrf; extract may be NULL. */
PRF_API int prf_hashrf_trees(prf_ctx* ctx, const int32_t* node_label, const uint32_t* node_parent,
                             const uint64_t* tree_begin, const uint32_t* num_leaves, uint32_t T, uint32_t W,
                             int32_t editdist, uint16_t* out_upper, uint32_t* out_mask, prf_stats* stats,
                             prf_extract_info* extract);
/* Same for the tolerant mode: prf_allbips (mode CLADES) + prf_tolerant_load + prf_tolerant_compute (+ fetch). */
PRF_API int prf_tolerant_trees(prf_ctx* ctx, const int32_t* node_label, const uint32_t* node_parent,
                               const uint64_t* tree_begin, const uint32_t* num_leaves, uint32_t T, uint32_t W,
                               int32_t editdist, uint16_t* out_upper, uint32_t* out_mask, prf_stats* stats,
                               prf_extract_info* extract);

/* ---- first consumer of the matrix: flat clusters at the --editdist threshold -------------------- */

/* Connected components of the graph { (a,b) : d(a,b) <= editdist } read from a threshold mask that
 * covers the whole triangle [0, P): d_mask = a device mask written by *_compute, or NULL for the
 * context-owned mask of the last compute.  labels (host, T entries) receives the smallest tree id of
 * every tree's component.  Single linkage: these are the reference's flat clusters (sliceDendro,
 * PhyBin.hs:415-422); complete linkage / UPGMA clusters are subsets of them. */
PRF_API int prf_mask_components(prf_ctx* ctx, const uint32_t* d_mask, uint32_t T, uint32_t* labels,
                                uint32_t* n_components);

/* Shard form (the mask is sharded like the matrix).  The mask covers [p_begin, p_end) -- a device pointer
 * whose first bit is pair p_begin, or NULL for the context-owned mask of the last compute, which must cover exactly
 * that range; an empty range reads no mask at all (pure merge of the seeds).
 * seed (host, n_seed x T, may be NULL): label arrays other shards / ranks returned; their components are merged in as
 * edges (t, seed[i][t]).  So every rank calls this on its own shard, the label arrays are exchanged (all-gather), and
 * one more call with an empty range and all of them as seeds yields the components of the whole graph -- the
 * all-reduce(min) of labels, done as a union-find on the device. */
PRF_API int prf_mask_components_shard(prf_ctx* ctx, const uint32_t* d_mask, uint64_t p_begin, uint64_t p_end, uint32_t T,
                                      const uint32_t* seed, uint32_t n_seed, uint32_t* labels, uint32_t* n_components);

/* Sub-matrix for m trees ids[0] < ids[1] < ... (host array): out (host, m*(m-1)/2 uint16) is the packed upper
 * triangle of the m x m matrix d(ids[i], ids[j]) -- what the host-side complete-linkage / UPGMA step needs for
 * one component (phb_cluster_component).  d_matrix: a DEVICE buffer holding the whole packed triangle of T trees
 * (as written by *_compute), or NULL for the context-owned result of the last compute (T is then ignored and the
 * result must cover every requested pair). */
PRF_API int prf_fetch_submatrix(prf_ctx* ctx, const uint16_t* d_matrix, uint32_t T, const uint32_t* ids, uint32_t m,
                                uint16_t* out);

/* ---- the agglomeration itself: C.dendrogram linkage trees dist (PhyBin.hs:358) ------------------- */

typedef struct prf_agglo_info {
  uint32_t n_merges;    /* merges performed (m - 1 for a full dendrogram) */
  uint32_t n_clusters;  /* m - n_merges */
  int32_t gpu_launches;
  float ms_kernels;     /* gather + agglomeration kernel */
} prf_agglo_info;

/* Agglomerative clustering of m trees on the device, replacing hierarchical-clustering-0.4.6's C.dendrogram
 * (stack.yaml:9, called at PhyBin.hs:358) and, through `thresh`, sliceDendro (PhyBin.hs:415-422).
 *   d_matrix, T   as for prf_fetch_submatrix: a DEVICE buffer with the whole packed triangle, or NULL for the
 *                 context-owned result of the last compute
 *   ids, m        ascending tree ids (host) -- e.g. one component of prf_mask_components; NULL = all T trees
 *   linkage       0 single, 1 complete, 2 UPGMA (C.SingleLinkage / C.CompleteLinkage / C.UPGMA, Main.hs:96-98)
 *   thresh        merging stops when the smallest distance exceeds it (sliceDendro's `dist > dstThresh`);
 *                 INFINITY = the full dendrogram (dendrogram.txt, PhyBin.hs:238-243)
 *   labels        out (host, m entries, may be NULL): smallest LOCAL index of every tree's cluster at the cut
 *   merge_a/b/d   out (host, m-1 entries each, may be NULL): step k merged the clusters whose smallest local
 *                 members are merge_a[k] < merge_b[k] at distance merge_d[k]; the merged cluster keeps merge_a[k]
 * Every step merges the pair at minimal distance, the first in (lower key, higher key) order among ties, and
 * updates distances by Lance-Williams in IEEE double -- the same rule and arithmetic as the host restatement
 * phb_cluster_component, so both produce the same merges.  One persistent cooperative kernel, one grid barrier
 * per merge; the working matrix is m x m uint16 (single / complete) or double (UPGMA) in HBM: PRF_E_NOMEM when
 * that does not fit. */
PRF_API int prf_agglomerate(prf_ctx* ctx, const uint16_t* d_matrix, uint32_t T, const uint32_t* ids, uint32_t m,
                            int linkage, double thresh, uint32_t* labels, uint32_t* merge_a, uint32_t* merge_b,
                            double* merge_d, prf_agglo_info* info);

/* ---- strict consensus of every cluster (replaces the intersection inside consensusTree) -------- */

typedef struct prf_consensus_info {
  uint64_t n_out;       /* bipartitions in all consensus sets together */
  uint32_t n_clusters;  /* distinct labels */
  int32_t gpu_launches;
  float ms_kernels;
  uint32_t reserved;
} prf_consensus_info;

/* For every cluster, the bipartitions ALL of its trees have: foldl' S.intersection (allBips hd) (map allBips tl)
 * (consensusTree, RFDistance.hs:396-400, called per cluster at PhyBin.hs:432-436).  bips / tree_off are laid out
 * as for prf_hashrf_load (`kind` says where they live; every tree's entries must be distinct, as prf_allbips and
 * phb_allbips produce them).  labels: HOST array, labels[t] = smallest tree id of t's cluster (what
 * prf_mask_components returns).  The result stays on the device: per tree t a range cons_off[t] .. cons_off[t+1]
 * that is empty unless t names a cluster, and then holds the consensus bipartitions in the order they have in
 * tree t.  bipsToTree (RFDistance.hs:410-445) on those few sets is host work (phb_consensus_newick). */
PRF_API int prf_consensus(prf_ctx* ctx, const uint64_t* bips, const uint64_t* tree_off, uint32_t T, uint32_t W,
                          int kind /* prf_memkind */, const uint32_t* labels, prf_consensus_info* info);
/* Copies the last prf_consensus result to host memory: keys n_out*W words, cons_off T+1 (either may be NULL). */
PRF_API int prf_consensus_fetch(prf_ctx* ctx, uint64_t* keys, uint64_t* cons_off);

/* ---- helpers -------------------------------------------------------------------------------- */

/* p(a,b) for a < b < T (see layout above); UINT64_MAX on bad arguments. */
PRF_API uint64_t prf_packed_index(uint32_t T, uint32_t a, uint32_t b);
/* Shard boundaries handed to *_compute must be multiples of this many entries (or equal P). */
PRF_API uint64_t prf_shard_align(void);
/* Balanced, aligned split of [0, P) over `world` ranks: writes rank's [*p_begin, *p_end). */
PRF_API int prf_shard_range(uint32_t T, uint32_t rank, uint32_t world, uint64_t* p_begin,
                            uint64_t* p_end);
/* Largest W and per-tree bipartition count the kernels accept. */
PRF_API void prf_limits(uint32_t* max_words, uint32_t* max_bips_per_tree);
/* Kernels launched by this context since it was created (bench.py reports it as gpu_launches). */
PRF_API uint64_t prf_launch_count(const prf_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* PHYBIN_RF_H */
