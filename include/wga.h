/* =============================================================================
 *  wga.h -- C ABI of libwgans: B200-native (sm_100a) webgraph-ans hot path.
 * =============================================================================
 *  Drop-in boundary for the ANS decode of BvGraph components and the encoder-side
 *  symbol-model construction of ciminilorenzo/webgraph-ans-rs.  The reference has
 *  no FFI today (one Rust process, webgraph-rs traits); its per-symbol pull
 *  interface (`impl Decode for ANSDecoder`, src/ans/decoder.rs:103-139) is too fine
 *  for a GPU, so the boundary sits one level up, at the graph API:
 *
 *    reference (Rust)                                         this library
 *    -------------------------------------------------------  -------------------------
 *    ANSBvGraph::load      src/bvgraph/random_access.rs:52    wga_open / wga_open_mem
 *    ANSBvGraphSeq::load   src/bvgraph/sequential.rs:29       wga_open (flags WGA_OPEN_SEQUENTIAL: .ans alone
 *                                                              is enough, the phases are rebuilt at load)
 *    ANSModel4Decoder::new src/ans/models/model4decoder.rs:18 done inside wga_open (packed tables)
 *    graph.iter()          examples/bench_seq_access.rs:24    wga_decode_range (+ _host)
 *    graph.successors(v)   examples/bench_random_access.rs:35 wga_successors_batch (+ _host)
 *    ANSModel4EncoderBuilder::push_symbol / build
 *                          src/ans/model4encoder_builder.rs:67,80
 *                                                             wga_model_* (histogram + normalise on GPU)
 *    ANSBvGraph::store     src/bvgraph/random_access.rs:91    wga_store (host front end + GPU model build)
 *
 *  Conventions: plain pointers and sizes only.  Names starting with d_ are DEVICE
 *  pointers, h_ are HOST pointers.  `stream` is a cudaStream_t passed as void*
 *  (NULL = default stream).  Every function returns 0 on success or a negative
 *  WGA_E_* code; wga_last_error() returns a thread-local message.  Successor ids
 *  are u32 (graphs with < 2^32 nodes), CSR offsets u64.  A handle is read-only after open, but it owns the
 *  small pinned block through which the kernels publish scalars to the host, so run ONE decode / random-access
 *  call at a time per handle (open a second handle, or a shard per rank, for concurrency).  There is NO CPU fallback:
 *  calls that need the GPU fail with WGA_E_CUDA when no device is usable.
 * ============================================================================= */
#ifndef WGA_H
#define WGA_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WGA_COMPONENTS 9 /* src/bvgraph/mod.rs:27 */

enum {
  WGA_OK = 0,
  WGA_E_IO = -1,        /* file missing / unreadable (anyhow error in load) */
  WGA_E_FORMAT = -2,    /* bad epserde header / layout */
  WGA_E_ARG = -3,       /* invalid argument */
  WGA_E_CUDA = -4,      /* CUDA error or no usable device */
  WGA_E_CORRUPT = -5,   /* decode hit an inconsistent stream/table (reference panics) */
  WGA_E_WORKSPACE = -6, /* workspace too small; see wga_last_error / wga_decode_workspace_size */
  WGA_E_UNSUPPORTED = -7
};

/* EncoderModelEntry, src/ans/models/component_model4encoder.rs:11-26 (repr(C), 8 bytes) */
typedef struct {
  uint32_t upperbound;
  uint16_t cumul_freq;
  uint16_t freq;
} wga_encoder_entry;

/* ANSComponentModel4Encoder, src/ans/models/component_model4encoder.rs:37-57 */
typedef struct {
  const wga_encoder_entry* table;
  uint64_t table_len;
  uint64_t frame_size; /* log2 of the frame */
  uint64_t radix;
  uint64_t fidelity;
  uint64_t folding_threshold;
  uint64_t folding_offset;
} wga_component_model;

/* Prelude (src/ans/mod.rs:31-54) + .states + expanded .pointers, all HOST memory */
typedef struct {
  wga_component_model tables[WGA_COMPONENTS];
  const uint16_t* stream;
  uint64_t stream_len;
  uint32_t state;
  uint64_t number_of_nodes;
  uint64_t compression_window;
  uint64_t min_interval_length;
  uint64_t number_of_arcs;
  const uint32_t* states;   /* entry i = node N-1-i (src/bvgraph/random_access.rs:202) */
  const uint64_t* pointers; /* entry i = node N-1-i, unit = u16 words */
} wga_prelude_view;

typedef struct wga_graph wga_graph;
typedef struct wga_model wga_model;

const char* wga_last_error(void);
/* 1 when a CUDA device is usable by this process, else 0 */
int wga_cuda_available(void);
/* number of kernels this library launched on the calling process so far (bench.py's gpu_launches) */
uint64_t wga_kernel_launches(void);

/* ---------------------------------------------------------------- load -------------------------- */
#define WGA_OPEN_DEFAULT 0
#define WGA_OPEN_HOST_ONLY 1 /* parse files, keep host copies, do not touch the GPU (format tests) */
#define WGA_OPEN_SEQUENTIAL 2 /* ANSBvGraphSeq::load (src/bvgraph/sequential.rs:29-51): the .ans alone is enough. When
                                 .pointers and .states are both absent, the per-node phases are recovered at load time by
                                 one walk of the stream from (stream.len(), prelude.state), as
                                 bvgraphseq_decoder_factory.rs:29-35 starts it */
/* Loads <basename>.ans/.pointers/.states (ANSBvGraph::load, src/bvgraph/random_access.rs:52-82) and uploads to the
 * current CUDA device. */
int wga_open(const char* basename, int flags, wga_graph** out);
/* Same from host arrays (copied). */
int wga_open_mem(const wga_prelude_view* view, int flags, wga_graph** out);
/* Same, but only nodes [first,last) plus whatever the caller says precedes (shard for one rank):
 * uploads only the stream words, states and pointers that range needs. */
int wga_open_shard(const char* basename, uint64_t first, uint64_t last, int flags, wga_graph** out);
void wga_close(wga_graph* g);

uint64_t wga_num_nodes(const wga_graph* g);
uint64_t wga_num_arcs(const wga_graph* g);
uint64_t wga_window(const wga_graph* g);
uint64_t wga_min_interval_length(const wga_graph* g);
uint64_t wga_stream_len(const wga_graph* g);
/* bytes of the inputs the decode reads from HBM: 2*stream_len + 4*N + bytes(.pointers payload as stored
 * on the device) -- the "compressed bytes" of the roofline */
uint64_t wga_compressed_bytes(const wga_graph* g);
/* host view of what was loaded (valid until wga_close) */
int wga_prelude(const wga_graph* g, wga_prelude_view* out);

/* ---------------------------------------------------------------- decode (graph.iter()) --------- */
/* Device workspace needed by wga_decode_range for nodes [first,last). */
uint64_t wga_decode_workspace_size(const wga_graph* g, uint64_t first, uint64_t last);
/* Decodes the successor lists of nodes [first,last) into CSR form:
 *   d_offsets[last-first+1] (u64, d_offsets[0]==0), d_succ[>= arcs of the range] (u32, ascending per node).
 * `succ_capacity` = elements available in d_succ; on return *h_arcs (optional, host) = arcs written.
 * References that leave the range on the left are resolved by re-decoding the needed predecessor nodes (halo)
 * inside the workspace.
 * The kernels run on `stream`, but the call is NOT asynchronous: it returns after they have completed, because
 * the host reads the totals and the error word (one stream synchronisation, two for sub-ranges with a halo and
 * for reference chains deeper than six levels) to size-check the output and to report WGA_E_* codes.  A too
 * small d_succ or workspace returns WGA_E_WORKSPACE ("need N elements" in wga_last_error) and writes nothing
 * behind the given capacities. */
int wga_decode_range(wga_graph* g, uint64_t first, uint64_t last, uint64_t* d_offsets, uint32_t* d_succ,
                     uint64_t succ_capacity, void* d_workspace, uint64_t workspace_bytes, uint64_t* h_arcs,
                     void* stream);
/* Only the outdegrees (first symbol of every record; bvgraph_decoder_factory.rs:46-58 + decoder.rs:104)
 * and their exclusive prefix sum. d_offsets[last-first+1]. */
int wga_outdegrees(wga_graph* g, uint64_t first, uint64_t last, uint64_t* d_offsets, void* d_workspace,
                   uint64_t workspace_bytes, void* stream);
/* End-to-end variant with HOST buffers (pinned memory recommended): D2H of offsets + successors.  Ranges
 * larger than one chunk (2^19 nodes) are pipelined: chunk i is decoded while the results of chunk i-1 travel
 * to the host and, after wga_upload(g, NULL), while the inputs of later chunks are still arriving. */
int wga_decode_range_host(wga_graph* g, uint64_t first, uint64_t last, uint64_t* h_offsets, uint32_t* h_succ,
                          uint64_t succ_capacity, uint64_t* h_arcs);

/* Re-copies the decode inputs of the resident range (stream words, states, pointers) from the handle's
 * pinned host copy to the device: the host->device leg of an end-to-end step.  stream == NULL: asynchronous,
 * in node-range chunks on a stream of the handle; the next wga_decode_range_host overlaps with it. */
int wga_upload(wga_graph* g, void* stream);
uint64_t wga_upload_bytes(const wga_graph* g);
/* Predecessor nodes the last wga_decode_range decoded in addition to its range, to resolve the references that leave
 * it on the left (0 when it started at the first resident node). */
uint64_t wga_last_halo_nodes(const wga_graph* g);
/* Per-stage device times of the last wga_decode_range (CUDA events on the caller's stream):
 * [outdegrees+scan, entropy decode (K1), levels+sort, resolve per level (K2)].
 * Returns the number of events. */
int wga_set_profiling(wga_graph* g, int on);
int wga_last_profile(const wga_graph* g, float* h_stage_ms8);

/* ---------------------------------------------------------------- random access (successors(v)) - */
/* graph.successors(v) for a batch of query nodes (examples/bench_random_access.rs:30-38): the reference
 * creates one decoder per query (bvgraph_decoder_factory.rs:46-58) and webgraph follows the reference chain
 * recursively; here the reference closure of all queries is decoded once and the query lists are gathered.
 *   d_nodes[n_queries]      query node ids (duplicates allowed, any order)
 *   d_offsets[n_queries+1]  CSR offsets of the answer (d_offsets[0] == 0)
 *   d_succ                  successors of query i at d_succ[d_offsets[i] .. d_offsets[i+1]); NULL = sizing call:
 *                           only d_offsets and *h_arcs are produced (one symbol per query is decoded)
 * Workspace: wga_successors_workspace_size(g, n_queries, max_total_arcs) with max_total_arcs >= the *h_arcs of
 * the sizing call (0 is enough for the sizing call itself).  The decoded closure may hold up to 4x that many
 * arcs and 8x n_queries nodes; deeper reference chains return WGA_E_WORKSPACE (retry with a larger bound). */
uint64_t wga_successors_workspace_size(const wga_graph* g, uint64_t n_queries, uint64_t max_total_arcs);
int wga_successors_batch(wga_graph* g, const uint64_t* d_nodes, uint64_t n_queries, uint64_t* d_offsets,
                         uint32_t* d_succ, uint64_t succ_capacity, void* d_workspace, uint64_t workspace_bytes,
                         uint64_t* h_arcs, void* stream);
/* Same with HOST arrays: copies the queries up, sizes, decodes into buffers owned by the handle (grow-only) and
 * copies h_offsets[n_queries+1] and the successors back.  h_succ == NULL or a too small succ_capacity: only
 * h_offsets and *h_arcs are produced and WGA_E_WORKSPACE is returned in the second case. */
int wga_successors_batch_host(wga_graph* g, const uint64_t* h_nodes, uint64_t n_queries, uint64_t* h_offsets,
                              uint32_t* h_succ, uint64_t succ_capacity, uint64_t* h_arcs);

/* ---------------------------------------------------------------- debug / parity hooks ---------- */
/* Expands the packed device tables of component c into the REFERENCE layout
 * (DecoderModelEntry, src/ans/models/component_model4decoder.rs:8-22: u16 freq, u16 cumul_freq, pad,
 * u64 quasi_folded; 16 bytes per slot, 2^frame_size slots) into h_out. Runs the device lookup per slot. */
int wga_debug_expand_table(wga_graph* g, int component, void* h_out, uint64_t n_slots);
/* Decodes `n` symbols of the given components (u8 each) from (ptr,state) [ptr==UINT64_MAX: sequential
 * start (stream_len, prelude.state)] with ONE device thread -- ANSDecoder::decode, src/ans/decoder.rs:58-87. */
int wga_debug_decode_symbols(wga_graph* g, const uint8_t* h_components, uint64_t n, uint64_t ptr, uint32_t state,
                             uint64_t* h_out, uint64_t* h_end_ptr, uint32_t* h_end_state);

/* Kernel tuning knobs of the decode path (process-wide, for tests and profiling: every decode call takes one
 * consistent snapshot of them under a lock; the parity tests shrink them so that small graphs cross unit
 * boundaries, refill lanes one by one and stride the grid).  Keys: "unit", "k1_blocks", "refill",
 * "k2_blocks", "k2_batch", "hub_min", "e2e_chunk", "reset". */
int wga_debug_set_tuning(const char* key, uint64_t value);

/* ---------------------------------------------------------------- model build -------------------- */
/* ANSModel4EncoderBuilder (src/ans/model4encoder_builder.rs:39-56) with device-resident histograms. */
int wga_model_create(wga_model** out);
void wga_model_destroy(wga_model* m);
/* dense canonical bins all-reduced across ranks: WGA_COMPONENTS * WGA_CANON_BINS u64 on the device */
#define WGA_CANON_BINS 20480
uint64_t* wga_model_bins(wga_model* m); /* device pointer (for ncclAllReduce / torch.distributed) */
/* push_symbol for n (component, raw symbol) pairs that live on the DEVICE (:67-78) */
int wga_model_accumulate(wga_model* m, const uint8_t* d_components, const uint64_t* d_symbols, uint64_t n,
                         void* stream);
/* same with HOST arrays (copies, then accumulates) */
int wga_model_accumulate_host(wga_model* m, const uint8_t* h_components, const uint64_t* h_symbols, uint64_t n);
/* The sparse part (distinct raw symbols >= 1024 with their counts), needed for the exact raw entropy
 * (:275-289). export/import let ranks exchange it (all-gather). Arrays are HOST memory. */
uint64_t wga_model_sparse_count(wga_model* m);
int wga_model_sparse_export(wga_model* m, uint8_t* h_components, uint64_t* h_symbols, uint64_t* h_counts);
int wga_model_sparse_merge(wga_model* m, const uint8_t* h_components, const uint64_t* h_symbols,
                           const uint64_t* h_counts, uint64_t n);
/* Sums the histograms of all ranks of an NCCL communicator (ncclComm_t passed as void*; `stream` = cudaStream_t):
 * ncclAllReduce of the dense bins, ncclAllGather + merge of the sparse tail of large raw symbols.  Collective: every
 * rank calls it once before wga_model_build, which then yields the identical tables on every rank.  NCCL is resolved
 * at run time from the process (libnccl.so.2); WGA_E_UNSUPPORTED when it is not there. */
int wga_model_allreduce(wga_model* m, void* nccl_comm, void* stream);
/* Convenience for callers without NCCL code of their own: ncclGetUniqueId (128 bytes, to be sent to the other ranks by
 * any means), ncclCommInitRank, ncclCommDestroy. */
int wga_nccl_get_unique_id(void* out128);
int wga_nccl_comm_init(int n_ranks, int rank, const void* id128, void** out_comm);
void wga_nccl_comm_destroy(void* comm);
/* build() (:80-271) on the GPU. out_tables[c].table points into memory owned by `m`. */
int wga_model_build(wga_model* m, wga_component_model out_tables[WGA_COMPONENTS], double* h_original_cost9,
                    double* h_final_cost9);

/* ---------------------------------------------------------------- bvcomp front end (host) -------- */
/* The host side of ANSBvGraph::store (src/bvgraph/random_access.rs:91-222). See INTEGRATION.md. */
typedef struct wga_symbols wga_symbols; /* (component, raw symbol) stream of one BvComp pass */
/* estimator: NULL tables => Log2Estimator (log2_estimator.rs:15-49), else EntropyEstimator built from
 * the given model (entropy_estimator.rs:33-113).  CSR input in host memory.  `chunk_nodes`==0 => one
 * sequential BvComp (bit-identical to the reference's single pass); >0 => independent BvComp per chunk of
 * nodes (webgraph's parallel compression: start_node = chunk start), run on `threads` host threads. */
int wga_bvcomp_symbols(const uint64_t* h_offsets, const uint32_t* h_succ, uint64_t n_nodes,
                       uint64_t compression_window, uint64_t max_ref_count, uint64_t min_interval_length,
                       const wga_component_model* estimator_tables, uint64_t chunk_nodes, int threads,
                       wga_symbols** out);
/* Same for nodes [first_node, first_node + n_nodes) only: the CSR arrays hold just these nodes (successor ids stay
 * global).  With chunk_nodes > 0 the chunks are those of the whole-graph call, so a rank of a sharded model build
 * gets exactly the whole graph's symbols of its node range (first_node a multiple of chunk_nodes). */
int wga_bvcomp_symbols_range(const uint64_t* h_offsets, const uint32_t* h_succ, uint64_t first_node, uint64_t n_nodes,
                             uint64_t compression_window, uint64_t max_ref_count, uint64_t min_interval_length,
                             const wga_component_model* estimator_tables, uint64_t chunk_nodes, int threads,
                             wga_symbols** out);
/* Same result, with the candidate-reference costing and the reference selection on the GPU (SURVEY.md 8f rank 2;
 * replaces the per-candidate estimator calls of BvComp, entropy_estimator.rs:81-113 via random_access.rs:108-125): one
 * device thread per (node, reference offset) pair, integer costs, so the chosen references -- and the symbols -- are
 * exactly those of wga_bvcomp_symbols_range.  Needs a CUDA device (no fallback); the host threads then compress only
 * the chosen candidate of every node. */
int wga_bvcomp_symbols_range_gpu(const uint64_t* h_offsets, const uint32_t* h_succ, uint64_t first_node, uint64_t n_nodes,
                                 uint64_t compression_window, uint64_t max_ref_count, uint64_t min_interval_length,
                                 const wga_component_model* estimator_tables, uint64_t chunk_nodes, int threads,
                                 wga_symbols** out);
/* Parity hook: the cost of every (node, reference offset) candidate record, h_costs[n_nodes * (window + 1)]
 * (UINT64_MAX: no such candidate), computed by the GPU kernel (use_gpu != 0) or by the host BvComp. */
int wga_debug_bvcomp_costs(const uint64_t* h_offsets, const uint32_t* h_succ, uint64_t first_node, uint64_t n_nodes,
                           uint64_t compression_window, uint64_t min_interval_length,
                           const wga_component_model* estimator_tables, uint64_t chunk_nodes, int use_gpu,
                           uint64_t* h_costs);
uint64_t wga_symbols_len(const wga_symbols* s);
const uint8_t* wga_symbols_components(const wga_symbols* s);
const uint64_t* wga_symbols_values(const wga_symbols* s);
void wga_symbols_free(wga_symbols* s);
/* ANSEncoder over the symbols in REVERSE order + a phase after every Outdegree
 * (src/ans/encoder.rs:39-78, src/bvgraph/writers/bvgraph_encoder.rs:159-174).  Outputs are malloc'ed by
 * the library; free with wga_free. */
int wga_ans_encode(const wga_component_model tables[WGA_COMPONENTS], const uint8_t* h_components,
                   const uint64_t* h_symbols, uint64_t n, uint16_t** out_stream, uint64_t* out_stream_len,
                   uint32_t* out_state, uint32_t** out_states, uint64_t** out_pointers, uint64_t* out_n_phases);
void wga_free(void* p);
/* Writes <basename>.ans/.pointers/.states in the reference's epserde layout (random_access.rs:198-221). */
int wga_write_files(const char* basename, const wga_prelude_view* view);
/* Full ANSBvGraph::store from a Java/webgraph BVGraph basename (.graph + .properties). Needs the GPU
 * (model build). */
int wga_store(const char* basename, const char* new_basename, uint64_t compression_window,
              uint64_t max_ref_count, uint64_t min_interval_length);
/* Same from a host CSR. */
int wga_store_csr(const uint64_t* h_offsets, const uint32_t* h_succ, uint64_t n_nodes, const char* new_basename,
                  uint64_t compression_window, uint64_t max_ref_count, uint64_t min_interval_length,
                  uint64_t chunk_nodes, int threads);
/* Java/webgraph BVGraph reader (what BvGraphSeq::with_basename(..).load() gives the reference,
 * random_access.rs:101-103). Two calls: sizes, then fill. */
int wga_bvgraph_read(const char* basename, uint64_t* n_nodes, uint64_t* n_arcs, uint64_t* h_offsets,
                     uint32_t* h_succ);
/* Elias-Fano (sux 0.4.6 layout) write/read used for .pointers and for the golden .ef test. */
int wga_ef_write(const char* path, const uint64_t* values, uint64_t n, uint64_t u);
int wga_ef_read(const char* path, uint64_t* n, uint64_t* h_values);

/* ---------------------------------------------------------------- synthetic graphs ---------------- */
/* Counter-based generators keyed by node id: kind 0 = web-like (locality, copying,
 * intervals), 1 = social-like (power-law degrees, large gaps).  Two calls: h_succ==NULL returns arcs. */
int wga_synth_graph(int kind, uint64_t n_nodes, double mean_degree, uint64_t seed, uint64_t first,
                    uint64_t last, int threads, uint64_t* h_offsets, uint32_t* h_succ, uint64_t* n_arcs);

#ifdef __cplusplus
}
#endif
#endif /* WGA_H */
