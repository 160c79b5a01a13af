/*
 * isocon_nn.h -- C ABI of libisocon_nn.so: the B200 (sm_100a) implementation of IsoCon's
 * nearest-neighbour-graph hot path.
 *
 * The reference has no FFI for this path: it is Python calling the third-party `edlib`
 * module once per pair from two scan loops.  The entry points below are what a binding for
 * that path binds instead (one call per GRAPH, not per pair); each cites the reference
 * interface it replaces (paths relative to the IsoCon repository):
 *
 *   isocon_nn_store_add     the sorted list `seq_to_acc_list_sorted` that every function of
 *     + isocon_nn_set_list  modules/nearest_neighbor_graph.py receives (built at :243-246 and
 *     (isocon_nn_set_reads) :202-208): sequences resident on the device, then named in list order.
 *   isocon_nn_graph_begin   get_nearest_neighbors(batch, global_index, start_index, list,
 *     + _run + _finalize    has_converged, depth) :110-198   (mode 1)  and
 *                           get_nearest_neighbors_2set(batch, start_index, list,
 *                           target_accessions, depth) :341-424 (mode 2), including their Pool
 *                           drivers get_exact_nearest_neighbor_graph[_2set] :19-82, :300-334.
 *   isocon_nn_ed_pairs      edlib_ed(x, y, mode="NW", task="distance", k) :104-107, batched.
 *
 * Conventions: every function returns 0 on success and a non-zero code on failure;
 * isocon_nn_last_error() then describes it.  All pointers are plain host pointers unless the
 * name ends in _dev.  Indices are positions in the list given to isocon_nn_set_reads.
 * There is no CPU fallback: without a CUDA device every call fails with ISOCON_ERR_CUDA.
 */
#ifndef ISOCON_NN_H
#define ISOCON_NN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct isocon_nn_ctx isocon_nn_ctx;

enum {
    ISOCON_OK = 0,
    ISOCON_ERR_CUDA = 1,      /* CUDA runtime / launch failure, or no device */
    ISOCON_ERR_ARG = 2,       /* bad argument */
    ISOCON_ERR_ALPHABET = 3,  /* a read holds a symbol outside the store's 4-symbol alphabet */
    ISOCON_ERR_STATE = 4,     /* call sequence violated */
    ISOCON_ERR_OVERFLOW = 5   /* candidate-edge buffer too small: reserve_edges + rebuild (the bindings do) */
};

enum { ISOCON_ALGO_AUTO = 0, ISOCON_ALGO_TILE = 1, ISOCON_ALGO_SCAN = 2 };

enum {
    ISOCON_PHASE_SEED = 1,    /* cheap upper bounds from length-adjacent targets */
    ISOCON_PHASE_MAIN = 2,    /* all pairs of this rank's row tiles.  After a PILOT pass the targets are first
                                 re-binned by threshold class (window words their pairs need).  One-sided graphs
                                 (2-set) climb a ladder of threshold caps: rows still unresolved after a pass are
                                 redone at the next cap.  world <= 1: all passes in one call; several ranks: ONE pass
                                 per call -- reduce best[] (MIN) and call again until isocon_nn_last_run_rows is 0 */
    ISOCON_PHASE_WIDE = 4,    /* rows still unresolved above the register-band limit: any threshold */
    ISOCON_PHASE_PILOT = 8,   /* symmetric 1-set graph: the first 10 % of the rows against everything behind them,
                                 so that best[] is a usable bound for every read (replaces SEED there) */
    ISOCON_PHASE_ALL = 15     /* run order: SEED, PILOT, MAIN, WIDE */
};

typedef struct {
    int32_t mode;             /* 1 = 1-set (reads vs reads), 2 = 2-set (reads vs candidates) */
    int32_t algo;             /* ISOCON_ALGO_*; AUTO = TILE unless (mode 2 and finite depth) */
    int64_t depth;            /* neighbor_search_depth (default of the reference: 2^32) */
    const uint8_t* is_query;  /* [n] 1 = this list entry is a query of the call.
                                 mode 1: entry in the batch range and sequence not in has_converged
                                 mode 2: entry in the batch range and accession not in target_accessions */
    const uint8_t* is_target; /* [n] mode 2: accession in target_accessions; mode 1: NULL (all entries) */
    int32_t symmetric;        /* mode 1, TILE: evaluate each unordered pair once (-1 = default on) */
    int32_t rank, world;      /* shard of the cost-balanced row tiles this context computes (0,1 = all) */
} isocon_nn_params;

/* Work counters of the last graph (device side, this rank). */
typedef struct {
    uint64_t pairs;           /* (query, target) pairs aligned */
    uint64_t word_columns;    /* 32-row x 1-column bit-vector updates executed (x32 lanes) */
    uint64_t groups;          /* warp tasks (query x 32 targets) executed */
    uint64_t wide_pairs;      /* pairs that needed the wide (global-memory) band */
    uint64_t items;           /* row tiles handed out */
    uint64_t edges_raw;       /* candidate edges appended before the tie filter */
    uint64_t launches;        /* kernels launched since graph_begin (begin, run, finalize) */
    uint64_t bins;            /* threshold-class bins of the target layout the MAIN pass used */
    uint64_t pilot_rows;      /* rows aligned by the PILOT pass */
    uint64_t unresolved_rows; /* rows the WIDE pass had to redo above the register-band limit */
    uint64_t useful_cells;    /* row kernel: sum over aligned pairs and over the columns until that pair's answer was
                                 known of the rows of the pair's own Ukkonen strip that the window in force still
                                 held: the DP cells the thresholds made necessary, without word padding, lock-step
                                 waiting or bookkeeping */
    uint64_t columns;         /* row kernel: DP columns walked by the warps (x32 lanes); word_columns / columns =
                                 mean window width in words (it shrinks along a walk, diag_band.cuh) */
    uint64_t main_passes;     /* MAIN passes launched (one-sided graphs climb a ladder of threshold caps) */
    uint64_t clusters;        /* similarity clusters the MAIN pass's targets were ordered by (0 = list order) */
} isocon_nn_stats;

/* Resident read store (isocon_nn_store_info). */
typedef struct {
    uint64_t slots;           /* sequences resident in the store */
    uint64_t arena_words;     /* 32-bit words of packed reads in use */
    uint64_t list_entries;    /* entries of the current list (isocon_nn_set_list) */
    uint64_t uploaded_reads;  /* cumulative since the context was created: sequences sent host -> device */
    uint64_t uploaded_bytes;  /*   ... and their ASCII bytes */
    uint64_t upload_calls;    /*   ... and the store_add calls that carried them */
    uint64_t resets;          /* store_reset calls */
    uint64_t lists;           /* set_list calls */
    uint64_t foreign_reads;   /* cumulative: uploaded sequences with a symbol outside the alphabet */
} isocon_nn_store_stats;

int isocon_nn_device_count(int* count);
int isocon_nn_create(int device, isocon_nn_ctx** out);
void isocon_nn_destroy(isocon_nn_ctx* ctx);
const char* isocon_nn_last_error(const isocon_nn_ctx* ctx); /* ctx may be NULL (creation errors) */

/* ---- Resident read store: reads live on the device 2-bit packed and STAY there across graph builds.
 *
 * The reference rebuilds the graph once per correction round on mostly unchanged sequences
 * (modules/isocon_get_candidates.py:141-214: only reads of unconverged partitions change, :186-201) and re-sends
 * the whole read list to every worker each time (nearest_neighbor_graph.py:34, :318).  Here a sequence is uploaded
 * ONCE: store_add packs new sequences into slots of a device arena (slots never move), set_list names the slots of
 * the sorted list `seq_to_acc_list_sorted` (:243-246, :202-208) the next graphs work on.  A round therefore costs
 * the upload of the sequences the correction changed (delta upload) plus 12 bytes per list entry.
 *
 *   isocon_nn_store_reset   forget every slot.  alphabet: the 4 symbols packed as codes 0..3 (NULL = "ACGT");
 *                           comparisons are exact on raw symbols like edlib's, so any 4 distinct bytes work.
 *   isocon_nn_store_add     n_new sequences (`ascii` concatenated, offsets[n_new + 1]) -> slots
 *                           [*first_slot, *first_slot + n_new).  ISOCON_ERR_ALPHABET (nothing added) when a
 *                           sequence holds a symbol outside the store's alphabet.
 *   isocon_nn_set_list      the sorted list: entry i is slot slots[i]; lengths must not decrease.  Indices of every
 *                           later call (is_query, edges, ed_pairs) are positions in this list.
 *   isocon_nn_host_buffer   pinned host memory of at least `bytes` owned by the context: sequences gathered there
 *                           go to the device in one asynchronous copy when handed to store_add (any other host
 *                           pointer is staged through pinned buffers internally).  Valid until the next call.
 *   isocon_nn_set_reads     reset + add + identity list in one call (the whole list is new). */
int isocon_nn_store_reset(isocon_nn_ctx* ctx, const uint8_t* alphabet /* [4] or NULL */);
int isocon_nn_store_add(isocon_nn_ctx* ctx, const uint8_t* ascii, const int64_t* offsets, int64_t n_new, int64_t* first_slot);
int isocon_nn_set_list(isocon_nn_ctx* ctx, const int32_t* slots, int64_t n);
int isocon_nn_host_buffer(isocon_nn_ctx* ctx, int64_t bytes, void** ptr);
int isocon_nn_store_info(isocon_nn_ctx* ctx, isocon_nn_store_stats* out);
int isocon_nn_set_reads(isocon_nn_ctx* ctx, const uint8_t* ascii, const int64_t* offsets, int64_t n);

/* Build a graph in three steps so a multi-GPU driver can reduce `best` across ranks between
 * them.  Single GPU: begin, run(ISOCON_PHASE_ALL), finalize, fetch. */
int isocon_nn_graph_begin(isocon_nn_ctx* ctx, const isocon_nn_params* params);
int isocon_nn_graph_run(isocon_nn_ctx* ctx, int phases);
/* Rows (queries) the last graph_run call scheduled, counted BEFORE the split across ranks: the same
 * number on every rank, so a multi-GPU driver can skip the reduction after a phase that ran nowhere. */
int isocon_nn_last_run_rows(isocon_nn_ctx* ctx, int64_t* rows);
/* Device pointer of best[n] (int32: running best distance per list entry; len(seq) when nothing
 * closer was found).  A multi-GPU driver all-reduces it (MIN) in place between phases. */
int isocon_nn_best_dev(isocon_nn_ctx* ctx, void** best_dev);
/* Several ranks: call right after the MIN-reduce of best[] that follows a phase, then let no rank start its next phase
 * before every rank has returned from this call (any collective will do).  It keeps a snapshot of the agreed best[]:
 * the host-side decisions of the next phase (threshold classes, ladder rows, WIDE rows) are taken from the snapshot,
 * so they are the same on every rank although peers that run ahead keep lowering the live best[] over NVLink. */
int isocon_nn_best_agree(isocon_nn_ctx* ctx);
/* After a PILOT phase that recorded them: device pointer of pnear[2n] (uint64: distance << 32 | pilot row; ~0 = none),
 * the two nearest pilot rows of every list entry, from which the MAIN phase orders its targets by similarity
 * (isocon_nn.cu: cluster_order).  Every rank must hand the MAIN phase the same values: a multi-GPU driver gathers
 * the ranks' arrays and writes the two smallest entries per read back on every rank.  count = 0: nothing to merge. */
int isocon_nn_pilot_near_dev(isocon_nn_ctx* ctx, void** dev, int64_t* count);
/* ---- The ranks of one box over NVLink peer memory (optional; one process per GPU).
 *
 * ipc_handles: this context's handle record (ISOCON_IPC_BYTES): three CUDA IPC handles -- best[], the block of tile
 * queues, and the "share" block (nearest-pilot-row records, barrier / result counters, final edges) with its layout
 * -- and a generation number that changes whenever one of the allocations moves (the records must then be exchanged
 * again).
 * set_peers: the records of all `world` ranks in rank order (this rank's own entry is ignored); world <= 1 closes the
 * peer mappings.  Call it on every rank, then let no rank go on before all have returned.  Afterwards
 *  (1) every improvement of best[x] found by the pair kernels is also applied to the peers' best[x] with system-scope
 *      atomicMin, so all ranks prune with the box-wide running best instead of their own share;
 *  (2) the PILOT / MAIN / WIDE launches of all ranks pull their row tiles from ONE queue in rank 0's memory
 *      (system-scope atomicAdd), so the GPUs of the box finish together instead of each draining a fixed share;
 *  (3) isocon_nn_graph_run(ISOCON_PHASE_ALL) runs the FUSED flow (isocon_nn_can_fuse): every phase in one call, the
 *      ranks meeting at device-side barriers (a counter per rank in the share block) instead of collectives --
 *      after a barrier all copies of best[] are equal because every improvement went to all of them -- and
 *      isocon_nn_graph_finalize delivers the surviving edges of every rank to every rank's share block, so
 *      isocon_nn_graph_fetch returns the whole graph on each rank.  No NCCL call on the data path.
 * Results do not depend on any of this (any threshold >= the final best is valid, every tile is computed by exactly
 * one rank).  Without mapped peers (several nodes, no peer access) the driver runs the phases one by one and connects
 * the ranks with collectives: MIN all-reduce of best[] + isocon_nn_best_agree after each phase, all-gather of the edges. */
#define ISOCON_IPC_BYTES 256
int isocon_nn_ipc_handles(isocon_nn_ctx* ctx, uint8_t handles[ISOCON_IPC_BYTES], uint64_t* generation);
int isocon_nn_set_peers(isocon_nn_ctx* ctx, const uint8_t* handles, int32_t world, int32_t rank);
/* 1 when the graph that was begun can run fused: several ranks, all peers mapped, pair-matrix algorithm. */
int isocon_nn_can_fuse(isocon_nn_ctx* ctx, int32_t* yes);
/* Free best[] allocations that were exported and later outgrown; call once every rank has re-run
 * set_peers (i.e. closed its mapping of them) and a barrier has passed. */
int isocon_nn_release_retired(isocon_nn_ctx* ctx);

/* Keep the edges whose distance equals best[query]; returns their number.  ISOCON_ERR_OVERFLOW: the candidate-edge
 * buffer (default max(2^20, 64 n) edges) was too small for this input (tie-heavy late rounds); stats.edges_raw
 * holds the number needed -- reserve more with isocon_nn_reserve_edges and build the graph again (best[] of the
 * failed build is exact, only edges were dropped). */
int isocon_nn_graph_finalize(isocon_nn_ctx* ctx, int64_t* n_edges);
/* Capacity (edges) of the candidate-edge buffer of the graphs begun from now on: > 0 at least that many,
 * 0 the default, < 0 exactly -capacity (memory-constrained callers, tests of the overflow path). */
int isocon_nn_reserve_edges(isocon_nn_ctx* ctx, int64_t capacity);
/* best[n] and the surviving edges (query index, neighbour index, distance), unordered. */
int isocon_nn_graph_fetch(isocon_nn_ctx* ctx, int32_t* best, int32_t* edge_q, int32_t* edge_t, int32_t* edge_d);
/* Device pointers of the finalized edge arrays (for a gather over NVLink). */
int isocon_nn_edges_dev(isocon_nn_ctx* ctx, void** q_dev, void** t_dev, void** d_dev);

/* Batched edlib_ed: out[p] = distance(read a[p], read b[p]) if <= k[p] else -1; k == NULL or
 * k[p] < 0 means unbounded. */
int isocon_nn_ed_pairs(isocon_nn_ctx* ctx, const int32_t* a, const int32_t* b, const int32_t* k,
                       int64_t n_pairs, int32_t* out);

int isocon_nn_get_stats(isocon_nn_ctx* ctx, isocon_nn_stats* out);
/* Device time (CUDA events on the library's stream) of the last call of: 0 = set_reads,
 * 1 = graph_begin + graph_run (accumulated since graph_begin), 2 = finalize, 3 = ed_pairs, 4 = int32 probe,
 * 5 = the pair kernels alone (SEED + PILOT + MAIN + WIDE launches, accumulated since graph_begin). */
int isocon_nn_last_ms(isocon_nn_ctx* ctx, int which, float* ms);
/* Block until the library's stream is idle. */
int isocon_nn_sync(isocon_nn_ctx* ctx);
/* Stop-watch on the library's stream (CUDA events): start records an event now, stop records a
 * second one, waits for it and returns the device time between the two -- the bracket bench.py
 * puts around its K timed steps. */
int isocon_nn_timer_start(isocon_nn_ctx* ctx);
int isocon_nn_timer_stop(isocon_nn_ctx* ctx, float* ms);

/* Dependency-free LOP3/IADD3 micro-kernel: measured INT32 ALU issue rate of this device in
 * lane-operations per second (the roofline denominator of SURVEY.md §8d). */
int isocon_nn_int32_peak(isocon_nn_ctx* ctx, double* lane_ops_per_s);

#ifdef __cplusplus
}
#endif
#endif
