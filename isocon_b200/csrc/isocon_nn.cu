// isocon_nn.cu -- host side of libisocon_nn.so (C ABI in include/isocon_nn.h).
//
// Owns the device-resident read store (2-bit packed, row-major + 32-way interleaved target
// groups), cuts the length-sorted pair matrix into row tiles (one query x 8 groups of 32
// targets), runs the phases SEED -> MAIN -> WIDE of the tile algorithm or the sequential
// scan emulation, and filters the candidate edges down to the ties at the final best.
// Reference semantics: IsoCon modules/nearest_neighbor_graph.py :19-82, :110-198, :300-334,
// :341-424.  No CPU path: every entry point needs a CUDA device.
#include <algorithm>
#include <climits>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>
#include <chrono>

#include "../../include/isocon_nn.h"
#include "nn_kernels.cuh"

using namespace isocon;

namespace {

static constexpr size_t STAGE_BYTES = 4u << 20;

std::string g_create_error;

template <class T>
struct DBuf {
    T* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        const size_t want = n + n / 8 + 64;
        cudaError_t e = cudaMalloc((void**)&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    // grow and keep the first `used` elements (the read arena: slots stay where they are)
    cudaError_t grow_keep(size_t used, size_t n, cudaStream_t st) {
        if (n <= cap) return cudaSuccess;
        const size_t want = std::max(n + n / 4 + 64, cap * 2);
        T* np = nullptr;
        cudaError_t e = cudaMalloc((void**)&np, want * sizeof(T));
        if (e != cudaSuccess) return e;
        if (p && used) e = cudaMemcpyAsync(np, p, used * sizeof(T), cudaMemcpyDeviceToDevice, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (p) cudaFree(p);
        p = np; cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// Pinned host memory handed out in pieces (bump allocation): every small table a graph uploads is copied here
// first, so the H2D copy is truly asynchronous and the caller's vectors may go away at once.  reset() when the
// stream is known to be idle (graph_begin); a piece that does not fit makes the caller synchronise and reset.
struct PinnedArena {
    uint8_t* p = nullptr;
    size_t cap = 0, used = 0;
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0; used = 0;
        const size_t want = n + n / 2 + 4096;
        cudaError_t e = cudaHostAlloc((void**)&p, want, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void* take(size_t n) {
        const size_t at = (used + 63) & ~(size_t)63;
        if (at + n > cap) return nullptr;
        used = at + n;
        return p + at;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = used = 0; }
};

// Tile table of one pass (see GraphArgs in nn_kernels.cuh).
struct ItemTable {
    int gpi = GROUPS_PER_ITEM;   // groups per row tile (upper bound; tiles of a row are equal)
    bool row_kernel = false;     // nn_row_kernel (one query per block) instead of nn_tile_kernel
    bool swapped = false;        // row kernel with the roles swapped (nn_row_swapped_kernel): rows are targets, lanes queries
    std::vector<int> qlist, segoff, gtotal, gsize, seg_g0, seg_n;
    std::vector<long long> item_off;
    long long total() const { return item_off.empty() ? 0 : item_off.back(); }
    void add_row(int q) { qlist.push_back(q); segoff.push_back((int)seg_g0.size()); gtotal.push_back(0); }
    void add_segment(int g0, int cnt) { seg_g0.push_back(g0); seg_n.push_back(cnt); gtotal.back() += cnt; }
};

}  // namespace

// Layout of a share block of capacities (n_cap entries, f_cap final edges).
enum { CT_BAR = 0, CT_FCOUNT = 1, CT_NEEDED = 2, CT_ERR = 3, CT_WORDS = 16 };
struct ShareView {
    unsigned long long* pnear = nullptr;      // [2 n_cap]
    unsigned long long* ctrl = nullptr;       // [CT_WORDS]
    int* fq = nullptr; int* ft = nullptr; int* fd = nullptr;   // [f_cap] each
    long long f_cap = 0;
};
static size_t share_bytes(long long n_cap, long long f_cap) {
    return (size_t)n_cap * 16 + CT_WORDS * 8 + (size_t)f_cap * 12 + 64;
}
static ShareView share_view(uint8_t* base, long long n_cap, long long f_cap) {
    ShareView v;
    v.pnear = (unsigned long long*)base;
    v.ctrl = v.pnear + 2 * n_cap;
    v.fq = (int*)(v.ctrl + CT_WORDS);
    v.ft = v.fq + f_cap;
    v.fd = v.ft + f_cap;
    v.f_cap = f_cap;
    return v;
}

struct isocon_nn_ctx {
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evt0 = nullptr, evt1 = nullptr;
    static constexpr int KEV = 16;           // event pairs around the pair-matrix kernel launches of one graph_run
    cudaEvent_t kev[2 * KEV] = {};
    int kev_used = 0;
    float ms[6] = {0, 0, 0, 0, 0, 0};
    std::string err;

    // options
    long long opt_edge_capacity = 0;
    int opt_kcap_main = KCAP_MAIN;
    int opt_seed = 1;
    int opt_blocks_per_sm = 0;

    // read store: every sequence ever added since the last reset lives in a SLOT of the packed arena d_rowpk
    // (2 bits per base, 16 bases per word, + 4 zero words); slots never move, so a later graph over mostly the
    // same sequences (the next correction round) only uploads the sequences it has not seen
    uint8_t alphabet[4] = {'A', 'C', 'G', 'T'};
    std::vector<int> s_len;                 // slot -> length
    std::vector<long long> s_off;           // slot -> first word in the arena
    long long arena_used = 0;               // words of d_rowpk in use
    // sequences with a symbol outside the alphabet ("foreign") also keep their raw bytes (general-alphabet path)
    std::vector<long long> s_foff;          // slot -> first byte in d_fascii, or -1
    std::vector<int> s_nsym;                // slot -> distinct symbols (foreign slots only)
    long long fascii_used = 0;
    DBuf<uint8_t> d_fascii;
    DBuf<long long> d_foff;                 // list entry -> byte offset or -1
    std::vector<uint8_t> h_foreign;         // list entry -> foreign?
    std::vector<uint8_t> h_ist_main;        // targets of the 2-bit pair kernels (foreign entries left out)
    std::vector<int> h_flist;               // foreign list entries that take part in the graph
    DBuf<int> d_flist;
    long long n_foreign = 0;                // foreign entries of the list
    int gen_syms = 4;                       // most distinct symbols of any list entry (general-alphabet table rows)
    int foreign_level = 0;                  // foreign passes launched for this graph (0..2)
    long long scr_stride = 0;               // scratch words per warp
    isocon_nn_store_stats store{};          // cumulative upload counters
    PinnedArena host_buf;                   // isocon_nn_host_buffer: the caller gathers its sequences here
    PinnedArena bounce;                     // small tables of a graph on their way to the device
    PinnedArena best_host;                  // best[] fetched for the host-side re-binning
    unsigned long long best_host_launches = ~0ull;   // value of `launches` when best_host was fetched
    DBuf<uint8_t> d_flag;
    // the list the graphs work on: entry i = slot h_slot[i], sorted by length
    long long n = 0;
    int max_len = 0, nbmax = 1, peq_words = 1 + PEQ_PAD_WORDS;
    std::vector<int> h_len, h_slot;
    DBuf<uint8_t> d_ascii;
    DBuf<long long> d_off, d_rowoff, d_newoff;
    DBuf<int> d_len;
    DBuf<uint32_t> d_rowpk;
    DBuf<unsigned long long> d_small;  // [0] bad symbol, [1] work counter, [2] ecount, [3] fcount, [8..] stats

    // graph
    bool graph_open = false, finalized = false;
    isocon_nn_params prm{};
    int algo = ISOCON_ALGO_TILE;
    int symmetric = 0;
    std::vector<uint8_t> h_isq, h_ist;
    std::vector<int> h_qlist;
    // target layout: slots of 32-target groups; a bin is a run of whole groups holding targets of one
    // threshold class in list (= length) order, padded with -1
    std::vector<int> h_tpos;                  // slot -> list index or -1
    std::vector<int> bin_first, bin_count;    // bin -> first slot (multiple of 32), valid targets
    int nT = 0, nG = 0;                       // slots, groups
    bool binned = false;
    bool all_queries = false;
    DBuf<uint8_t> d_isq, d_ist;
    DBuf<int> d_tpos, d_best, d_qlist, d_segoff, d_gtotal, d_gsize, d_seg_g0, d_seg_n;
    DBuf<long long> d_goff, d_item_off;
    DBuf<uint32_t> d_il, d_scratch;
    DBuf<int> d_eq, d_et, d_ed;
    // The "share" block: what the other ranks of a box write into over NVLink besides best[] and the tile queues --
    // nearest-pilot-row records (pnear), barrier / result counters (ctrl) and the graph's final edges (every rank
    // pushes the edges it found to ALL ranks, so each ends up with the whole graph without a collective).  One
    // allocation, one IPC handle; a single GPU uses it the same way with no peers.
    DBuf<uint8_t> d_share;
    long long share_n = 0, share_f = 0;       // capacities: list entries, final edges
    ShareView sv{};                           // own block
    ShareView peer_sv[7] = {};                // the peers' blocks (same order as peer_best)
    uint8_t* peer_share[7] = {};
    unsigned long long bar_seq = 0;           // barriers enqueued since the peers were connected
    bool fused = false;                       // this graph runs all phases in one call with device-side barriers
    // the last pilot rows run as a second launch queued right behind the first one, so the GPU has work while the
    // host turns the first launch's results into the MAIN pass's layout and tile table
    int opt_primer = 0;                       // PILOT: the first row as a launch of its own from this many ranks on (0 = never)
    int opt_bridge = 40;                      // pilot rows per GPU in the second launch (0 = one launch)
    cudaEvent_t ev_pilot = nullptr;           // best[] (and pnear) of the first PILOT launch are on the host
    bool pilot_prefetched = false;
    PinnedArena fetch_host;                   // finalize: best[] and the first edges, fetched with the counters
    long long spec_edges = 0;                 // edges already on the host after finalize
    long long ecap = 0, n_final = 0;
    long long edge_reserve = 0;               // isocon_nn_reserve_edges: capacity a caller asked for after an overflow
    int grid = 0;
    size_t smem = 0;
    // row kernel (diagonal band, one query per block): 0 grid = unavailable (reads too long)
    int row_grid = 0, row_padbits = 0, row_xmax = 0;
    size_t row_smem = 0;
    int opt_row_kernel = 1;
    // symmetric 1-set graph: pilot rows, then targets re-binned by threshold class (see graph_run)
    int opt_bins = 1;
    int opt_pilot_div = 10;       // the pilot is 1/opt_pilot_div of the rows
    int opt_narrow = 4;           // row kernel: shrink the diagonal window every N chunks of 32 columns (0 = never)
    uint8_t* stage[2] = {nullptr, nullptr};   // pinned staging buffers of the ASCII upload (STAGE_BYTES each)
    cudaEvent_t stage_ev[2] = {};
    int opt_class_gran = 32;      // threshold classes of the target bins: best / gran (32 = one window word)
    int opt_ladder_first = 0;     // > 0: first cap of the ladder (tests: forces several passes)
    int opt_ladder = 1;           // one-sided MAIN passes climb a ladder of threshold caps (0 = one pass at kcap)
    int ladder_prev = -1;         // cap of the last MAIN pass of this graph (-1: none yet)
    int ladder_level = 0;         // MAIN passes launched so far
    size_t seed_rows = 0;         // queries the SEED phase sampled
    bool main_done = false;       // the single-pass MAIN phase has run
    int opt_debug = 0;
    // similarity order of the MAIN pass's targets (see cluster_order)
    int opt_cluster = 1;
    int opt_order_best = 2;       // minor sort key inside a cluster: 0 nearest pilot row, 1 own bound, 2 pilot row then bound
    int opt_fuse = 1;             // several ranks with mapped peers: all phases in one call, device-side barriers
    bool cluster_pilot = false;   // the PILOT launch records every entry's two nearest pilot rows
    bool clustered = false;       // the target layout is in similarity order and unordered pairs are split by rank
    bool bins_unsorted = false;   // the bins are not in length order: a row takes them whole (the lanes prune by length)
    DBuf<unsigned long long> d_sig;   // min-hash signatures of the targets (one-sided passes)
    PinnedArena sig_host;
    std::vector<int> h_hint_g0, h_hint_n;   // per query: the groups of the cluster it is probably related to (-1: no hint)
    std::vector<int> h_hint_rep;            // per query: representative of the cluster the SEED pass covered entirely (-1: none)
    DBuf<unsigned long long> d_sigkeys; DBuf<int> d_sigvals, d_hint;
    // two-level one-sided passes (see sketch_order / two_level_pass)
    int opt_two_level = 1;
    int opt_seed_sample = 1;     // hinted SEED rows: a sample picks the first cap
    int opt_swap = 1;            // hinted SEED rows: candidates as rows, reads as lanes (nn_row_swapped_kernel)
    std::vector<int> h_root;     // sketch_order: representative of every clustered target's cluster (-1: none)
    long long opt_surv_cap = 0;             // tests: capacity of the survivor buffer (forces the fall-back)
    bool two_level = false;
    std::vector<int> h_tposA, h_tposB;      // layouts: all targets cluster after cluster / the cluster representatives
    std::vector<int> h_slack;               // list index -> radius of its cluster if it is a representative, else 0
    std::vector<int> cl_g0, cl_ng;          // representative's list index -> its cluster's groups in layout A
    DBuf<int> d_slack, d_sq, d_st;          // slack on the device; survivors (query, representative) of level 1
    DBuf<uint32_t> d_qgram;                 // q-gram bit sets of the representatives (layout B), see nn_kernels.cuh
    bool qgram_ready = false;
    int opt_qgram = 1;
    long long surv_cap = 0;
    DBuf<int> d_rank, d_snap;
    bool snap_valid = false;      // d_snap holds the best[] all ranks agreed on after the last phase
    std::vector<int> h_rank;
    PinnedArena pnear_host;
    size_t pilot_rows = 0;        // leading rows aligned by the PILOT pass
    isocon_nn_stats stats{};
    unsigned long long launches = 0;

    // peers (NVLink sharing of best[] between the ranks of a box)
    unsigned long long best_generation = 0;   // bumps when d_best moves
    int* peer_best[7] = {};
    int n_peers = 0;
    unsigned long long* peer_small[7] = {};   // the peers' d_small (same order as peer_best)
    unsigned long long* root_small = nullptr; // rank 0's d_small when peers are connected (own copy on rank 0)
    long long last_run_rows = 0;              // rows scheduled by the last graph_run (before sharding)
    bool best_exported = false;               // IPC handles of d_best / d_share were handed out
    std::vector<void*> retired_best;          // exported allocations that peers may still have mapped

    // pairs
    DBuf<int> d_pa, d_pb, d_pk, d_pout;
    DBuf<long long> d_runoff;
};

namespace {

struct DebugLap {   // ISOCON_NN_DEBUG=2: host wall time between laps, to stderr
    bool on; const char* what; std::chrono::steady_clock::time_point t;
    DebugLap(bool on_, const char* w) : on(on_), what(w), t(std::chrono::steady_clock::now()) {}
    void lap(const char* name) {
        if (!on) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[isocon_nn]   %s/%s %.3f ms\n", what, name, std::chrono::duration<double, std::milli>(now - t).count());
        t = now;
    }
};

void close_peers(isocon_nn_ctx* c) {
    for (int p = 0; p < c->n_peers; ++p) {
        if (c->peer_best[p]) cudaIpcCloseMemHandle(c->peer_best[p]);
        if (c->peer_small[p]) cudaIpcCloseMemHandle(c->peer_small[p]);
        if (c->peer_share[p]) cudaIpcCloseMemHandle(c->peer_share[p]);
        c->peer_best[p] = nullptr; c->peer_small[p] = nullptr; c->peer_share[p] = nullptr; c->peer_sv[p] = ShareView{};
    }
    c->n_peers = 0;
    c->root_small = nullptr;
}

int fail(isocon_nn_ctx* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}

#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(ctx, ISOCON_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// d_small: [0] bad symbol, [1] tile counter of a launch, [2] edge count, [3] filtered edge count,
// [4..6] box-wide tile queues of the PILOT / MAIN / WIDE launches (rank 0's copy is the one all ranks pull
// from over NVLink; zeroed at the END of a graph so no rank can race the reset), [8..] work counters
// SM_QUEUE: box-wide tile queues (rank 0's copy is the one in use): 0 PILOT, 2 WIDE, 1 and 3..7 the MAIN passes
enum { SM_SURV = 0, SM_COUNTER = 1, SM_ECOUNT = 2, SM_FCOUNT = 3, SM_QUEUE = 4, SM_NQUEUE = 8, SM_STATS = 12, SM_WORDS = 12 + ST_COUNT };

int configure_launch(isocon_nn_ctx* ctx) {
    ctx->smem = (size_t)WARPS_PER_BLOCK * ctx->peq_words * 4 * sizeof(uint32_t);
    if (ctx->smem > 48 * 1024) {
        CU(cudaFuncSetAttribute(nn_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem));
        CU(cudaFuncSetAttribute(nn_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem));
        CU(cudaFuncSetAttribute(ed_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem));
    }
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nn_tile_kernel, WARPS_PER_BLOCK * 32, ctx->smem));
    if (per_sm < 1) return fail(ctx, ISOCON_ERR_ARG, "reads of %d bases need %zu B of shared memory per block: too long",
                                ctx->max_len, ctx->smem);
    if (ctx->opt_blocks_per_sm > 0) per_sm = std::min(per_sm, ctx->opt_blocks_per_sm);
    ctx->grid = ctx->num_sms * per_sm;  // persistent: a multiple of the SM count
    // row kernel: the 32-way shifted mask table of one query per block
    ctx->row_padbits = 32 * ((ctx->opt_kcap_main + 31) / 32);
    ctx->row_xmax = ((ctx->row_padbits + ctx->max_len) >> 5) + TAB_TAIL_WORDS;
    ctx->row_smem = ((size_t)ctx->row_xmax * 128 + (size_t)(ctx->row_xmax + 1) * 4) * sizeof(uint32_t);
    ctx->row_grid = 0;
    int max_optin = 0;
    CU(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
    if (ctx->opt_row_kernel && ctx->row_smem + 64 <= (size_t)max_optin) {
        CU(cudaFuncSetAttribute(nn_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->row_smem));
        CU(cudaFuncSetAttribute(nn_row_swapped_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->row_smem));
        int row_per_sm = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&row_per_sm, nn_row_kernel, ROW_WARPS * 32, ctx->row_smem));
        if (ctx->opt_blocks_per_sm > 0) row_per_sm = std::min(row_per_sm, ctx->opt_blocks_per_sm);
        ctx->row_grid = ctx->num_sms * row_per_sm;
    }
    const size_t warps = std::max((size_t)ctx->grid * WARPS_PER_BLOCK, (size_t)ctx->row_grid * ROW_WARPS);
    // per warp: band state of the global-memory fall-back (96 nbmax words) + the general-alphabet match table
    ctx->scr_stride = 96ll * ctx->nbmax + (ctx->n_foreign ? (long long)(ctx->gen_syms + 1) * ctx->nbmax : 0);
    CU(ctx->d_scratch.ensure(warps * (size_t)ctx->scr_stride));
    return ISOCON_OK;
}

// (Re)allocate the share block for n list entries and f final edges.  A block that was exported stays alive (parked)
// until the peers have dropped their mappings (isocon_nn_release_retired); the handles must then be exchanged again.
int ensure_share(isocon_nn_ctx* ctx, long long n, long long f) {
    if (n + 1 <= ctx->share_n && f <= ctx->share_f && ctx->d_share.p) return ISOCON_OK;
    CU(cudaStreamSynchronize(ctx->stream));
    const long long n_cap = std::max(ctx->share_n, n + n / 8 + 64), f_cap = std::max(ctx->share_f, f);
    if (ctx->d_share.p) {
        if (ctx->best_exported) ctx->retired_best.push_back(ctx->d_share.p); else cudaFree(ctx->d_share.p);
        ctx->d_share.p = nullptr; ctx->d_share.cap = 0;
    }
    close_peers(ctx);
    ++ctx->best_generation;
    CU(ctx->d_share.ensure(share_bytes(n_cap, f_cap)));
    ctx->share_n = n_cap; ctx->share_f = f_cap;
    ctx->sv = share_view(ctx->d_share.p, n_cap, f_cap);
    CU(cudaMemsetAsync(ctx->sv.ctrl, 0, CT_WORDS * sizeof(unsigned long long), ctx->stream));
    ctx->bar_seq = 0;
    return ISOCON_OK;
}

// Device-side barrier of the ranks of a box (all ranks enqueue the same sequence of barriers).
int enqueue_barrier(isocon_nn_ctx* ctx) {
    if (ctx->n_peers <= 0) return ISOCON_OK;
    BarrierArgs B{};
    B.own = ctx->sv.ctrl + CT_BAR; B.err = ctx->sv.ctrl + CT_ERR; B.n_peers = ctx->n_peers;
    for (int p = 0; p < ctx->n_peers; ++p) B.peer[p] = ctx->peer_sv[p].ctrl + CT_BAR;
    B.target = ++ctx->bar_seq * (unsigned long long)(ctx->n_peers + 1);
    peer_barrier_kernel<<<1, 1, 0, ctx->stream>>>(B);
    CU(cudaGetLastError());
    ++ctx->launches;
    return ISOCON_OK;
}

// Fused multi-rank flow: all ranks have finished the phase; keep the (now identical) best[] as the basis of the next
// host decisions, and let nobody lower it before everyone has its copy.
int agree_on_best(isocon_nn_ctx* ctx) {
    int rc = enqueue_barrier(ctx);
    if (rc) return rc;
    CU(ctx->d_snap.ensure((size_t)ctx->n + 1));
    CU(cudaMemcpyAsync(ctx->d_snap.p, ctx->d_best.p, (size_t)ctx->n * sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->snap_valid = true;
    ctx->best_host_launches = ~0ull;
    return enqueue_barrier(ctx);
}

// Asynchronous H2D of a small host table: through the pinned bounce arena, so neither the copy nor the caller waits.
int h2d(isocon_nn_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (!bytes) return ISOCON_OK;
    void* b = ctx->bounce.take(bytes);
    if (!b) {
        CU(cudaStreamSynchronize(ctx->stream));      // every earlier piece has been copied: start over
        ctx->bounce.used = 0;
        CU(ctx->bounce.ensure(std::max(bytes + 4096, ctx->bounce.cap * 2)));
        b = ctx->bounce.take(bytes);
    }
    memcpy(b, src, bytes);
    CU(cudaMemcpyAsync(dst, b, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return ISOCON_OK;
}

// Upload the target layout (h_tpos, bins) and interleave the packed targets group by group.
int apply_layout(isocon_nn_ctx* ctx) {
    ctx->nT = (int)ctx->h_tpos.size();
    ctx->nG = ctx->nT / 32;
    std::vector<long long> goff((size_t)ctx->nG + 1, 0);
    for (int g = 0; g < ctx->nG; ++g) {
        int longest = 0;   // (similarity-ordered bins are not in length order: look at every lane)
        for (int l = 0; l < 32; ++l) {
            const int t = ctx->h_tpos[(size_t)32 * g + l];
            if (t >= 0) longest = std::max(longest, ctx->h_len[t]);
        }
        goff[g + 1] = goff[g] + 32ll * (((longest + 15) >> 4) + 4);
    }
    CU(ctx->d_tpos.ensure((size_t)ctx->nT + 1));
    CU(ctx->d_goff.ensure((size_t)ctx->nG + 1));
    CU(ctx->d_il.ensure((size_t)goff[ctx->nG] + 64));
    int rc = h2d(ctx, ctx->d_tpos.p, ctx->h_tpos.data(), (size_t)ctx->nT * sizeof(int));
    if (!rc) rc = h2d(ctx, ctx->d_goff.p, goff.data(), goff.size() * sizeof(long long));
    if (rc) return rc;
    if (ctx->nG) {
        interleave_kernel<<<ctx->nG, 256, 0, ctx->stream>>>(ctx->d_rowpk.p, ctx->d_rowoff.p, ctx->d_len.p, ctx->d_tpos.p,
                                                           ctx->nT, ctx->d_goff.p, ctx->nG, ctx->d_il.p);
        CU(cudaGetLastError());
        ++ctx->launches;
    }
    return ISOCON_OK;
}

// Layout from a class per list entry (targets only): bins in ascending class order.
void set_layout(isocon_nn_ctx* ctx, const std::vector<int>& cls, int n_classes) {
    std::vector<std::vector<int>> bins((size_t)n_classes);
    for (long long i = 0; i < ctx->n; ++i)
        if (ctx->h_ist_main[i]) bins[(size_t)cls[i]].push_back((int)i);
    ctx->h_tpos.clear(); ctx->bin_first.clear(); ctx->bin_count.clear();
    for (const auto& b : bins) {
        if (b.empty()) continue;
        ctx->bin_first.push_back((int)ctx->h_tpos.size());
        ctx->bin_count.push_back((int)b.size());
        ctx->h_tpos.insert(ctx->h_tpos.end(), b.begin(), b.end());
        while (ctx->h_tpos.size() % 32) ctx->h_tpos.push_back(-1);
    }
    ctx->binned = ctx->bin_first.size() > 1;
}

// Row tiles of one pass.  kw[i] = half-width of query i's length window.  Per query one row whose
// segments are the group ranges of its window in every bin.  For the row kernel the tile size
// follows the amount of work: every block of every rank should get about a dozen tiles (short
// tail at the end of the launch) but a tile should keep each of the block's warps busy for
// several groups (the block synchronises between tiles).
// fine_rows: the first rows get the smallest tiles (one group per warp), whatever work is left.
void build_items(const isocon_nn_ctx* c, const std::vector<int>& queries, const std::vector<int>& kw,
                 bool upper_only, ItemTable& T, size_t fine_rows = 0) {
    const size_t nq = queries.size();
    const std::vector<int>& tp = c->h_tpos;
    const size_t nb = c->bin_first.size();
    long long total_groups = 0;
    // rows normally come in list order with a fixed window width: the window bounds inside every bin then
    // only move forward (one sweep over the bin); any other row falls back to binary searches
    std::vector<int> plo(nb, 0), phi(nb, 0), pup(nb, 0);
    int prev_lo = INT_MIN, prev_hi = INT_MIN, prev_q = INT_MIN;
    T.qlist.reserve(nq); T.segoff.reserve(nq + 1); T.gtotal.reserve(nq);
    T.seg_g0.reserve(nq * nb); T.seg_n.reserve(nq * nb);
    for (size_t i = 0; c->bins_unsorted && i < nq; ++i) {
        // similarity order: a bin is sorted by rank, not by length -- a row takes every bin whole (the lanes prune by
        // length themselves), or, when each unordered pair is aligned once, the part of the bin behind its own rank
        const int q = queries[i];
        T.add_row(q);
        const bool by_rank = upper_only && c->clustered;     // (one-sided passes never split pairs between rows)
        const bool ascending = by_rank && c->h_rank[q] >= prev_q;      // rows normally come in rank order: one sweep per bin
        if (ascending) prev_q = c->h_rank[q];
        for (size_t b = 0; b < nb; ++b) {
            const int* first = tp.data() + c->bin_first[b];
            const int* last = first + c->bin_count[b];
            const int* lo = first;
            if (by_rank && ascending) {
                while (pup[b] < c->bin_count[b] && c->h_rank[first[pup[b]]] <= c->h_rank[q]) ++pup[b];
                lo = first + pup[b];
            } else if (by_rank) {
                lo = std::upper_bound(first, last, c->h_rank[q], [&](int v, int t) { return v < c->h_rank[t]; });
            }
            if (last > lo) {
                const int g0 = (int)((lo - tp.data()) / 32), g1 = (int)((last - 1 - tp.data()) / 32);
                T.add_segment(g0, g1 - g0 + 1);
            }
        }
        total_groups += T.gtotal.back();
    }
    for (size_t i = 0; !c->bins_unsorted && i < nq; ++i) {
        const int q = queries[i];
        const long long m = c->h_len[q];
        const int len_lo = (int)std::max<long long>(m - kw[i], 0), len_hi = (int)std::min<long long>(m + kw[i], INT_MAX);
        const bool sweep = len_lo >= prev_lo && len_hi >= prev_hi && q >= prev_q && !(c->prm.mode == 1 && c->prm.depth < c->n);
        if (sweep) { prev_lo = len_lo; prev_hi = len_hi; prev_q = q; }
        T.add_row(q);
        for (size_t b = 0; b < nb; ++b) {
            const int* first = tp.data() + c->bin_first[b];
            const int cnt = c->bin_count[b];
            const int* last = first + cnt;
            const int* lo;
            const int* hi;
            if (sweep) {
                while (plo[b] < cnt && c->h_len[first[plo[b]]] < len_lo) ++plo[b];
                while (phi[b] < cnt && c->h_len[first[phi[b]]] <= len_hi) ++phi[b];
                lo = first + plo[b]; hi = first + phi[b];
                if (upper_only) {
                    while (pup[b] < cnt && first[pup[b]] <= q) ++pup[b];
                    lo = std::max(lo, first + pup[b]);
                }
            } else {
                lo = std::lower_bound(first, last, len_lo, [&](int t, int v) { return c->h_len[t] < v; });
                hi = std::upper_bound(first, last, len_hi, [&](int v, int t) { return v < c->h_len[t]; });
                if (c->prm.mode == 1) {
                    if (c->prm.depth < c->n) {   // offsets 1..depth only (:190); list indices ascend inside a bin
                        lo = std::max(lo, std::lower_bound(first, last, (int)std::max<long long>(q - c->prm.depth, 0)));
                        hi = std::min(hi, std::upper_bound(first, last, (int)std::min<long long>(q + c->prm.depth, INT_MAX)));
                    }
                    if (upper_only) lo = std::max(lo, std::upper_bound(first, last, q));
                }
            }
            if (hi > lo) {
                const int g0 = (int)((lo - tp.data()) / 32), g1 = (int)((hi - 1 - tp.data()) / 32);
                T.add_segment(g0, g1 - g0 + 1);
            }
        }
        total_groups += T.gtotal.back();
    }
    T.segoff.push_back((int)T.seg_g0.size());
    // Tile sizes of the row kernel shrink with the work that is left (guided self-scheduling): a tile is a
    // quarter of a block's share of the remaining groups -- large tiles first (few block-wide
    // synchronisations), one group per warp at the very end (short tail) -- and the tiles of one row are equal,
    // so tiles dealt round-robin to the ranks have smoothly varying cost.
    const long long blocks = (long long)std::max(1, c->row_grid) * std::max(1, c->prm.world);
    long long remaining = total_groups;
    T.gsize.assign(nq, T.gpi);
    T.item_off.assign(nq + 1, 0);
    for (size_t i = 0; i < nq; ++i) {
        int gpi = T.gpi;
        if (T.row_kernel)
            gpi = i < fine_rows ? ROW_WARPS
                                : (int)std::min<long long>(ROW_GROUPS_PER_ITEM, std::max<long long>(ROW_WARPS, (remaining / (blocks * 4) + 7) / 8 * 8));
        remaining -= T.gtotal[i];
        if (T.gtotal[i] <= gpi) {            // the whole row is one tile (or none): no divisions
            if (T.gtotal[i] > 0) T.gsize[i] = T.gtotal[i];
            T.item_off[i + 1] = T.item_off[i] + (T.gtotal[i] > 0 ? 1 : 0);
            continue;
        }
        const int tiles = (T.gtotal[i] + gpi - 1) / gpi;
        if (tiles > 0) T.gsize[i] = (T.gtotal[i] + tiles - 1) / tiles;
        T.item_off[i + 1] = T.item_off[i] + (tiles > 0 ? (T.gtotal[i] + T.gsize[i] - 1) / T.gsize[i] : 0);
    }
}

int upload_items(isocon_nn_ctx* ctx, const ItemTable& T) {
    const size_t nq = T.qlist.size(), ns = T.seg_g0.size();
    CU(ctx->d_qlist.ensure(nq + 1)); CU(ctx->d_segoff.ensure(nq + 2)); CU(ctx->d_gtotal.ensure(nq + 1));
    CU(ctx->d_gsize.ensure(nq + 1)); CU(ctx->d_item_off.ensure(nq + 2));
    CU(ctx->d_seg_g0.ensure(ns + 1)); CU(ctx->d_seg_n.ensure(ns + 1));
    int rc = h2d(ctx, ctx->d_qlist.p, T.qlist.data(), nq * sizeof(int));
    if (!rc) rc = h2d(ctx, ctx->d_gtotal.p, T.gtotal.data(), nq * sizeof(int));
    if (!rc) rc = h2d(ctx, ctx->d_gsize.p, T.gsize.data(), nq * sizeof(int));
    if (!rc) rc = h2d(ctx, ctx->d_seg_g0.p, T.seg_g0.data(), ns * sizeof(int));
    if (!rc) rc = h2d(ctx, ctx->d_seg_n.p, T.seg_n.data(), ns * sizeof(int));
    if (!rc) rc = h2d(ctx, ctx->d_segoff.p, T.segoff.data(), (nq + 1) * sizeof(int));
    if (!rc) rc = h2d(ctx, ctx->d_item_off.p, T.item_off.data(), (nq + 1) * sizeof(long long));
    return rc;
}

// Similarity order of the targets inside their threshold-class bins.
//
// The 32 lanes of a warp walk 32 targets in lock step until the LAST of them is done, under a window that is the
// union of their alive intervals: a group of targets at very different distances from the query makes most lanes
// wait (c3: half of the issued work).  Reads of one gene copy behave alike towards any query, so the targets of a
// bin are ordered by cluster: the PILOT pass recorded the two nearest pilot rows of every entry; pilot rows that share
// a neighbourhood are merged (union-find over "p's nearest pilot rows" and "the two pilot rows nearest to x"), and
// an entry belongs to the cluster of its nearest pilot row.  Results never depend on this order: every pair is
// still aligned exactly once with a valid threshold; only who shares a warp changes.
//   pnear : [2n] (distance << 32 | pilot row), ~0 = none
//   cls   : threshold class per entry
// Fills h_tpos / bins (class-major, cluster order inside) and h_rank (pilot rows first, then layout order).
void cluster_order(isocon_nn_ctx* ctx, const unsigned long long* pnear, const std::vector<int>& cls, int n_classes,
                   const int* best) {
    const long long n = ctx->n;
    std::vector<int> parent((size_t)n);
    for (long long i = 0; i < n; ++i) parent[(size_t)i] = (int)i;
    auto find = [&](int x) { while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; } return x; };
    auto unite = [&](int a, int b) { a = find(a); b = find(b); if (a != b) parent[std::max(a, b)] = std::min(a, b); };
    const unsigned long long none = ~0ull;
    for (long long x = 0; x < n; ++x) {
        const unsigned long long v0 = pnear[x], v1 = pnear[n + x];
        if (v0 == none) continue;
        const int a0 = (int)(v0 & 0xffffffffu);
        const long long d0 = (long long)(v0 >> 32);
        const bool is_pilot = ctx->h_rank[(size_t)x] >= 0;           // (h_rank holds the pilot marks on entry)
        if (is_pilot) unite((int)x, a0);
        if (v1 != none && (long long)(v1 >> 32) * 4 <= d0 * 5) {     // the second one is about as near: same cluster
            const int a1 = (int)(v1 & 0xffffffffu);
            unite(a0, a1);
        }
    }
    // label: cluster of the nearest pilot row (a pilot row: its own); none: a cluster of its own behind the others
    std::vector<int> label((size_t)n), near((size_t)n);
    for (long long x = 0; x < n; ++x) {
        const bool is_pilot = ctx->h_rank[(size_t)x] >= 0;
        const unsigned long long v0 = pnear[x];
        near[(size_t)x] = is_pilot ? (int)x : (v0 == none ? INT_MAX : (int)(v0 & 0xffffffffu));
        label[(size_t)x] = near[(size_t)x] == INT_MAX ? INT_MAX : find(near[(size_t)x]);
    }
    // order = (class, pilot rows first, cluster, nearest pilot row, list index): three stable counting sorts over the
    // targets in list order (O(n); this runs between two kernel launches on every rank)
    std::vector<int> order, tmp;
    order.reserve((size_t)n);
    for (long long i = 0; i < n; ++i)
        if (ctx->h_ist_main[i]) order.push_back((int)i);
    tmp.resize(order.size());
    std::vector<int> count;
    auto counting_sort = [&](int n_keys, auto key_of) {
        count.assign((size_t)n_keys + 1, 0);
        for (int x : order) ++count[(size_t)key_of(x) + 1];
        for (int k = 0; k < n_keys; ++k) count[(size_t)k + 1] += count[(size_t)k];
        for (int x : order) tmp[(size_t)count[(size_t)key_of(x)]++] = x;
        order.swap(tmp);
    };
    const int n_keys = (int)n + 1;                                         // INT_MAX (none) -> n
    // minor keys inside a cluster: the nearest pilot row (reads around one pilot row are each other's neighbours),
    // optionally the read's own bound (ISOCON_NN_ORDER_BEST: 1 = instead of, 2 = below the pilot row)
    const int bcap = ctx->opt_kcap_main + 1;
    if (ctx->opt_order_best) counting_sort(bcap + 1, [&](int x) { return std::min(best[(size_t)x], bcap); });
    if (ctx->opt_order_best != 1)
        counting_sort(n_keys, [&](int x) { return near[(size_t)x] == INT_MAX ? (int)n : near[(size_t)x]; });
    counting_sort(n_keys, [&](int x) {
        if (ctx->h_rank[(size_t)x] >= 0) return 0;                         // pilot rows: in front, in list order
        return label[(size_t)x] == INT_MAX ? (int)n : label[(size_t)x];
    });
    // (a pilot row's label is <= its own index, and label 0 can only be pilot row 0's cluster: pilot rows sorted by key
    // 0 stay in list order because near[] = own index ascends; non-pilot members of cluster 0 share the key -- split
    // them off with one more stable pass on "is pilot")
    counting_sort(2, [&](int x) { return ctx->h_rank[(size_t)x] >= 0 ? 0 : 1; });
    counting_sort(n_classes, [&](int x) { return cls[(size_t)x]; });
    ctx->h_tpos.clear(); ctx->bin_first.clear(); ctx->bin_count.clear();
    ctx->h_tpos.reserve(order.size() + 32 * (size_t)n_classes);
    int next_rank = 0;
    std::vector<int> rank((size_t)n, -1);
    // pilot rows keep the lowest ranks in list order: all their pairs were aligned by the PILOT pass
    for (long long i = 0; i < n; ++i)
        if (ctx->h_rank[(size_t)i] >= 0) rank[(size_t)i] = next_rank++;
    long long clusters = 0;
    std::vector<char> seen((size_t)n, 0);
    for (size_t i = 0; i < order.size();) {
        size_t j = i;
        while (j < order.size() && cls[(size_t)order[j]] == cls[(size_t)order[i]]) ++j;
        ctx->bin_first.push_back((int)ctx->h_tpos.size());
        ctx->bin_count.push_back((int)(j - i));
        for (size_t k = i; k < j; ++k) {
            const int x = order[k];
            if (rank[(size_t)x] < 0) rank[(size_t)x] = next_rank++;
            if (label[(size_t)x] != INT_MAX && !seen[(size_t)label[(size_t)x]]) { seen[(size_t)label[(size_t)x]] = 1; ++clusters; }
            ctx->h_tpos.push_back(x);
        }
        while (ctx->h_tpos.size() % 32) ctx->h_tpos.push_back(-1);
        i = j;
    }
    for (long long i = 0; i < n; ++i)
        if (rank[(size_t)i] < 0) rank[(size_t)i] = next_rank++;           // entries that are no targets
    ctx->h_rank.swap(rank);
    ctx->binned = true;
    ctx->stats.clusters = (uint64_t)clusters;
}

// Make `tpos` (slots of whole 32-target groups, -1 = pad, ONE bin, not in length order) the target layout in force.
int use_layout(isocon_nn_ctx* ctx, const std::vector<int>& tpos) {
    ctx->h_tpos = tpos;
    int count = (int)tpos.size();
    while (count > 0 && tpos[(size_t)count - 1] < 0) --count;
    ctx->bin_first.assign(1, 0); ctx->bin_count.assign(1, count);
    ctx->binned = false;
    return apply_layout(ctx);
}

// Similarity order for ONE-SIDED passes (2-set graph: reads against candidates).  There is no PILOT pass to learn
// clusters from, but the targets of such a graph are few and clean (candidate transcripts), so four min-hash values
// over their 16-mers (minhash_kernel) tell relatives apart from strangers: targets that share any of the four values
// are merged (union-find), the layout lists cluster after cluster.  Why it pays: a read is related to a handful of
// candidates -- pairs that run the whole length -- and a stranger to the rest (early exit after ~300 columns); in
// length order every relative sits in a different group of 32 and makes 31 lanes wait (c5: 10 relatives per read =
// 10 slow groups; clustered: one).  A heuristic only: any order gives the same graph.
int sketch_order(isocon_nn_ctx* ctx) {
    const long long n = ctx->n;
    DebugLap lap(ctx->opt_debug >= 2, "sketch");
    CU(ctx->d_sig.ensure(4 * (size_t)n + 4));
    CU(ctx->sig_host.ensure(4 * (size_t)n * sizeof(unsigned long long) + 64));
    CU(ctx->d_ist.ensure((size_t)n + 1));
    // d_ist holds the caller's targets; foreign entries are no targets of the 2-bit kernels: mask from h_ist_main
    DBuf<uint8_t>& mask = ctx->d_flag;
    CU(mask.ensure((size_t)n + 1));
    std::vector<uint8_t> pick((size_t)n);
    for (long long i = 0; i < n; ++i) pick[(size_t)i] = (ctx->h_ist_main[(size_t)i] || (ctx->h_isq[(size_t)i] && !ctx->h_foreign[(size_t)i])) ? 1 : 0;
    int rc = h2d(ctx, mask.p, pick.data(), (size_t)n);
    if (rc) return rc;
    minhash_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(ctx->d_rowpk.p, ctx->d_rowoff.p, ctx->d_len.p, mask.p, (int)n, ctx->d_sig.p);
    CU(cudaGetLastError());
    ++ctx->launches;
    CU(cudaMemcpyAsync(ctx->sig_host.p, ctx->d_sig.p, 4 * (size_t)n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    const unsigned long long* sig = (const unsigned long long*)ctx->sig_host.p;
    lap.lap("minhash+d2h");
    std::vector<int> targets;
    for (long long i = 0; i < n; ++i) if (ctx->h_ist_main[(size_t)i]) targets.push_back((int)i);
    std::vector<int> parent((size_t)n);
    for (long long i = 0; i < n; ++i) parent[(size_t)i] = (int)i;
    auto find = [&](int x) { while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; } return x; };
    const size_t nt = targets.size();
    std::vector<std::pair<unsigned long long, int>> keyed(nt);
    std::vector<unsigned long long> skeys(4 * nt);     // per hash: the targets' values in ascending order ...
    std::vector<int> svals(4 * nt);                    // ... and their owners (for the hint lookup on the device)
    for (int h = 0; h < 4; ++h) {
        for (size_t k = 0; k < nt; ++k) keyed[k] = std::make_pair(sig[4ll * targets[k] + h], targets[k]);
        std::sort(keyed.begin(), keyed.end());
        for (size_t k = 0; k < nt; ++k) { skeys[h * nt + k] = keyed[k].first; svals[h * nt + k] = keyed[k].second; }
        for (size_t k = 1; k < nt; ++k)
            if (keyed[k].first == keyed[k - 1].first && keyed[k].first != ~0ull) {
                const int a = find(keyed[k].second), b = find(keyed[k - 1].second);
                if (a != b) parent[(size_t)std::max(a, b)] = std::min(a, b);
            }
    }
    // hint lookup on the device while the host goes on: every query's first target with a common min-hash value
    DBuf<unsigned long long>& d_keys = ctx->d_sigkeys;
    DBuf<int>& d_vals = ctx->d_sigvals;
    CU(d_keys.ensure(4 * nt + 4)); CU(d_vals.ensure(4 * nt + 4)); CU(ctx->d_hint.ensure((size_t)n + 1));
    rc = h2d(ctx, d_keys.p, skeys.data(), 4 * nt * sizeof(unsigned long long));
    if (!rc) rc = h2d(ctx, d_vals.p, svals.data(), 4 * nt * sizeof(int));
    if (rc) return rc;
    hint_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_sig.p, mask.p, (int)n, d_keys.p, d_vals.p, (int)nt, ctx->d_hint.p);
    CU(cudaGetLastError());
    ++ctx->launches;
    std::vector<int> hint_t((size_t)n);
    CU(cudaMemcpyAsync(hint_t.data(), ctx->d_hint.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));

    lap.lap("sort+union+hint launch");
    std::vector<int> root((size_t)n, -1);
    long long clusters = 0;
    for (int t : targets) { root[(size_t)t] = find(t); if (root[(size_t)t] == t) ++clusters; }
    if (clusters * 2 > (long long)nt) { CU(cudaStreamSynchronize(ctx->stream)); return ISOCON_OK; }   // mostly singletons: keep length order
    // Representatives and radii.  Inside a cluster of <= 64 members every pair is aligned (explicit-pairs kernel,
    // unbounded, exact): the representative is the member whose farthest fellow is nearest (1-centre), the radius that
    // distance.  Larger clusters: the first member, radius from the distances to it.  A member farther than RMAX from
    // the representative (a chance merge, a distant relative) becomes a cluster of its own, so the slack added to
    // the thresholds of level 1 stays within a few window words.
    constexpr int RMAX = 128;
    ctx->h_slack.assign((size_t)n, 0);
    {
        std::vector<std::vector<int>> members((size_t)n);
        std::vector<int> first_roots;                       // (root[] is rewritten below: keep the clusters found above)
        for (int t : targets) { members[(size_t)root[(size_t)t]].push_back(t); if (root[(size_t)t] == t) first_roots.push_back(t); }
        std::vector<int> pa, pb;
        for (int t : first_roots) {
            const std::vector<int>& m = members[(size_t)t];
            if (m.size() <= 64) { for (size_t i = 0; i < m.size(); ++i) for (size_t j = i + 1; j < m.size(); ++j) { pa.push_back(m[i]); pb.push_back(m[j]); } }
            else for (size_t j = 1; j < m.size(); ++j) { pa.push_back(m[0]); pb.push_back(m[j]); }
        }
        std::vector<int> dist(pa.size()), bound(pa.size(), RMAX);       // (-1 = farther than RMAX: all we need to know)
        if (!pa.empty()) {
            rc = isocon_nn_ed_pairs(ctx, pa.data(), pb.data(), bound.data(), (int64_t)pa.size(), dist.data());
            if (rc) return rc;
            for (int& d : dist) if (d < 0) d = RMAX + 1;
        }
        size_t at = 0;
        for (int t : first_roots) {
            const std::vector<int>& m = members[(size_t)t];
            const size_t sz = m.size();
            if (sz <= 1) continue;
            int rep = m[0];
            std::vector<int> to_rep(sz, 0);
            if (sz <= 64) {
                std::vector<int> far(sz, 0), d(sz * sz, 0);
                for (size_t i = 0; i < sz; ++i) for (size_t j = i + 1; j < sz; ++j) { d[i * sz + j] = d[j * sz + i] = dist[at++]; }
                size_t best_i = 0;
                for (size_t i = 0; i < sz; ++i) { for (size_t j = 0; j < sz; ++j) far[i] = std::max(far[i], d[i * sz + j]); if (far[i] < far[best_i]) best_i = i; }
                rep = m[best_i];
                for (size_t j = 0; j < sz; ++j) to_rep[j] = d[best_i * sz + j];
            } else {
                for (size_t j = 1; j < sz; ++j) to_rep[j] = dist[at++];
            }
            for (size_t j = 0; j < sz; ++j) {
                if (to_rep[j] > RMAX) { root[(size_t)m[j]] = m[j]; ++clusters; }
                else { root[(size_t)m[j]] = rep; ctx->h_slack[(size_t)rep] = std::max(ctx->h_slack[(size_t)rep], to_rep[j]); }
            }
        }
    }
    lap.lap("representatives (ed_pairs)");
    // layout A: cluster after cluster (by representative = smallest list index), list order inside a cluster
    std::vector<std::pair<int, int>> order(targets.size());
    for (size_t k = 0; k < targets.size(); ++k) order[k] = std::make_pair(root[(size_t)targets[k]], targets[k]);
    std::sort(order.begin(), order.end());
    ctx->h_tposA.clear(); ctx->h_tposB.clear();
    for (const auto& e : order) ctx->h_tposA.push_back(e.second);
    while (ctx->h_tposA.size() % 32) ctx->h_tposA.push_back(-1);
    ctx->cl_g0.assign((size_t)n, -1); ctx->cl_ng.assign((size_t)n, 0);
    {
        std::vector<int> first_slot((size_t)n, -1), last_slot((size_t)n, -1);
        for (size_t k = 0; k < order.size(); ++k) {
            if (first_slot[(size_t)order[k].first] < 0) first_slot[(size_t)order[k].first] = (int)k;
            last_slot[(size_t)order[k].first] = (int)k;
        }
        for (int t : targets)
            if (root[(size_t)t] == t) {
                ctx->h_tposB.push_back(t);                      // layout B: the representatives, in list order
                ctx->cl_g0[(size_t)t] = first_slot[(size_t)t] / 32;
                ctx->cl_ng[(size_t)t] = last_slot[(size_t)t] / 32 - first_slot[(size_t)t] / 32 + 1;
            }
    }
    while (ctx->h_tposB.size() % 32) ctx->h_tposB.push_back(-1);
    ctx->h_root = root;
    ctx->bins_unsorted = true;
    ctx->stats.clusters = (uint64_t)clusters;
    ctx->two_level = ctx->opt_two_level && clusters * 2 <= (long long)targets.size();
    if (ctx->two_level) {
        CU(ctx->d_slack.ensure((size_t)n + 1));
        rc = h2d(ctx, ctx->d_slack.p, ctx->h_slack.data(), (size_t)n * sizeof(int));
        if (rc) return rc;
    }
    // Hints: a query that shares a min-hash value with a target is most likely related to its cluster (a read with
    // 3 % errors keeps a candidate's value with probability ~0.4 per hash).  The SEED pass aligns every hinted query
    // against that cluster first -- with edges, so level 2 need not meet the cluster again -- and the MAIN pass starts
    // from bounds near the final ones instead of the cap.
    CU(cudaStreamSynchronize(ctx->stream));              // hint_t has arrived
    ctx->h_hint_g0.assign((size_t)n, -1); ctx->h_hint_n.assign((size_t)n, 0); ctx->h_hint_rep.assign((size_t)n, -1);
    for (int q : ctx->h_qlist) {
        const int t = hint_t[(size_t)q];
        if (t < 0 || root[(size_t)t] < 0) continue;
        const int rep = root[(size_t)t];
        ctx->h_hint_g0[(size_t)q] = ctx->cl_g0[(size_t)rep];
        ctx->h_hint_n[(size_t)q] = std::min(ctx->cl_ng[(size_t)rep], GROUPS_PER_ITEM);
        if (ctx->cl_ng[(size_t)rep] <= GROUPS_PER_ITEM) ctx->h_hint_rep[(size_t)q] = rep;   // the SEED pass covers the whole cluster
    }
    if (ctx->two_level) {
        CU(ctx->d_hint.ensure((size_t)n + 1));
        rc = h2d(ctx, ctx->d_hint.p, ctx->h_hint_rep.data(), (size_t)n * sizeof(int));   // (the lookup result is on the host by now)
        if (rc) return rc;
    }
    lap.lap("layouts+hints");
    rc = use_layout(ctx, ctx->h_tposA);
    lap.lap("use_layout");
    return rc;
}

// best[] on the host (pinned): the one synchronisation the host-side re-binning / row selection needs.
int fetch_best(isocon_nn_ctx* ctx, const int** out) {
    if (ctx->best_host_launches == ctx->launches && ctx->best_host.p) {   // nothing ran since the last fetch
        *out = (const int*)ctx->best_host.p;
        return ISOCON_OK;
    }
    CU(ctx->best_host.ensure((size_t)ctx->n * sizeof(int) + 64));
    // Several ranks: every host decision (threshold classes, ladder rows, WIDE rows) must come out the same on all of
    // them -- they build the same tile table and drain one queue.  The live best[] is no basis for that: a peer that
    // is already running its next pass lowers it over NVLink.  The driver therefore takes a snapshot on every rank
    // right after its MIN-reduce (isocon_nn_best_agree) and lets no rank go on before all have theirs.
    const int* src = (ctx->prm.world > 1 && ctx->snap_valid) ? ctx->d_snap.p : ctx->d_best.p;
    CU(cudaMemcpyAsync(ctx->best_host.p, src, (size_t)ctx->n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->best_host_launches = ctx->launches;
    *out = (const int*)ctx->best_host.p;
    return ISOCON_OK;
}

GraphArgs base_args(isocon_nn_ctx* c) {
    GraphArgs A{};
    A.mode = c->prm.mode; A.symmetric = 0; A.pass = PASS_MAIN; A.kcap = INT_MAX; A.kprev = -1; A.append = 1;
    A.depth = c->prm.depth;
    A.n = (int)c->n; A.nT = c->nT; A.nG = c->nG;
    A.len = c->d_len.p; A.rowoff = c->d_rowoff.p; A.rowpk = c->d_rowpk.p;
    A.tpos = c->d_tpos.p; A.goff = c->d_goff.p; A.il = c->d_il.p;
    A.isq = c->d_isq.p; A.ist = c->d_ist.p;
    A.best = c->d_best.p;
    // peers only take part in a graph that was begun as one of several ranks: a context whose mappings outlive
    // the process group must not write into the other GPUs' best[] while it builds a graph alone
    A.n_peers = c->prm.world > 1 ? c->n_peers : 0;
    for (int p = 0; p < 7; ++p) A.peer_best[p] = p < A.n_peers ? c->peer_best[p] : nullptr;
    A.qlist = c->d_qlist.p; A.item_off = c->d_item_off.p; A.segoff = c->d_segoff.p; A.gtotal = c->d_gtotal.p;
    A.gsize = c->d_gsize.p; A.seg_g0 = c->d_seg_g0.p; A.seg_n = c->d_seg_n.p;
    A.counter = c->d_small.p + SM_COUNTER;
    A.eq = c->d_eq.p; A.et = c->d_et.p; A.ed = c->d_ed.p; A.ecount = c->d_small.p + SM_ECOUNT; A.ecap = c->ecap;
    A.scratch = c->d_scratch.p; A.nbmax = c->nbmax; A.peq_words = c->peq_words;
    A.narrow = c->opt_narrow;
    A.stats = c->d_small.p + SM_STATS;
    A.foff = c->n_foreign ? c->d_foff.p : nullptr; A.fascii = c->d_fascii.p;
    A.abc = (uint32_t)c->alphabet[0] | ((uint32_t)c->alphabet[1] << 8) | ((uint32_t)c->alphabet[2] << 16) | ((uint32_t)c->alphabet[3] << 24);
    A.gen_syms = c->gen_syms; A.scr_stride = c->scr_stride;
    A.rank = c->clustered ? c->d_rank.p : nullptr; A.pnear = nullptr; A.pilot_last = -1;
    A.slack = nullptr; A.surv_q = nullptr; A.surv_t = nullptr; A.surv_count = nullptr; A.surv_cap = 0;
    A.qgram = nullptr;
    return A;
}

// Row tiles are dealt out to the ranks round-robin: tile i belongs to rank i mod world.  Tiles
// of one row have equal cost and neighbouring rows differ by a few targets, so every rank gets
// the same share of every part of the pair matrix (and of every read's candidates).
void shard(long long total, int rank, int world, GraphArgs& A) {
    A.item_end = total;
    if (world <= 1) { A.item_begin = 0; A.item_stride = 1; return; }
    A.item_begin = rank; A.item_stride = world;
}

// item_lo / item_hi: the tiles of the table this launch covers (default: all); table_resident: the table was uploaded
// by an earlier launch of the same pass.
int launch_tile(isocon_nn_ctx* ctx, GraphArgs A, const ItemTable& T, bool sharded, int queue = -1,
                long long item_lo = 0, long long item_hi = -1, bool table_resident = false) {
    if (T.total() == 0) return ISOCON_OK;
    if (item_hi < 0) item_hi = T.total();
    if (item_hi <= item_lo) return ISOCON_OK;
    int rc = ISOCON_OK;
    if (!table_resident) {
        ctx->last_run_rows += (long long)T.qlist.size();
        rc = upload_items(ctx, T);
        if (rc) return rc;
    }
    A.nQ = (int)T.qlist.size();
    A.qlist = ctx->d_qlist.p; A.item_off = ctx->d_item_off.p;   // (re)allocated by upload_items
    A.segoff = ctx->d_segoff.p; A.gtotal = ctx->d_gtotal.p; A.gsize = ctx->d_gsize.p;
    A.seg_g0 = ctx->d_seg_g0.p; A.seg_n = ctx->d_seg_n.p;
    // Which tiles this rank computes: alone, all; with the peers' memory mapped, whatever it pulls from the
    // box-wide queue in rank 0's memory (all GPUs of the box drain one queue: no rank finishes early);
    // otherwise every world-th tile.
    const bool box_queue = sharded && ctx->prm.world > 1 && ctx->root_small && queue >= 0;
    if (box_queue) {
        shard(item_hi, 0, 1, A);
        A.item_begin = item_lo;
        A.counter = ctx->root_small + SM_QUEUE + queue;
    } else {
        if (sharded) shard(item_hi, ctx->prm.rank, ctx->prm.world, A);
        else shard(item_hi, 0, 1, A);
        A.item_begin += item_lo;
        if (A.item_end <= A.item_begin) return ISOCON_OK;
        CU(cudaMemsetAsync(ctx->d_small.p + SM_COUNTER, 0, sizeof(unsigned long long), ctx->stream));
    }
    const bool timed = ctx->kev_used < isocon_nn_ctx::KEV;
    if (timed) CU(cudaEventRecord(ctx->kev[2 * ctx->kev_used], ctx->stream));
    if (T.row_kernel && T.swapped)
        nn_row_swapped_kernel<<<ctx->row_grid, ROW_WARPS * 32, ctx->row_smem, ctx->stream>>>(A, ctx->row_padbits, ctx->row_xmax);
    else if (T.row_kernel)
        nn_row_kernel<<<ctx->row_grid, ROW_WARPS * 32, ctx->row_smem, ctx->stream>>>(A, ctx->row_padbits, ctx->row_xmax);
    else
        nn_tile_kernel<<<ctx->grid, WARPS_PER_BLOCK * 32, ctx->smem, ctx->stream>>>(A);
    CU(cudaGetLastError());
    if (timed) { CU(cudaEventRecord(ctx->kev[2 * ctx->kev_used + 1], ctx->stream)); ++ctx->kev_used; }
    ++ctx->launches;
    return ISOCON_OK;
}

// SWAPPED launch (nn_row_swapped_kernel) of a one-sided pass over clustered targets: for every pair (read, cluster
// representative) of the list, every member of that cluster is aligned with the read -- with the MEMBERS as the rows of
// the diagonal-band row kernel and the READS as its lanes (edit distance is symmetric; each lane uses its own read's
// threshold min(best, cap), lowers its own read's best and reports the edge (read, member)).  Layout C: the reads
// cluster after cluster, each cluster's reads padded to whole groups of 32; one row per member over the groups of its
// cluster's reads.  Leaves the candidates' cluster layout (h_tposA) in force, like it found it.
int launch_swapped(isocon_nn_ctx* ctx, const std::vector<int>& reads, const std::vector<int>& reps, int cap, bool sharded,
                   bool* launched) {
    *launched = false;
    // counting sort of the reads by representative (a comparison sort of 100 000 pairs costs more than the launch);
    // inside a cluster list order = length order, so that the lanes of a group are alike
    std::vector<int> first((size_t)ctx->n + 1, 0);
    for (int rep : reps) ++first[(size_t)rep + 1];
    for (long long r = 0; r < ctx->n; ++r) first[(size_t)r + 1] += first[(size_t)r];
    std::vector<int> byrep(reads.size());
    {
        std::vector<int> at(first.begin(), first.end() - 1);
        for (size_t i = 0; i < reads.size(); ++i) byrep[(size_t)at[(size_t)reps[i]]++] = reads[i];
    }
    std::vector<int> tposC, c_g0((size_t)ctx->n, -1), c_ng((size_t)ctx->n, 0);
    tposC.reserve(byrep.size() + byrep.size() / 4 + 64);
    for (long long r = 0; r < ctx->n; ++r) {
        const int a = first[(size_t)r], b = first[(size_t)r + 1];
        if (a == b) continue;
        if (!std::is_sorted(byrep.begin() + a, byrep.begin() + b)) std::sort(byrep.begin() + a, byrep.begin() + b);
        c_g0[(size_t)r] = (int)(tposC.size() / 32);
        tposC.insert(tposC.end(), byrep.begin() + a, byrep.begin() + b);
        while (tposC.size() % 32) tposC.push_back(-1);
        c_ng[(size_t)r] = (int)(tposC.size() / 32) - c_g0[(size_t)r];
    }
    ItemTable S;
    S.row_kernel = true; S.swapped = true;
    const int tile_groups = 2 * ROW_WARPS;   // a cluster with more reads than that: several tiles per member
    long long items = 0;
    for (int t : ctx->h_tposA) {             // the members, cluster after cluster
        if (t < 0) continue;
        const int rep = ctx->h_root[(size_t)t];
        if (rep < 0 || c_ng[(size_t)rep] == 0) continue;
        S.add_row(t); S.add_segment(c_g0[(size_t)rep], c_ng[(size_t)rep]);
        S.gsize.push_back(std::min(tile_groups, c_ng[(size_t)rep]));
        S.item_off.push_back(items);
        items += (c_ng[(size_t)rep] + S.gsize.back() - 1) / S.gsize.back();
    }
    S.item_off.push_back(items);
    S.segoff.push_back((int)S.seg_g0.size());
    if (items == 0) return ISOCON_OK;
    int rc = use_layout(ctx, tposC);
    if (rc) return rc;
    GraphArgs B = base_args(ctx);
    B.pass = PASS_MAIN; B.kcap = cap; B.kprev = -1; B.append = 1; B.symmetric = 0;
    const long long rows_before = ctx->last_run_rows;
    rc = launch_tile(ctx, B, S, sharded, -1);
    if (rc) return rc;
    ctx->last_run_rows = rows_before;        // (rows = queries served; these rows are targets)
    *launched = true;
    return use_layout(ctx, ctx->h_tposA);
}

}  // namespace

extern "C" {

int isocon_nn_device_count(int* count) {
    isocon_nn_ctx* ctx = nullptr;
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) { *count = 0; return fail(ctx, ISOCON_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e)); }
    *count = c;
    return ISOCON_OK;
}

const char* isocon_nn_last_error(const isocon_nn_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int isocon_nn_create(int device, isocon_nn_ctx** out) {
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, ISOCON_ERR_CUDA, "no CUDA device (%s); libisocon_nn has no CPU path",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= count) return fail(nullptr, ISOCON_ERR_ARG, "device %d out of range [0,%d)", device, count);
    isocon_nn_ctx* ctx = new isocon_nn_ctx();
    ctx->device = device;
    e = cudaSetDevice(device);
    cudaDeviceProp prop{};
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev1);
    for (int i = 0; i < 2 * isocon_nn_ctx::KEV && e == cudaSuccess; ++i) e = cudaEventCreate(&ctx->kev[i]);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->evt0);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->evt1);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_pilot, cudaEventDisableTiming);
    if (e == cudaSuccess) e = ctx->d_small.ensure(SM_WORDS);
    if (e != cudaSuccess) {
        fail(nullptr, ISOCON_ERR_CUDA, "context creation on device %d: %s", device, cudaGetErrorString(e));
        delete ctx;
        return ISOCON_ERR_CUDA;
    }
    ctx->num_sms = prop.multiProcessorCount;
    if (const char* s = getenv("ISOCON_NN_EDGE_CAPACITY")) ctx->opt_edge_capacity = atoll(s);
    if (const char* s = getenv("ISOCON_NN_KCAP")) ctx->opt_kcap_main = std::max(1, atoi(s));
    if (const char* s = getenv("ISOCON_NN_SEED")) ctx->opt_seed = atoi(s);
    if (const char* s = getenv("ISOCON_NN_BLOCKS_PER_SM")) ctx->opt_blocks_per_sm = atoi(s);
    if (const char* s = getenv("ISOCON_NN_ROW_KERNEL")) ctx->opt_row_kernel = atoi(s);
    if (const char* s = getenv("ISOCON_NN_BINS")) ctx->opt_bins = atoi(s);
    if (const char* s = getenv("ISOCON_NN_PILOT_DIV")) ctx->opt_pilot_div = std::max(2, atoi(s));
    if (const char* s = getenv("ISOCON_NN_NARROW")) ctx->opt_narrow = std::max(0, atoi(s));
    if (const char* s = getenv("ISOCON_NN_LADDER")) ctx->opt_ladder = atoi(s);
    if (const char* s = getenv("ISOCON_NN_CLASS_GRAN")) ctx->opt_class_gran = std::max(1, atoi(s));
    if (const char* s = getenv("ISOCON_NN_LADDER_FIRST")) ctx->opt_ladder_first = atoi(s);
    if (const char* s = getenv("ISOCON_NN_DEBUG")) ctx->opt_debug = atoi(s);
    if (const char* s = getenv("ISOCON_NN_CLUSTER")) ctx->opt_cluster = atoi(s);
    if (const char* s = getenv("ISOCON_NN_ORDER_BEST")) ctx->opt_order_best = atoi(s);
    if (const char* s = getenv("ISOCON_NN_TWO_LEVEL")) ctx->opt_two_level = atoi(s);
    if (const char* s = getenv("ISOCON_NN_SEED_SAMPLE")) ctx->opt_seed_sample = atoi(s);
    if (const char* s = getenv("ISOCON_NN_SWAP")) ctx->opt_swap = atoi(s);    // 1: SEED pass, 2: level 2 as well
    if (const char* s = getenv("ISOCON_NN_QGRAM")) ctx->opt_qgram = atoi(s);
    if (const char* s = getenv("ISOCON_NN_SURV_CAP")) ctx->opt_surv_cap = atoll(s);
    if (const char* s = getenv("ISOCON_NN_FUSE")) ctx->opt_fuse = atoi(s);
    if (const char* s = getenv("ISOCON_NN_BRIDGE")) ctx->opt_bridge = atoi(s);
    if (const char* s = getenv("ISOCON_NN_PRIMER")) ctx->opt_primer = atoi(s);
    *out = ctx;
    return ISOCON_OK;
}

void isocon_nn_destroy(isocon_nn_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    close_peers(ctx);
    for (void* p : ctx->retired_best) cudaFree(p);
    ctx->d_ascii.release(); ctx->d_off.release(); ctx->d_rowoff.release(); ctx->d_len.release();
    ctx->d_rowpk.release(); ctx->d_small.release(); ctx->d_isq.release(); ctx->d_ist.release();
    ctx->d_tpos.release(); ctx->d_best.release(); ctx->d_qlist.release();
    ctx->d_segoff.release(); ctx->d_gtotal.release(); ctx->d_gsize.release(); ctx->d_seg_g0.release();
    ctx->d_seg_n.release(); ctx->d_goff.release(); ctx->d_item_off.release(); ctx->d_il.release();
    ctx->d_scratch.release(); ctx->d_eq.release(); ctx->d_et.release(); ctx->d_ed.release();
    ctx->d_share.release();
    ctx->d_pa.release(); ctx->d_pb.release(); ctx->d_pk.release(); ctx->d_pout.release(); ctx->d_runoff.release();
    ctx->d_flag.release(); ctx->d_newoff.release(); ctx->d_fascii.release(); ctx->d_foff.release(); ctx->d_flist.release();
    ctx->host_buf.release(); ctx->bounce.release(); ctx->best_host.release(); ctx->pnear_host.release();
    ctx->d_rank.release(); ctx->d_snap.release(); ctx->d_sig.release(); ctx->sig_host.release();
    ctx->d_slack.release(); ctx->d_sq.release(); ctx->d_st.release();
    ctx->d_sigkeys.release(); ctx->d_sigvals.release(); ctx->d_hint.release(); ctx->d_qgram.release();
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    for (int i = 0; i < 2 * isocon_nn_ctx::KEV; ++i) if (ctx->kev[i]) cudaEventDestroy(ctx->kev[i]);
    if (ctx->evt0) cudaEventDestroy(ctx->evt0);
    if (ctx->evt1) cudaEventDestroy(ctx->evt1);
    if (ctx->ev_pilot) cudaEventDestroy(ctx->ev_pilot);
    ctx->fetch_host.release();
    for (int b = 0; b < 2; ++b) {
        if (ctx->stage[b]) cudaFreeHost(ctx->stage[b]);
        if (ctx->stage_ev[b]) cudaEventDestroy(ctx->stage_ev[b]);
    }
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int isocon_nn_sync(isocon_nn_ctx* ctx) {
    if (!ctx) return ISOCON_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    return ISOCON_OK;
}

int isocon_nn_timer_start(isocon_nn_ctx* ctx) {
    if (!ctx) return ISOCON_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventRecord(ctx->evt0, ctx->stream));
    return ISOCON_OK;
}

int isocon_nn_timer_stop(isocon_nn_ctx* ctx, float* ms) {
    if (!ctx || !ms) return ISOCON_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventRecord(ctx->evt1, ctx->stream));
    CU(cudaEventSynchronize(ctx->evt1));
    CU(cudaEventElapsedTime(ms, ctx->evt0, ctx->evt1));
    return ISOCON_OK;
}

int isocon_nn_host_buffer(isocon_nn_ctx* ctx, int64_t bytes, void** ptr) {
    if (!ctx || !ptr || bytes < 0) return ISOCON_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));          // an upload from the old buffer may be in flight
    CU(ctx->host_buf.ensure((size_t)bytes + 64));
    *ptr = ctx->host_buf.p;
    return ISOCON_OK;
}

int isocon_nn_store_reset(isocon_nn_ctx* ctx, const uint8_t* alphabet) {
    if (!ctx) return ISOCON_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    static const uint8_t dna[4] = {'A', 'C', 'G', 'T'};
    const uint8_t* abc = alphabet ? alphabet : dna;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < i; ++j)
            if (abc[i] == abc[j]) return fail(ctx, ISOCON_ERR_ARG, "store_reset: the four alphabet symbols must differ");
    memcpy(ctx->alphabet, abc, 4);
    ctx->s_len.clear(); ctx->s_off.clear(); ctx->arena_used = 0;
    ctx->s_foff.clear(); ctx->s_nsym.clear(); ctx->fascii_used = 0;
    ctx->n = 0; ctx->h_len.clear(); ctx->h_slot.clear(); ctx->h_foreign.clear(); ctx->n_foreign = 0;
    ctx->graph_open = false; ctx->finalized = false;
    ++ctx->store.resets;
    return ISOCON_OK;
}

int isocon_nn_store_add(isocon_nn_ctx* ctx, const uint8_t* ascii, const int64_t* offsets, int64_t n_new, int64_t* first_slot) {
    if (!ctx) return ISOCON_ERR_ARG;
    if (n_new < 0 || !offsets || (long long)ctx->s_len.size() + n_new > INT_MAX - 64 || (!ascii && n_new > 0 && offsets[n_new] > offsets[0]))
        return fail(ctx, ISOCON_ERR_ARG, "store_add: bad arguments (n=%lld)", (long long)n_new);
    CU(cudaSetDevice(ctx->device));
    const long long slot0 = (long long)ctx->s_len.size();
    if (first_slot) *first_slot = slot0;
    if (n_new == 0) return ISOCON_OK;
    std::vector<long long> off0((size_t)n_new + 1), newoff((size_t)n_new);
    long long words = ctx->arena_used;
    for (int64_t i = 0; i < n_new; ++i) {
        const int64_t l = offsets[i + 1] - offsets[i];
        if (l < 0 || l > (1 << 28)) return fail(ctx, ISOCON_ERR_ARG, "store_add: read %lld has length %lld", (long long)i, (long long)l);
        off0[i] = offsets[i] - offsets[0];
        newoff[i] = words;
        words += ((l + 15) >> 4) + 4;  // 4 zero words of padding per read
    }
    off0[n_new] = offsets[n_new] - offsets[0];
    const int64_t total = off0[n_new];
    CU(cudaEventRecord(ctx->ev0, ctx->stream));
    CU(ctx->d_rowpk.grow_keep((size_t)ctx->arena_used, (size_t)words + 64, ctx->stream));
    CU(ctx->d_ascii.ensure((size_t)total + 16));
    CU(ctx->d_off.ensure((size_t)n_new + 1)); CU(ctx->d_newoff.ensure((size_t)n_new + 1)); CU(ctx->d_flag.ensure((size_t)n_new + 1));
    const uint8_t* src = ascii + offsets[0];
    if (ctx->host_buf.p && src >= ctx->host_buf.p && src + total <= ctx->host_buf.p + ctx->host_buf.cap) {
        // gathered straight into our pinned buffer (isocon_nn_host_buffer): one asynchronous copy
        if (total) CU(cudaMemcpyAsync(ctx->d_ascii.p, src, (size_t)total, cudaMemcpyHostToDevice, ctx->stream));
    } else {
        // The caller's buffer is pageable: a plain cudaMemcpyAsync stages it through the driver at 1-3 GB/s.  Two
        // pinned buffers of our own, filled by memcpy while the other one is in flight, move it at host-memcpy speed.
        for (int b = 0; b < 2; ++b)
            if (!ctx->stage[b]) {
                CU(cudaHostAlloc((void**)&ctx->stage[b], STAGE_BYTES, cudaHostAllocDefault));
                CU(cudaEventCreateWithFlags(&ctx->stage_ev[b], cudaEventDisableTiming));
            }
        size_t done = 0;
        int turn = 0;
        bool used[2] = {false, false};
        while (done < (size_t)total) {
            const size_t len = std::min<size_t>(STAGE_BYTES, (size_t)total - done);
            if (used[turn]) CU(cudaEventSynchronize(ctx->stage_ev[turn]));
            memcpy(ctx->stage[turn], src + done, len);
            CU(cudaMemcpyAsync(ctx->d_ascii.p + done, ctx->stage[turn], len, cudaMemcpyHostToDevice, ctx->stream));
            CU(cudaEventRecord(ctx->stage_ev[turn], ctx->stream));
            used[turn] = true;
            done += len; turn ^= 1;
        }
    }
    CU(cudaMemcpyAsync(ctx->d_off.p, off0.data(), (size_t)(n_new + 1) * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_newoff.p, newoff.data(), (size_t)n_new * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(ctx->d_flag.p, 0, (size_t)n_new, ctx->stream));
    const uint32_t abc = (uint32_t)ctx->alphabet[0] | ((uint32_t)ctx->alphabet[1] << 8) | ((uint32_t)ctx->alphabet[2] << 16) |
                         ((uint32_t)ctx->alphabet[3] << 24);
    pack_rows_kernel<<<(unsigned)n_new, 64, 0, ctx->stream>>>(ctx->d_ascii.p, ctx->d_off.p, ctx->d_newoff.p, (int)n_new, abc,
                                                           ctx->d_rowpk.p, ctx->d_flag.p);
    CU(cudaGetLastError());
    std::vector<uint8_t> flag((size_t)n_new);
    CU(cudaMemcpyAsync(flag.data(), ctx->d_flag.p, (size_t)n_new, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaEventRecord(ctx->ev1, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaEventElapsedTime(&ctx->ms[0], ctx->ev0, ctx->ev1));
    // foreign sequences keep their raw bytes on the device too
    std::vector<long long> dst((size_t)n_new, -1);
    std::vector<int> nsym((size_t)n_new, 0);
    long long fbytes = ctx->fascii_used, n_for = 0;
    for (int64_t r = 0; r < n_new; ++r) {
        if (!flag[r]) continue;
        dst[r] = fbytes;
        fbytes += off0[r + 1] - off0[r];
        bool seen[256] = {};
        for (long long pos = off0[r]; pos < off0[r + 1]; ++pos)
            if (!seen[src[pos]]) { seen[src[pos]] = true; ++nsym[r]; }
        ++n_for;
    }
    if (n_for) {
        CU(ctx->d_fascii.grow_keep((size_t)ctx->fascii_used, (size_t)fbytes + 64, ctx->stream));
        CU(cudaMemcpyAsync(ctx->d_newoff.p, dst.data(), (size_t)n_new * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
        keep_ascii_kernel<<<(unsigned)n_new, 128, 0, ctx->stream>>>(ctx->d_ascii.p, ctx->d_off.p, ctx->d_newoff.p, (int)n_new,
                                                                  ctx->d_fascii.p);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(ctx->stream));
        ctx->fascii_used = fbytes;
        ctx->store.foreign_reads += (uint64_t)n_for;
    }
    for (int64_t i = 0; i < n_new; ++i) {
        ctx->s_len.push_back((int)(off0[i + 1] - off0[i]));
        ctx->s_off.push_back(newoff[i]);
        ctx->s_foff.push_back(dst[i]);
        ctx->s_nsym.push_back(nsym[i]);
    }
    ctx->arena_used = words;
    ctx->store.uploaded_reads += (uint64_t)n_new;
    ctx->store.uploaded_bytes += (uint64_t)total;
    ++ctx->store.upload_calls;
    return ISOCON_OK;
}

int isocon_nn_set_list(isocon_nn_ctx* ctx, const int32_t* slots, int64_t n) {
    if (!ctx) return ISOCON_ERR_ARG;
    if (n < 0 || n > INT_MAX - 64 || (n > 0 && !slots)) return fail(ctx, ISOCON_ERR_ARG, "set_list: bad arguments (n=%lld)", (long long)n);
    CU(cudaSetDevice(ctx->device));
    ctx->graph_open = false; ctx->finalized = false;
    const long long n_slots = (long long)ctx->s_len.size();
    std::vector<int> len((size_t)n);
    int max_len = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (slots[i] < 0 || slots[i] >= n_slots) return fail(ctx, ISOCON_ERR_ARG, "set_list: entry %lld names slot %d of %lld", (long long)i, slots[i], n_slots);
        len[i] = ctx->s_len[slots[i]];
        if (i > 0 && len[i] < len[i - 1])
            return fail(ctx, ISOCON_ERR_ARG, "set_list: list is not sorted by length at entry %lld", (long long)i);
        max_len = std::max(max_len, len[i]);
    }
    ctx->n = n;
    ctx->h_len.swap(len);
    ctx->h_slot.assign(slots, slots + n);
    ctx->h_foreign.assign((size_t)n, 0);
    ctx->n_foreign = 0; ctx->gen_syms = 4;
    for (int64_t i = 0; i < n; ++i)
        if (ctx->s_foff[slots[i]] >= 0) {
            ctx->h_foreign[i] = 1; ++ctx->n_foreign;
            ctx->gen_syms = std::max(ctx->gen_syms, ctx->s_nsym[slots[i]]);
        }
    ctx->max_len = max_len;
    ctx->nbmax = std::max(1, (ctx->max_len + 31) >> 5);
    ctx->peq_words = ctx->nbmax + PEQ_PAD_WORDS;
    CU(ctx->d_rowoff.ensure((size_t)n + 1)); CU(ctx->d_len.ensure((size_t)n + 1));
    if ((size_t)n + 1 > ctx->d_best.cap) {
        // best[] moves: the peers' mappings go stale.  An exported allocation must outlive those mappings,
        // so it is parked until isocon_nn_set_peers(world <= 1 or new handles) + release on every rank.
        CU(cudaStreamSynchronize(ctx->stream));
        if (ctx->best_exported && ctx->d_best.p) {
            ctx->retired_best.push_back((void*)ctx->d_best.p);
            ctx->d_best.p = nullptr; ctx->d_best.cap = 0;
        }
        close_peers(ctx);
        ++ctx->best_generation;
        ctx->best_exported = false;
        CU(ctx->d_best.ensure((size_t)n + 1));
    }
    {
        int rc = ensure_share(ctx, n, ctx->share_f);
        if (rc) return rc;
    }
    if (n) {
        CU(cudaStreamSynchronize(ctx->stream));
        ctx->bounce.used = 0;
        CU(ctx->bounce.ensure((size_t)n * 64 + (1u << 20)));
        long long* ro = (long long*)ctx->bounce.take((size_t)n * sizeof(long long));
        int* ln = (int*)ctx->bounce.take((size_t)n * sizeof(int));
        for (int64_t i = 0; i < n; ++i) { ro[i] = ctx->s_off[slots[i]]; ln[i] = ctx->h_len[i]; }
        CU(cudaMemcpyAsync(ctx->d_rowoff.p, ro, (size_t)n * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->d_len.p, ln, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        if (ctx->n_foreign) {
            CU(ctx->d_foff.ensure((size_t)n + 1));
            long long* fo = (long long*)ctx->bounce.take((size_t)n * sizeof(long long));
            for (int64_t i = 0; i < n; ++i) fo[i] = ctx->s_foff[slots[i]];
            CU(cudaMemcpyAsync(ctx->d_foff.p, fo, (size_t)n * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
        }
    }
    ++ctx->store.lists;
    return configure_launch(ctx);
}

int isocon_nn_set_reads(isocon_nn_ctx* ctx, const uint8_t* ascii, const int64_t* offsets, int64_t n) {
    if (!ctx) return ISOCON_ERR_ARG;
    if (n < 0 || n > INT_MAX - 64 || !offsets) return fail(ctx, ISOCON_ERR_ARG, "set_reads: bad arguments (n=%lld)", (long long)n);
    uint8_t abc[4];
    memcpy(abc, ctx->alphabet, 4);
    int rc = isocon_nn_store_reset(ctx, abc);
    if (rc) return rc;
    int64_t first = 0;
    rc = isocon_nn_store_add(ctx, ascii, offsets, n, &first);
    if (rc) return rc;
    std::vector<int32_t> ident((size_t)n);
    for (int64_t i = 0; i < n; ++i) ident[i] = (int32_t)i;
    return isocon_nn_set_list(ctx, ident.data(), n);
}

int isocon_nn_store_info(isocon_nn_ctx* ctx, isocon_nn_store_stats* out) {
    if (!ctx || !out) return ISOCON_ERR_ARG;
    ctx->store.slots = (uint64_t)ctx->s_len.size();
    ctx->store.arena_words = (uint64_t)ctx->arena_used;
    ctx->store.list_entries = (uint64_t)ctx->n;
    *out = ctx->store;
    return ISOCON_OK;
}

int isocon_nn_graph_begin(isocon_nn_ctx* ctx, const isocon_nn_params* P) {
    if (!ctx || !P) return ISOCON_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    const long long n = ctx->n;
    if (P->mode != 1 && P->mode != 2) return fail(ctx, ISOCON_ERR_ARG, "graph_begin: mode must be 1 or 2");
    if (n > 0 && !P->is_query) return fail(ctx, ISOCON_ERR_ARG, "graph_begin: is_query is required");
    if (P->mode == 2 && n > 0 && !P->is_target) return fail(ctx, ISOCON_ERR_ARG, "graph_begin: is_target is required in mode 2");
    if (P->depth < 0) return fail(ctx, ISOCON_ERR_ARG, "graph_begin: depth must be >= 0");
    if (P->world < 0 || (P->world > 0 && (P->rank < 0 || P->rank >= P->world)))
        return fail(ctx, ISOCON_ERR_ARG, "graph_begin: bad rank/world %d/%d", P->rank, P->world);
    ctx->prm = *P;
    if (ctx->prm.world == 0) { ctx->prm.world = 1; ctx->prm.rank = 0; }
    // 1-set: the scan tests `j >= depth` after offset j = 1 (:190), so depth <= 1 all mean "offset 1 only";
    // 2-set: `processed >= depth` after every offset (:416), so depth 0 stops after offset 1 whatever it aligned
    if (P->mode == 1 && ctx->prm.depth < 1) ctx->prm.depth = 1;
    CU(cudaStreamSynchronize(ctx->stream));   // normally idle already; the bounce arena starts over
    ctx->bounce.used = 0;
    ctx->graph_open = false; ctx->finalized = false; ctx->n_final = 0;
    ctx->pilot_rows = 0; ctx->ms[5] = 0.f; ctx->stats.unresolved_rows = 0; ctx->stats.bins = 1;
    ctx->ladder_prev = -1; ctx->ladder_level = 0; ctx->main_done = false; ctx->seed_rows = 0; ctx->stats.main_passes = 0;
    ctx->cluster_pilot = false; ctx->clustered = false; ctx->bins_unsorted = false; ctx->stats.clusters = 0; ctx->snap_valid = false;
    ctx->two_level = false; ctx->qgram_ready = false;
    ctx->pilot_prefetched = false; ctx->spec_edges = 0;
    ctx->h_isq.assign(P->is_query, P->is_query + n);
    if (P->mode == 2) ctx->h_ist.assign(P->is_target, P->is_target + n); else ctx->h_ist.assign((size_t)n, 1);
    ctx->prm.is_query = nullptr; ctx->prm.is_target = nullptr;
    long long n_targets = 0;
    for (long long i = 0; i < n; ++i) {
        if (ctx->h_ist[i]) ++n_targets;
        if (ctx->h_isq[i] && P->mode == 2 && ctx->h_ist[i])
            return fail(ctx, ISOCON_ERR_ARG, "graph_begin: entry %lld is both query and target", i);
    }
    // algorithm: the closed form needs the whole window; the 2-set depth counts alignments
    int algo = P->algo;
    if (const char* s = getenv("ISOCON_NN_ALGO")) { if (atoi(s) > 0) algo = atoi(s); }
    if (algo == ISOCON_ALGO_AUTO)
        algo = (P->mode == 2 && P->depth < n_targets) ? ISOCON_ALGO_SCAN : ISOCON_ALGO_TILE;
    if (algo == ISOCON_ALGO_TILE && P->mode == 2 && P->depth < n_targets)
        return fail(ctx, ISOCON_ERR_ARG, "graph_begin: the tile algorithm cannot honour a finite 2-set depth; use SCAN");
    ctx->algo = algo;
    ctx->symmetric = (P->mode == 1 && algo == ISOCON_ALGO_TILE) ? (P->symmetric != 0) : 0;
    if (const char* s = getenv("ISOCON_NN_SYMMETRIC")) { if (P->mode == 1 && algo == ISOCON_ALGO_TILE) ctx->symmetric = atoi(s) != 0; }
    // Foreign entries (a symbol outside the store's alphabet): the scan emulation aligns their pairs itself, lane
    // by lane; the pair-matrix algorithm leaves them out of its rows and target layout and gives them a pass of
    // their own (nn_foreign_kernel) after the MAIN passes.
    const bool split_foreign = algo == ISOCON_ALGO_TILE && ctx->n_foreign > 0;
    ctx->h_qlist.clear(); ctx->h_flist.clear(); ctx->foreign_level = 0;
    ctx->h_ist_main = ctx->h_ist;
    long long clean_entries = 0;
    for (long long i = 0; i < n; ++i) {
        const bool foreign = split_foreign && ctx->h_foreign[i];
        if (foreign) {
            ctx->h_ist_main[i] = 0;
            if (ctx->h_isq[i] || ctx->h_ist[i]) ctx->h_flist.push_back((int)i);
        } else {
            ++clean_entries;
            if (ctx->h_isq[i]) ctx->h_qlist.push_back((int)i);
        }
    }
    set_layout(ctx, std::vector<int>((size_t)n, 0), 1);   // one bin: all targets in list order
    ctx->nT = (int)ctx->h_tpos.size();
    ctx->nG = ctx->nT / 32;
    ctx->all_queries = (long long)ctx->h_qlist.size() == clean_entries;

    // device state
    CU(ctx->d_isq.ensure((size_t)n + 1)); CU(ctx->d_ist.ensure((size_t)n + 1));
    ctx->ecap = ctx->opt_edge_capacity > 0 ? ctx->opt_edge_capacity : std::max<long long>(1 << 20, 64 * n);
    if (ctx->edge_reserve > 0) ctx->ecap = std::max(ctx->ecap, ctx->edge_reserve);
    if (ctx->edge_reserve < 0) ctx->ecap = -ctx->edge_reserve;
    CU(ctx->d_eq.ensure((size_t)ctx->ecap)); CU(ctx->d_et.ensure((size_t)ctx->ecap)); CU(ctx->d_ed.ensure((size_t)ctx->ecap));
    {
        int rc = ensure_share(ctx, n, ctx->ecap);     // final edges of ALL ranks land here: same capacity as the candidates
        if (rc) return rc;
    }
    ctx->fused = false;
    ctx->launches = 0;
    ctx->best_host_launches = ~0ull;
    CU(cudaEventRecord(ctx->ev0, ctx->stream));
    if (n) {
        int rc = h2d(ctx, ctx->d_isq.p, ctx->h_isq.data(), (size_t)n);
        if (!rc) rc = h2d(ctx, ctx->d_ist.p, ctx->h_ist.data(), (size_t)n);
        if (!rc) rc = apply_layout(ctx);
        if (rc) return rc;
        init_best_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_len.p, (int)n, ctx->d_best.p);
        CU(cudaGetLastError());
        ++ctx->launches;
    }
    // everything but the box-wide tile queues, which a peer may already be pulling from (they were zeroed when
    // the previous graph was finalized)
    CU(cudaMemsetAsync(ctx->d_small.p, 0, SM_QUEUE * sizeof(unsigned long long), ctx->stream));
    CU(cudaMemsetAsync(ctx->d_small.p + SM_STATS, 0, ST_COUNT * sizeof(unsigned long long), ctx->stream));
    if (ctx->prm.world <= 1) CU(cudaMemsetAsync(ctx->d_small.p + SM_QUEUE, 0, SM_NQUEUE * sizeof(unsigned long long), ctx->stream));
    // own share block: no nearest-pilot-row records yet, no final edges, no overflow (peers write here only between the
    // barriers of a fused run, which starts with one)
    if (n) CU(cudaMemsetAsync(ctx->sv.pnear, 0xff, 2 * (size_t)n * sizeof(unsigned long long), ctx->stream));
    CU(cudaMemsetAsync(ctx->sv.ctrl + CT_FCOUNT, 0, 2 * sizeof(unsigned long long), ctx->stream));
    ctx->ms[1] = 0.f;
    ctx->graph_open = true;
    return ISOCON_OK;                          // no synchronisation: the first graph_run queues behind this
}

}  // extern "C"

namespace {

// The phases of one graph_run call (see isocon_nn.h).  final_sync: wait for the device and read the kernel timers
// (the fused multi-rank flow strings several calls together and waits once).
int run_phases(isocon_nn_ctx* ctx, int phases, bool final_sync) {
    const auto host_t0 = std::chrono::steady_clock::now();
    ctx->last_run_rows = 0;
    if (ctx->n == 0) return ISOCON_OK;
    if (final_sync) CU(cudaEventRecord(ctx->ev0, ctx->stream));
    int rc = ISOCON_OK;
    const bool no_pairs = ctx->h_qlist.empty() || ctx->nT == 0;     // nothing for the 2-bit pair kernels
    if (no_pairs) {
    } else if (ctx->algo == ISOCON_ALGO_SCAN) {
        if ((phases & ISOCON_PHASE_MAIN) && !ctx->main_done) {   // ONE pass: a driver that calls MAIN until no rows are left stops here
            ctx->main_done = true;
            ItemTable T;   // one item per query; only qlist is used by the scan kernel
            for (int q : ctx->h_qlist) T.add_row(q);
            T.segoff.push_back(0);
            T.gsize.assign(T.qlist.size(), 1);
            T.item_off.resize(T.qlist.size() + 1);
            for (size_t i = 0; i <= T.qlist.size(); ++i) T.item_off[i] = (long long)i;
            rc = upload_items(ctx, T);
            if (rc) return rc;
            GraphArgs A = base_args(ctx);   // after upload_items: it may reallocate the item arrays
            A.nQ = (int)T.qlist.size();
            ctx->last_run_rows += (long long)T.qlist.size();
            shard(T.total(), ctx->prm.rank, ctx->prm.world, A);
            if (A.item_end > A.item_begin) {
                CU(cudaMemsetAsync(ctx->d_small.p + SM_COUNTER, 0, sizeof(unsigned long long), ctx->stream));
                nn_scan_kernel<<<ctx->grid, WARPS_PER_BLOCK * 32, ctx->smem, ctx->stream>>>(A);
                CU(cudaGetLastError());
                ++ctx->launches;
            }
        }
    } else {
        const int kcap = ctx->opt_kcap_main;
        const size_t nq = ctx->h_qlist.size();
        const bool upper_only = ctx->symmetric && ctx->all_queries;
        // The symmetric graph seeds itself with a PILOT pass (the first 5 % of the rows against everything
        // behind them: every pair is aligned once anyway, and afterwards each read has met a sample of
        // its candidates) and then re-bins the targets by threshold class for the MAIN pass.
        const bool pilot = ctx->symmetric && ctx->opt_bins && ctx->row_grid > 0 && nq >= 20 && ctx->prm.depth >= ctx->n;
        if ((phases & ISOCON_PHASE_SEED) && ctx->opt_seed && !pilot) {
            // each query against the (up to) 3 groups around its own position in the target list
            // With the MAIN ladder the seeds only pick the first cap (90th percentile of the seeded bests): an evenly
            // spaced sample of the queries tells as much as all of them (c5: the SEED passes were 19 % of the step).
            const bool ladder = ctx->opt_ladder && !ctx->symmetric && ctx->row_grid > 0 && (nq >= 64 || ctx->opt_ladder_first > 0);
            const size_t step = ladder ? std::max<size_t>(1, nq / std::max<size_t>(2048, nq / 32)) : 1;
            // similarity order of the targets + hints (sketch_order), when every row's length window holds (nearly)
            // all targets anyway -- then taking the bins whole costs nothing
            if (ladder && ctx->opt_cluster && nq >= 512 && ctx->nT >= 256 && ctx->nT <= (1 << 18) && !ctx->bins_unsorted) {
                int lmin = INT_MAX, lmax = 0;
                for (int t : ctx->h_tpos) if (t >= 0) { lmin = std::min(lmin, ctx->h_len[(size_t)t]); lmax = std::max(lmax, ctx->h_len[(size_t)t]); }
                if (lmax - lmin <= 127) { rc = sketch_order(ctx); if (rc) return rc; }
            }
            ItemTable T;
            size_t n_sample = 0;                        // hinted rows only: the rows that pick the first cap (below)
            size_t n_uncovered = 0;                     // hinted rows the swapped launch (below) does not serve
            bool swap = false;
            if (ctx->bins_unsorted) {
                std::vector<int> hq;                    // every hinted query against its cluster
                for (size_t i = 0; i < nq; ++i) if (ctx->h_hint_g0[(size_t)ctx->h_qlist[i]] >= 0) hq.push_back(ctx->h_qlist[i]);
                if (ctx->opt_seed_sample && hq.size() >= 1024 && kcap > 127)
                    n_sample = std::min<size_t>(2048, std::max<size_t>(256, hq.size() / 64));
                n_sample -= n_sample % (size_t)std::max(1, ctx->prm.world);   // a row stays with one rank through all launches
                // the swapped launch needs every rank to see every result before the launches behind it (shared best[]
                // + a device-side barrier: the fused flow), and it serves the reads whose hinted cluster fits one tile
                swap = ctx->opt_swap && n_sample > 0 && ctx->row_grid > 0 && (ctx->prm.world <= 1 || ctx->fused);
                // an evenly spaced sample first, then the rows the swapped launch leaves out, then the rest
                std::vector<uint8_t> part(hq.size(), 2);
                for (size_t i = 0; i < hq.size(); ++i) if (swap && ctx->h_hint_rep[(size_t)hq[i]] < 0) part[i] = 1;
                for (size_t i = 0; i < n_sample; ++i) part[i * hq.size() / n_sample] = 0;
                for (int round = 0; round < 3; ++round)
                    for (size_t i = 0; i < hq.size(); ++i) {
                        if (part[i] != round) continue;
                        const int q = hq[i];
                        if (round == 1) ++n_uncovered;
                        T.add_row(q); T.add_segment(ctx->h_hint_g0[(size_t)q], ctx->h_hint_n[(size_t)q]);
                    }
            } else {
                for (size_t i = 0; i < nq; i += step) {
                    const int q = ctx->h_qlist[i];
                    const long long ord = std::lower_bound(ctx->h_tpos.begin(), ctx->h_tpos.begin() + ctx->bin_count[0], q) - ctx->h_tpos.begin();
                    const int g = (int)std::min<long long>(ord / 32, ctx->nG - 1);
                    const int a = std::max(0, g - 1), b = std::min(ctx->nG - 1, g + 1);
                    T.add_row(q); T.add_segment(a, b - a + 1);
                }
            }
            T.segoff.push_back((int)T.seg_g0.size());
            const size_t ns = T.qlist.size();
            ctx->seed_rows = ns;
            DebugLap seedlap(ctx->opt_debug >= 2, "seed");
            if (ctx->opt_debug >= 2) fprintf(stderr, "[isocon_nn]   seed: %zu rows of %zu queries\n", ns, nq);
            T.gsize.assign(ns, GROUPS_PER_ITEM);
            T.item_off.resize(ns + 1);
            for (size_t i = 0; i <= ns; ++i) T.item_off[i] = (long long)i;
            int prev = -1;
            // hinted rows meet their relatives: a pair that runs its whole length whatever the cap, so the small caps
            // would only repeat it
            bool resident = false;
            if (n_sample) {
                // Hinted rows align relatives: pairs that run their whole length in a band as wide as the cap, whatever
                // the distance turns out to be.  How far a read is from its candidate is a property of the data (the
                // error rate), the same for all rows: a sample of them at cap 127 tells which cap most rows need, and
                // the rest starts there (c5: distances 75 +- 9 -> cap 96, 4-word bands instead of 5; reads with 1 %
                // errors: 2 words).  Rows that fail at the chosen cap repeat at 127.  Every rank chooses from the rows
                // of the sample it ran itself -- ranks need not agree: a row's caps stay with the rank that owns it.
                GraphArgs A = base_args(ctx);
                A.pass = PASS_SEED; A.kcap = 127; A.kprev = -1; A.append = 1; A.symmetric = ctx->symmetric;
                rc = launch_tile(ctx, A, T, true, -1, 0, (long long)n_sample);
                if (rc) return rc;
                resident = true;
                // The swapped launch deals the members of a cluster to different ranks: its cap must be the same on
                // all of them (a read counts as done at that cap only if all its pairs were tried there), so the
                // ranks look at the whole sample in an agreed snapshot.  Otherwise a private look at the live best[].
                const bool agreed = swap && ctx->fused;
                const int* b = nullptr;
                if (agreed) {
                    rc = agree_on_best(ctx); if (rc) return rc;
                    rc = fetch_best(ctx, &b); if (rc) return rc;
                } else {
                    CU(ctx->best_host.ensure((size_t)ctx->n * sizeof(int) + 64));
                    ctx->best_host_launches = ~0ull;
                    CU(cudaMemcpyAsync(ctx->best_host.p, ctx->d_best.p, (size_t)ctx->n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
                    CU(cudaStreamSynchronize(ctx->stream));
                    b = (const int*)ctx->best_host.p;
                }
                const int world = agreed ? 1 : std::max(1, ctx->prm.world);
                long long seen = 0, within[3] = {0, 0, 0};      // caps 32 / 64 / 96: bands of 2 / 3 / 4 words (5 at 127)
                for (size_t i = (size_t)(world > 1 ? ctx->prm.rank : 0); i < n_sample; i += (size_t)world) {
                    const int d = b[(size_t)T.qlist[i]];
                    ++seen;
                    for (int c = 0; c < 3; ++c) if (d <= 32 * (c + 1)) ++within[c];
                }
                int first_cap = 127;
                if (!swap) {
                    double cost = 5.0;
                    for (int c = 2; c >= 0 && seen > 0; --c) {
                        const double cc = (double)(c + 2) + 5.0 * (double)(seen - within[c]) / (double)seen;
                        if (cc < cost) { cost = cc; first_cap = 32 * (c + 1); }
                    }
                } else {
                    // swapped launch: diagonal bands of (cap + 32) / 32 words, 32 reads per warp; a read that fails there
                    // repeats as a row of its own in the block band (about 25 times the cost of its lane)
                    long long in[3] = {0, 0, 0};
                    for (size_t i = (size_t)(world > 1 ? ctx->prm.rank : 0); i < n_sample; i += (size_t)world)
                        for (int c = 0; c < 3; ++c) if (b[(size_t)T.qlist[i]] <= 32 * (c + 1) - 1) ++in[c];
                    double cost = 4.0;
                    for (int c = 2; c >= 0 && seen > 0; --c) {
                        const double cc = (double)(c + 1) + 25.0 * (double)(seen - in[c]) / (double)seen;
                        if (cc < cost) { cost = cc; first_cap = 32 * (c + 1) - 1; }
                    }
                }
                if (ctx->opt_debug >= 2) fprintf(stderr, "[isocon_nn]   seed: sample of %zu rows -> first cap %d%s\n", n_sample, first_cap, swap ? " (swapped launch)" : "");
                seedlap.lap("sample");
                if (swap) {
                    // SWAPPED launch: every member of a cluster against the reads hinted at it (launch_swapped); the
                    // sample rows are done (complete at 127)
                    std::vector<int> pr((size_t)(ns - n_sample - n_uncovered)), pc(pr.size());
                    for (size_t i = n_sample + n_uncovered; i < ns; ++i) {
                        pr[i - n_sample - n_uncovered] = T.qlist[i];
                        pc[i - n_sample - n_uncovered] = ctx->h_hint_rep[(size_t)T.qlist[i]];
                    }
                    bool launched = false;
                    rc = launch_swapped(ctx, pr, pc, first_cap, true, &launched);
                    if (rc) return rc;
                    if (launched) {
                        // several ranks: every rank's results must have landed in every best[] before the rows behind
                        // decide from best[] whether they still have to run
                        if (ctx->fused) { rc = enqueue_barrier(ctx); if (rc) return rc; }
                        resident = false;                                     // (the tile table on the device is another now)
                        A = base_args(ctx);                                   // (the layout buffers may have moved)
                        A.pass = PASS_SEED; A.append = 1; A.symmetric = ctx->symmetric;
                    }
                    seedlap.lap("swapped launch");
                    // rows the swapped launch left out: from scratch at 127; its own rows: only those still above the cap
                    if (n_uncovered) {
                        A.kcap = 127; A.kprev = -1;
                        rc = launch_tile(ctx, A, T, true, -1, (long long)n_sample, (long long)(n_sample + n_uncovered), resident);
                        if (rc) return rc;
                        resident = true;
                    }
                    if (first_cap < 127) {
                        A.kcap = 127; A.kprev = first_cap;
                        rc = launch_tile(ctx, A, T, true, -1, (long long)(n_sample + n_uncovered), -1, resident);
                        if (rc) return rc;
                        resident = true;
                    }
                    prev = 127;
                } else {
                    for (int cap : {first_cap, 127}) {
                        if (cap <= prev) continue;
                        A.kcap = cap; A.kprev = prev;
                        rc = launch_tile(ctx, A, T, true, -1, (long long)n_sample, -1, true);
                        if (rc) return rc;
                        prev = cap;
                    }
                }
            }
            for (int cap : {63, 127, 255, kcap}) {
                if (cap > kcap || cap <= prev) continue;
                if (ctx->bins_unsorted && (cap == 63 || cap == 255)) continue;
                GraphArgs A = base_args(ctx);
                A.pass = PASS_SEED; A.kcap = cap; A.kprev = prev; A.append = ctx->bins_unsorted ? 1 : 0; A.symmetric = ctx->symmetric;
                rc = launch_tile(ctx, A, T, true, -1, 0, -1, resident);   // ranks seed disjoint shares; best is MIN-reduced next
                if (rc) return rc;
                prev = cap; resident = true;
                seedlap.lap("table+launch");
            }
        }
        if ((phases & ISOCON_PHASE_PILOT) && pilot) {
            // The pilot rows go out in TWO launches when the MAIN pass follows in this very flow: the host starts
            // turning best[] / pnear into the MAIN pass's layout and tile table as soon as the first launch is done,
            // while the GPU works on the last pilot rows (a fixed number per GPU: the host work they cover and
            // their own cost both grow with the number of reads).  The layout is then made from slightly older
            // bounds -- classes are upper bounds, clusters a heuristic: both stay valid.
            const size_t na_all = std::max<size_t>(1, nq / ctx->opt_pilot_div);
            size_t nb = 0;
            if (ctx->opt_bridge && nq >= 2048 && (ctx->fused || (ctx->prm.world <= 1 && (phases & ISOCON_PHASE_MAIN))))
                nb = std::min<size_t>((size_t)ctx->opt_bridge * (size_t)std::max(1, ctx->prm.world), na_all / 4);
            const size_t na = na_all - nb;
            std::vector<int> qs(ctx->h_qlist.begin(), ctx->h_qlist.begin() + na + nb), kw(na + nb, kcap);
            // primer (below): off by default -- measured on 8 GPUs it removes 3 % of the executed work and gives all of
            // it back as latency (one wide alignment on a nearly idle box + a barrier): c2 16.82 -> 17.27 ms,
            // profiles/r02l_*; kept as an option (ISOCON_NN_PRIMER = number of ranks from which it is used)
            const bool primer = (ctx->opt_primer > 0 ? ctx->prm.world >= ctx->opt_primer : false) && nq >= 2048 && na > 1;
            ItemTable T;
            T.row_kernel = true;
            build_items(ctx, qs, kw, upper_only, T, primer ? 1 : 0);
            GraphArgs A = base_args(ctx);
            A.pass = PASS_MAIN; A.kcap = kcap; A.append = 1; A.symmetric = 1;
            // similarity order needs every row's window to hold every target (then a row takes whole bins)
            ctx->cluster_pilot = ctx->opt_cluster && nq >= 512 && ctx->h_len[(size_t)ctx->n - 1] - ctx->h_len[0] <= kcap;
            if (ctx->cluster_pilot) {
                A.pnear = ctx->sv.pnear; A.pilot_last = qs.back();     // (reset to "none" by graph_begin)
                if (ctx->fused) for (int p = 0; p < A.n_peers; ++p) A.peer_pnear[p] = ctx->peer_sv[p].pnear;
            }
            // Primer: the very first row alone.  Until a read has met its first partner its bound is the cap (400:
            // 13-word windows); a full first wave of tiles would align every read dozens of times at that width
            // before the first result lands (8 GPUs: 600 k wide pairs, +4 % work).  One row touches every read once;
            // the launches behind it start from bounds near the final ones.
            long long primer_end = 0;
            if (primer) {
                primer_end = T.item_off[1];
                rc = launch_tile(ctx, A, T, true, 4, 0, primer_end);
                if (rc) return rc;
                if (ctx->fused) { rc = enqueue_barrier(ctx); if (rc) return rc; }   // the peers' results have landed too
            }
            rc = launch_tile(ctx, A, T, true, 0, primer_end, T.item_off[na], primer_end > 0);
            if (rc) return rc;
            ctx->pilot_rows = na + nb;
            if (nb) {
                // what the MAIN pass's layout is made from goes to the host now; the bridge rows run meanwhile
                if (ctx->fused) { rc = agree_on_best(ctx); if (rc) return rc; }
                CU(ctx->best_host.ensure((size_t)ctx->n * sizeof(int) + 64));
                ctx->best_host_launches = ~0ull;
                CU(cudaMemcpyAsync(ctx->best_host.p, ctx->fused ? ctx->d_snap.p : ctx->d_best.p, (size_t)ctx->n * sizeof(int),
                                   cudaMemcpyDeviceToHost, ctx->stream));
                if (ctx->cluster_pilot) {
                    CU(ctx->pnear_host.ensure(2 * (size_t)ctx->n * sizeof(unsigned long long) + 64));
                    CU(cudaMemcpyAsync(ctx->pnear_host.p, ctx->sv.pnear, 2 * (size_t)ctx->n * sizeof(unsigned long long),
                                       cudaMemcpyDeviceToHost, ctx->stream));
                }
                CU(cudaEventRecord(ctx->ev_pilot, ctx->stream));
                ctx->pilot_prefetched = true;
                GraphArgs B = A;
                B.pnear = nullptr;
                for (int p = 0; p < 7; ++p) B.peer_pnear[p] = nullptr;
                rc = launch_tile(ctx, B, T, true, 3, T.item_off[na], T.total(), true);
                if (rc) return rc;
            }
        }
        if (phases & ISOCON_PHASE_MAIN) {
            DebugLap lap(ctx->opt_debug >= 2, "main");
            if (pilot && ctx->pilot_rows > 0 && !ctx->main_done) {
                // Threshold class of a read = window words its pairs need, ceil((best + 1) / 32).  best only
                // falls, so a read never outgrows its class: grouping the targets by class keeps a read that is
                // far from everything (or merely above a word boundary) from widening the band of the 31 reads
                // that would otherwise share its group.  Reads that are not queries never raise a threshold.
                const int* best = nullptr;
                if (ctx->pilot_prefetched) {
                    CU(cudaEventSynchronize(ctx->ev_pilot));          // the bridge rows keep the GPU busy meanwhile
                    best = (const int*)ctx->best_host.p;
                } else {
                    rc = fetch_best(ctx, &best); if (rc) return rc;
                }
                const int gran = ctx->opt_class_gran;
                const int n_classes = (kcap + gran) / gran + 1;
                std::vector<int> cls((size_t)ctx->n, 0);
                for (long long i = 0; i < ctx->n; ++i)
                    if (ctx->h_isq[i]) cls[(size_t)i] = (std::min(best[(size_t)i], kcap) + gran) / gran;
                lap.lap("best_d2h+classes");
                if (ctx->cluster_pilot) {
                    if (!ctx->pilot_prefetched) {
                        CU(ctx->pnear_host.ensure(2 * (size_t)ctx->n * sizeof(unsigned long long) + 64));
                        CU(cudaMemcpyAsync(ctx->pnear_host.p, ctx->sv.pnear, 2 * (size_t)ctx->n * sizeof(unsigned long long),
                                           cudaMemcpyDeviceToHost, ctx->stream));
                        CU(cudaStreamSynchronize(ctx->stream));
                    }
                    ctx->h_rank.assign((size_t)ctx->n, -1);          // marks the pilot rows for cluster_order
                    for (size_t i = 0; i < ctx->pilot_rows; ++i) ctx->h_rank[(size_t)ctx->h_qlist[i]] = (int)i;
                    cluster_order(ctx, (const unsigned long long*)ctx->pnear_host.p, cls, n_classes, best);
                    ctx->clustered = true; ctx->bins_unsorted = true;
                    CU(ctx->d_rank.ensure((size_t)ctx->n + 1));
                    rc = h2d(ctx, ctx->d_rank.p, ctx->h_rank.data(), (size_t)ctx->n * sizeof(int));
                    if (rc) return rc;
                } else {
                    set_layout(ctx, cls, n_classes);
                }
                ctx->stats.bins = ctx->bin_first.size();
                lap.lap("set_layout");
                if (ctx->binned) { rc = apply_layout(ctx); if (rc) return rc; }
                lap.lap("apply_layout");
            }
            // One-sided passes (2-set, or symmetric off) have no pilot: a row starts at best = len, i.e. at the widest
            // window, and keeps it until it meets a true neighbour -- for reads among unrelated candidates (c5:
            // 500 families) that is most of the row.  Such passes climb a LADDER of caps instead: every pair of
            // the rows that are still unresolved (best > previous cap) is aligned with min(best, cap); a row
            // whose best ends <= cap is complete (every pair at its final distance was aligned with a threshold
            // >= that distance), the others are redone at the next cap.  The first cap comes from the rows the
            // SEED pass resolved (90th percentile of their best), later caps double.  A symmetric pass cannot
            // skip rows (a row also serves the reads below it), so it keeps the single pass at kcap.
            const bool ladder = ctx->opt_ladder && !ctx->symmetric && ctx->row_grid > 0 && (nq >= 64 || ctx->opt_ladder_first > 0);
            for (;;) {
                if (!ladder && ctx->main_done) break;
                std::vector<int> qs;
                int cap = kcap;
                if (ladder) {
                    if (ctx->ladder_prev >= kcap) break;
                    const int* best = nullptr;
                    rc = fetch_best(ctx, &best); if (rc) return rc;
                    if (ctx->ladder_prev < 0) {
                        std::vector<int> seeded;
                        for (int q : ctx->h_qlist) if (best[q] < ctx->h_len[q]) seeded.push_back(best[q]);
                        cap = 63;
                        if (seeded.size() >= std::max<size_t>(16, ctx->seed_rows / 50)) {
                            const size_t k90 = seeded.size() * 9 / 10;
                            std::nth_element(seeded.begin(), seeded.begin() + k90, seeded.end());
                            cap = seeded[k90];
                        }
                        cap = std::max(31, (cap + 32) / 32 * 32 - 1);
                        if (ctx->opt_ladder_first > 0) cap = ctx->opt_ladder_first;
                        qs = ctx->h_qlist;
                    } else {
                        cap = 2 * ctx->ladder_prev + 1;
                        for (int q : ctx->h_qlist) if (best[q] > ctx->ladder_prev) qs.push_back(q);
                    }
                    if (cap * 2 > kcap) cap = kcap;       // no pass for a last small step
                    if (qs.empty()) break;
                } else {
                    qs.assign(ctx->h_qlist.begin() + ctx->pilot_rows, ctx->h_qlist.end());
                    if (ctx->clustered) {  // rows in layout order: neighbouring tiles belong to one cluster, work falls along the rows
                        qs.clear();
                        for (int t : ctx->h_tpos)
                            if (t >= 0 && ctx->h_isq[(size_t)t] && ctx->h_rank[(size_t)t] >= (int)ctx->pilot_rows) qs.push_back(t);
                    }
                }
                std::vector<int> kw(qs.size());
                for (size_t i = 0; i < qs.size(); ++i)
                    kw[i] = ctx->symmetric ? cap : std::min(cap, ctx->h_len[qs[i]]);
                const int queue = ctx->ladder_level == 0 ? 1 : (ctx->ladder_level + 2 < SM_NQUEUE ? ctx->ladder_level + 2 : -1);
                bool done_two_level = false;
                if (ladder && ctx->two_level) {
                    // TWO-LEVEL pass (exact, triangle inequality).  Level 1: every row against the cluster
                    // REPRESENTATIVES only, with the cluster's radius added to the row's threshold:
                    // d(q, rep) > k + radius  =>  d(q, c) >= d(q, rep) - d(rep, c) > k for every member c -- one
                    // alignment dismisses the whole cluster.  Level 2: the rows meet the members of the clusters that
                    // survived (typically the one family a read belongs to).  c5: 500 representatives instead of 5 000
                    // candidates per read.
                    rc = use_layout(ctx, ctx->h_tposB); if (rc) return rc;
                    // Level 1 comes in two kinds.  The q-gram FILTER (no alignment) for the first pass: most rows carry a
                    // tight bound from the SEED pass and (k + radius) * q stays well below the read length, so the
                    // count dismisses all but the related clusters.  Later passes hold the few rows that are far from
                    // everything they met; at their thresholds the count proves nothing (c5, cap 191: a third of all
                    // clusters survived -- 5.4 ms of level 2 for 498 rows) while ALIGNING those rows against the
                    // representatives costs little.  ISOCON_NN_QGRAM: 0 never filter, 2 always.
                    const bool use_filter = ctx->opt_qgram >= 2 || (ctx->opt_qgram == 1 && ctx->ladder_level == 0);
                    if (use_filter && !ctx->qgram_ready) {
                        CU(ctx->d_qgram.ensure((size_t)ctx->nG * 32 * QG_WORDS + 64));
                        qgram_targets_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(ctx->d_rowpk.p, ctx->d_rowoff.p, ctx->d_len.p,
                                                                                      ctx->d_tpos.p, ctx->nT, ctx->d_qgram.p);
                        CU(cudaGetLastError());
                        ++ctx->launches;
                        ctx->qgram_ready = true;
                    }
                    ItemTable T1;
                    T1.row_kernel = true;
                    if (!use_filter) build_items(ctx, qs, kw, false, T1);
                    ctx->surv_cap = ctx->opt_surv_cap > 0 ? ctx->opt_surv_cap : std::max<long long>(1 << 20, 8 * (long long)qs.size());
                    CU(ctx->d_sq.ensure((size_t)ctx->surv_cap)); CU(ctx->d_st.ensure((size_t)ctx->surv_cap));
                    CU(cudaMemsetAsync(ctx->d_small.p + SM_SURV, 0, sizeof(unsigned long long), ctx->stream));
                    GraphArgs A1 = base_args(ctx);
                    A1.pass = PASS_MAIN; A1.kcap = cap; A1.append = 1; A1.symmetric = 0;
                    A1.slack = ctx->d_slack.p; A1.surv_q = ctx->d_sq.p; A1.surv_t = ctx->d_st.p;
                    A1.surv_count = ctx->d_small.p + SM_SURV; A1.surv_cap = ctx->surv_cap;
                    if (use_filter) {
                        // level 1 as a pure filter: no alignment, the q-gram count decides which clusters survive
                        A1.qgram = ctx->d_qgram.p;
                        CU(ctx->d_qlist.ensure(qs.size() + 1));
                        rc = h2d(ctx, ctx->d_qlist.p, qs.data(), qs.size() * sizeof(int));
                        if (rc) return rc;
                        const bool timed = ctx->kev_used < isocon_nn_ctx::KEV;
                        if (timed) CU(cudaEventRecord(ctx->kev[2 * ctx->kev_used], ctx->stream));
                        qgram_level1_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(
                            A1, ctx->d_qlist.p, (int)qs.size(), ctx->prm.world > 1 ? ctx->prm.rank : 0, std::max(1, ctx->prm.world),
                            ctx->opt_seed ? ctx->d_hint.p : nullptr);
                        CU(cudaGetLastError());
                        if (timed) { CU(cudaEventRecord(ctx->kev[2 * ctx->kev_used + 1], ctx->stream)); ++ctx->kev_used; }
                        ++ctx->launches;
                        ctx->last_run_rows += (long long)qs.size();
                    } else {
                        rc = launch_tile(ctx, A1, T1, true, queue);
                        if (rc) return rc;
                    }
                    lap.lap("level1");
                    unsigned long long ns = 0;
                    CU(cudaMemcpyAsync(&ns, ctx->d_small.p + SM_SURV, sizeof ns, cudaMemcpyDeviceToHost, ctx->stream));
                    CU(cudaStreamSynchronize(ctx->stream));
                    rc = use_layout(ctx, ctx->h_tposA); if (rc) return rc;
                    if ((long long)ns <= ctx->surv_cap) {
                        std::vector<int> sq((size_t)ns), st((size_t)ns);
                        if (ns) {
                            CU(cudaMemcpyAsync(sq.data(), ctx->d_sq.p, (size_t)ns * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
                            CU(cudaMemcpyAsync(st.data(), ctx->d_st.p, (size_t)ns * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
                            CU(cudaStreamSynchronize(ctx->stream));
                        }
                        // many survivors: level 2 as a swapped launch (members as rows, surviving reads as lanes)
                        bool swapped2 = false;
                        if (ctx->opt_swap >= 2 && ns >= 2048 && !ctx->h_root.empty()) {
                            std::vector<int> pr, pc;
                            for (size_t k = 0; k < (size_t)ns; ++k)
                                if (!(ctx->opt_seed && ctx->h_hint_rep[(size_t)sq[k]] == st[k])) { pr.push_back(sq[k]); pc.push_back(st[k]); }
                            const long long rows_so_far = ctx->last_run_rows;
                            rc = launch_swapped(ctx, pr, pc, cap, false, &swapped2);
                            if (rc) return rc;
                            ctx->last_run_rows = rows_so_far;
                            if (ctx->opt_debug >= 2)
                                fprintf(stderr, "[isocon_nn]   two-level pass at cap %d: %zu rows, %llu survivors, level 2 swapped (%zu pairs)\n",
                                        cap, qs.size(), ns, pr.size());
                            if (pr.empty()) swapped2 = true;     // nothing left to align
                        }
                        ItemTable T2;      // one tile per (row, <= 8 groups of a surviving cluster); this rank's own survivors
                        for (size_t k = 0; k < (size_t)(swapped2 ? 0 : ns); ++k) {
                            // the SEED pass aligned the query with its hinted cluster at thresholds up to the MAIN cap:
                            // every member within min(final best, cap) was found there, with its edge
                            if (ctx->opt_seed && ctx->h_hint_rep[(size_t)sq[k]] == st[k]) continue;
                            const int g0 = ctx->cl_g0[(size_t)st[k]], ng = ctx->cl_ng[(size_t)st[k]];
                            for (int o = 0; o < ng; o += GROUPS_PER_ITEM) { T2.add_row(sq[k]); T2.add_segment(g0 + o, std::min(GROUPS_PER_ITEM, ng - o)); }
                        }
                        T2.segoff.push_back((int)T2.seg_g0.size());
                        T2.gsize.assign(T2.qlist.size(), GROUPS_PER_ITEM);
                        T2.item_off.resize(T2.qlist.size() + 1);
                        for (size_t i = 0; i <= T2.qlist.size(); ++i) T2.item_off[i] = (long long)i;
                        GraphArgs A2 = base_args(ctx);
                        A2.pass = PASS_MAIN; A2.kcap = cap; A2.append = 1; A2.symmetric = 0;
                        const long long rows_so_far = ctx->last_run_rows;
                        if (ctx->opt_debug >= 2)
                            fprintf(stderr, "[isocon_nn]   two-level pass at cap %d: %zu rows, %llu survivors, %zu level-2 tiles\n",
                                    cap, qs.size(), ns, T2.qlist.size());
                        rc = launch_tile(ctx, A2, T2, false);
                        if (rc) return rc;
                        ctx->last_run_rows = rows_so_far;      // (the survivors differ from rank to rank; the rows do not)
                        done_two_level = true;
                        lap.lap("level2");
                    }   // else: more survivors than the buffer holds -- the plain pass below covers everything
                }
                if (!done_two_level) {
                ItemTable T;
                T.row_kernel = ctx->row_grid > 0;   // diagonal-band row kernel
                build_items(ctx, qs, kw, upper_only, T);
                lap.lap("build_items");
                GraphArgs A = base_args(ctx);
                A.pass = PASS_MAIN; A.kcap = cap; A.append = 1; A.symmetric = ctx->symmetric;
                // (after an overflowed level 1 the pass's box-wide queue is spent: deal the tiles round-robin)
                rc = launch_tile(ctx, A, T, true, (ladder && ctx->two_level) ? -1 : queue);
                if (rc) return rc;
                lap.lap("upload+launch");
                }
                ctx->main_done = true; ctx->ladder_prev = cap; ++ctx->ladder_level; ++ctx->stats.main_passes;
                // several ranks: the driver MIN-reduces best[] and calls MAIN again until no rows are left
                if (!ladder || ctx->prm.world > 1) break;
            }
        }
        if (phases & ISOCON_PHASE_WIDE) {
            // rows whose best is still above the register-band limit: full windows, any threshold
            // (the row kernel falls back to the block band / the global-memory band per group)
            const int* best = nullptr;
            rc = fetch_best(ctx, &best); if (rc) return rc;
            std::vector<int> qs, kw;
            for (size_t i = 0; i < nq; ++i) {
                const int q = ctx->h_qlist[i];
                if (best[q] > kcap) { qs.push_back(q); kw.push_back(best[q]); }
            }
            ctx->stats.unresolved_rows = qs.size();
            if (!qs.empty()) {
                ItemTable T;
                T.row_kernel = ctx->row_grid > 0;
                build_items(ctx, qs, kw, false, T);
                GraphArgs A = base_args(ctx);
                A.pass = PASS_WIDE; A.kcap = INT_MAX; A.append = 1; A.symmetric = 0;
                rc = launch_tile(ctx, A, T, true, 2);
                if (rc) return rc;
            }
        }
    }
    // Foreign passes (pair-matrix algorithm only): after the MAIN passes of the 2-bit kernels have all been launched
    // -- this call launched none, or a single rank runs everything in one call.  Several ranks: one pass per call,
    // like the MAIN ladder (the driver MIN-reduces best[] in between and calls MAIN until no rows are left).
    if ((phases & ISOCON_PHASE_MAIN) && ctx->algo == ISOCON_ALGO_TILE && !ctx->h_flist.empty() &&
        (ctx->last_run_rows == 0 || ctx->prm.world <= 1)) {
        const int kcap = ctx->opt_kcap_main;
        while (ctx->foreign_level < 2) {
            const size_t nF = ctx->h_flist.size();
            if (ctx->foreign_level == 0) {
                CU(ctx->d_flist.ensure(nF + 1));
                rc = h2d(ctx, ctx->d_flist.p, ctx->h_flist.data(), nF * sizeof(int));
                if (rc) return rc;
            }
            GraphArgs A = base_args(ctx);
            A.pass = PASS_MAIN; A.append = 1;
            A.item_end = (long long)nF * ((ctx->n + 31) >> 5);
            A.item_begin = ctx->prm.world > 1 ? ctx->prm.rank : 0;
            A.item_stride = ctx->prm.world > 1 ? ctx->prm.world : 1;
            CU(cudaMemsetAsync(ctx->d_small.p + SM_COUNTER, 0, sizeof(unsigned long long), ctx->stream));
            nn_foreign_kernel<<<ctx->grid, WARPS_PER_BLOCK * 32, 0, ctx->stream>>>(
                A, ctx->d_flist.p, (int)nF, ctx->foreign_level == 0 ? kcap : INT_MAX, ctx->foreign_level == 0 ? -1 : kcap);
            CU(cudaGetLastError());
            ++ctx->launches;
            ++ctx->foreign_level;
            ctx->last_run_rows += (long long)nF;
            if (ctx->prm.world > 1) break;
        }
    }
    if (!final_sync) return ISOCON_OK;
    CU(cudaEventRecord(ctx->ev1, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->ms[1] += ms;
    for (int i = 0; i < ctx->kev_used; ++i) {
        float k_ms = 0.f;
        CU(cudaEventElapsedTime(&k_ms, ctx->kev[2 * i], ctx->kev[2 * i + 1]));
        ctx->ms[5] += k_ms;
        if (ctx->opt_debug >= 2) fprintf(stderr, "[isocon_nn]   pair-kernel launch %d of this call: %.3f ms\n", i, k_ms);
    }
    if (ctx->opt_debug) {
        const double host_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count();
        fprintf(stderr, "[isocon_nn] graph_run phases=%d rank=%d: host %.3f ms, device %.3f ms, pair kernels so far %.3f ms (%d launches), %zu bins\n",
                phases, ctx->prm.rank, host_ms, ms, ctx->ms[5], ctx->kev_used, ctx->bin_first.size());
    }
    ctx->kev_used = 0;
    return ISOCON_OK;
}

bool can_fuse(const isocon_nn_ctx* ctx) {
    return ctx->graph_open && ctx->prm.world > 1 && ctx->n_peers == ctx->prm.world - 1 && ctx->algo == ISOCON_ALGO_TILE &&
           ctx->n > 0 && ctx->opt_fuse;
}

}  // namespace

extern "C" {

int isocon_nn_can_fuse(isocon_nn_ctx* ctx, int32_t* yes) {
    if (!ctx || !yes) return ISOCON_ERR_ARG;
    *yes = can_fuse(ctx) ? 1 : 0;
    return ISOCON_OK;
}

int isocon_nn_graph_run(isocon_nn_ctx* ctx, int phases) {
    if (!ctx) return ISOCON_ERR_ARG;
    if (!ctx->graph_open) return fail(ctx, ISOCON_ERR_STATE, "graph_run: call graph_begin first");
    CU(cudaSetDevice(ctx->device));
    if (!(phases == ISOCON_PHASE_ALL && can_fuse(ctx))) return run_phases(ctx, phases, true);
    // Fused multi-rank flow (the ranks of one box, peers mapped): every phase in this one call.  The ranks meet at
    // device-side barriers over NVLink peer memory instead of collectives -- best[] and the nearest-pilot-row records
    // are already everywhere (the pair kernels push every improvement to all copies), so after a barrier all copies
    // are equal; agree_on_best keeps that state for the host decisions of the next phase.
    ctx->fused = true;
    const auto host_t0 = std::chrono::steady_clock::now();
    CU(cudaEventRecord(ctx->ev0, ctx->stream));
    int rc = enqueue_barrier(ctx);                       // every rank has initialised its best[] and counters
    long long rows = 0;
    if (!rc) rc = run_phases(ctx, ISOCON_PHASE_SEED, false);
    if (!rc && ctx->last_run_rows) { rows += ctx->last_run_rows; rc = agree_on_best(ctx); }
    if (!rc) rc = run_phases(ctx, ISOCON_PHASE_PILOT, false);
    if (!rc && ctx->last_run_rows) { rows += ctx->last_run_rows; if (!ctx->pilot_prefetched) rc = agree_on_best(ctx); }
    while (!rc) {                                        // one MAIN / foreign pass per round, like the collective driver
        rc = run_phases(ctx, ISOCON_PHASE_MAIN, false);
        if (rc || ctx->last_run_rows == 0) break;
        rows += ctx->last_run_rows;
        rc = agree_on_best(ctx);
    }
    if (!rc) rc = run_phases(ctx, ISOCON_PHASE_WIDE, false);
    if (rc) return rc;
    rows += ctx->last_run_rows;
    ctx->last_run_rows = rows;
    CU(cudaEventRecord(ctx->ev1, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->ms[1] += ms;
    for (int i = 0; i < ctx->kev_used; ++i) {
        float k_ms = 0.f;
        CU(cudaEventElapsedTime(&k_ms, ctx->kev[2 * i], ctx->kev[2 * i + 1]));
        ctx->ms[5] += k_ms;
    }
    ctx->kev_used = 0;
    if (ctx->opt_debug) {
        const double host_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count();
        fprintf(stderr, "[isocon_nn] fused graph_run rank=%d: host %.3f ms, device %.3f ms, pair kernels %.3f ms, %llu barriers so far\n",
                ctx->prm.rank, host_ms, ms, ctx->ms[5], ctx->bar_seq);
    }
    return ISOCON_OK;
}

int isocon_nn_last_run_rows(isocon_nn_ctx* ctx, int64_t* rows) {
    if (!ctx || !rows) return ISOCON_ERR_ARG;
    *rows = ctx->last_run_rows;
    return ISOCON_OK;
}

int isocon_nn_best_dev(isocon_nn_ctx* ctx, void** best_dev) {
    if (!ctx || !best_dev) return ISOCON_ERR_ARG;
    *best_dev = ctx->d_best.p;
    return ISOCON_OK;
}

int isocon_nn_best_agree(isocon_nn_ctx* ctx) {
    if (!ctx) return ISOCON_ERR_ARG;
    if (!ctx->graph_open) return fail(ctx, ISOCON_ERR_STATE, "best_agree: call graph_begin first");
    CU(cudaSetDevice(ctx->device));
    if (ctx->n) {
        CU(ctx->d_snap.ensure((size_t)ctx->n + 1));
        CU(cudaMemcpyAsync(ctx->d_snap.p, ctx->d_best.p, (size_t)ctx->n * sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    ctx->snap_valid = true;
    ctx->best_host_launches = ~0ull;
    return ISOCON_OK;
}

int isocon_nn_pilot_near_dev(isocon_nn_ctx* ctx, void** dev, int64_t* count) {
    if (!ctx || !dev || !count) return ISOCON_ERR_ARG;
    const bool on = ctx->graph_open && ctx->cluster_pilot && !ctx->main_done && ctx->pilot_rows > 0;
    *dev = on ? (void*)ctx->sv.pnear : nullptr;
    *count = on ? 2 * ctx->n : 0;
    return ISOCON_OK;
}

int isocon_nn_ipc_handles(isocon_nn_ctx* ctx, uint8_t handles[ISOCON_IPC_BYTES], uint64_t* generation) {
    if (!ctx || !handles || !generation) return ISOCON_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    static_assert(ISOCON_IPC_BYTES >= 3 * 64 + 16, "handle record too small");
    if (!ctx->d_best.p || !ctx->d_share.p) return fail(ctx, ISOCON_ERR_STATE, "ipc_handles: call set_list first");
    memset(handles, 0, ISOCON_IPC_BYTES);
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, ctx->d_best.p));
    memcpy(handles, &h, 64);
    CU(cudaIpcGetMemHandle(&h, ctx->d_small.p));
    memcpy(handles + 64, &h, 64);
    CU(cudaIpcGetMemHandle(&h, ctx->d_share.p));
    memcpy(handles + 128, &h, 64);
    const long long caps[2] = {ctx->share_n, ctx->share_f};     // the layout of the share block
    memcpy(handles + 192, caps, sizeof caps);
    *generation = ctx->best_generation;
    ctx->best_exported = true;
    return ISOCON_OK;
}

int isocon_nn_release_retired(isocon_nn_ctx* ctx) {
    if (!ctx) return ISOCON_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    for (void* p : ctx->retired_best) cudaFree(p);
    ctx->retired_best.clear();
    return ISOCON_OK;
}

int isocon_nn_set_peers(isocon_nn_ctx* ctx, const uint8_t* handles, int32_t world, int32_t rank) {
    if (!ctx) return ISOCON_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    close_peers(ctx);
    if (world <= 1) return ISOCON_OK;
    if (!handles || world > 8 || rank < 0 || rank >= world)
        return fail(ctx, ISOCON_ERR_ARG, "set_peers: need 2..8 ranks of one box (world %d, rank %d)", world, rank);
    if (!ctx->d_share.p) return fail(ctx, ISOCON_ERR_STATE, "set_peers: call set_list first");
    for (int r = 0; r < world; ++r) {
        if (r == rank) continue;
        void* pb = nullptr;
        void* ps = nullptr;
        void* pv = nullptr;
        const uint8_t* rec = handles + (size_t)ISOCON_IPC_BYTES * (size_t)r;
        cudaIpcMemHandle_t h;
        memcpy(&h, rec, 64);
        cudaError_t e = cudaIpcOpenMemHandle(&pb, h, cudaIpcMemLazyEnablePeerAccess);
        if (e == cudaSuccess) {
            memcpy(&h, rec + 64, 64);
            e = cudaIpcOpenMemHandle(&ps, h, cudaIpcMemLazyEnablePeerAccess);
        }
        if (e == cudaSuccess) {
            memcpy(&h, rec + 128, 64);
            e = cudaIpcOpenMemHandle(&pv, h, cudaIpcMemLazyEnablePeerAccess);
        }
        if (e != cudaSuccess) {
            if (pb) cudaIpcCloseMemHandle(pb);
            if (ps) cudaIpcCloseMemHandle(ps);
            close_peers(ctx);
            return fail(ctx, ISOCON_ERR_CUDA, "set_peers: cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e));
        }
        long long caps[2];
        memcpy(caps, rec + 192, sizeof caps);
        ctx->peer_best[ctx->n_peers] = (int*)pb;
        ctx->peer_small[ctx->n_peers] = (unsigned long long*)ps;
        ctx->peer_share[ctx->n_peers] = (uint8_t*)pv;
        ctx->peer_sv[ctx->n_peers] = share_view((uint8_t*)pv, caps[0], caps[1]);
        if (r == 0) ctx->root_small = (unsigned long long*)ps;
        ++ctx->n_peers;
    }
    // the barrier counters start over with the new mappings (the driver lets no rank go on before all are here)
    CU(cudaMemsetAsync(ctx->sv.ctrl, 0, CT_WORDS * sizeof(unsigned long long), ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->bar_seq = 0;
    if (rank == 0) {
        ctx->root_small = ctx->d_small.p;
        CU(cudaMemsetAsync(ctx->d_small.p + SM_QUEUE, 0, SM_NQUEUE * sizeof(unsigned long long), ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return ISOCON_OK;
}

int isocon_nn_graph_finalize(isocon_nn_ctx* ctx, int64_t* n_edges) {
    if (!ctx || !n_edges) return ISOCON_ERR_ARG;
    if (!ctx->graph_open) return fail(ctx, ISOCON_ERR_STATE, "graph_finalize: call graph_begin first");
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventRecord(ctx->ev0, ctx->stream));
    // tie filter: the kernel reads the number of candidates itself and delivers the survivors to this rank's share
    // block -- in a fused multi-rank run to every rank's, between two barriers: all pair kernels of the box are done
    // (best[] is final and the same everywhere), then all edges have arrived
    FilterArgs F{};
    F.n_dst = 1;
    F.dst[0] = FilterDst{ctx->sv.fq, ctx->sv.ft, ctx->sv.fd, ctx->sv.ctrl, ctx->sv.f_cap};
    if (ctx->fused) {
        for (int p = 0; p < ctx->n_peers; ++p)
            F.dst[F.n_dst++] = FilterDst{ctx->peer_sv[p].fq, ctx->peer_sv[p].ft, ctx->peer_sv[p].fd, ctx->peer_sv[p].ctrl, ctx->peer_sv[p].f_cap};
        int rc = enqueue_barrier(ctx);
        if (rc) return rc;
    }
    const unsigned grid = (unsigned)std::min<long long>((ctx->ecap + 255) / 256, 8ll * ctx->num_sms);
    filter_edges_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->d_eq.p, ctx->d_et.p, ctx->d_ed.p, ctx->d_small.p + SM_ECOUNT, ctx->ecap,
                                                       ctx->d_best.p, F);
    CU(cudaGetLastError());
    ++ctx->launches;
    if (ctx->fused) { int rc = enqueue_barrier(ctx); if (rc) return rc; }
    // one round trip for everything the caller will fetch: the counters, best[] and the first edges (an NN graph
    // has about one edge per read; the rest, if any, follows in graph_fetch)
    ctx->spec_edges = std::min<long long>(ctx->sv.f_cap, 2 * ctx->n + 1024);
    CU(ctx->fetch_host.ensure(((size_t)ctx->n + 3 * (size_t)ctx->spec_edges + 2 * (SM_WORDS + CT_WORDS)) * sizeof(int) + 256));
    unsigned long long* small = (unsigned long long*)ctx->fetch_host.p;
    unsigned long long* ctrl = small + SM_WORDS;
    int* hb = (int*)(ctrl + CT_WORDS);
    int* he = hb + ctx->n;
    CU(cudaMemcpyAsync(small, ctx->d_small.p, SM_WORDS * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(ctrl, ctx->sv.ctrl, CT_WORDS * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    if (ctx->n) CU(cudaMemcpyAsync(hb, ctx->d_best.p, (size_t)ctx->n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (ctx->spec_edges) {
        const size_t b = (size_t)ctx->spec_edges * sizeof(int);
        CU(cudaMemcpyAsync(he, ctx->sv.fq, b, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(he + ctx->spec_edges, ctx->sv.ft, b, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(he + 2 * ctx->spec_edges, ctx->sv.fd, b, cudaMemcpyDeviceToHost, ctx->stream));
    }
    // the box-wide tile queues start from zero in the next graph (collective mode: every rank has left its pair
    // kernels -- the driver reduced best[] since -- and the driver's edge gather orders this before any peer's next
    // launch; fused mode: the barrier above)
    CU(cudaMemsetAsync(ctx->d_small.p + SM_QUEUE, 0, SM_NQUEUE * sizeof(unsigned long long), ctx->stream));
    CU(cudaEventRecord(ctx->ev1, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaEventElapsedTime(&ctx->ms[2], ctx->ev0, ctx->ev1));
    const long long ne = (long long)small[SM_ECOUNT];
    ctx->stats.pairs = small[SM_STATS + ST_PAIRS];
    ctx->stats.word_columns = small[SM_STATS + ST_WORDCOLS];
    ctx->stats.groups = small[SM_STATS + ST_GROUPS];
    ctx->stats.wide_pairs = small[SM_STATS + ST_WIDE];
    ctx->stats.items = small[SM_STATS + ST_ITEMS];
    ctx->stats.useful_cells = small[SM_STATS + ST_CELLS];
    ctx->stats.columns = small[SM_STATS + ST_COLS];
    ctx->stats.edges_raw = (uint64_t)ne;
    ctx->stats.launches = ctx->launches;
    ctx->stats.pilot_rows = ctx->pilot_rows;
    if (ctrl[CT_ERR])
        return fail(ctx, ISOCON_ERR_CUDA, "a rank of the box did not reach a barrier within 5 s (did it fail?)");
    // overflow: the candidate buffer of this rank (collective mode) / of any rank (fused mode: CT_NEEDED came from
    // all), or more final edges than the share block holds.  Edges were dropped: reserve more, build the graph again.
    long long needed = std::max<long long>(ne > ctx->ecap ? ne : 0, (long long)ctrl[CT_NEEDED]);
    if ((long long)ctrl[CT_FCOUNT] > ctx->sv.f_cap) needed = std::max<long long>(needed, (long long)ctrl[CT_FCOUNT]);
    if (needed > 0) {
        ctx->stats.edges_raw = (uint64_t)needed;
        return fail(ctx, ISOCON_ERR_OVERFLOW, "candidate edge buffer overflow (%lld > %lld): call isocon_nn_reserve_edges and rebuild the graph", needed, ctx->ecap);
    }
    ctx->n_final = (long long)ctrl[CT_FCOUNT];
    ctx->finalized = true;
    *n_edges = ctx->n_final;
    return ISOCON_OK;
}

int isocon_nn_reserve_edges(isocon_nn_ctx* ctx, int64_t capacity) {
    if (!ctx) return ISOCON_ERR_ARG;
    ctx->edge_reserve = capacity;
    return ISOCON_OK;
}

int isocon_nn_graph_fetch(isocon_nn_ctx* ctx, int32_t* best, int32_t* eq, int32_t* et, int32_t* ed) {
    if (!ctx) return ISOCON_ERR_ARG;
    if (!ctx->finalized) return fail(ctx, ISOCON_ERR_STATE, "graph_fetch: call graph_finalize first");
    CU(cudaSetDevice(ctx->device));
    // graph_finalize already brought best[] and the first spec_edges edges to the host
    const int* hb = (const int*)((const unsigned long long*)ctx->fetch_host.p + SM_WORDS + CT_WORDS);
    const int* he = hb + ctx->n;
    if (best && ctx->n) memcpy(best, hb, (size_t)ctx->n * sizeof(int));
    const long long head = std::min(ctx->n_final, ctx->spec_edges), rest = ctx->n_final - head;
    if (head) {
        if (eq) memcpy(eq, he, (size_t)head * sizeof(int));
        if (et) memcpy(et, he + ctx->spec_edges, (size_t)head * sizeof(int));
        if (ed) memcpy(ed, he + 2 * ctx->spec_edges, (size_t)head * sizeof(int));
    }
    if (rest > 0) {
        const size_t b = (size_t)rest * sizeof(int);
        if (eq) CU(cudaMemcpyAsync(eq + head, ctx->sv.fq + head, b, cudaMemcpyDeviceToHost, ctx->stream));
        if (et) CU(cudaMemcpyAsync(et + head, ctx->sv.ft + head, b, cudaMemcpyDeviceToHost, ctx->stream));
        if (ed) CU(cudaMemcpyAsync(ed + head, ctx->sv.fd + head, b, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return ISOCON_OK;
}

int isocon_nn_edges_dev(isocon_nn_ctx* ctx, void** q_dev, void** t_dev, void** d_dev) {
    if (!ctx || !ctx->finalized) return ctx ? fail(ctx, ISOCON_ERR_STATE, "edges_dev: call graph_finalize first") : ISOCON_ERR_ARG;
    if (q_dev) *q_dev = ctx->sv.fq;
    if (t_dev) *t_dev = ctx->sv.ft;
    if (d_dev) *d_dev = ctx->sv.fd;
    return ISOCON_OK;
}

int isocon_nn_ed_pairs(isocon_nn_ctx* ctx, const int32_t* a, const int32_t* b, const int32_t* k, int64_t np, int32_t* out) {
    if (!ctx || np < 0 || (np > 0 && (!a || !b || !out))) return ctx ? fail(ctx, ISOCON_ERR_ARG, "ed_pairs: bad arguments") : ISOCON_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    if (np == 0) return ISOCON_OK;
    for (int64_t p = 0; p < np; ++p)
        if (a[p] < 0 || a[p] >= ctx->n || b[p] < 0 || b[p] >= ctx->n)
            return fail(ctx, ISOCON_ERR_ARG, "ed_pairs: pair %lld refers to read outside [0,%lld)", (long long)p, ctx->n);
    std::vector<long long> order((size_t)np);
    for (int64_t p = 0; p < np; ++p) order[p] = p;
    std::stable_sort(order.begin(), order.end(), [&](long long x, long long y) { return a[x] < a[y]; });
    std::vector<int> sa((size_t)np), sb((size_t)np), sk((size_t)np);
    std::vector<long long> runs;
    for (int64_t i = 0; i < np; ++i) {
        sa[i] = a[order[i]]; sb[i] = b[order[i]]; sk[i] = k ? k[order[i]] : -1;
        if (i == 0 || sa[i] != sa[i - 1]) runs.push_back(i);
    }
    runs.push_back(np);
    const long long n_runs = (long long)runs.size() - 1;
    CU(ctx->d_pa.ensure((size_t)np)); CU(ctx->d_pb.ensure((size_t)np)); CU(ctx->d_pk.ensure((size_t)np));
    CU(ctx->d_pout.ensure((size_t)np)); CU(ctx->d_runoff.ensure(runs.size()));
    CU(cudaEventRecord(ctx->ev0, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_pa.p, sa.data(), (size_t)np * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_pb.p, sb.data(), (size_t)np * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_pk.p, sk.data(), (size_t)np * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_runoff.p, runs.data(), runs.size() * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(ctx->d_small.p + SM_COUNTER, 0, sizeof(unsigned long long), ctx->stream));
    GraphArgs A = base_args(ctx);
    ed_pairs_kernel<<<ctx->grid, WARPS_PER_BLOCK * 32, ctx->smem, ctx->stream>>>(A, ctx->d_pa.p, ctx->d_pb.p, ctx->d_pk.p,
                                                                              ctx->d_runoff.p, n_runs, ctx->d_pout.p);
    CU(cudaGetLastError());
    std::vector<int> so((size_t)np);
    CU(cudaMemcpyAsync(so.data(), ctx->d_pout.p, (size_t)np * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaEventRecord(ctx->ev1, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaEventElapsedTime(&ctx->ms[3], ctx->ev0, ctx->ev1));
    for (int64_t i = 0; i < np; ++i) out[order[i]] = so[i];
    return ISOCON_OK;
}

int isocon_nn_get_stats(isocon_nn_ctx* ctx, isocon_nn_stats* out) {
    if (!ctx || !out) return ISOCON_ERR_ARG;
    *out = ctx->stats;
    return ISOCON_OK;
}

int isocon_nn_last_ms(isocon_nn_ctx* ctx, int which, float* ms) {
    if (!ctx || !ms || which < 0 || which > 5) return ISOCON_ERR_ARG;
    *ms = ctx->ms[which];
    return ISOCON_OK;
}

int isocon_nn_int32_peak(isocon_nn_ctx* ctx, double* lane_ops_per_s) {
    if (!ctx || !lane_ops_per_s) return ISOCON_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    const int blocks = ctx->num_sms * 8, threads = 256, iters = 20000;
    DBuf<uint32_t> outbuf;
    CU(outbuf.ensure((size_t)blocks * threads));
    float best_ms = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CU(cudaEventRecord(ctx->ev0, ctx->stream));
        int32_probe_kernel<<<blocks, threads, 0, ctx->stream>>>(outbuf.p, iters, 12345u + rep);
        CU(cudaGetLastError());
        CU(cudaEventRecord(ctx->ev1, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        if (rep > 0) best_ms = std::min(best_ms, ms);
    }
    outbuf.release();
    ctx->ms[4] = best_ms;
    *lane_ops_per_s = (double)blocks * threads * (double)iters * PROBE_OPS_PER_ITER / (best_ms * 1e-3);
    return ISOCON_OK;
}

}  // extern "C"
