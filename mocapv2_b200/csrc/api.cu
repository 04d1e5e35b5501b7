// C-ABI glue of the detection path: workspace carving and the mocap_detect_batch / mocap_filter_batch entry points
// (include/mocap_b200.h).  _find_dot of the reference (lib/ImageOperations.py:33-78) = launch_filter + launch_blobs.
#include "common.cuh"
#include <string.h>

// detect_filter.cu
int table_view(const void* table_dev, int H, int W, TableView* tv);
int launch_scan(const uint8_t* frames, int n, int H, int W, int64_t fstride, const TableView& tv, int thresh,
                const FilterWs& ws, cudaStream_t s, StageTimer* timer);
int launch_tiles(const uint8_t* frames, int n, int H, int W, int64_t fstride, const TableView& tv, int thresh,
                 const FilterWs& ws, int max_fg, int* flags, const int* need_general, cudaStream_t s);
// detect_cluster.cu
size_t cluster_ws_bytes(int n, int H, int W, int max_contours, size_t* offs);
bool cluster_path_supported(int H, int W);
int launch_cluster_path(const uint8_t* frames, int n, int H, int W, int64_t fstride, const TableView& tv, int thresh,
                        const uint32_t* cellbox, char* ws_base, const size_t* offs, int* need_general,
                        int max_contours, int max_blobs, double min_area, double min_circ,
                        int32_t* out_xy, int32_t* out_count, int32_t* out_flags, double* out_contours, int32_t* out_contour_count,
                        cudaStream_t s, StageTimer* timer, const ClusterLaunch& how);
int launch_materialize_bits(const FilterWs& ws, int n, int H, int W, const TableView& tv, uint32_t* out, cudaStream_t s);
// detect_scan_tma.cu
bool scan_tma_supported(const uint8_t* frames, int n, int H, int W, int64_t fstride, int thresh);
size_t scan_tma_ctrl_bytes(int chunks);
int launch_scan_tma(const uint8_t* frames, int n, int H, int W, int64_t fstride, const TableView& tv, int thresh, uint32_t* cellbox,
                    int* ctrl, int chunks, int chunk_frames, int widx, int item_begin, int item_end, int stages, cudaStream_t s);
int stream_wait_geq(cudaStream_t s, const int* addr_dev, int value);
bool stream_wait_supported();
// detect_blobs.cu
int launch_tiles_from_bits(const uint32_t* bits, int n, int H, int TX, int TY, uint32_t* fg_tiles, int* n_fg, int max_fg, cudaStream_t s);
size_t blob_ws_stride(int H, int max_runs, int max_contours);
int launch_blobs(const uint32_t* bits, const uint32_t* fg_tiles, const int* n_fg, int n, int H, int W, int TX,
                 int max_fg, int max_runs, int max_blobs, int max_contours, double min_area, double min_circ,
                 char* ws, size_t ws_stride,
                 int32_t* out_xy, int32_t* out_count, int32_t* out_flags,
                 int64_t* out_blob_sums, int32_t* out_blob_count, double* out_contours, int32_t* out_contour_count,
                 int32_t* out_labels, const int* need_general, cudaStream_t s);

extern "C" const char* mocap_status_string(int status)
{
    switch (status) {
        case MOCAP_OK: return "ok";
        case MOCAP_ERR_INVALID: return "invalid argument";
        case MOCAP_ERR_WORKSPACE: return "workspace too small";
        case MOCAP_ERR_CUDA: return "CUDA runtime error";
        case MOCAP_ERR_UNSUPPORTED: return "shape outside the supported limits";
        default: return "unknown status";
    }
}

extern "C" int mocap_abi_version(void) { return MOCAP_ABI_VERSION; }

unsigned long long g_mocap_launches = 0;
extern "C" unsigned long long mocap_kernel_launch_count(void) { return __atomic_load_n(&g_mocap_launches, __ATOMIC_RELAXED); }

extern "C" const char* mocap_stage_name(int stage)
{
    static const char* names[MOCAP_N_STAGES] = {"scan", "group", "filter", "borders", "finish"};
    return (stage >= 0 && stage < MOCAP_N_STAGES) ? names[stage] : "?";
}

extern "C" void* mocap_stage_timer_create(void)
{
    StageTimer* t = new StageTimer();
    for (int i = 0; i < 2 * MOCAP_N_STAGES; ++i)
        if (cudaEventCreate(&t->ev[i]) != cudaSuccess) { delete t; return nullptr; }
    for (int i = 0; i < MOCAP_N_STAGES; ++i) t->recorded[i] = 0;
    return t;
}

extern "C" void mocap_stage_timer_destroy(void* timer)
{
    StageTimer* t = (StageTimer*)timer;
    if (!t) return;
    for (int i = 0; i < 2 * MOCAP_N_STAGES; ++i) cudaEventDestroy(t->ev[i]);
    delete t;
}

extern "C" int mocap_stage_timer_read(void* timer, float* ms_out)
{
    StageTimer* t = (StageTimer*)timer;
    if (!t || !ms_out) return MOCAP_ERR_INVALID;
    for (int i = 0; i < MOCAP_N_STAGES; ++i) {
        ms_out[i] = -1.0f;
        if (!t->recorded[i]) continue;
        CUDA_TRY(cudaEventSynchronize(t->ev[2 * i + 1]));
        CUDA_TRY(cudaEventElapsedTime(&ms_out[i], t->ev[2 * i], t->ev[2 * i + 1]));
    }
    return MOCAP_OK;
}

struct DetectLayout {
    size_t off_active, off_list, off_counters, off_bits, off_fg, off_nfg, off_flags, off_cellbox, off_blob, off_cluster;
    size_t blob_stride, total;
    size_t cl_offs[16];
    int TX, TY, TXW, max_fg;
};

static int detect_layout(int n, int H, int W, int max_contours, int max_runs, bool with_blobs, DetectLayout* L)
{
    if (n <= 0 || H <= 0 || W <= 0) return MOCAP_ERR_INVALID;
    if (H > 16384 || W > 16384) return MOCAP_ERR_UNSUPPORTED;
    L->TX = cdiv(W, TILE); L->TY = cdiv(H, TILE); L->TXW = cdiv(L->TX, 32);
    L->max_fg = L->TX * L->TY;
    if ((long long)n * L->TX * L->TY >= (1LL << 31)) return MOCAP_ERR_UNSUPPORTED;   // tile codes are 32-bit
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t r = off; off += align_up(bytes, 256); return r; };
    L->off_active = take((size_t)n * L->TY * L->TXW * 4);
    L->off_list = take((size_t)n * L->TX * L->TY * 4);
    L->off_counters = take(64);
    L->off_bits = take((size_t)n * H * L->TX * 4);
    L->off_fg = take((size_t)n * L->max_fg * 4);
    L->off_nfg = take((size_t)n * 4);
    L->off_flags = take((size_t)n * 4);
    L->off_cellbox = take((size_t)n * L->TX * L->TY * 4);
    L->blob_stride = with_blobs ? blob_ws_stride(H, max_runs, max_contours) : 0;
    L->off_blob = take(L->blob_stride * (size_t)n);
    L->off_cluster = take(cluster_ws_bytes(n, H, W, max_contours, L->cl_offs));
    L->total = off;
    return MOCAP_OK;
}

static FilterWs filter_ws(char* base, const DetectLayout& L)
{
    FilterWs ws;
    ws.active = (uint32_t*)(base + L.off_active);
    ws.list = (uint32_t*)(base + L.off_list);
    ws.counters = (int*)(base + L.off_counters);
    ws.bits = (uint32_t*)(base + L.off_bits);
    ws.fg_tiles = (uint32_t*)(base + L.off_fg);
    ws.n_fg = (int*)(base + L.off_nfg);
    ws.cellbox = (uint32_t*)(base + L.off_cellbox);
    return ws;
}

extern "C" size_t mocap_detect_workspace_bytes(int n_frames, int H, int W, int max_blobs, int max_contours, int max_runs)
{
    (void)max_blobs;
    DetectLayout L;
    if (max_contours <= 0 || max_runs <= 0) return 0;
    if (detect_layout(n_frames, H, W, max_contours, max_runs, true, &L) != MOCAP_OK) return 0;
    return L.total;
}

extern "C" int mocap_filter_batch(const uint8_t* frames_dev, int n_frames, int H, int W, int64_t frame_stride,
                                  const void* table_dev, int thresh, uint32_t* out_bits,
                                  void* workspace, size_t workspace_bytes, void* stream)
{
    if (!frames_dev || !table_dev || !out_bits || !workspace) return MOCAP_ERR_INVALID;
    if (frame_stride < (int64_t)H * W) return MOCAP_ERR_INVALID;
    DetectLayout L;
    int st = detect_layout(n_frames, H, W, 1, 1, false, &L);
    if (st != MOCAP_OK) return st;
    if (workspace_bytes < L.total) return MOCAP_ERR_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    TableView tv; table_view(table_dev, H, W, &tv);
    FilterWs ws = filter_ws((char*)workspace, L);
    int* flags = (int*)((char*)workspace + L.off_flags);
    CUDA_TRY(cudaMemsetAsync(flags, 0, (size_t)n_frames * 4, s));
    st = launch_scan(frames_dev, n_frames, H, W, frame_stride, tv, thresh, ws, s, nullptr);
    if (st != MOCAP_OK) return st;
    st = launch_tiles(frames_dev, n_frames, H, W, frame_stride, tv, thresh, ws, L.max_fg, flags, nullptr, s);
    if (st != MOCAP_OK) return st;
    return launch_materialize_bits(ws, n_frames, H, W, tv, out_bits, s);
}

extern "C" int mocap_detect_batch(const uint8_t* frames_dev, int n_frames, int H, int W, int64_t frame_stride,
                                  const void* table_dev, int thresh, double min_area, double min_circ,
                                  int max_blobs, int max_contours, int max_runs,
                                  int32_t* out_xy, int32_t* out_count, int32_t* out_flags,
                                  uint32_t* out_bits, int32_t* out_labels, int64_t* out_blob_sums, int32_t* out_blob_count,
                                  double* out_contours, int32_t* out_contour_count,
                                  void* workspace, size_t workspace_bytes, void* stream, void* stage_timer)
{
    StageTimer* timer = (StageTimer*)stage_timer;
    if (!frames_dev || !table_dev || !out_xy || !out_count || !out_flags || !workspace) return MOCAP_ERR_INVALID;
    if (max_blobs <= 0 || max_contours <= 0 || max_runs <= 0) return MOCAP_ERR_INVALID;
    if (frame_stride < (int64_t)H * W) return MOCAP_ERR_INVALID;
    DetectLayout L;
    int st = detect_layout(n_frames, H, W, max_contours, max_runs, true, &L);
    if (st != MOCAP_OK) return st;
    if (workspace_bytes < L.total) return MOCAP_ERR_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    TableView tv; table_view(table_dev, H, W, &tv);
    FilterWs ws = filter_ws((char*)workspace, L);
    CUDA_TRY(cudaMemsetAsync(out_flags, 0, (size_t)n_frames * 4, s));
    if (out_labels) CUDA_TRY(cudaMemsetAsync(out_labels, 0, (size_t)n_frames * H * W * 4, s));
    // Product path (no label / bit-image outputs wanted): scan -> groups of hot cells -> per-cluster units in shared memory
    // -> per-frame ordering.  Frames it cannot finish locally, and every frame when the parity outputs are requested,
    // take the general per-frame path (tiles -> packed image -> runs, labels, contour tree).
    const bool extras = out_bits || out_labels || out_blob_sums || out_blob_count;
    const bool use_cluster = !extras && cluster_path_supported(H, W);
    char* cl_base = (char*)workspace + L.off_cluster;
    int* need_general = (int*)(cl_base + L.cl_offs[0]);
    st = launch_scan(frames_dev, n_frames, H, W, frame_stride, tv, thresh, ws, s, timer);
    if (st != MOCAP_OK) return st;
    CUDA_TRY(cudaMemsetAsync(need_general, use_cluster ? 0 : 1, (size_t)n_frames * 4, s));
    if (use_cluster) {
        st = launch_cluster_path(frames_dev, n_frames, H, W, frame_stride, tv, thresh, ws.cellbox, cl_base, L.cl_offs, need_general,
                                 max_contours, max_blobs, min_area, min_circ, out_xy, out_count, out_flags, out_contours,
                                 out_contour_count, s, timer, ClusterLaunch());
        if (st != MOCAP_OK) return st;
    }
    stage_begin(timer, 4, s);
    st = launch_tiles(frames_dev, n_frames, H, W, frame_stride, tv, thresh, ws, L.max_fg, out_flags, need_general, s);
    if (st != MOCAP_OK) return st;
    if (out_bits) {
        st = launch_materialize_bits(ws, n_frames, H, W, tv, out_bits, s);
        if (st != MOCAP_OK) return st;
    }
    st = launch_blobs(ws.bits, ws.fg_tiles, ws.n_fg, n_frames, H, W, L.TX, L.max_fg, max_runs, max_blobs, max_contours,
                      min_area, min_circ, (char*)workspace + L.off_blob, L.blob_stride,
                      out_xy, out_count, out_flags, out_blob_sums, out_blob_count, out_contours, out_contour_count,
                      out_labels, need_general, s);
    stage_end(timer, 4, s);
    return st;
}

// findContours -> filter -> moments on a given packed binary image (lib/ImageOperations.py:41-65), stage parity entry
extern "C" int mocap_blobs_batch(const uint32_t* bits_dev, int n_frames, int H, int W,
                                 double min_area, double min_circ, int max_blobs, int max_contours, int max_runs,
                                 int32_t* out_xy, int32_t* out_count, int32_t* out_flags,
                                 int32_t* out_labels, int64_t* out_blob_sums, int32_t* out_blob_count,
                                 double* out_contours, int32_t* out_contour_count,
                                 void* workspace, size_t workspace_bytes, void* stream)
{
    if (!bits_dev || !out_xy || !out_count || !out_flags || !workspace) return MOCAP_ERR_INVALID;
    if (max_blobs <= 0 || max_contours <= 0 || max_runs <= 0) return MOCAP_ERR_INVALID;
    DetectLayout L;
    int st = detect_layout(n_frames, H, W, max_contours, max_runs, true, &L);
    if (st != MOCAP_OK) return st;
    if (workspace_bytes < L.total) return MOCAP_ERR_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    FilterWs ws = filter_ws((char*)workspace, L);
    CUDA_TRY(cudaMemsetAsync(out_flags, 0, (size_t)n_frames * 4, s));
    if (out_labels) CUDA_TRY(cudaMemsetAsync(out_labels, 0, (size_t)n_frames * H * W * 4, s));
    st = launch_tiles_from_bits(bits_dev, n_frames, H, L.TX, L.TY, ws.fg_tiles, ws.n_fg, L.max_fg, s);
    if (st != MOCAP_OK) return st;
    return launch_blobs(bits_dev, ws.fg_tiles, ws.n_fg, n_frames, H, W, L.TX, L.max_fg, max_runs, max_blobs, max_contours,
                        min_area, min_circ, (char*)workspace + L.off_blob, L.blob_stride,
                        out_xy, out_count, out_flags, out_blob_sums, out_blob_count, out_contours, out_contour_count,
                        out_labels, nullptr, s);
}


// =========================================================================================================
// Overlapped detection: the batch is cut into chunks of frames; the HBM-bound streaming scan of chunk k+1 runs
// beside the instruction-bound stages (group / piece filter / borders) of chunk k on the same SMs.
//   sync_mode 1: ONE TMA-fed scan kernel over the whole batch on its own (high-priority) stream; it publishes a flag per
//                finished chunk and the chunk's stages, on a worker stream, wait for it with a stream memory operation
//                (cuStreamWaitValue32) -- the scan CTA of an SM stays resident from the first byte to the last.
//   sync_mode 0: one scan launch per chunk + CUDA events (same kernels; the fallback when stream memory operations are
//                not available, and the A/B for the design note).
// Whatever the cluster path hands to the general path is finished for the whole batch after the join.
// =========================================================================================================
#define PIPE_MAX_PROC 6
#define PIPE_CTRL_WORK 64            // == SCAN_CTRL_WORK of detect_scan_tma.cu: work-counter slots ahead of chunk_done[] / chunk_flag[]
#define PIPE_MAX_CHUNKS 64
#define PIPE_TL_PER_CHUNK 4          // timeline marks per chunk: scan seen, grouped, filtered, borders done

// Store-to-peer epilogue of the multi-GPU pipeline: every frame's record (count + centroids) is copied to the address its consumer
// reads it from -- for the frame-sets of another rank that is a slot of THAT rank's receive buffer, mapped into this process
// (symmetric memory over NVLink), so the exchange step needs no collective, only a barrier.  One warp per frame.  phase 0 (after a
// chunk's border stage): frames the cluster path finished; phase 1 (after the general path): the frames it handed over.
__global__ void scatter_records_kernel(const int32_t* __restrict__ xy, const int32_t* __restrict__ count, const unsigned long long* __restrict__ xy_dst,
                                       const unsigned long long* __restrict__ count_dst, int n, int max_blobs, const int* __restrict__ need_general, int phase)
{
    const int lane = threadIdx.x & 31;
    const int f = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (f >= n) return;
    if ((need_general[f] != 0) != (phase != 0)) return;
    int c = count[f];
    c = c < 0 ? 0 : (c > max_blobs ? max_blobs : c);
    const uint2* src = (const uint2*)(xy + (size_t)f * max_blobs * 2);
    uint2* dst = (uint2*)xy_dst[f];
    for (int k = lane; k < c; k += 32) dst[k] = src[k];
    if (lane == 0) *(int32_t*)count_dst[f] = count[f];
}

static int launch_scatter(const int32_t* xy, const int32_t* count, const unsigned long long* xy_dst, const unsigned long long* count_dst,
                          int n, int max_blobs, const int* need_general, int phase, cudaStream_t s)
{
    LAUNCH(scatter_records_kernel, cdiv(n * 32, 128), 128, 0, s, xy, count, xy_dst, count_dst, n, max_blobs, need_general, phase);
    CUDA_TRY(cudaGetLastError());
    return MOCAP_OK;
}

struct DetectPipe {
    const unsigned long long* sc_xy = nullptr;      // per-frame destinations of the store-to-peer epilogue (device arrays [n]), or null
    const unsigned long long* sc_count = nullptr;
    void* tok_wait = nullptr;                       // scan token (mocap_detect_pipe_set_scan_token): cudaEvent_t the scan waits for / records
    void* tok_done = nullptr;
    const uint32_t* pre_cellbox = nullptr;          // hot cell boxes of the next calls' frames, computed by the caller (mocap_detect_pipe_set_cellbox)
#ifndef MOCAP_EMU
    cudaStream_t s_scan = nullptr, s_proc[PIPE_MAX_PROC] = {};
    cudaEvent_t ev_fork = nullptr, ev_scan_done = nullptr, ev_proc_done[PIPE_MAX_PROC] = {}, ev_join = nullptr;
    cudaEvent_t ev_scan[PIPE_MAX_CHUNKS] = {};                       // sync_mode 0
    cudaEvent_t ev_tl[PIPE_MAX_CHUNKS][PIPE_TL_PER_CHUNK] = {};      // timeline (timing enabled)
#endif
    int n_proc = 0;
    int tl_chunks = 0;               // chunks of the last call that recorded a timeline
    int last_scan = -1, last_chunks = 0, last_mode = -1;
};

extern "C" void* mocap_detect_pipe_create(int n_proc_streams, int prio_mode)
{
    if (n_proc_streams < 1) n_proc_streams = 1;
    if (n_proc_streams > PIPE_MAX_PROC) n_proc_streams = PIPE_MAX_PROC;
    DetectPipe* p = new DetectPipe();
    p->n_proc = n_proc_streams;
#ifndef MOCAP_EMU
    int lo = 0, hi = 0;
    bool ok = cudaDeviceGetStreamPriorityRange(&lo, &hi) == cudaSuccess;      // hi = numerically smallest = highest priority
    ok = ok && cudaStreamCreateWithPriority(&p->s_scan, cudaStreamNonBlocking, hi) == cudaSuccess;
    for (int i = 0; ok && i < n_proc_streams; ++i) {
        // prio_mode 0: every worker at the lowest priority.  1 / 2 (stage plan: worker 0 groups, worker 1 filters, the others take
        // the borders): 1 = earlier stages first (group > filter > borders), 2 = later stages first (borders > filter > group).
        int pr = lo;
        const int stage = i < 2 ? i : 2;
        if (prio_mode == 1) pr = hi + 1 + stage;
        if (prio_mode == 2) pr = hi + 3 - stage;
        if (pr > lo) pr = lo;
        ok = cudaStreamCreateWithPriority(&p->s_proc[i], cudaStreamNonBlocking, pr) == cudaSuccess;
    }
    ok = ok && cudaEventCreateWithFlags(&p->ev_fork, cudaEventDefault) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&p->ev_scan_done, cudaEventDefault) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&p->ev_join, cudaEventDefault) == cudaSuccess;
    for (int i = 0; ok && i < n_proc_streams; ++i) ok = cudaEventCreateWithFlags(&p->ev_proc_done[i], cudaEventDisableTiming) == cudaSuccess;
    for (int c = 0; ok && c < PIPE_MAX_CHUNKS; ++c) {
        ok = cudaEventCreateWithFlags(&p->ev_scan[c], cudaEventDisableTiming) == cudaSuccess;
        for (int k = 0; ok && k < PIPE_TL_PER_CHUNK; ++k) ok = cudaEventCreateWithFlags(&p->ev_tl[c][k], cudaEventDefault) == cudaSuccess;
    }
    if (!ok) { delete p; return nullptr; }
#endif
    return p;
}

extern "C" int mocap_detect_pipe_set_scatter(void* pipe, const uint64_t* xy_dst_dev, const uint64_t* count_dst_dev)
{
    DetectPipe* p = (DetectPipe*)pipe;
    if (!p || ((xy_dst_dev == nullptr) != (count_dst_dev == nullptr))) return MOCAP_ERR_INVALID;
    p->sc_xy = (const unsigned long long*)xy_dst_dev;
    p->sc_count = (const unsigned long long*)count_dst_dev;
    return MOCAP_OK;
}

extern "C" int mocap_detect_pipe_set_scan_token(void* pipe, void* wait_event, void* done_event)
{
    DetectPipe* p = (DetectPipe*)pipe;
    if (!p) return MOCAP_ERR_INVALID;
    p->tok_wait = wait_event;
    p->tok_done = done_event;
    return MOCAP_OK;
}

extern "C" int mocap_detect_pipe_set_cellbox(void* pipe, const uint32_t* cellbox_dev)
{
    DetectPipe* p = (DetectPipe*)pipe;
    if (!p) return MOCAP_ERR_INVALID;
    p->pre_cellbox = cellbox_dev;
    return MOCAP_OK;
}

extern "C" void mocap_detect_pipe_destroy(void* pipe)
{
    DetectPipe* p = (DetectPipe*)pipe;
    if (!p) return;
#ifndef MOCAP_EMU
    if (p->s_scan) cudaStreamDestroy(p->s_scan);
    for (int i = 0; i < PIPE_MAX_PROC; ++i) { if (p->s_proc[i]) cudaStreamDestroy(p->s_proc[i]); if (p->ev_proc_done[i]) cudaEventDestroy(p->ev_proc_done[i]); }
    if (p->ev_fork) cudaEventDestroy(p->ev_fork);
    if (p->ev_scan_done) cudaEventDestroy(p->ev_scan_done);
    if (p->ev_join) cudaEventDestroy(p->ev_join);
    for (int c = 0; c < PIPE_MAX_CHUNKS; ++c) {
        if (p->ev_scan[c]) cudaEventDestroy(p->ev_scan[c]);
        for (int k = 0; k < PIPE_TL_PER_CHUNK; ++k) if (p->ev_tl[c][k]) cudaEventDestroy(p->ev_tl[c][k]);
    }
#endif
    delete p;
}

struct PipeLayout {
    DetectLayout L;                  // general-path arrays of the whole batch (its cluster part serves chunk 0 .. see below)
    size_t off_ctrl, off_need, off_chunks, chunk_bytes, total;
    size_t cl_offs[16];              // offsets inside one chunk's cluster workspace
    int chunks, chunk_frames;
};

static int pipe_layout(int n, int H, int W, int max_contours, int max_runs, int chunk_frames, PipeLayout* P)
{
    if (chunk_frames <= 0 || chunk_frames > n) chunk_frames = n;
    int chunks = cdiv(n, chunk_frames);
    if (chunks > PIPE_MAX_CHUNKS) { chunk_frames = cdiv(n, PIPE_MAX_CHUNKS); chunks = cdiv(n, chunk_frames); }
    P->chunks = chunks; P->chunk_frames = chunk_frames;
    int st = detect_layout(n, H, W, max_contours, max_runs, true, &P->L);
    if (st != MOCAP_OK) return st;
    size_t off = P->L.off_cluster;                                  // the whole-batch cluster workspace is replaced by the per-chunk ones
    auto take = [&](size_t bytes) { size_t r = off; off += align_up(bytes, 256); return r; };
    P->off_ctrl = take(scan_tma_ctrl_bytes(chunks));
    P->off_need = take((size_t)n * 4);
    P->chunk_bytes = align_up(cluster_ws_bytes(chunk_frames, H, W, max_contours, P->cl_offs), 256);
    P->off_chunks = take(P->chunk_bytes * (size_t)chunks);
    P->total = off > P->L.total ? off : P->L.total;
    return MOCAP_OK;
}

extern "C" size_t mocap_detect_pipelined_workspace_bytes(int n_frames, int H, int W, int max_blobs, int max_contours, int max_runs, int chunk_frames)
{
    (void)max_blobs;
    if (max_contours <= 0 || max_runs <= 0) return 0;
    PipeLayout P;
    if (pipe_layout(n_frames, H, W, max_contours, max_runs, chunk_frames, &P) != MOCAP_OK) return 0;
    return P.total;
}

extern "C" int mocap_detect_batch_pipelined(void* pipe, const uint8_t* frames_dev, int n_frames, int H, int W, int64_t frame_stride,
                                            const void* table_dev, int thresh, double min_area, double min_circ,
                                            int max_blobs, int max_contours, int max_runs,
                                            int32_t* out_xy, int32_t* out_count, int32_t* out_flags,
                                            double* out_contours, int32_t* out_contour_count,
                                            void* workspace, size_t workspace_bytes, void* stream, const MocapPipeOpts* opts)
{
    DetectPipe* dp = (DetectPipe*)pipe;
    if (!dp || !frames_dev || !table_dev || !out_xy || !out_count || !out_flags || !workspace || !opts) return MOCAP_ERR_INVALID;
    if (max_blobs <= 0 || max_contours <= 0 || max_runs <= 0) return MOCAP_ERR_INVALID;
    if (frame_stride < (int64_t)H * W) return MOCAP_ERR_INVALID;
    PipeLayout P;
    int st = pipe_layout(n_frames, H, W, max_contours, max_runs, opts->chunk_frames, &P);
    if (st != MOCAP_OK) return st;
    if (workspace_bytes < P.total) return MOCAP_ERR_WORKSPACE;
    if (!cluster_path_supported(H, W)) return MOCAP_ERR_UNSUPPORTED;         // such shapes take mocap_detect_batch
    cudaStream_t s = (cudaStream_t)stream;
    TableView tv; table_view(table_dev, H, W, &tv);
    char* base = (char*)workspace;
    FilterWs ws = filter_ws(base, P.L);
    // hot cell boxes handed in (the Bayer front step computes them while it writes the grey frames): no streaming scan in this call
    const bool prescanned = dp->pre_cellbox != nullptr;
    if (prescanned) ws.cellbox = (uint32_t*)dp->pre_cellbox;
    int* ctrl = (int*)(base + P.off_ctrl);
    int* need_general = (int*)(base + P.off_need);
    const int chunks = P.chunks, cf = P.chunk_frames;
    const int per_frame_items = tv.TY * cdiv(W, 256);
    const bool tma = !prescanned && opts->scan_variant != 0 && scan_tma_supported(frames_dev, n_frames, H, W, frame_stride, thresh);
    int mode = opts->sync_mode;
    if (mode == 1 && !(tma && stream_wait_supported())) mode = 0;
    dp->last_scan = tma ? 1 : 0; dp->last_chunks = chunks; dp->last_mode = mode;
    dp->tl_chunks = 0;
    ClusterLaunch how;
    how.zero = 1;
    how.filter_ctas_per_sm = opts->filter_ctas_per_sm > 0 ? opts->filter_ctas_per_sm : (tma ? 6 : 8);
    how.cand_ctas_per_sm = opts->cand_ctas_per_sm > 0 ? opts->cand_ctas_per_sm : 12;
    CUDA_TRY(cudaMemsetAsync(out_flags, 0, (size_t)n_frames * 4, s));
    CUDA_TRY(cudaMemsetAsync(need_general, 0, (size_t)n_frames * 4, s));
    CUDA_TRY(cudaMemsetAsync(ctrl, 0, scan_tma_ctrl_bytes(chunks), s));
#ifdef MOCAP_EMU
    // CPU emulation build (tests/emu): no streams; the chunks run one after the other through the same launchers
    for (int c = 0; c < chunks; ++c) {
        const int f0 = c * cf, nc = (n_frames - f0) < cf ? (n_frames - f0) : cf;
        FilterWs wc = ws; wc.cellbox = ws.cellbox + (size_t)f0 * tv.TX * tv.TY;
        st = prescanned ? MOCAP_OK : launch_scan(frames_dev + (size_t)f0 * frame_stride, nc, H, W, frame_stride, tv, thresh, wc, s, nullptr);
        if (st != MOCAP_OK) return st;
        st = launch_cluster_path(frames_dev + (size_t)f0 * frame_stride, nc, H, W, frame_stride, tv, thresh, wc.cellbox,
                                 base + P.off_chunks + (size_t)c * P.chunk_bytes, P.cl_offs, need_general + f0, max_contours, max_blobs, min_area, min_circ,
                                 out_xy + (size_t)f0 * max_blobs * 2, out_count + f0, out_flags + f0,
                                 out_contours ? out_contours + (size_t)f0 * max_contours * 8 : nullptr, out_contour_count ? out_contour_count + f0 : nullptr,
                                 s, nullptr, how);
        if (st != MOCAP_OK) return st;
        if (dp->sc_xy) {
            st = launch_scatter(out_xy + (size_t)f0 * max_blobs * 2, out_count + f0, dp->sc_xy + f0, dp->sc_count + f0, nc, max_blobs, need_general + f0, 0, s);
            if (st != MOCAP_OK) return st;
        }
    }
#else
    const bool tl = opts->record_timeline != 0;
    CUDA_TRY(cudaEventRecord(dp->ev_fork, s));
    CUDA_TRY(cudaStreamWaitEvent(dp->s_scan, dp->ev_fork, 0));
    if (dp->tok_wait) CUDA_TRY(cudaStreamWaitEvent(dp->s_scan, (cudaEvent_t)dp->tok_wait, 0));    // scans of several pipes one after the other
    if (mode == 1) {
        st = launch_scan_tma(frames_dev, n_frames, H, W, frame_stride, tv, thresh, ws.cellbox, ctrl, chunks, cf, 0, 0, -1, opts->scan_stages, dp->s_scan);
        if (st != MOCAP_OK) return st;
    }
    // Stream plans (which worker stream runs which stage of which chunk):
    //   0  a chunk's stages as one chain on worker (chunk mod workers)
    //   1  stage streams: worker 0 groups, worker 1 filters, workers 2.. take the borders of alternate chunks
    //   2  worker 0 groups and filters chunk after chunk; workers 1.. take the borders (they fill the filter kernels' tails)
    //   3  like 2, grouping one chunk ahead of the filter (group(k+1) is issued before filter(k): never waits for filter CTAs to leave)
    int plan = opts->stream_plan;
    if ((plan == 1 && dp->n_proc < 3) || ((plan == 2 || plan == 3) && dp->n_proc < 2) || plan < 0 || plan > 3) plan = 0;
    for (int i = 0; i < dp->n_proc; ++i) CUDA_TRY(cudaStreamWaitEvent(dp->s_proc[i], dp->ev_fork, 0));
    auto group_stream = [&](int c) { return plan == 0 ? dp->s_proc[c % dp->n_proc] : dp->s_proc[0]; };
    auto run_stages = [&](int c, int stages) -> int {
        const int f0 = c * cf, nc = (n_frames - f0) < cf ? (n_frames - f0) : cf;
        char* cws = base + P.off_chunks + (size_t)c * P.chunk_bytes;
        uint32_t* cb = ws.cellbox + (size_t)f0 * tv.TX * tv.TY;
        cudaStream_t ps = group_stream(c);
        ClusterLaunch hc = how;
        hc.zero = 0;
        hc.stages = stages;
        hc.ev_group = dp->ev_tl[c][1]; hc.ev_filter = dp->ev_tl[c][2]; hc.ev_borders = dp->ev_tl[c][3];
        if (plan == 1) { hc.s_filter = dp->s_proc[1]; hc.s_borders = dp->s_proc[2 + c % (dp->n_proc - 2)]; }
        if (plan == 2 || plan == 3) { hc.s_filter = dp->s_proc[0]; hc.s_borders = dp->s_proc[1 + c % (dp->n_proc - 1)]; }
        int rc = launch_cluster_path(frames_dev + (size_t)f0 * frame_stride, nc, H, W, frame_stride, tv, thresh, cb, cws, P.cl_offs, need_general + f0,
                                     max_contours, max_blobs, min_area, min_circ,
                                     out_xy + (size_t)f0 * max_blobs * 2, out_count + f0, out_flags + f0,
                                     out_contours ? out_contours + (size_t)f0 * max_contours * 8 : nullptr, out_contour_count ? out_contour_count + f0 : nullptr,
                                     ps, nullptr, hc);
        if (rc == MOCAP_OK && (stages & 4) && dp->sc_xy) {      // the chunk's records go where their consumers read them, off the critical path
            cudaStream_t sb = hc.s_borders ? hc.s_borders : (hc.s_filter ? hc.s_filter : ps);
            rc = launch_scatter(out_xy + (size_t)f0 * max_blobs * 2, out_count + f0, dp->sc_xy + f0, dp->sc_count + f0, nc, max_blobs, need_general + f0, 0, sb);
        }
        return rc;
    };
    for (int c = 0; c < chunks; ++c) {
        const int f0 = c * cf, nc = (n_frames - f0) < cf ? (n_frames - f0) : cf;
        cudaStream_t ps = group_stream(c);
        char* cws = base + P.off_chunks + (size_t)c * P.chunk_bytes;
        CUDA_TRY(cudaMemsetAsync(cws + P.cl_offs[14], 0, P.cl_offs[15], ps));     // the chunk's counters / lists, off the critical path
        if (prescanned) {
            // nothing to wait for: the boxes were complete before the fork
        } else if (mode == 1) {
            st = stream_wait_geq(ps, ctrl + PIPE_CTRL_WORK + chunks + c, 1);
            if (st != MOCAP_OK) return st;
        } else {
            if (tma) {
                // all launches share the batch-wide tensor map and item numbering; launch c covers the boxes of chunk c
                const int ib = f0 * per_frame_items, ie = (f0 + nc) * per_frame_items;
                st = launch_scan_tma(frames_dev, n_frames, H, W, frame_stride, tv, thresh, ws.cellbox, ctrl, chunks, cf, c, ib, ie, opts->scan_stages, dp->s_scan);
            } else {
                FilterWs wc = ws; wc.cellbox = ws.cellbox + (size_t)f0 * tv.TX * tv.TY;
                st = launch_scan(frames_dev + (size_t)f0 * frame_stride, nc, H, W, frame_stride, tv, thresh, wc, dp->s_scan, nullptr);
            }
            if (st != MOCAP_OK) return st;
            CUDA_TRY(cudaEventRecord(dp->ev_scan[c], dp->s_scan));
            CUDA_TRY(cudaStreamWaitEvent(ps, dp->ev_scan[c], 0));
        }
        if (tl) CUDA_TRY(cudaEventRecord(dp->ev_tl[c][0], ps));
        if (plan == 3) {
            st = run_stages(c, 1);
            if (st == MOCAP_OK && c > 0) st = run_stages(c - 1, 2 | 4);
            if (st == MOCAP_OK && c == chunks - 1) st = run_stages(c, 2 | 4);
        } else {
            st = run_stages(c, 1 | 2 | 4);
        }
        if (st != MOCAP_OK) return st;
    }
    CUDA_TRY(cudaEventRecord(dp->ev_scan_done, dp->s_scan));
    if (dp->tok_done) CUDA_TRY(cudaEventRecord((cudaEvent_t)dp->tok_done, dp->s_scan));
    CUDA_TRY(cudaStreamWaitEvent(s, dp->ev_scan_done, 0));
    for (int i = 0; i < dp->n_proc; ++i) {
        CUDA_TRY(cudaEventRecord(dp->ev_proc_done[i], dp->s_proc[i]));
        CUDA_TRY(cudaStreamWaitEvent(s, dp->ev_proc_done[i], 0));
    }
    if (tl) { CUDA_TRY(cudaEventRecord(dp->ev_join, s)); dp->tl_chunks = chunks; }
#endif
    // general path for the frames the cluster units could not finish (rare), whole batch
    st = launch_tiles(frames_dev, n_frames, H, W, frame_stride, tv, thresh, ws, P.L.max_fg, out_flags, need_general, s);
    if (st != MOCAP_OK) return st;
    st = launch_blobs(ws.bits, ws.fg_tiles, ws.n_fg, n_frames, H, W, P.L.TX, P.L.max_fg, max_runs, max_blobs, max_contours,
                      min_area, min_circ, base + P.L.off_blob, P.L.blob_stride,
                      out_xy, out_count, out_flags, nullptr, nullptr, out_contours, out_contour_count, nullptr, need_general, s);
    if (st == MOCAP_OK && dp->sc_xy) st = launch_scatter(out_xy, out_count, dp->sc_xy, dp->sc_count, n_frames, max_blobs, need_general, 1, s);
    return st;
}

// Timeline of the last mocap_detect_batch_pipelined call that had record_timeline set (read it after the stream has been
// synchronised): ms since the fork for [scan kernel(s) done, join] followed by, per chunk, [scan seen, grouped, filtered,
// borders done].  Returns the number of floats written (2 + 4 chunks), 0 if there is no timeline, or a negative status.
extern "C" int mocap_detect_pipe_timeline(void* pipe, float* ms_out, int cap)
{
    DetectPipe* dp = (DetectPipe*)pipe;
    if (!dp || !ms_out) return MOCAP_ERR_INVALID;
#ifdef MOCAP_EMU
    (void)cap;
    return 0;
#else
    const int need = 2 + PIPE_TL_PER_CHUNK * dp->tl_chunks;
    if (dp->tl_chunks == 0) return 0;
    if (cap < need) return MOCAP_ERR_INVALID;
    CUDA_TRY(cudaEventSynchronize(dp->ev_join));
    CUDA_TRY(cudaEventElapsedTime(&ms_out[0], dp->ev_fork, dp->ev_scan_done));
    CUDA_TRY(cudaEventElapsedTime(&ms_out[1], dp->ev_fork, dp->ev_join));
    for (int c = 0; c < dp->tl_chunks; ++c)
        for (int k = 0; k < PIPE_TL_PER_CHUNK; ++k)
            CUDA_TRY(cudaEventElapsedTime(&ms_out[2 + c * PIPE_TL_PER_CHUNK + k], dp->ev_fork, dp->ev_tl[c][k]));
    return need;
#endif
}

// what the last pipelined call actually ran: [0] scan kernel (1 = TMA ring, 0 = classic), [1] chunks, [2] sync mode used
extern "C" int mocap_detect_pipe_info(void* pipe, int* out3)
{
    DetectPipe* dp = (DetectPipe*)pipe;
    if (!dp || !out3) return MOCAP_ERR_INVALID;
    out3[0] = dp->last_scan; out3[1] = dp->last_chunks; out3[2] = dp->last_mode;
    return MOCAP_OK;
}

// the streaming scan alone (stage parity): cellbox_out [n][TY][TX] packed hot boxes (x0 | x1 << 8 | y0 << 16 | y1 << 24,
// 0xffffffff = no pixel > thresh in the cell).  variant 0: classic register-staged kernels, 1: TMA ring.
extern "C" int mocap_scan_cells_batch(const uint8_t* frames_dev, int n_frames, int H, int W, int64_t frame_stride, const void* table_dev,
                                      int thresh, int variant, uint32_t* cellbox_out, void* workspace, size_t workspace_bytes, void* stream)
{
    if (!frames_dev || !table_dev || !cellbox_out || n_frames <= 0 || H <= 0 || W <= 0) return MOCAP_ERR_INVALID;
    if (frame_stride < (int64_t)H * W) return MOCAP_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    TableView tv; table_view(table_dev, H, W, &tv);
    if (variant == 0) {
        FilterWs ws; memset(&ws, 0, sizeof(ws));
        ws.cellbox = cellbox_out;
        return launch_scan(frames_dev, n_frames, H, W, frame_stride, tv, thresh, ws, s, nullptr);
    }
    if (!scan_tma_supported(frames_dev, n_frames, H, W, frame_stride, thresh)) return MOCAP_ERR_UNSUPPORTED;
    if (!workspace || workspace_bytes < scan_tma_ctrl_bytes(1)) return MOCAP_ERR_WORKSPACE;
    CUDA_TRY(cudaMemsetAsync(workspace, 0, scan_tma_ctrl_bytes(1), s));
    return launch_scan_tma(frames_dev, n_frames, H, W, frame_stride, tv, thresh, cellbox_out, (int*)workspace, 1, n_frames, 0, 0, -1, variant == 2 ? 3 : 6, s);
}
