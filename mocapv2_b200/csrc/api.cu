// C-ABI glue of the detection path: workspace carving and the mocap_detect_batch / mocap_filter_batch entry points
// (include/mocap_b200.h).  _find_dot of the reference (lib/ImageOperations.py:33-78) = launch_filter + launch_blobs.
#include "common.cuh"

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
                        const uint32_t* cellbox, char* ws_base, const size_t* offs,
                        int max_contours, int max_blobs, double min_area, double min_circ,
                        int32_t* out_xy, int32_t* out_count, int32_t* out_flags, double* out_contours, int32_t* out_contour_count,
                        cudaStream_t s, StageTimer* timer);
int launch_materialize_bits(const FilterWs& ws, int n, int H, int W, const TableView& tv, uint32_t* out, cudaStream_t s);
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
        st = launch_cluster_path(frames_dev, n_frames, H, W, frame_stride, tv, thresh, ws.cellbox, cl_base, L.cl_offs,
                                 max_contours, max_blobs, min_area, min_circ, out_xy, out_count, out_flags, out_contours,
                                 out_contour_count, s, timer);
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
