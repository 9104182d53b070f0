// Model container of libiq_b200: folded weights on the device + forward orchestration.
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "common.cuh"
#include "kernels.cuh"

namespace iq {

struct HostTensor {
    const float *data = nullptr;
    int64_t numel = 0;
};
typedef std::map<std::string, HostTensor> StateDict;

// bump allocator over the caller-provided workspace (no cudaMalloc on the forward path)
struct Workspace {
    char *base = nullptr;
    int64_t size = 0, off = 0;
    bool dry = false;             // dry run: only measure
    template <typename T>
    T *take(int64_t n)
    {
        off = round_up(off, 256);
        T *p = dry ? nullptr : reinterpret_cast<T *>(base + off);
        off += n * (int64_t)sizeof(T);
        return p;
    }
    bool ok() const { return dry || off <= size; }
};

// A dense layer with eval-mode BatchNorm folded in: y = act(W' x + b'), W' (cout, cin) row-major.
struct Dense {
    float *w = nullptr;
    float *b = nullptr;
    float *w_hi = nullptr, *w_lo = nullptr;   // tf32 hi/lo split of w for the tcgen05 path
    __half *w_h_hi = nullptr, *w_h_lo = nullptr;   // optional: two-term fp16 split of w * h_scale (kind::f16 products)
    float h_scale = 1.0f;
    int cout = 0, cin = 0;
};

// tf32 hi/lo split on the host (same rounding as split_tf32_kernel)
void split_tf32_host(const std::vector<float> &w, std::vector<float> &hi, std::vector<float> &lo);
// two-term fp16 split (common.cuh, split_f16) of w * scale, scale = the power of two that brings max|w| into [2^9, 2^10):
// far from the fp16 overflow, and the low terms stay normal numbers down to |w| = 2^-12 max|w|
float split_f16_host(const std::vector<float> &w, std::vector<__half> &hi, std::vector<__half> &lo);

struct DeviceArena {              // owns every device buffer of a model
    std::vector<void *> ptrs;
    ~DeviceArena();
    int upload(const std::vector<float> &host, float **dev);
    int upload_bytes(const void *host, size_t bytes, void **dev);
};

class Model {
public:
    virtual ~Model();
    virtual const char *kind() const = 0;
    int num_classes = 10;
    // largest number of clouds one internal pass handles; forward() loops over chunks
    int chunk = 64;
    // chunks in flight: 1 = the caller's stream only, n = dealt round-robin over it and n-1 internal side streams
    // (own scratch per lane); the results do not depend on it
    static constexpr int MAX_LANES = 4;
    int lanes = 2;
    // GEMM engine for the large products: 1 = tcgen05 3xTF32 (default), 0 = exact fp32 SIMT
    int engine = 1;
    // bytes of workspace needed to run `B` clouds of `N` points (already capped by chunk)
    int64_t workspace_bytes(int64_t B, int64_t N);
    // collapse_loc (optional, device, 3 floats): the location the coalition masks moved the absent regions to; clouds
    // are then evaluated on their kept points + a few copies of it (collapse.cu) -- same logits, fewer rows
    int forward(const float *x, int layout_point_major, int64_t B, int64_t N, float *logits, void *ws, int64_t ws_bytes,
                float *aux_trans_feat, int64_t *aux_crt, const float *collapse_loc, cudaStream_t st);
    // statistics of the last collapsed forward: rows evaluated / rows of the uncollapsed batch (1 when not collapsed)
    double last_row_fraction = 1.0;
    // clouds of the last forward per evaluated size: last_buckets[t-1] = clouds run at 128*t points (plain: all at N)
    std::vector<int64_t> last_buckets;

    DeviceArena arena_;

private:
    cudaStream_t side_[MAX_LANES - 1] = {nullptr, nullptr, nullptr};
    cudaEvent_t fork_ev_ = nullptr, join_ev_[MAX_LANES - 1] = {nullptr, nullptr, nullptr};
    // The scratch, the lane events and the pinned staging buffer belong to the handle, so forwards of one handle are ordered
    // even when the caller switches streams: a forward on another stream than the previous one first waits for its end.
    cudaEvent_t done_ev_ = nullptr;
    cudaStream_t last_stream_ = nullptr;
    bool has_last_ = false;
    int32_t *host_meta_ = nullptr;     // pinned staging of the collapse plan (kept counts in, sorted order out)
    int64_t host_meta_cap_ = 0;
    int plan_collapsed(Workspace &ws, const float *x, int point_major, int64_t B, int64_t N, float *logits,
                       const float *collapse_loc, cudaStream_t st);
    int fork_lanes(int nl, cudaStream_t st);
    int join_lanes(int nl, cudaStream_t st);
    int64_t lane_bytes_for(int64_t head_mark, int64_t B, int64_t N, bool collapsed, cudaStream_t st);

protected:
    // Set by plan() for the chunk run_body is about to enqueue (read on the host at launch time only):
    // pool_extra (per cloud of the chunk, device) = weight of the chunk's last point beyond 1 for average pooling,
    // null when the clouds are not collapsed.
    struct ChunkInfo {
        const float *pool_extra = nullptr;
    } cur_;
    // Collapse support: < 0 = the forward depends on multiplicities in a way collapse.cu does not carry (FPS /
    // ball query / density models); otherwise the number of copies of the collapsed location a cloud must keep
    // (k for a k-nearest-neighbour graph, 1 for point-wise networks with max pooling).
    virtual int collapse_copies() const { return -1; }
    // width of the pooled per-cloud feature the body hands to the head
    virtual int pooled_dim() const = 0;
    // body: point-wise layers + pooling for a chunk of Bc clouds -> pooled (Bc, pooled_dim())
    virtual int run_body(Workspace &ws, const float *x, int point_major, int64_t Bc, int64_t N, float *pooled,
                         float *aux_trans_feat, int64_t *aux_crt, cudaStream_t st) = 0;
    // head: the classifier MLP over all B clouds at once (M = B keeps the GEMM grid filled)
    virtual int run_head(Workspace &ws, const float *pooled, int64_t B, float *logits, cudaStream_t st) = 0;
    int plan(Workspace &ws, const float *x, int point_major, int64_t B, int64_t N, float *logits, float *aux_trans_feat,
             int64_t *aux_crt, cudaStream_t st);
};

Model *create_edgeconv_model(const StateDict &sd, bool dynamic_graph, int k, int num_classes, std::string &err);
Model *create_pointnet_model(const StateDict &sd, int num_classes, std::string &err);
Model *create_pointnet2_model(const StateDict &sd, int num_classes, std::string &err);
Model *create_pointconv_model(const StateDict &sd, int num_classes, std::string &err);

// host half of the collapsed forward (models_common.cu), exported as iq_collapse_plan for CPU tests
void collapse_plan(const int32_t *kept, int64_t B, int64_t N, int copies, int32_t *src, int32_t *size, float *extra,
                   int64_t *count);

// host-side folding helpers (models_common.cu)
bool fold_dense(const StateDict &sd, const std::string &w_key, const std::string &b_key, const std::string &bn_prefix,
                int cout, int cin, std::vector<float> &w, std::vector<float> &b, std::string &err);

}  // namespace iq
