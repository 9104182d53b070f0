// Shared model plumbing: error state, device arena, BN folding, chunked forward.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <mutex>

#include "model.cuh"

namespace iq {

std::atomic<unsigned long long> g_launch_count{0};

static thread_local std::string t_error;
void set_error(const std::string &msg) { t_error = msg; }
const char *last_error() { return t_error.c_str(); }

int ensure_dynamic_smem(const void *func, int bytes)
{
    static std::mutex mu;
    static std::map<std::pair<const void *, int>, int> granted;
    int dev = 0;
    IQ_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    int &have = granted[std::make_pair(func, dev)];
    if (bytes > have) {
        IQ_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        have = bytes;
    }
    return 0;
}

static std::mutex g_env_mu;
static std::map<std::string, int> g_env_seen;
void env_forget()
{
    std::lock_guard<std::mutex> lock(g_env_mu);
    g_env_seen.clear();
}

int env_int(const char *name, int fallback)
{
    std::mutex &mu = g_env_mu;
    std::map<std::string, int> &seen = g_env_seen;
    std::lock_guard<std::mutex> lock(mu);
    auto it = seen.find(name);
    if (it != seen.end()) return it->second;
    const char *v = getenv(name);
    const int r = v ? (*v ? atoi(v) : 1) : fallback;
    seen[name] = r;
    return r;
}

// ---- profiling (shared by every thread that uses the library: guarded by one mutex; the launch counter is atomic)
namespace {
struct ProfEntry { const char *name; cudaEvent_t a, b; };
std::mutex g_prof_mu;
bool g_prof_on = false;
std::vector<ProfEntry> g_prof;
constexpr size_t PROF_CAP = 1 << 16;            // events kept per profiling session; later launches are not timed
}  // namespace

ProfileScope::ProfileScope(const char *name, cudaStream_t s) : slot(-1), st(s)
{
    std::lock_guard<std::mutex> lock(g_prof_mu);
    if (!g_prof_on || g_prof.size() >= PROF_CAP) return;
    ProfEntry e;
    e.name = name;
    if (cudaEventCreate(&e.a) != cudaSuccess) return;
    if (cudaEventCreate(&e.b) != cudaSuccess) { cudaEventDestroy(e.a); return; }
    cudaEventRecord(e.a, st);
    g_prof.push_back(e);
    slot = (int)g_prof.size() - 1;
}
ProfileScope::~ProfileScope()
{
    if (slot < 0) return;
    std::lock_guard<std::mutex> lock(g_prof_mu);
    if (slot < (int)g_prof.size()) cudaEventRecord(g_prof[slot].b, st);
}
void profile_enable(bool on)
{
    std::lock_guard<std::mutex> lock(g_prof_mu);
    for (auto &e : g_prof) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    g_prof.clear();
    g_prof_on = on;
}
int profile_report(const char **names, double *ms, long long *counts, int cap)
{
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lock(g_prof_mu);
    std::map<std::string, std::pair<double, long long>> acc;
    std::map<std::string, const char *> keep;
    for (auto &e : g_prof) {
        float t = 0.0f;
        if (cudaEventElapsedTime(&t, e.a, e.b) != cudaSuccess) continue;
        auto &slot = acc[e.name];
        slot.first += t;
        slot.second += 1;
        keep[e.name] = e.name;
    }
    int n = 0;
    for (auto &kv : acc) {
        if (n < cap) { names[n] = keep[kv.first]; ms[n] = kv.second.first; counts[n] = kv.second.second; }
        ++n;
    }
    return n;
}

int f16_paths()
{
    int m = 0;
    if (env_int("IQ_F16_CONV5", (F16_DEFAULT_PATHS & F16_CONV5) ? 1 : 0)) m |= F16_CONV5;
    if ((m & F16_CONV5) && env_int("IQ_F16_STORE", (F16_DEFAULT_PATHS & F16_STORE) ? 1 : 0)) m |= F16_STORE;
    if ((m & F16_CONV5) && env_int("IQ_F16_GRAM", (F16_DEFAULT_PATHS & F16_GRAM) ? 1 : 0)) m |= F16_GRAM;
    return m;
}

DeviceArena::~DeviceArena()
{
    for (void *p : ptrs) cudaFree(p);
}

int DeviceArena::upload_bytes(const void *host, size_t bytes, void **dev)
{
    void *p = nullptr;
    IQ_CUDA(cudaMalloc(&p, std::max<size_t>(bytes, 16)));
    ptrs.push_back(p);
    IQ_CUDA(cudaMemcpy(p, host, bytes, cudaMemcpyHostToDevice));
    *dev = p;
    return 0;
}

int DeviceArena::upload(const std::vector<float> &host, float **dev)
{
    void *p = nullptr;
    if (int rc = upload_bytes(host.data(), sizeof(float) * host.size(), &p)) return rc;
    *dev = reinterpret_cast<float *>(p);
    return 0;
}

// y = bn(W x + b) with eval statistics  ->  W' = s*W, b' = s*b + (beta - mean*s), s = gamma / sqrt(var + 1e-5).
// (BatchNorm eps 1e-5 is the nn.BatchNorm default the reference uses everywhere.)  Folded in double, rounded once.
bool fold_dense(const StateDict &sd, const std::string &w_key, const std::string &b_key, const std::string &bn_prefix,
                int cout, int cin, std::vector<float> &w, std::vector<float> &b, std::string &err)
{
    auto need = [&](const std::string &k, int64_t n) -> const float * {
        auto it = sd.find(k);
        if (it == sd.end()) { err = "state dict is missing '" + k + "'"; return nullptr; }
        if (it->second.numel != n) {
            err = "'" + k + "' has " + std::to_string(it->second.numel) + " elements, expected " + std::to_string(n);
            return nullptr;
        }
        return it->second.data;
    };
    const float *W = need(w_key, (int64_t)cout * cin);
    if (!W) return false;
    const float *bias = nullptr;
    if (!b_key.empty()) { bias = need(b_key, cout); if (!bias) return false; }
    const float *gamma = nullptr, *beta = nullptr, *mean = nullptr, *var = nullptr;
    if (!bn_prefix.empty()) {
        gamma = need(bn_prefix + ".weight", cout);
        beta = need(bn_prefix + ".bias", cout);
        mean = need(bn_prefix + ".running_mean", cout);
        var = need(bn_prefix + ".running_var", cout);
        if (!gamma || !beta || !mean || !var) return false;
    }
    w.resize((size_t)cout * cin);
    b.resize(cout);
    for (int o = 0; o < cout; ++o) {
        double s = 1.0, t = 0.0;
        if (gamma) {
            s = (double)gamma[o] / sqrt((double)var[o] + 1e-5);
            t = (double)beta[o] - (double)mean[o] * s;
        }
        for (int i = 0; i < cin; ++i) w[(size_t)o * cin + i] = (float)(s * (double)W[(size_t)o * cin + i]);
        b[o] = (float)(s * (bias ? (double)bias[o] : 0.0) + t);
    }
    return true;
}

void split_tf32_host(const std::vector<float> &w, std::vector<float> &hi, std::vector<float> &lo)
{
    auto rnd = [](float x) {
        uint32_t u;
        memcpy(&u, &x, 4);
        u = (u + 0x1000u) & 0xffffe000u;
        float r;
        memcpy(&r, &u, 4);
        return r;
    };
    hi.resize(w.size());
    lo.resize(w.size());
    for (size_t i = 0; i < w.size(); ++i) {
        hi[i] = rnd(w[i]);
        lo[i] = rnd(w[i] - hi[i]);
    }
}

float split_f16_host(const std::vector<float> &w, std::vector<__half> &hi, std::vector<__half> &lo)
{
    float mx = 0.0f;
    for (float v : w) mx = std::max(mx, std::fabs(v));
    float scale = 1.0f;
    if (mx > 0.0f && std::isfinite(mx)) {
        int e = 0;
        frexpf(mx, &e);                                 // mx = f * 2^e, f in [0.5, 1)
        scale = ldexpf(1.0f, std::min(std::max(10 - e, -60), 60));
    }
    hi.resize(w.size());
    lo.resize(w.size());
    for (size_t i = 0; i < w.size(); ++i) split_f16(w[i], scale, hi[i], lo[i]);
    return scale;
}

Model::~Model()
{
    for (int i = 0; i < MAX_LANES - 1; ++i) {
        if (side_[i]) cudaStreamDestroy(side_[i]);
        if (join_ev_[i]) cudaEventDestroy(join_ev_[i]);
    }
    if (fork_ev_) cudaEventDestroy(fork_ev_);
    if (done_ev_) cudaEventDestroy(done_ev_);
    if (host_meta_) cudaFreeHost(host_meta_);
}

// ---- chunk lanes: chunks of a forward are independent, so they are dealt round-robin to the caller's stream and
// internal side streams, each lane with its own scratch; one chunk's kernel tails and pipeline fill overlap another's
// steady state.
static int lanes_wanted(int lanes)
{
    const int lanes_env = env_int("IQ_LANES", 0);
    return std::min(std::max(lanes_env > 0 ? lanes_env : lanes, 1), (int)Model::MAX_LANES);
}

int Model::fork_lanes(int nl, cudaStream_t st)
{
    if (nl <= 1) return 0;
    if (!fork_ev_) IQ_CUDA(cudaEventCreateWithFlags(&fork_ev_, cudaEventDisableTiming));
    IQ_CUDA(cudaEventRecord(fork_ev_, st));                         // the side lanes see everything queued so far
    for (int l = 0; l < nl - 1; ++l) {
        if (!side_[l]) {
            IQ_CUDA(cudaStreamCreateWithFlags(&side_[l], cudaStreamNonBlocking));
            IQ_CUDA(cudaEventCreateWithFlags(&join_ev_[l], cudaEventDisableTiming));
        }
        IQ_CUDA(cudaStreamWaitEvent(side_[l], fork_ev_, 0));
    }
    return 0;
}

// the head (and the caller) wait for every lane; also run on the error path so that no side stream is left
// working on the scratch unobserved
int Model::join_lanes(int nl, cudaStream_t st)
{
    for (int l = 0; l < nl - 1; ++l) {
        if (!side_[l]) continue;
        IQ_CUDA(cudaEventRecord(join_ev_[l], side_[l]));
        IQ_CUDA(cudaStreamWaitEvent(st, join_ev_[l], 0));
    }
    return 0;
}

// clouds per chunk when the clouds have n instead of N points: the chunk keeps its number of rows
static int64_t chunk_for(int64_t chunk, int64_t N, int64_t n) { return std::max<int64_t>(1, chunk * N / n); }

// scratch of one lane: the largest chunk this forward can enqueue
int64_t Model::lane_bytes_for(int64_t head_mark, int64_t B, int64_t N, bool collapsed, cudaStream_t st)
{
    int64_t most = 0;
    for (int64_t n = collapsed ? 128 : N; n <= N; n += 128) {
        Workspace probe;
        probe.dry = true;
        probe.off = head_mark;
        if (run_body(probe, nullptr, 1, std::min<int64_t>(chunk_for(chunk, N, n), B), n, nullptr, nullptr, nullptr, st) != 0)
            return -1;
        most = std::max(most, round_up(probe.off, 256) - head_mark);
        if (!collapsed) break;
    }
    return most;
}

int Model::plan(Workspace &ws, const float *x, int point_major, int64_t B, int64_t N, float *logits,
                float *aux_trans_feat, int64_t *aux_crt, cudaStream_t st)
{
    float *pooled = ws.take<float>(B * pooled_dim());
    const int64_t head_mark = round_up(ws.off, 256);
    Workspace probe;
    probe.dry = true;
    probe.off = head_mark;
    if (int rc = run_body(probe, nullptr, point_major, std::min<int64_t>(chunk, B), N, nullptr, nullptr, nullptr, st)) return rc;
    const int64_t lane_bytes = round_up(probe.off, 256) - head_mark;
    const int nl = (int)std::min<int64_t>(lanes_wanted(lanes), ceil_div(B, chunk));
    const int64_t body_end = head_mark + nl * lane_bytes;
    if (!ws.dry) {
        IQ_CHECK(body_end <= ws.size, "forward: workspace too small");
        if (int rc = fork_lanes(nl, st)) return rc;
        cur_ = ChunkInfo();
        int rc = 0;
        int64_t i = 0;
        for (int64_t b0 = 0; b0 < B && rc == 0; b0 += chunk, ++i) {
            const int64_t Bc = std::min<int64_t>(chunk, B - b0);
            const int lane = (int)(i % nl);
            Workspace lw = ws;
            lw.off = head_mark + lane * lane_bytes;
            lw.size = head_mark + (lane + 1) * lane_bytes;
            rc = run_body(lw, x + b0 * N * 3, point_major, Bc, N, pooled + b0 * pooled_dim(),
                          aux_trans_feat ? aux_trans_feat + b0 * 64 * 64 : nullptr,
                          aux_crt ? aux_crt + b0 * 1024 : nullptr, lane == 0 ? st : side_[lane - 1]);
        }
        const int jrc = join_lanes(nl, st);
        if (rc != 0) return rc;
        if (jrc != 0) return jrc;
    }
    ws.off = head_mark;
    const int rc = run_head(ws, pooled, B, logits, st);
    if (rc != 0) return rc;
    ws.off = std::max(body_end, ws.off);
    return 0;
}

// Host half of the collapsed forward: from the kept-point counts U[b] of B clouds of N points, the compacted size of every
// cloud, n = N when nothing is masked, else round_up(U + min(M, copies), 128) with M = N - U (never more than N: U + M = N is
// a multiple of 128), the clouds ordered by size, largest first (stable: original order inside a size), and the weight the
// average pool still owes the last copy, extra = M - (n - U).  count[t] = clouds of size 128 * t.
void collapse_plan(const int32_t *kept, int64_t B, int64_t N, int copies, int32_t *src, int32_t *size, float *extra,
                   int64_t *count)
{
    const int nb = (int)(N / 128);
    auto size_of = [&](int32_t U) -> int64_t {
        const int64_t M = N - U;
        return M <= 0 ? N : std::min<int64_t>(N, round_up(U + std::min<int64_t>(M, copies), 128));
    };
    for (int t = 0; t <= nb; ++t) count[t] = 0;
    for (int64_t b = 0; b < B; ++b) count[size_of(kept[b]) / 128] += 1;
    std::vector<int64_t> fill(nb + 2, 0);
    for (int t = nb; t >= 1; --t) fill[t - 1] = fill[t] + count[t];
    for (int64_t b = 0; b < B; ++b) {
        const int64_t n = size_of(kept[b]);
        const int64_t s = fill[n / 128]++;
        src[s] = (int32_t)b;
        size[s] = (int32_t)n;
        extra[s] = (float)((N - kept[b]) - (n - kept[b]));                // M - m: copies not materialised
    }
}

// The collapsed forward (collapse.cu): count the kept points per cloud, group the clouds by compacted size, rewrite
// them, run every group through run_body at its own number of points, the head over all clouds, and hand the logits
// back in the caller's order.  One host round trip (the kept counts) decides the grouping.
int Model::plan_collapsed(Workspace &ws, const float *x, int point_major, int64_t B, int64_t N, float *logits,
                          const float *loc, cudaStream_t st)
{
    const int copies = std::max(collapse_copies(), 1);
    const int pd = pooled_dim();
    float *pooled = ws.take<float>(B * pd);
    float *slog = ws.take<float>(B * num_classes);
    int32_t *kept = ws.take<int32_t>(B);
    int32_t *meta = ws.take<int32_t>(3 * B);                         // src | row_off | size, in sorted order
    float *extra = ws.take<float>(B);
    float *xc = ws.take<float>(B * N * 3);
    const int64_t head_mark = round_up(ws.off, 256);
    const int64_t lane_bytes = lane_bytes_for(head_mark, B, N, true, st);
    if (lane_bytes < 0) return -1;
    const int want = lanes_wanted(lanes);
    if (ws.dry) {
        ws.off = head_mark;
        if (int rc = run_head(ws, nullptr, B, nullptr, st)) return rc;
        ws.off = std::max(head_mark + std::min<int64_t>(want, B) * lane_bytes, ws.off);
        return 0;
    }
    IQ_CHECK(B * N < ((int64_t)1 << 31), "forward: too many rows for one collapsed call");
    if (host_meta_cap_ < 5 * B) {
        if (host_meta_) cudaFreeHost(host_meta_);
        host_meta_ = nullptr;
        host_meta_cap_ = 0;
        IQ_CUDA(cudaMallocHost(reinterpret_cast<void **>(&host_meta_), sizeof(int32_t) * 5 * B));
        host_meta_cap_ = 5 * B;
    }
    if (int rc = launch_collapse_count(x, point_major, B, N, loc, kept, st)) return rc;
    int32_t *h_kept = host_meta_, *h_src = host_meta_ + B, *h_off = host_meta_ + 2 * B, *h_size = host_meta_ + 3 * B;
    float *h_extra = reinterpret_cast<float *>(host_meta_ + 4 * B);
    IQ_CUDA(cudaMemcpyAsync(h_kept, kept, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, st));
    IQ_CUDA(cudaStreamSynchronize(st));
    // counting sort by compacted size, largest first
    const int nb = (int)(N / 128);
    std::vector<int64_t> count(nb + 1, 0), start(nb + 2, 0);
    collapse_plan(h_kept, B, N, copies, h_src, h_size, h_extra, count.data());
    for (int t = nb; t >= 1; --t) start[t - 1] = start[t] + count[t];     // start[t] = first sorted position of size 128*t
    int64_t rows = 0;
    for (int64_t s = 0; s < B; ++s) { h_off[s] = (int32_t)rows; rows += h_size[s]; }
    last_row_fraction = (double)rows / (double)(B * N);
    last_buckets.assign(count.begin() + 1, count.end());
    IQ_CUDA(cudaMemcpyAsync(meta, h_src, sizeof(int32_t) * 3 * B, cudaMemcpyHostToDevice, st));
    IQ_CUDA(cudaMemcpyAsync(extra, h_extra, sizeof(float) * B, cudaMemcpyHostToDevice, st));
    if (int rc = launch_collapse_compact(x, point_major, B, N, loc, meta, meta + B, meta + 2 * B, xc, st)) return rc;

    int64_t nchunks = 0;
    for (int t = nb; t >= 1; --t) nchunks += ceil_div(count[t], chunk_for(chunk, N, 128 * t));
    const int nl = (int)std::min<int64_t>(std::min<int64_t>(want, B), nchunks);
    IQ_CHECK(head_mark + nl * lane_bytes <= ws.size, "forward: workspace too small");
    if (int rc = fork_lanes(nl, st)) return rc;
    int rc = 0;
    int64_t i = 0;
    for (int t = nb; t >= 1 && rc == 0; --t) {
        const int64_t n = 128 * t, cn = chunk_for(chunk, N, n);
        for (int64_t s0 = start[t]; s0 < start[t] + count[t] && rc == 0; s0 += cn, ++i) {
            const int64_t Bc = std::min<int64_t>(cn, start[t] + count[t] - s0);
            const int lane = (int)(i % nl);
            Workspace lw = ws;
            lw.off = head_mark + lane * lane_bytes;
            lw.size = head_mark + (lane + 1) * lane_bytes;
            cur_.pool_extra = extra + s0;
            rc = run_body(lw, xc + (int64_t)h_off[s0] * 3, 1, Bc, n, pooled + s0 * pd, nullptr, nullptr,
                          lane == 0 ? st : side_[lane - 1]);
        }
    }
    cur_ = ChunkInfo();
    const int jrc = join_lanes(nl, st);
    if (rc != 0) return rc;
    if (jrc != 0) return jrc;
    ws.off = head_mark;
    if (int hrc = run_head(ws, pooled, B, slog, st)) return hrc;
    return launch_scatter_rows(slog, meta, B, num_classes, logits, st);
}

int64_t Model::workspace_bytes(int64_t B, int64_t N)
{
    B = std::max<int64_t>(B, 1);
    Workspace ws;
    ws.dry = true;
    if (plan(ws, nullptr, 1, B, N, nullptr, nullptr, nullptr, nullptr) != 0) return -1;
    int64_t need = ws.off;
    if (collapse_copies() >= 0 && N % 128 == 0 && N >= 128) {          // a collapsed forward may be asked for as well
        Workspace wc;
        wc.dry = true;
        if (plan_collapsed(wc, nullptr, 1, B, N, nullptr, nullptr, nullptr) != 0) return -1;
        need = std::max(need, wc.off);
    }
    return round_up(need, 256) + 256;
}

int Model::forward(const float *x, int point_major, int64_t B, int64_t N, float *logits, void *wsp, int64_t ws_bytes,
                   float *aux_trans_feat, int64_t *aux_crt, const float *collapse_loc, cudaStream_t st)
{
    if (B == 0) return 0;                                          // an empty batch has null tensors and nothing to do
    IQ_CHECK(x && logits, "forward: null input or output");
    IQ_CHECK(wsp, "forward: null workspace");
    Workspace ws;
    ws.base = reinterpret_cast<char *>(wsp);
    ws.size = ws_bytes;
    last_row_fraction = 1.0;
    last_buckets.assign((size_t)ceil_div(N, 128), 0);
    last_buckets.back() = B;
    if (has_last_ && last_stream_ != st) IQ_CUDA(cudaStreamWaitEvent(st, done_ev_, 0));   // see model.cuh: one handle, any stream
    const bool no_collapse = env_int("IQ_NO_COLLAPSE", 0) != 0;              // A/B switch for scripts and tests
    int rc;
    if (collapse_loc && !no_collapse && collapse_copies() >= 0 && N % 128 == 0 && N >= 128 && !aux_trans_feat && !aux_crt)
        rc = plan_collapsed(ws, x, point_major, B, N, logits, collapse_loc, st);
    else
        rc = plan(ws, x, point_major, B, N, logits, aux_trans_feat, aux_crt, st);
    if (!done_ev_ && cudaEventCreateWithFlags(&done_ev_, cudaEventDisableTiming) != cudaSuccess) done_ev_ = nullptr;
    if (done_ev_ && cudaEventRecord(done_ev_, st) == cudaSuccess) {
        last_stream_ = st;
        has_last_ = true;
    }
    return rc;
}

}  // namespace iq
