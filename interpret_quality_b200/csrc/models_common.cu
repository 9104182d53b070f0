// Shared model plumbing: error state, device arena, BN folding, chunked forward.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "model.cuh"

namespace iq {

unsigned long long g_launch_count = 0;

static thread_local std::string t_error;
void set_error(const std::string &msg) { t_error = msg; }
const char *last_error() { return t_error.c_str(); }

// ---- profiling
namespace {
struct ProfEntry { const char *name; cudaEvent_t a, b; };
bool g_prof_on = false;
std::vector<ProfEntry> g_prof;
std::vector<const char *> g_rep_names;
}  // namespace

ProfileScope::ProfileScope(const char *name, cudaStream_t s) : slot(-1), st(s)
{
    if (!g_prof_on) return;
    ProfEntry e;
    e.name = name;
    if (cudaEventCreate(&e.a) != cudaSuccess || cudaEventCreate(&e.b) != cudaSuccess) return;
    cudaEventRecord(e.a, st);
    g_prof.push_back(e);
    slot = (int)g_prof.size() - 1;
}
ProfileScope::~ProfileScope()
{
    if (slot >= 0) cudaEventRecord(g_prof[slot].b, st);
}
void profile_enable(bool on)
{
    for (auto &e : g_prof) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    g_prof.clear();
    g_prof_on = on;
}
int profile_report(const char **names, double *ms, long long *counts, int cap)
{
    cudaDeviceSynchronize();
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

DeviceArena::~DeviceArena()
{
    for (void *p : ptrs) cudaFree(p);
}

int DeviceArena::upload(const std::vector<float> &host, float **dev)
{
    void *p = nullptr;
    IQ_CUDA(cudaMalloc(&p, sizeof(float) * std::max<size_t>(host.size(), 1)));
    ptrs.push_back(p);
    IQ_CUDA(cudaMemcpy(p, host.data(), sizeof(float) * host.size(), cudaMemcpyHostToDevice));
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

Model::~Model()
{
    for (int i = 0; i < MAX_LANES - 1; ++i) {
        if (side_[i]) cudaStreamDestroy(side_[i]);
        if (join_ev_[i]) cudaEventDestroy(join_ev_[i]);
    }
    if (fork_ev_) cudaEventDestroy(fork_ev_);
}

int Model::plan(Workspace &ws, const float *x, int point_major, int64_t B, int64_t N, float *logits,
                float *aux_trans_feat, int64_t *aux_crt, cudaStream_t st)
{
    float *pooled = ws.take<float>(B * pooled_dim());
    const int64_t head_mark = round_up(ws.off, 256);
    // scratch of one chunk, measured on the largest one
    Workspace probe;
    probe.dry = true;
    probe.off = head_mark;
    if (int rc = run_body(probe, nullptr, point_major, std::min<int64_t>(chunk, B), N, nullptr, nullptr, nullptr, st)) return rc;
    const int64_t lane_bytes = round_up(probe.off, 256) - head_mark;
    // Chunks are independent: with lanes > 1 they are dealt round-robin to the caller's stream and internal side
    // streams, each lane with its own scratch, so one chunk's kernel tails and pipeline fill overlap another's
    // steady state.
    static const int lanes_env = getenv("IQ_LANES") ? atoi(getenv("IQ_LANES")) : 0;
    const int want = std::min(std::max(lanes_env > 0 ? lanes_env : lanes, 1), (int)MAX_LANES);
    const int nl = (int)std::min<int64_t>(want, ceil_div(B, chunk));
    const int64_t body_end = head_mark + nl * lane_bytes;
    if (!ws.dry) {
        IQ_CHECK(body_end <= ws.size, "forward: workspace too small");
        if (nl > 1) {
            if (!fork_ev_) IQ_CUDA(cudaEventCreateWithFlags(&fork_ev_, cudaEventDisableTiming));
            IQ_CUDA(cudaEventRecord(fork_ev_, st));                 // the side lanes see everything queued so far
            for (int l = 0; l < nl - 1; ++l) {
                if (!side_[l]) {
                    IQ_CUDA(cudaStreamCreateWithFlags(&side_[l], cudaStreamNonBlocking));
                    IQ_CUDA(cudaEventCreateWithFlags(&join_ev_[l], cudaEventDisableTiming));
                }
                IQ_CUDA(cudaStreamWaitEvent(side_[l], fork_ev_, 0));
            }
        }
        int64_t i = 0;
        for (int64_t b0 = 0; b0 < B; b0 += chunk, ++i) {
            const int64_t Bc = std::min<int64_t>(chunk, B - b0);
            const int lane = (int)(i % nl);
            Workspace lw = ws;
            lw.off = head_mark + lane * lane_bytes;
            lw.size = head_mark + (lane + 1) * lane_bytes;
            const int rc = run_body(lw, x + b0 * N * 3, point_major, Bc, N, pooled + b0 * pooled_dim(),
                                    aux_trans_feat ? aux_trans_feat + b0 * 64 * 64 : nullptr,
                                    aux_crt ? aux_crt + b0 * 1024 : nullptr, lane == 0 ? st : side_[lane - 1]);
            if (rc != 0) return rc;
        }
        for (int l = 0; l < nl - 1; ++l) {                            // the head (and the caller) wait for every lane
            IQ_CUDA(cudaEventRecord(join_ev_[l], side_[l]));
            IQ_CUDA(cudaStreamWaitEvent(st, join_ev_[l], 0));
        }
    }
    ws.off = head_mark;
    const int rc = run_head(ws, pooled, B, logits, st);
    if (rc != 0) return rc;
    ws.off = std::max(body_end, ws.off);
    return 0;
}

int64_t Model::workspace_bytes(int64_t B, int64_t N)
{
    Workspace ws;
    ws.dry = true;
    if (plan(ws, nullptr, 1, std::max<int64_t>(B, 1), N, nullptr, nullptr, nullptr, nullptr) != 0) return -1;
    return round_up(ws.off, 256) + 256;
}

int Model::forward(const float *x, int point_major, int64_t B, int64_t N, float *logits, void *wsp, int64_t ws_bytes,
                   float *aux_trans_feat, int64_t *aux_crt, cudaStream_t st)
{
    if (B == 0) return 0;                                          // an empty batch has null tensors and nothing to do
    IQ_CHECK(x && logits, "forward: null input or output");
    IQ_CHECK(wsp, "forward: null workspace");
    Workspace ws;
    ws.base = reinterpret_cast<char *>(wsp);
    ws.size = ws_bytes;
    return plan(ws, x, point_major, B, N, logits, aux_trans_feat, aux_crt, st);
}

}  // namespace iq
