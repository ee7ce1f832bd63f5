// libespnet_b200.so -- host side of the C ABI declared in include/espnet_b200.h:
// weight packing (BN folding in fp64), workspace layout, stage scheduling and kernel launches.
// No torch, no cuDNN, no CPU fallback: if a CUDA call fails the error is returned to the caller.
#include "../../include/espnet_b200.h"

#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <array>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "kernels_fp32.cuh"
#include "kernels_fp32_tma.cuh"
#include "kernels_tc.cuh"
#include "kernels_tc_down.cuh"
#include "kernels_wsi.cuh"
#include "kernels_frontend.cuh"
#include "kernels_dec.cuh"
#include "kernels_tail_generic.cuh"

using namespace espnet;

static std::atomic<unsigned long long> g_launches{0};
static thread_local std::string g_create_error;

#define LAUNCH_COUNT() g_launches.fetch_add(1, std::memory_order_relaxed)

namespace {

// kernel classes of the "pdl" option
constexpr int kPdlStem = 1, /* 2: unused */ kPdlDown = 4, kPdlBranch = 8, kPdlReduce = 16, kPdlTail = 32, kPdlLast = 64, kPdlAll = 127;

struct HostTensor {
    const float* data;
    std::vector<int64_t> shape;
    size_t numel() const {
        size_t n = 1;
        for (auto s : shape) n *= (size_t)s;
        return n;
    }
};

// offsets (in floats) of the packed, kernel-ready parameter arrays inside the device blob
struct BlockW {
    size_t c1 = 0, d1 = 0, chain = 0, s = 0, t = 0, a = 0;
    size_t tc = 0;   // offset (bytes) into the fp16 blob for the tensor-core path: branch weights
    size_t tc_c1 = 0;   // ... and the c1 reduce weights (1x1: [CIN/8][NOUT][8]; 3x3 s2: [ks][tap][2][NOUT][8])
    size_t tc3 = 0, tc3_c1 = 0;   // 3-term split copies (fp32-equivalent tensor-core path): branch weights (merged / streamed-group layouts, see the packer), 1x1 reduce [hi|lo][CIN/8][NOUT][8]
};

struct Packed {
    size_t w1, l1_s, l1_t, l1_a, b1_s, b1_t, b1_a;
    size_t b2_s, b2_t, b2_a, b3_s, b3_t, b3_a;
    BlockW l2_0, l3_0;
    std::vector<BlockW> l2, l3;
    size_t cls_w;                        // encoder.classifier [256][NC]
    size_t br_s, br_t, up3_w;            // br BN, up_l3 ConvT
    size_t l3c_w;                        // level3_C [131][NC]
    size_t c0_s, c0_t, c0_a;             // combine_l2_l3.0 BR (2NC)
    size_t c1_w, c1_s, c1_t, c1_a;       // combine_l2_l3.1 CBR
    size_t up2_w, u2_s, u2_t, u2_a;      // up_l2
    size_t cv_w, cv_s, cv_t, cv_a;       // conv CBR
    size_t clsT_w;                       // classifier ConvT
};

struct Workspace {
    size_t out0cat, o1, l2a, l2b, out1cat, l3a, l3b, out2cat, enc, up3, t10, comb, total;
};

struct StageRef {
    const float* ptr = nullptr;
    size_t count = 0;
};

}  // namespace

struct espnet_handle {
    int classes = 0, p = 0, q = 0, net = 0, device = 0, mode = ESPNET_MODE_FP32;
    int num_sms = 148;
    std::string err;
    bool packed = false;
    float* dparams = nullptr;
    size_t nparams = 0;
    uint8_t* dparams_h = nullptr;   // fp16 tensor-core weights (BlockW::tc byte offsets)
    size_t nparams_h = 0;
    Packed pk;
    std::map<std::string, StageRef> stages;
    int fp32_impl = 1;     // fp32 mode: 1 (default) = tensor cores with 3-term fp16 operand splits (fp32-equivalent), 0 = CUDA-core FMA kernels ("fp32_impl")
    int dbg = 0;           // timing experiments inside the level-3 split branch kernel (wrong results), see BranchTcParams::dbg
    int down_impl = 1;     // tensor-core 3x3-s2 reduce: 1 = TMA-staged input regions (default), 0 = per-thread global loads ("down_impl")
    int tail_impl = 0;     // 1 = run the generic run-time-class-count tail kernels even for 5 / 20 classes ("tail_impl", cross-check)
    int dec_impl = 1;      // decoder tail: 1 = 4 pixels per thread (dec_c4_kernel), 0 = 1 pixel per thread ("dec_impl")
    int pdl = -1;          // programmatic dependent launch between the kernels of a forward (kernels_fp32.cuh): -1 auto, else a kPdl* mask
    int pdl_eff = 0;       // the mask of the forward being enqueued
    int l2_reverse = 1;    // 1x1 reduce walks its tiles against the order of the kernel that produced its input (L2 reuse)
    int tc_reduce = 1;     // f16tc mode: 1 = 1x1 reduce on tensor cores, 0 = CUDA-core fp32 reduce rounded to fp16 ("tc_reduce")
    int branch_impl = 0;   // 0 auto, 1 per-thread global loads, 2 TMA-staged (espnet_set_option "branch_impl")
    // per-kernel CUDA-event timing (espnet_set_profiling)
    bool profiling = false;
    struct ProfRec { const char* name; cudaEvent_t e0, e1; };
    std::vector<ProfRec> prof;
    // captured forwards (espnet_graph_capture): fixed buffers, one cudaGraphLaunch per forward
    struct GraphRec { cudaGraphExec_t exec = nullptr; };
    std::vector<GraphRec> graphs;
    std::map<std::array<uint64_t, 9>, CUtensorMap> tmaps;   // encoded tensor maps by (kind, base, shape, box): a pure function of the key
    std::set<const void*> smem_done;   // kernels whose dynamic shared-memory limit has been raised on this device
    // host-convenience path (espnet_segment_host)
    cudaStream_t own_stream = nullptr;
    void* hb_in = nullptr; void* hb_mask = nullptr; void* hb_ws = nullptr;
    size_t hb_in_sz = 0, hb_mask_sz = 0, hb_ws_sz = 0;
};

namespace {

int fail(espnet_t* h, int code, const std::string& msg) {
    if (h) h->err = msg; else g_create_error = msg;
    return code;
}

#define CUDA_TRY(h, expr)                                                                         \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess)                                                                    \
            return fail((h), ESPNET_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));    \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// Handle-less entry points launch on the device that OWNS their first device pointer, whatever device is current.
struct PtrDeviceGuard {
    int prev = -1;
    explicit PtrDeviceGuard(const void* p) {
        cudaPointerAttributes at{};
        if (p && cudaPointerGetAttributes(&at, p) == cudaSuccess && (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged)) {
            cudaGetDevice(&prev);
            if (prev != at.device) cudaSetDevice(at.device); else prev = -1;
        } else {
            cudaGetLastError();     // a host / unknown pointer is the callee's problem, not a sticky error
        }
    }
    ~PtrDeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// Brackets one kernel launch with CUDA events on the launching stream when profiling is on.
struct ProfScope {
    espnet_t* h; cudaStream_t st; cudaEvent_t e0 = nullptr, e1 = nullptr; const char* name;
    ProfScope(espnet_t* h_, const char* name_, cudaStream_t st_) : h(h_), st(st_), name(name_) {
        if (!h->profiling) return;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0, st);
    }
    ~ProfScope() {
        if (!e0) return;
        cudaEventRecord(e1, st);
        h->prof.push_back({name, e0, e1});
    }
};

inline bool pdl_on(const espnet_t* h, int cls) { return (h->pdl_eff & cls) && !h->profiling; }

// "pdl" = -1: measured on B200 (profiles/r02_pdl_ab.json), full ESPNet, 512 x 512 crops: chaining the kernels takes 16 % off a
// batch-1 forward (0.318 -> 0.264 ms) and 3 % off batch 16, nothing at batch 64 -- and a chain that is never broken (every kernel
// of every forward, back to back) costs 15 % there.  So: small forwards chain everything but the stem, large ones nothing.
constexpr long long kPdlAutoPixels = 32LL * 512 * 512;
inline int pdl_mask_for(const espnet_t* h, long long pixels) {
    if (h->pdl >= 0) return h->pdl;
    return pixels <= kPdlAutoPixels ? (kPdlAll & ~kPdlStem) : 0;
}

// A persistent kernel (one CTA per SM, static tile split) launched that way must not fit twice on an SM: its CTAs are placed
// as SMs drain, and the SMs that drain first would take two of them while others stay empty.
constexpr size_t kOneCtaSmem = 116 * 1024;
inline size_t one_cta_smem(size_t bytes) { return bytes > kOneCtaSmem ? bytes : kOneCtaSmem; }

// Launch of a forward kernel that follows the pdl_trigger / pdl_wait contract (kernels_fp32.cuh): with the "pdl" option on
// (default) its CTAs may be scheduled while the kernel before it drains.  Off while profiling: the per-kernel events would
// otherwise time overlapping prologues.
template <typename... KArgs, typename... Args>
cudaError_t launch_k(espnet_t* h, int cls, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_on(h, cls) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// ----------------------------------------------------------------------------------------------
// packing
// ----------------------------------------------------------------------------------------------
struct Packer {
    std::vector<float> blob;
    std::vector<uint16_t> blob_h;     // fp16 bits, tensor-core operand layout
    const std::map<std::string, HostTensor>* sd = nullptr;
    std::string missing;

    size_t add(const std::vector<float>& v) {
        while (blob.size() % 4) blob.push_back(0.f);      // 16 B alignment for float4 / LDS.128 rows
        const size_t off = blob.size();
        blob.insert(blob.end(), v.begin(), v.end());
        return off;
    }
    const HostTensor* get(const std::string& name, std::initializer_list<int64_t> shape) {
        auto it = sd->find(name);
        if (it == sd->end()) { if (missing.empty()) missing = "missing tensor '" + name + "'"; return nullptr; }
        if (it->second.shape != std::vector<int64_t>(shape)) {
            if (missing.empty()) missing = "tensor '" + name + "' has the wrong shape";
            return nullptr;
        }
        return &it->second;
    }
    // eval-mode BN (eps 1e-3) folded to y = x*s + t, in fp64 (running_var holds denormals)
    bool bn(const std::string& key, int c, size_t& s_off, size_t& t_off) {
        const HostTensor *g = get(key + ".weight", {c}), *b = get(key + ".bias", {c});
        const HostTensor *m = get(key + ".running_mean", {c}), *v = get(key + ".running_var", {c});
        if (!g || !b || !m || !v) return false;
        std::vector<float> s(c), t(c);
        for (int i = 0; i < c; ++i) {
            const double sc = (double)g->data[i] / std::sqrt((double)v->data[i] + 1e-3);
            s[i] = (float)sc;
            t[i] = (float)((double)b->data[i] - (double)m->data[i] * sc);
        }
        s_off = add(s);
        t_off = add(t);
        return true;
    }
    bool vec(const std::string& name, int c, size_t& off) {
        const HostTensor* a = get(name, {c});
        if (!a) return false;
        off = add(std::vector<float>(a->data, a->data + c));
        return true;
    }
    // conv weight [co][ci][k][k] -> [tap][ci][pad4(co)]
    bool conv_tap_ci_co(const std::string& name, int co, int ci, int k, size_t& off) {
        const HostTensor* w = get(name, {co, ci, k, k});
        if (!w) return false;
        const int cp = pad4(co), taps = k * k;
        std::vector<float> v((size_t)taps * ci * cp, 0.f);
        for (int o = 0; o < co; ++o)
            for (int c = 0; c < ci; ++c)
                for (int t = 0; t < taps; ++t) v[((size_t)t * ci + c) * cp + o] = w->data[((size_t)o * ci + c) * taps + t];
        off = add(v);
        return true;
    }
    // conv weight [co][ci][k][k] -> [ci][tap][co] (unpadded; decoder kernels)
    bool conv_ci_tap_co(const std::string& name, int co, int ci, int k, size_t& off) {
        const HostTensor* w = get(name, {co, ci, k, k});
        if (!w) return false;
        const int taps = k * k;
        std::vector<float> v((size_t)taps * ci * co, 0.f);
        for (int o = 0; o < co; ++o)
            for (int c = 0; c < ci; ++c)
                for (int t = 0; t < taps; ++t) v[((size_t)c * taps + t) * co + o] = w->data[((size_t)o * ci + c) * taps + t];
        off = add(v);
        return true;
    }
    bool raw(const std::string& name, std::initializer_list<int64_t> shape, size_t& off) {
        const HostTensor* w = get(name, shape);
        if (!w) return false;
        off = add(std::vector<float>(w->data, w->data + w->numel()));
        return true;
    }
    // one DownSamplerB / ESP block: c1 (k = 3 or 1), d1, chain, BN + PReLU
    bool block(const std::string& key, int cin, int cout, bool down, BlockW& bw) {
        const int n = cout / 5, n1 = cout - 4 * n;
        bool ok = conv_tap_ci_co(key + ".c1.conv.weight", n, cin, down ? 3 : 1, bw.c1);
        ok = ok && conv_tap_ci_co(key + ".d1.conv.weight", n1, n, 3, bw.d1);
        // chain: [4][9][n][pad4(n)]
        const int cp = pad4(n);
        std::vector<float> v((size_t)4 * 9 * n * cp, 0.f);
        const int ds[4] = {2, 4, 8, 16};
        for (int b = 0; b < 4 && ok; ++b) {
            const HostTensor* w = get(key + ".d" + std::to_string(ds[b]) + ".conv.weight", {n, n, 3, 3});
            if (!w) { ok = false; break; }
            for (int o = 0; o < n; ++o)
                for (int c = 0; c < n; ++c)
                    for (int t = 0; t < 9; ++t) v[(((size_t)b * 9 + t) * n + c) * cp + o] = w->data[((size_t)o * n + c) * 9 + t];
        }
        if (!ok) return false;
        bw.chain = add(v);
        {   // tensor-core copy [9 taps][NKC][5 branches][NOUT][8] fp16: element (n, ci) of tap t = W[co = n][ci][t], zero padded
            // (the five branches are contiguous along N so that the centre tap of all of them is ONE wide MMA)
            const int nkc = 2 * ((n + 15) / 16), nout = n1 <= 16 ? 16 : 32;
            while (blob_h.size() % 64) blob_h.push_back(0);          // 128 B alignment (cp.async.bulk needs 16 B)
            bw.tc = blob_h.size() * sizeof(uint16_t);
            blob_h.resize(blob_h.size() + (size_t)5 * 9 * nkc * nout * 8, 0);
            uint16_t* dst = blob_h.data() + bw.tc / sizeof(uint16_t);
            const int dd[5] = {1, 2, 4, 8, 16};
            for (int b = 0; b < 5; ++b) {
                const int co_n = b == 0 ? n1 : n;
                const HostTensor* w = get(key + ".d" + std::to_string(dd[b]) + ".conv.weight", {co_n, n, 3, 3});
                if (!w) return false;
                for (int t = 0; t < 9; ++t)
                    for (int o = 0; o < co_n; ++o)
                        for (int c = 0; c < n; ++c) {
                            const __half hv = __float2half_rn(w->data[((size_t)o * n + c) * 9 + t]);
                            uint16_t bits;
                            std::memcpy(&bits, &hv, 2);
                            dst[((((size_t)t * nkc + c / 8) * 5 + b) * nout + o) * 8 + (c % 8)] = bits;
                        }
            }
        }
        if (down) {    // tensor-core 3x3 stride-2 reduce: [ks][tap][2 chunks][NOUT][8] fp16, element = W[o][16 ks + 8 kc + j][tap]
            const int nout = 8 * (2 * ((n + 15) / 16)), ks_n = (cin + 15) / 16;
            const HostTensor* w = get(key + ".c1.conv.weight", {n, cin, 3, 3});
            if (!w) return false;
            while (blob_h.size() % 64) blob_h.push_back(0);
            bw.tc_c1 = blob_h.size() * sizeof(uint16_t);
            blob_h.resize(blob_h.size() + (size_t)ks_n * 9 * 2 * nout * 8, 0);
            uint16_t* dst = blob_h.data() + bw.tc_c1 / sizeof(uint16_t);
            for (int o = 0; o < n; ++o)
                for (int c = 0; c < cin; ++c)
                    for (int t = 0; t < 9; ++t) {
                        const __half hv = __float2half_rn(w->data[((size_t)o * cin + c) * 9 + t]);
                        uint16_t bits;
                        std::memcpy(&bits, &hv, 2);
                        dst[(((((size_t)(c / 16) * 9 + t) * 2 + (c % 16) / 8) * nout + o) * 8) + (c % 8)] = bits;
                    }
        }
        if (!down) {   // tensor-core 1x1 reduce: [CIN/8][NOUT][8] fp16, element (kc, o, j) = W1[o][8 kc + j]
            const int nout = 8 * (2 * ((n + 15) / 16));     // = 8 * NKC of the branch kernel's A operand
            const HostTensor* w = get(key + ".c1.conv.weight", {n, cin, 1, 1});
            if (!w) return false;
            while (blob_h.size() % 64) blob_h.push_back(0);
            bw.tc_c1 = blob_h.size() * sizeof(uint16_t);
            blob_h.resize(blob_h.size() + (size_t)(cin / 8) * nout * 8, 0);
            uint16_t* dst = blob_h.data() + bw.tc_c1 / sizeof(uint16_t);
            for (int o = 0; o < n; ++o)
                for (int c = 0; c < cin; ++c) {
                    const __half hv = __float2half_rn(w->data[(size_t)o * cin + c]);
                    uint16_t bits;
                    std::memcpy(&bits, &hv, 2);
                    dst[((size_t)(c / 8) * nout + o) * 8 + (c % 8)] = bits;
                }
        }
        {   // 3-term fp16 split copies: w_hi = fp16(4 w), w_lo = fp16(4 w - w_hi)
            auto split = [](float wv, uint16_t& hi, uint16_t& lo) {
                const float sw = wv * 4.0f;
                const __half h = __float2half_rn(sw);
                const __half l = __float2half_rn(sw - __half2float(h));
                std::memcpy(&hi, &h, 2); std::memcpy(&lo, &l, 2);
            };
            const int nkc = 2 * ((n + 15) / 16), ks_n = nkc / 2, nout = n1 <= 16 ? 16 : 32;
            const size_t unit = (size_t)5 * 9 * 2 * nout * 8;           // halves per (K step, hi|lo)
            while (blob_h.size() % 64) blob_h.push_back(0);
            bw.tc3 = blob_h.size() * sizeof(uint16_t);
            blob_h.resize(blob_h.size() + 2 * ks_n * unit, 0);
            uint16_t* dst = blob_h.data() + bw.tc3 / sizeof(uint16_t);
            const int dd[5] = {1, 2, 4, 8, 16};
            for (int b = 0; b < 5; ++b) {
                const int co_n = b == 0 ? n1 : n;
                const HostTensor* w = get(key + ".d" + std::to_string(dd[b]) + ".conv.weight", {co_n, n, 3, 3});
                if (!w) return false;
                for (int t = 0; t < 9; ++t)
                    for (int o = 0; o < co_n; ++o)
                        for (int c = 0; c < n; ++c) {
                            uint16_t hi, lo;
                            split(w->data[((size_t)o * n + c) * 9 + t], hi, lo);
                            if (ks_n == 1) {
                                // one K step: merged unit [tap][2][branch][hi NOUT | lo NOUT][8] -- A_hi x [W_hi | W_lo] is one N = 2 NOUT MMA
                                const size_t idx = ((((size_t)t * 2 + c / 8) * 5 + b) * 2 * nout + o) * 8 + (c % 8);
                                dst[idx] = hi;
                                dst[idx + (size_t)nout * 8] = lo;
                            } else {
                                // streamed layout (kernels_tc_branch.cuh, BranchTcCfg): per K step [W_hi A | W_hi B | W_lo A | W_lo B], a unit =
                                // [tap][2 chunks][branches of the group][NOUT][8]; group A = (d1, d2), group B = (d4, d8, d16)
                                const int ga = 2, gb = 3;
                                const size_t ua = (size_t)9 * 2 * ga * nout * 8, ub = (size_t)9 * 2 * gb * nout * 8;     // halves
                                const bool in_a = b < ga;
                                const int gn = in_a ? ga : gb, bl = in_a ? b : b - ga;
                                const size_t base = (size_t)(c / 16) * 2 * unit + (in_a ? 0 : ua);
                                const size_t idx = ((((size_t)t * 2 + (c % 16) / 8) * gn + bl) * nout + o) * 8 + (c % 8);
                                dst[base + idx] = hi;
                                dst[base + ua + ub + idx] = lo;
                            }
                        }
            }
            if (down) {    // 3x3 stride-2 reduce: [hi|lo][ks][tap][2][NOUT][8]
                const int nout1 = 8 * nkc, ks1 = (cin + 15) / 16;
                const HostTensor* w = get(key + ".c1.conv.weight", {n, cin, 3, 3});
                if (!w) return false;
                const size_t part = (size_t)ks1 * 9 * 2 * nout1 * 8;
                while (blob_h.size() % 64) blob_h.push_back(0);
                bw.tc3_c1 = blob_h.size() * sizeof(uint16_t);
                blob_h.resize(blob_h.size() + 2 * part, 0);
                uint16_t* d1p = blob_h.data() + bw.tc3_c1 / sizeof(uint16_t);
                for (int o = 0; o < n; ++o)
                    for (int c = 0; c < cin; ++c)
                        for (int t = 0; t < 9; ++t) {
                            uint16_t hi, lo;
                            split(w->data[((size_t)o * cin + c) * 9 + t], hi, lo);
                            const size_t idx = (((((size_t)(c / 16) * 9 + t) * 2 + (c % 16) / 8) * nout1 + o) * 8) + (c % 8);
                            d1p[idx] = hi;
                            d1p[part + idx] = lo;
                        }
            }
            if (!down) {
                const int nout1 = 8 * nkc;
                const HostTensor* w = get(key + ".c1.conv.weight", {n, cin, 1, 1});
                if (!w) return false;
                const size_t part = (size_t)(cin / 8) * nout1 * 8;
                while (blob_h.size() % 64) blob_h.push_back(0);
                bw.tc3_c1 = blob_h.size() * sizeof(uint16_t);
                blob_h.resize(blob_h.size() + 2 * part, 0);
                uint16_t* d1p = blob_h.data() + bw.tc3_c1 / sizeof(uint16_t);
                for (int o = 0; o < n; ++o)
                    for (int c = 0; c < cin; ++c) {
                        uint16_t hi, lo;
                        split(w->data[(size_t)o * cin + c], hi, lo);
                        const size_t idx = ((size_t)(c / 8) * nout1 + o) * 8 + (c % 8);
                        d1p[idx] = hi;
                        d1p[part + idx] = lo;
                    }
            }
        }
        const std::string bnk = down ? key + ".bn" : key + ".bn.bn";
        const std::string ak = down ? key + ".act.weight" : key + ".bn.act.weight";
        return bn(bnk, cout, bw.s, bw.t) && vec(ak, cout, bw.a);
    }
};

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

Workspace layout(const espnet_t* h, int B, int H, int W) {
    Workspace w{};
    const size_t P2 = (size_t)(H / 2) * (W / 2), P4 = (size_t)(H / 4) * (W / 4), P8 = (size_t)(H / 8) * (W / 8);
    const size_t NC = (size_t)h->classes;
    size_t off = 0;
    auto take = [&](size_t elems) { const size_t o = off; off = align_up(off + elems, 64); return o; };
    w.out0cat = take((size_t)B * 19 * P2);
    {   // o1 rows are padded to a multiple of 4 floats so that the map can be a TMA tensor (16 B strides)
        const size_t p4 = (size_t)(H / 4) * (size_t)pad4(W / 4), p8 = (size_t)(H / 8) * (size_t)pad4(W / 8);
        size_t need = (size_t)B * 12 * p4 > (size_t)B * 25 * p8 ? (size_t)B * 12 * p4 : (size_t)B * 25 * p8;
        // fp16 chunk-plane hi / lo pair of the split tensor-core path: 2 x 16 ch x 2 B (level 2), 2 x 32 ch x 2 B (level 3) per pixel
        const size_t split4 = (size_t)B * 16 * P4, split8 = (size_t)B * 32 * P8;
        if (split4 > need) need = split4;
        if (split8 > need) need = split8;
        w.o1 = take(need);
    }
    w.l2a = take((size_t)B * 64 * P4);
    w.l2b = take((size_t)B * 64 * P4);
    w.out1cat = take((size_t)B * 131 * P4);
    w.l3a = take((size_t)B * 128 * P8);
    w.l3b = take((size_t)B * 128 * P8);
    w.out2cat = take((size_t)B * 256 * P8);
    w.enc = take((size_t)B * NC * P8);
    w.up3 = take((size_t)B * NC * P4);
    w.t10 = take((size_t)B * 2 * NC * P4);
    w.comb = take((size_t)B * NC * P2);
    w.total = off;
    return w;
}

template <typename K>
int set_smem(espnet_t* h, K kernel, size_t bytes) {
    // once per (handle = device, kernel): the attribute is sticky, and the call costs a microsecond per launch otherwise
    if (h && h->smem_done.count((const void*)kernel)) return ESPNET_OK;
    CUDA_TRY(h, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    if (h) h->smem_done.insert((const void*)kernel);
    return ESPNET_OK;
}

// a launch whose threads do not even fill the GPU once: latency, not bandwidth, is what its time is made of
inline bool small_launch(const espnet_t* h, size_t threads) { return threads <= (size_t)h->num_sms * 256; }

int grid_for(const espnet_t* h, long long items) {
    long long g = items < (long long)h->num_sms ? items : (long long)h->num_sms;
    return (int)(g < 1 ? 1 : g);
}

long long tile_items(int B, int H, int W) { return (long long)B * ((H + kRows - 1) / kRows) * ((W + 31) / 32); }

template <int CIN, int CO>
int run_reduce1x1(espnet_t* h, const float* in, size_t w_off, float* o1, int B, int H, int W, cudaStream_t st) {
    const int HW = H * W;
    const size_t smem = (size_t)CIN * pad4(CO) * sizeof(float);
    const long long items = (long long)B * ((HW + 127) / 128);
    const int threads = 256;
    long long ctas = (items + 7) / 8;
    const long long cap = 2LL * h->num_sms;
    if (ctas > cap) ctas = cap;
    { ProfScope _ps(h, CIN == 64 ? "reduce1x1_l2" : "reduce1x1_l3", st); reduce1x1_kernel<CIN, CO><<<(int)ctas, threads, smem, st>>>(in, h->dparams + w_off, o1, B, HW, W, pad4(W)); }
    LAUNCH_COUNT();
    CUDA_TRY(h, cudaPeekAtLastError());
    return ESPNET_OK;
}

template <int CIN, int CO>
int run_reduce3x3(espnet_t* h, const float* in, size_t w_off, float* o1, int B, int Hi, int Wi, cudaStream_t st) {
    const size_t smem = ((size_t)9 * CIN + 4) * pad4(CO) * sizeof(float);
    int rc = set_smem(h, reduce3x3s2_kernel<CIN, CO>, smem);
    if (rc) return rc;
    const int grid = grid_for(h, tile_items(B, Hi / 2, Wi / 2));
    { ProfScope _ps(h, CIN == 19 ? "reduce3x3s2_l2" : "reduce3x3s2_l3", st); reduce3x3s2_kernel<CIN, CO><<<grid, kHeavyThreads, smem, st>>>(in, h->dparams + w_off, o1, B, Hi, Wi, pad4(Wi / 2)); }
    LAUNCH_COUNT();
    CUDA_TRY(h, cudaPeekAtLastError());
    return ESPNET_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
    }
    return fn;
}

// fp32 tensor map over o1 [B][N][H][Wp] with a [1][CH][64][64] box, zero OOB fill, no swizzle
// Encoded tensor maps are cached in the handle: a descriptor is a pure function of (base address, shape, box), a forward at
// an unchanged workspace re-uses all 16 of its maps instead of paying cuTensorMapEncodeTiled per launch (batch-1 loop).
inline bool tmap_lookup(espnet_t* h, const std::array<uint64_t, 9>& key, CUtensorMap* map) {
    if (!h) return false;
    auto it = h->tmaps.find(key);
    if (it == h->tmaps.end()) return false;
    *map = it->second;
    return true;
}
inline void tmap_store(espnet_t* h, const std::array<uint64_t, 9>& key, const CUtensorMap* map) {
    if (!h) return;
    if (h->tmaps.size() >= 512) h->tmaps.clear();      // callers that keep changing workspaces / shapes: bounded
    h->tmaps[key] = *map;
}

int make_o1_map(espnet_t* h, CUtensorMap* map, const float* o1, int B, int N, int H, int W, int Wp, int CH) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return fail(h, ESPNET_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)Wp * 4, (cuuint64_t)H * Wp * 4, (cuuint64_t)N * H * Wp * 4};
    cuuint32_t box[4] = {(cuuint32_t)kTmaBox, (cuuint32_t)kTmaBox, (cuuint32_t)CH, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)o1, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(h, ESPNET_ECUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return ESPNET_OK;
}

template <int N, int CO1, int CO>
int run_branch(espnet_t* h, const BlockW& bw, const float* o1, const float* res, float* out, float* out2, int C2, int c2_off,
               size_t s2, size_t t2, size_t a2, int B, int H, int W, cudaStream_t st) {
    constexpr int C = CO1 + 4 * CO;
    BranchParams p{};
    p.o1 = o1;
    p.w_d1 = h->dparams + bw.d1;
    p.w_chain = h->dparams + bw.chain;
    p.res = res;
    p.s = h->dparams + bw.s; p.t = h->dparams + bw.t; p.a = h->dparams + bw.a;
    p.out = out;
    p.s2 = h->dparams + s2; p.t2 = h->dparams + t2; p.a2 = h->dparams + a2;
    p.out2 = out2; p.C2 = C2; p.c2_off = c2_off;
    p.B = B; p.H = H; p.W = W; p.Wp = pad4(W);

    const long long cta_tiles = (long long)B * ((H + kTmaTile - 1) / kTmaTile) * ((W + kTmaTile - 1) / kTmaTile);
    // the TMA-staged kernel works on 32x32 CTA tiles: it needs about one tile per SM to fill the machine
    const bool use_tma = h->branch_impl == 2 || (h->branch_impl == 0 && cta_tiles >= (long long)h->num_sms);
    if (use_tma) {
        constexpr int CH = (N == 25) ? 2 : 3;
        constexpr int NS = 3;
        using Cfg = BranchTmaCfg<N, CO1, CO, CH, NS>;
        int rc = set_smem(h, esp_branch_tma_kernel<N, CO1, CO, CH, NS>, Cfg::SMEM);
        if (rc) return rc;
        CUtensorMap map;
        rc = make_o1_map(h, &map, o1, B, N, H, W, p.Wp, CH);
        if (rc) return rc;
        const int grid = grid_for(h, cta_tiles);
        { ProfScope _ps(h, N == 12 ? "esp_branch_l2" : "esp_branch_l3", st);
          esp_branch_tma_kernel<N, CO1, CO, CH, NS><<<grid, 256, Cfg::SMEM, st>>>(map, p); }
        LAUNCH_COUNT();
        CUDA_TRY(h, cudaPeekAtLastError());
        return ESPNET_OK;
    }
    const size_t smem = ((size_t)9 * N * pad4(CO1) + (size_t)36 * N * pad4(CO) + 16 * pad4(CO1) + 6 * C) * sizeof(float);
    int rc = set_smem(h, esp_branch_kernel<N, CO1, CO>, smem);
    if (rc) return rc;
    const int grid = grid_for(h, tile_items(B, H, W));
    { ProfScope _ps(h, N == 12 ? "esp_branch_l2" : "esp_branch_l3", st); esp_branch_kernel<N, CO1, CO><<<grid, kHeavyThreads, smem, st>>>(p); }
    LAUNCH_COUNT();
    CUDA_TRY(h, cudaPeekAtLastError());
    return ESPNET_OK;
}


// ---------------------------------------------------------------------------------------------- tcgen05 path
// tensor map over the fp16 chunk-plane map o1h [B][NKC][H][W][8]: [W][8 halves] is presented as W*4 32-bit words so that
// a box row is one contiguous 768 B run (a 16 B innermost box dimension makes TMA crawl); box {48*4, 48, NKC, 1},
// zero OOB fill (= the conv zero padding), no swizzle
int make_o1h_map(espnet_t* h, CUtensorMap* map, const __half* o1h, int B, int NKC, int H, int W, int box_w, int box_h, int box_planes) {
    const std::array<uint64_t, 9> key = {1, (uint64_t)(uintptr_t)o1h, (uint64_t)B, (uint64_t)NKC, (uint64_t)H, (uint64_t)W, (uint64_t)box_w, (uint64_t)box_h, (uint64_t)box_planes};
    if (tmap_lookup(h, key, map)) return ESPNET_OK;
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return fail(h, ESPNET_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[4] = {(cuuint64_t)W * 4, (cuuint64_t)H, (cuuint64_t)NKC, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)NKC * H * W * 16};
    cuuint32_t box[4] = {(cuuint32_t)box_w * 4, (cuuint32_t)box_h, (cuuint32_t)box_planes, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, (void*)o1h, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(h, ESPNET_ECUDA, "cuTensorMapEncodeTiled (o1h) failed with CUresult " + std::to_string((int)r));
    tmap_store(h, key, map);
    return ESPNET_OK;
}

template <int CIN, int CO, int NKC, bool SPLIT = false>
int run_reduce1x1_f16(espnet_t* h, const float* in, size_t w_off, __half* o1h, int B, int H, int W, cudaStream_t st) {
    const int HW = H * W;
    const size_t smem = (size_t)CIN * pad4(CO) * sizeof(float);
    const long long items = (long long)B * ((HW + 127) / 128);
    long long ctas = (items + 7) / 8;
    const long long cap = 2LL * h->num_sms;
    if (ctas > cap) ctas = cap;
    { ProfScope _ps(h, CIN == 64 ? "reduce1x1_f16_l2" : "reduce1x1_f16_l3", st);
      reduce1x1_f16_kernel<CIN, CO, NKC, SPLIT><<<(int)ctas, 256, smem, st>>>(in, h->dparams + w_off, o1h, B, HW); }
    LAUNCH_COUNT();
    CUDA_TRY(h, cudaPeekAtLastError());
    return ESPNET_OK;
}

template <int CIN, int NOUT, int NKC, bool SPLIT = false>
int run_reduce1x1_tc(espnet_t* h, const float* in, const BlockW& bw, __half* o1h, int B, int H, int W, cudaStream_t st) {
    using Cfg = ReduceTcCfg<CIN, NOUT, SPLIT>;
    const int HW = H * W;
    int rc = set_smem(h, reduce1x1_tc_kernel<CIN, NOUT, NKC, SPLIT>, one_cta_smem(Cfg::SMEM));
    if (rc) return rc;
    const int grid = grid_for(h, (long long)B * ((HW + 127) / 128));
    { ProfScope _ps(h, SPLIT ? (CIN == 64 ? "reduce1x1_tc3_l2" : "reduce1x1_tc3_l3") : (CIN == 64 ? "reduce1x1_tc_l2" : "reduce1x1_tc_l3"), st);
      // the producer of `in` (a branch kernel) wrote its tiles in ascending order: walking them in DESCENDING order starts on
      // the part that is still in the 126 MB L2 ("l2_reverse" option, default on)
      launch_k(h, kPdlReduce, reduce1x1_tc_kernel<CIN, NOUT, NKC, SPLIT>, grid, kRedThreads, one_cta_smem(Cfg::SMEM), st, in, reinterpret_cast<const __half*>(h->dparams_h + (SPLIT ? bw.tc3_c1 : bw.tc_c1)), o1h, B, HW, h->l2_reverse); }
    LAUNCH_COUNT();
    CUDA_TRY(h, cudaPeekAtLastError());
    return ESPNET_OK;
}

// fp32 tensor map over a planar activation tensor [B][C][H][W] with a {20 cols, 33 rows, 16 channels, 1} box (zero OOB fill):
// the input region of one K = 16 step of the TMA-staged 3x3-s2 reduce (kernels_tc_down.cuh)
int make_down_map(espnet_t* h, CUtensorMap* map, const float* in, int B, int C, int H, int W, int box_c) {
    const std::array<uint64_t, 9> key = {2, (uint64_t)(uintptr_t)in, (uint64_t)B, (uint64_t)C, (uint64_t)H, (uint64_t)W, (uint64_t)box_c, 0, 0};
    if (tmap_lookup(h, key, map)) return ESPNET_OK;
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return fail(h, ESPNET_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)C * H * W * 4};
    cuuint32_t box[4] = {(cuuint32_t)kDownBoxCols, (cuuint32_t)kDownRows, (cuuint32_t)box_c, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)in, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(h, ESPNET_ECUDA, "cuTensorMapEncodeTiled (3x3-s2 input) failed with CUresult " + std::to_string((int)r));
    tmap_store(h, key, map);
    return ESPNET_OK;
}

template <int CIN, int NOUT, int NKC, bool SPLIT = false>
int run_reduce3x3_tc(espnet_t* h, const float* in, const BlockW& bw, __half* o1h, int B, int Hi, int Wi, cudaStream_t st) {
    const int Ho = Hi / 2, Wo = Wi / 2;
    const int grid = grid_for(h, (long long)B * ((Ho + 15) / 16) * ((Wo + 7) / 8));
    const char* name = SPLIT ? (CIN == 19 ? "reduce3x3s2_tc3_l2" : "reduce3x3s2_tc3_l3") : (CIN == 19 ? "reduce3x3s2_tc_l2" : "reduce3x3s2_tc_l3");
    const __half* wp = reinterpret_cast<const __half*>(h->dparams_h + (SPLIT ? bw.tc3_c1 : bw.tc_c1));
    // TMA needs 16 B row pitches and a 16 B aligned base; "down_impl" = 0 forces the per-thread loader kernel (cross-check)
    if (h->down_impl != 0 && (Wi % 4) == 0 && ((uintptr_t)in % 16) == 0) {
        using TCfg = DownTmaCfg<CIN, NOUT, SPLIT>;
        int rc = set_smem(h, reduce3x3s2_tma_kernel<CIN, NOUT, NKC, SPLIT>, one_cta_smem(TCfg::SMEM));
        if (rc) return rc;
        CUtensorMap map, map_tail;
        rc = make_down_map(h, &map, in, B, CIN, Hi, Wi, 16);
        if (rc) return rc;
        rc = make_down_map(h, &map_tail, in, B, CIN, Hi, Wi, (CIN % 16) ? (CIN % 16) : 16);     // the last K step's real channels only
        if (rc) return rc;
        { ProfScope _ps(h, name, st); launch_k(h, kPdlDown, reduce3x3s2_tma_kernel<CIN, NOUT, NKC, SPLIT>, grid, kDownThreads, one_cta_smem(TCfg::SMEM), st, map, map_tail, wp, o1h, B, Hi, Wi); }
        LAUNCH_COUNT();
        CUDA_TRY(h, cudaPeekAtLastError());
        return ESPNET_OK;
    }
    using Cfg = DownTcCfg<CIN, NOUT, SPLIT>;
    int rc = set_smem(h, reduce3x3s2_tc_kernel<CIN, NOUT, NKC, SPLIT>, one_cta_smem(Cfg::SMEM));
    if (rc) return rc;
    { ProfScope _ps(h, name, st); launch_k(h, kPdlDown, reduce3x3s2_tc_kernel<CIN, NOUT, NKC, SPLIT>, grid, kDownThreads, one_cta_smem(Cfg::SMEM), st, in, wp, o1h, B, Hi, Wi); }
    LAUNCH_COUNT();
    CUDA_TRY(h, cudaPeekAtLastError());
    return ESPNET_OK;
}

template <int CIN, int CO, int NKC, bool SPLIT = false>
int run_reduce3x3_f16(espnet_t* h, const float* in, size_t w_off, __half* o1h, int B, int Hi, int Wi, cudaStream_t st) {
    const size_t smem = ((size_t)9 * CIN + 4) * pad4(CO) * sizeof(float);
    int rc = set_smem(h, reduce3x3s2_f16_kernel<CIN, CO, NKC, SPLIT>, smem);
    if (rc) return rc;
    const int grid = grid_for(h, tile_items(B, Hi / 2, Wi / 2));
    { ProfScope _ps(h, CIN == 19 ? "reduce3x3s2_f16_l2" : "reduce3x3s2_f16_l3", st);
      reduce3x3s2_f16_kernel<CIN, CO, NKC, SPLIT><<<grid, kHeavyThreads, smem, st>>>(in, h->dparams + w_off, o1h, B, Hi, Wi); }
    LAUNCH_COUNT();
    CUDA_TRY(h, cudaPeekAtLastError());
    return ESPNET_OK;
}

template <int NKC, int NOUT, int CO1, int CO, bool SPLIT = false>
int run_branch_tc(espnet_t* h, const BlockW& bw, const __half* o1h, const float* res, float* out, float* out2, int C2, int c2_off,
                  size_t s2, size_t t2, size_t a2, int B, int H, int W, cudaStream_t st) {
    using Cfg = BranchTcCfg<NKC, NOUT, SPLIT>;
    BranchTcParams p{};
    p.w = reinterpret_cast<const __half*>(h->dparams_h + (SPLIT ? bw.tc3 : bw.tc));
    p.res = res;
    p.s = h->dparams + bw.s; p.t = h->dparams + bw.t; p.a = h->dparams + bw.a;
    p.out = out;
    p.s2 = h->dparams + s2; p.t2 = h->dparams + t2; p.a2 = h->dparams + a2;
    p.out2 = out2; p.C2 = C2; p.c2_off = c2_off;
    p.B = B; p.H = H; p.W = W;
    p.dbg = h->dbg;
    const int var = res == nullptr ? 0 : (out != nullptr ? 1 : 2);
    if ((var == 0 && (!out || !out2)) || (var == 2 && !out2)) return fail(h, ESPNET_EINVAL, "run_branch_tc: unsupported output combination");
    CUtensorMap map;
    int rc = make_o1h_map(h, &map, o1h, SPLIT ? 2 * B : B, NKC, H, W, kTcBoxW, kTcBoxH, 2);
    if (rc) return rc;
    const long long tiles = (long long)B * ((H + kTcTileH - 1) / kTcTileH) * ((W + kTcTileW - 1) / kTcTileW);
    auto k0 = esp_branch_tc_kernel<NKC, NOUT, CO1, CO, 0, SPLIT>;
    auto k1 = esp_branch_tc_kernel<NKC, NOUT, CO1, CO, 1, SPLIT>;
    auto k2 = esp_branch_tc_kernel<NKC, NOUT, CO1, CO, 2, SPLIT>;
    auto kern = var == 0 ? k0 : (var == 1 ? k1 : k2);
    rc = set_smem(h, kern, one_cta_smem(Cfg::SMEM));
    if (rc) return rc;
    const int grid = grid_for(h, tiles);
    { ProfScope _ps(h, SPLIT ? (NKC == 2 ? "esp_branch_tc3_l2" : "esp_branch_tc3_l3") : (NKC == 2 ? "esp_branch_tc_l2" : "esp_branch_tc_l3"), st);
      launch_k(h, kPdlBranch, kern, grid, kTcThreads, one_cta_smem(Cfg::SMEM), st, map, p); }
    LAUNCH_COUNT();
    CUDA_TRY(h, cudaPeekAtLastError());
    return ESPNET_OK;
}

template <int NC>
int run_tail(espnet_t* h, const espnet_forward_args* a, const Workspace& L, float* ws, cudaStream_t st) {
    const int B = a->B, H = a->H, W = a->W;
    const int H2 = H / 2, W2 = W / 2, H4 = H / 4, W4 = W / 4, H8 = H / 8, W8 = W / 8;
    const Packed& pk = h->pk;
    const float* P = h->dparams;
    const bool full = h->net == ESPNET_NET_FULL;
    {
        Head3Params<NC> p{};
        p.in = ws + L.out2cat;
        p.w = P + pk.cls_w;
        p.B = B; p.H8 = H8; p.W8 = W8;
        if (full) {
            p.bn_s = P + pk.br_s; p.bn_t = P + pk.br_t; p.wt = P + pk.up3_w;
            p.up_out = ws + L.up3;
            p.enc_out = nullptr;
        } else {
            p.enc_out = a->logits ? a->logits : ws + L.enc;
        }
        const size_t n = (size_t)B * H8 * W8;
        int grid = (int)((n + 255) / 256);
        if (grid > 8 * h->num_sms) grid = 8 * h->num_sms;
        const bool v4 = NC == 5 && h->dec_impl != 0 && (((size_t)H8 * W8) % 4 == 0) && ((uintptr_t)p.enc_out % 16 == 0);
        if (v4) {
            int g4 = (int)((n / 4 + 255) / 256);
            if (g4 > 8 * h->num_sms) g4 = 8 * h->num_sms;
            if (g4 < 1) g4 = 1;
            if constexpr (NC == 5) {      // (v4 implies NC == 5: the 4-pixel kernels are only instantiated for 5 classes)
                ProfScope _ps(h, "head3", st);
                if (small_launch(h, n / 4)) launch_k(h, kPdlTail, head3v_kernel<NC, 32>, (int)((n / 4 + 63) / 64), 64, 0, st, p);
                else launch_k(h, kPdlTail, head3v_kernel<NC, 8>, g4, 256, 0, st, p);
            }
        } else {
            { ProfScope _ps(h, "head3", st); head3_kernel<NC><<<grid, 256, 0, st>>>(p); }
        }
        LAUNCH_COUNT();
        CUDA_TRY(h, cudaPeekAtLastError());
        if (!full) {
            h->stages["encoder.classifier"] = {p.enc_out, (size_t)B * NC * H8 * W8};
            if (a->mask) {
                dim3 g((W / 4 + 31) / 32, (H8 + 1 + 7) / 8, B);     // one thread = 4 output pixels x the 8 rows of a source-row pair
                { ProfScope _ps(h, "upsample8_argmax", st); launch_k(h, kPdlLast, upsample8_argmax_kernel<NC>, g, 256, 0, st, (const float*)p.enc_out, B, H8, W8, a->mask, (float*)nullptr); }
                LAUNCH_COUNT();
                CUDA_TRY(h, cudaPeekAtLastError());
            }
            return ESPNET_OK;
        }
        h->stages["up_l3"] = {ws + L.up3, (size_t)B * NC * H4 * W4};
    }
    {
        DecAParams<NC> p{};
        p.out1cat = ws + L.out1cat; p.up3 = ws + L.up3; p.w = P + pk.l3c_w;
        p.s = P + pk.c0_s; p.t = P + pk.c0_t; p.a = P + pk.c0_a;
        p.tout = ws + L.t10; p.B = B; p.H4 = H4; p.W4 = W4;
        const size_t n = (size_t)B * H4 * W4;
        int grid = (int)((n + 255) / 256);
        if (grid > 8 * h->num_sms) grid = 8 * h->num_sms;
        if (NC == 5 && h->dec_impl != 0 && (((size_t)H4 * W4) % 4 == 0)) {
            int g4 = (int)((n / 4 + 255) / 256);
            if (g4 > 8 * h->num_sms) g4 = 8 * h->num_sms;
            if (g4 < 1) g4 = 1;
            if constexpr (NC == 5) {
                ProfScope _ps(h, "dec_a", st);
                if (small_launch(h, n / 4)) launch_k(h, kPdlTail, dec_av_kernel<NC, 32>, (int)((n / 4 + 63) / 64), 64, 0, st, p);
                else launch_k(h, kPdlTail, dec_av_kernel<NC, 8>, g4, 256, 0, st, p);
            }
        } else {
            { ProfScope _ps(h, "dec_a", st); dec_a_kernel<NC><<<grid, 256, 0, st>>>(p); }
        }
        LAUNCH_COUNT();
        CUDA_TRY(h, cudaPeekAtLastError());
        h->stages["combine_l2_l3.0"] = {ws + L.t10, (size_t)B * 2 * NC * H4 * W4};
    }
    {
        DecBParams<NC> p{};
        p.tin = ws + L.t10; p.w = P + pk.c1_w;
        p.s = P + pk.c1_s; p.t = P + pk.c1_t; p.a = P + pk.c1_a;
        p.wt = P + pk.up2_w;
        p.s2 = P + pk.u2_s; p.t2 = P + pk.u2_t; p.a2 = P + pk.u2_a;
        p.comb = ws + L.comb; p.B = B; p.H4 = H4; p.W4 = W4;
        const size_t n = (size_t)B * H4 * W4;
        int grid = (int)((n + 255) / 256);
        if (grid > 8 * h->num_sms) grid = 8 * h->num_sms;
        // 4 pixels per thread (kernels_dec.cuh) when the vector loads / stores are aligned
        const bool b4 = NC == 5 && h->dec_impl != 0 && (W4 % 4 == 0) && (((uintptr_t)p.tin | (uintptr_t)p.comb) % 16 == 0);
        if (b4) {
            dim3 g4((W4 / 4 + 31) / 32, (H4 + 7) / 8, B);
            if constexpr (NC == 5) { ProfScope _ps(h, "dec_b", st); launch_k(h, kPdlTail, dec_b4_kernel<NC, 2>, g4, 256, 0, st, p); }   // deeper unrolling is slower here, also at batch 1
        } else {
            ProfScope _ps(h, "dec_b", st); dec_b_kernel<NC><<<grid, 256, 0, st>>>(p);
        }
        LAUNCH_COUNT();
        CUDA_TRY(h, cudaPeekAtLastError());
        h->stages["up_l2"] = {ws + L.comb, (size_t)B * NC * H2 * W2};
    }
    {
        DecCParams<NC> p{};
        p.comb = ws + L.comb; p.out0cat = ws + L.out0cat; p.w = P + pk.cv_w;
        p.s = P + pk.cv_s; p.t = P + pk.cv_t; p.a = P + pk.cv_a;
        p.wt = P + pk.clsT_w;
        p.logits = a->logits; p.mask = a->mask; p.prob_acc = a->prob_acc;
        p.prob_init = a->prob_init; p.mask_from_prob = a->mask_from_prob;
        p.B = B; p.H2 = H2; p.W2 = W2;
        // 4 pixels per thread (kernels_dec.cuh) when the vector stores are aligned; the one-pixel kernel otherwise / for 20 classes
        const bool vec_ok = NC == 5 && (W2 % 4 == 0) && (((uintptr_t)p.logits | (uintptr_t)p.prob_acc) % 16 == 0) && ((uintptr_t)p.mask % 8 == 0);
        if (vec_ok && h->dec_impl != 0) {
            dim3 g((W2 / 4 + 31) / 32, (H2 + 7) / 8, B);
            if constexpr (NC == 5) {
                ProfScope _ps(h, "dec_c", st);
                if (small_launch(h, (size_t)B * H2 * W2 / 4)) launch_k(h, kPdlLast, dec_c4_kernel<NC, 8>, g, 256, 0, st, p);
                else launch_k(h, kPdlLast, dec_c4_kernel<NC, 2>, g, 256, 0, st, p);
            }
        } else {
            dim3 g((W2 + 31) / 32, (H2 + 7) / 8, B);
            { ProfScope _ps(h, "dec_c", st); dec_c_kernel<NC><<<g, 256, 0, st>>>(p); }
        }
        LAUNCH_COUNT();
        CUDA_TRY(h, cudaPeekAtLastError());
    }
    return ESPNET_OK;
}

// The same tail for any class count (kernels_tail_generic.cuh): classes other than 5 / 20, or option "tail_impl" = 1.
int run_tail_generic(espnet_t* h, const espnet_forward_args* a, const Workspace& L, float* ws, cudaStream_t st) {
    const int B = a->B, H = a->H, W = a->W, nc = h->classes;
    const int H2 = H / 2, W2 = W / 2, H4 = H / 4, W4 = W / 4, H8 = H / 8, W8 = W / 8;
    const Packed& pk = h->pk;
    const float* P = h->dparams;
    const bool full = h->net == ESPNET_NET_FULL;
    auto grid1d = [&](size_t n) { int g = (int)((n + 255) / 256); if (g > 8 * h->num_sms) g = 8 * h->num_sms; return g < 1 ? 1 : g; };
    {
        Head3Params<0> p{};
        p.in = ws + L.out2cat; p.w = P + pk.cls_w; p.B = B; p.H8 = H8; p.W8 = W8;
        if (full) { p.bn_s = P + pk.br_s; p.bn_t = P + pk.br_t; p.wt = P + pk.up3_w; p.up_out = ws + L.up3; p.enc_out = nullptr; }
        else p.enc_out = a->logits ? a->logits : ws + L.enc;
        const size_t smem = ((size_t)256 * nc + (size_t)nc * nc * 4 + 2 * nc) * sizeof(float);
        int rc = set_smem(h, g_head3_kernel, 227 * 1024);
        if (rc) return rc;
        { ProfScope _ps(h, "head3", st); g_head3_kernel<<<grid1d((size_t)B * H8 * W8), 256, smem, st>>>(p, nc); }
        LAUNCH_COUNT();
        CUDA_TRY(h, cudaPeekAtLastError());
        if (!full) {
            h->stages["encoder.classifier"] = {p.enc_out, (size_t)B * nc * H8 * W8};
            if (a->mask) {
                dim3 g((W + 31) / 32, (H + 7) / 8, B);
                { ProfScope _ps(h, "upsample8_argmax", st); g_upsample8_argmax_kernel<<<g, 256, 0, st>>>(p.enc_out, nc, B, H8, W8, a->mask); }
                LAUNCH_COUNT();
                CUDA_TRY(h, cudaPeekAtLastError());
            }
            return ESPNET_OK;
        }
        h->stages["up_l3"] = {ws + L.up3, (size_t)B * nc * H4 * W4};
    }
    {
        DecAParams<0> p{};
        p.out1cat = ws + L.out1cat; p.up3 = ws + L.up3; p.w = P + pk.l3c_w;
        p.s = P + pk.c0_s; p.t = P + pk.c0_t; p.a = P + pk.c0_a;
        p.tout = ws + L.t10; p.B = B; p.H4 = H4; p.W4 = W4;
        const size_t smem = ((size_t)131 * nc + 6 * nc) * sizeof(float);
        int rc = set_smem(h, g_dec_a_kernel, 227 * 1024);
        if (rc) return rc;
        { ProfScope _ps(h, "dec_a", st); g_dec_a_kernel<<<grid1d((size_t)B * H4 * W4), 256, smem, st>>>(p, nc); }
        LAUNCH_COUNT();
        CUDA_TRY(h, cudaPeekAtLastError());
        h->stages["combine_l2_l3.0"] = {ws + L.t10, (size_t)B * 2 * nc * H4 * W4};
    }
    {
        DecBParams<0> p{};
        p.tin = ws + L.t10; p.w = P + pk.c1_w;
        p.s = P + pk.c1_s; p.t = P + pk.c1_t; p.a = P + pk.c1_a;
        p.wt = P + pk.up2_w;
        p.s2 = P + pk.u2_s; p.t2 = P + pk.u2_t; p.a2 = P + pk.u2_a;
        p.comb = ws + L.comb; p.B = B; p.H4 = H4; p.W4 = W4;
        const size_t smem = ((size_t)2 * nc * 9 * nc + (size_t)nc * nc * 4 + 6 * nc) * sizeof(float);
        int rc = set_smem(h, g_dec_b_kernel, 227 * 1024);
        if (rc) return rc;
        { ProfScope _ps(h, "dec_b", st); g_dec_b_kernel<<<grid1d((size_t)B * H4 * W4), 256, smem, st>>>(p, nc); }
        LAUNCH_COUNT();
        CUDA_TRY(h, cudaPeekAtLastError());
        h->stages["up_l2"] = {ws + L.comb, (size_t)B * nc * H2 * W2};
    }
    {
        DecCParams<0> p{};
        p.comb = ws + L.comb; p.out0cat = ws + L.out0cat; p.w = P + pk.cv_w;
        p.s = P + pk.cv_s; p.t = P + pk.cv_t; p.a = P + pk.cv_a;
        p.wt = P + pk.clsT_w;
        p.logits = a->logits; p.mask = a->mask; p.prob_acc = a->prob_acc;
        p.prob_init = a->prob_init; p.mask_from_prob = a->mask_from_prob;
        p.B = B; p.H2 = H2; p.W2 = W2;
        const size_t smem = ((size_t)(nc + 19) * 9 * nc + (size_t)nc * nc * 4 + 3 * nc) * sizeof(float);
        int rc = set_smem(h, g_dec_c_kernel, 227 * 1024);
        if (rc) return rc;
        dim3 g((W2 + 31) / 32, (H2 + 7) / 8, B);
        { ProfScope _ps(h, "dec_c", st); g_dec_c_kernel<<<g, 256, smem, st>>>(p, nc); }
        LAUNCH_COUNT();
        CUDA_TRY(h, cudaPeekAtLastError());
    }
    return ESPNET_OK;
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

int espnet_version(void) { return 100; }
unsigned long long espnet_launch_count(void) { return g_launches.load(); }

const char* espnet_last_error(const espnet_t* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int espnet_create(int classes, int p, int q, int net, int device, espnet_t** out) {
    if (!out) return fail(nullptr, ESPNET_EINVAL, "espnet_create: out is NULL");
    *out = nullptr;
    if (classes < 1 || classes > kMaxClasses)
        return fail(nullptr, ESPNET_ESHAPE, "espnet_create: classes must be in [1, " + std::to_string(kMaxClasses) + "] (5 and 20 run compile-time "
                                            "specialised tail kernels, every other count the run-time generic ones)");
    if (p < 1 || q < 1) return fail(nullptr, ESPNET_EINVAL, "espnet_create: p and q must be >= 1 (the reference forward needs one block per level)");
    if (net != ESPNET_NET_FULL && net != ESPNET_NET_ENCODER) return fail(nullptr, ESPNET_EINVAL, "espnet_create: bad net");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, ESPNET_ECUDA, std::string("espnet_create: no CUDA device (") + cudaGetErrorString(e) + "); there is no CPU fallback");
    if (device < 0 || device >= ndev) return fail(nullptr, ESPNET_EINVAL, "espnet_create: bad device index");
    cudaDeviceProp prop{};
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return fail(nullptr, ESPNET_ECUDA, std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e));
    if (prop.major != 10)
        return fail(nullptr, ESPNET_ECUDA, "espnet_create: this library holds sm_100a code only (B200); device is sm_" +
                                               std::to_string(prop.major) + std::to_string(prop.minor));
    espnet_t* h = new espnet_t();
    h->classes = classes; h->p = p; h->q = q; h->net = net; h->device = device;
    h->num_sms = prop.multiProcessorCount;
    if (const char* e = std::getenv("ESPNET_B200_PDL")) { const int v = std::atoi(e); h->pdl = v < 0 ? -1 : (v & kPdlAll); }   // default of the "pdl" option
    {
        DeviceGuard g(device);
        e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { delete h; return fail(nullptr, ESPNET_ECUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(e)); }
    }
    *out = h;
    return ESPNET_OK;
}

void espnet_destroy(espnet_t* h) {
    if (!h) return;
    DeviceGuard g(h->device);
    if (h->dparams) cudaFree(h->dparams);
    if (h->dparams_h) cudaFree(h->dparams_h);
    if (h->hb_in) cudaFree(h->hb_in);
    if (h->hb_mask) cudaFree(h->hb_mask);
    if (h->hb_ws) cudaFree(h->hb_ws);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    for (auto& gr : h->graphs) { if (gr.exec) cudaGraphExecDestroy(gr.exec); }
    for (auto& r : h->prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    delete h;
}

int espnet_set_mode(espnet_t* h, int mode) {
    if (!h) return ESPNET_EINVAL;
    if (mode != ESPNET_MODE_FP32 && mode != ESPNET_MODE_F16TC) return fail(h, ESPNET_EINVAL, "espnet_set_mode: unknown mode");
    h->mode = mode;
    return ESPNET_OK;
}

int espnet_set_option(espnet_t* h, const char* key, int value) {
    if (!h || !key) return ESPNET_EINVAL;
    if (std::strcmp(key, "branch_impl") == 0 && value >= 0 && value <= 2) { h->branch_impl = value; return ESPNET_OK; }
    if (std::strcmp(key, "fp32_impl") == 0 && value >= 0 && value <= 1) { h->fp32_impl = value; return ESPNET_OK; }
    if (std::strcmp(key, "dec_impl") == 0 && value >= 0 && value <= 1) { h->dec_impl = value; return ESPNET_OK; }
    if (std::strcmp(key, "dbg") == 0 && value >= 0 && value <= 7) {
        if (value != 0 && !ESPNET_TC_TIMING_EXPERIMENTS)
            return fail(h, ESPNET_EINVAL, "espnet_set_option: \"dbg\" (timing experiments with wrong results) needs a library built with -DESPNET_TC_TIMING_EXPERIMENTS=1");
        h->dbg = value;
        return ESPNET_OK;
    }
    if (std::strcmp(key, "down_impl") == 0 && value >= 0 && value <= 1) { h->down_impl = value; return ESPNET_OK; }
    if (std::strcmp(key, "tail_impl") == 0 && value >= 0 && value <= 1) { h->tail_impl = value; return ESPNET_OK; }
    if (std::strcmp(key, "pdl") == 0 && value >= -1 && value <= kPdlAll) { h->pdl = value; return ESPNET_OK; }
    if (std::strcmp(key, "l2_reverse") == 0 && value >= 0 && value <= 1) { h->l2_reverse = value; return ESPNET_OK; }
    if (std::strcmp(key, "tc_reduce") == 0 && value >= 0 && value <= 2) { h->tc_reduce = value; return ESPNET_OK; }
    return fail(h, ESPNET_EINVAL, std::string("espnet_set_option: unknown option or bad value: ") + key);
}

int espnet_pack_weights(espnet_t* h, const espnet_tensor_desc* tensors, int n) {
    if (!h || !tensors || n <= 0) return fail(h, ESPNET_EINVAL, "espnet_pack_weights: bad arguments");
    std::map<std::string, HostTensor> sd;
    for (int i = 0; i < n; ++i) {
        if (!tensors[i].name || !tensors[i].data || tensors[i].ndim < 0 || tensors[i].ndim > 4)
            return fail(h, ESPNET_EINVAL, "espnet_pack_weights: malformed tensor descriptor");
        HostTensor t;
        t.data = tensors[i].data;
        t.shape.assign(tensors[i].shape, tensors[i].shape + tensors[i].ndim);
        sd[tensors[i].name] = t;
    }
    const int NC = h->classes;
    const std::string e = h->net == ESPNET_NET_FULL ? "encoder." : "";
    Packer pk;
    pk.sd = &sd;
    Packed o{};                // built locally and swapped in only when everything succeeded: a failed repack leaves the old weights intact
    o.l2.assign(h->p, BlockW());
    o.l3.assign(h->q, BlockW());
    bool ok = true;
    {   // level1: [16][3][3][3] -> [(ci*9+tap)][16]
        const HostTensor* w = pk.get(e + "level1.conv.weight", {16, 3, 3, 3});
        if (w) {
            std::vector<float> v(27 * 16);
            for (int co = 0; co < 16; ++co)
                for (int k = 0; k < 27; ++k) v[k * 16 + co] = w->data[co * 27 + k];
            o.w1 = pk.add(v);
        } else ok = false;
    }
    ok = ok && pk.bn(e + "level1.bn", 16, o.l1_s, o.l1_t) && pk.vec(e + "level1.act.weight", 16, o.l1_a);
    ok = ok && pk.bn(e + "b1.bn", 19, o.b1_s, o.b1_t) && pk.vec(e + "b1.act.weight", 19, o.b1_a);
    ok = ok && pk.bn(e + "b2.bn", 131, o.b2_s, o.b2_t) && pk.vec(e + "b2.act.weight", 131, o.b2_a);
    ok = ok && pk.bn(e + "b3.bn", 256, o.b3_s, o.b3_t) && pk.vec(e + "b3.act.weight", 256, o.b3_a);
    ok = ok && pk.block(e + "level2_0", 19, 64, true, o.l2_0);
    for (int i = 0; i < h->p && ok; ++i) ok = pk.block(e + "level2." + std::to_string(i), 64, 64, false, o.l2[i]);
    ok = ok && pk.block(e + "level3_0", 131, 128, true, o.l3_0);
    for (int i = 0; i < h->q && ok; ++i) ok = pk.block(e + "level3." + std::to_string(i), 128, 128, false, o.l3[i]);
    ok = ok && pk.conv_ci_tap_co(e + "classifier.conv.weight", NC, 256, 1, o.cls_w);
    if (ok && h->net == ESPNET_NET_FULL) {
        ok = ok && pk.bn("br", NC, o.br_s, o.br_t);
        ok = ok && pk.raw("up_l3.0.weight", {NC, NC, 2, 2}, o.up3_w);
        ok = ok && pk.conv_ci_tap_co("level3_C.conv.weight", NC, 131, 1, o.l3c_w);
        ok = ok && pk.bn("combine_l2_l3.0.bn", 2 * NC, o.c0_s, o.c0_t) && pk.vec("combine_l2_l3.0.act.weight", 2 * NC, o.c0_a);
        ok = ok && pk.conv_ci_tap_co("combine_l2_l3.1.conv.weight", NC, 2 * NC, 3, o.c1_w);
        ok = ok && pk.bn("combine_l2_l3.1.bn", NC, o.c1_s, o.c1_t) && pk.vec("combine_l2_l3.1.act.weight", NC, o.c1_a);
        ok = ok && pk.raw("up_l2.0.weight", {NC, NC, 2, 2}, o.up2_w);
        ok = ok && pk.bn("up_l2.1.bn", NC, o.u2_s, o.u2_t) && pk.vec("up_l2.1.act.weight", NC, o.u2_a);
        ok = ok && pk.conv_ci_tap_co("conv.conv.weight", NC, NC + 19, 3, o.cv_w);
        ok = ok && pk.bn("conv.bn", NC, o.cv_s, o.cv_t) && pk.vec("conv.act.weight", NC, o.cv_a);
        ok = ok && pk.raw("classifier.weight", {NC, NC, 2, 2}, o.clsT_w);
    }
    if (!ok) return fail(h, ESPNET_EMISSING, "espnet_pack_weights: " + (pk.missing.empty() ? std::string("packing failed") : pk.missing));
    DeviceGuard g(h->device);
    // forwards may still be running on non-blocking streams (HostPipeline, own_stream) with the old blobs: wait for the device
    // before they are overwritten or freed; from here on a failure leaves the handle unpacked, never half-packed
    CUDA_TRY(h, cudaDeviceSynchronize());
    h->packed = false;
    for (auto& gr : h->graphs) { if (gr.exec) cudaGraphExecDestroy(gr.exec); }
    h->graphs.clear();
    if (h->dparams && h->nparams < pk.blob.size()) { cudaFree(h->dparams); h->dparams = nullptr; }
    if (!h->dparams) CUDA_TRY(h, cudaMalloc(&h->dparams, pk.blob.size() * sizeof(float)));
    h->nparams = pk.blob.size();
    // synchronous copy: the host buffers belong to the caller and may go away after we return
    CUDA_TRY(h, cudaMemcpy(h->dparams, pk.blob.data(), pk.blob.size() * sizeof(float), cudaMemcpyHostToDevice));
    if (h->dparams_h && h->nparams_h < pk.blob_h.size() * 2) { cudaFree(h->dparams_h); h->dparams_h = nullptr; }
    if (!h->dparams_h) CUDA_TRY(h, cudaMalloc(&h->dparams_h, pk.blob_h.size() * 2));
    h->nparams_h = pk.blob_h.size() * 2;
    CUDA_TRY(h, cudaMemcpy(h->dparams_h, pk.blob_h.data(), pk.blob_h.size() * 2, cudaMemcpyHostToDevice));
    h->pk = o;
    h->packed = true;
    return ESPNET_OK;
}

size_t espnet_workspace_bytes(const espnet_t* h, int B, int H, int W) {
    if (!h || B <= 0 || H <= 0 || W <= 0) return 0;
    return layout(h, B, H, W).total * sizeof(float);
}

int espnet_forward(espnet_t* h, const espnet_forward_args* a) {
    if (!h || !a) return fail(h, ESPNET_EINVAL, "espnet_forward: NULL argument");
    if (!h->packed) return fail(h, ESPNET_ESTATE, "espnet_forward: weights have not been packed (espnet_pack_weights)");
    if (!a->x || !a->workspace) return fail(h, ESPNET_EINVAL, "espnet_forward: x / workspace is NULL");
    if (a->B <= 0 || a->H <= 0 || a->W <= 0) return fail(h, ESPNET_ESHAPE, "espnet_forward: empty batch or crop");
    if (a->B > 65535) return fail(h, ESPNET_ESHAPE, "espnet_forward: B > 65535, split the batch");
    if ((a->H % 8) || (a->W % 8))
        return fail(h, ESPNET_ESHAPE, "espnet_forward: H and W must be multiples of 8 (the reference's concat needs it)");
    if (a->in_fmt < 0 || a->in_fmt > 2) return fail(h, ESPNET_EINVAL, "espnet_forward: unknown in_fmt");
    if (a->in_fmt == ESPNET_IN_U8_SLIDE && (!a->origins || a->slide_h <= 0 || a->slide_w <= 0))
        return fail(h, ESPNET_EINVAL, "espnet_forward: slide input needs origins and slide size");
    if (h->net == ESPNET_NET_ENCODER && a->prob_acc) return fail(h, ESPNET_EINVAL, "espnet_forward: prob_acc needs the full net");
    const Workspace L = layout(h, a->B, a->H, a->W);
    if (a->workspace_bytes < L.total * sizeof(float)) return fail(h, ESPNET_ESTATE, "espnet_forward: workspace too small");
    if (((uintptr_t)a->workspace & 255) != 0) return fail(h, ESPNET_EINVAL, "espnet_forward: workspace must be 256-byte aligned (TMA tensor maps, 16 B vector accesses)");
    if ((size_t)a->H * a->W > ((size_t)1 << 30)) return fail(h, ESPNET_ESHAPE, "espnet_forward: crop too large for 32-bit plane offsets");
    // tensor-core kernels address channel planes as 32-bit byte offsets inside one crop; the tightest one is the level-3
    // 3x3-s2 reduce loader: 144 channels x (H/4 x W/4) x 4 B < 2^32  ->  H*W < 1.19e8; 2^26 (8192 x 8192) keeps a margin
    if ((h->mode == ESPNET_MODE_F16TC || h->fp32_impl != 0) && (size_t)a->H * a->W > ((size_t)1 << 26))
        return fail(h, ESPNET_ESHAPE, "espnet_forward: crop larger than 2^26 pixels: the tensor-core kernels use 32-bit channel offsets "
                                      "(use set_option(\"fp32_impl\", 0) or tile the crop)");

    DeviceGuard g(h->device);
    h->pdl_eff = pdl_mask_for(h, (long long)a->B * a->H * a->W);
    cudaStream_t st = (cudaStream_t)a->stream;
    float* ws = (float*)a->workspace;
    const float* P = h->dparams;
    const Packed& pk = h->pk;
    const int B = a->B, H = a->H, W = a->W;
    const int H2 = H / 2, W2 = W / 2, H4 = H / 4, W4 = W / 4, H8 = H / 8, W8 = W / 8;
    h->stages.clear();

    const bool tcm = h->mode == ESPNET_MODE_F16TC;
    __half* o1h = reinterpret_cast<__half*>(ws + L.o1);

    // ---- S1 stem -----------------------------------------------------------------------------------
    {
        StemParams p{};
        p.x = a->x; p.in_fmt = a->in_fmt; p.B = B; p.H = H; p.W = W;
        for (int c = 0; c < 3; ++c) { p.mean[c] = a->mean[c]; p.stdv[c] = a->std_[c]; }
        p.origins = a->origins; p.slide_h = a->slide_h; p.slide_w = a->slide_w;
        p.w1 = P + pk.w1;
        p.l1_s = P + pk.l1_s; p.l1_t = P + pk.l1_t; p.l1_a = P + pk.l1_a;
        p.b1_s = P + pk.b1_s; p.b1_t = P + pk.b1_t; p.b1_a = P + pk.b1_a;
        p.b2_s = P + pk.b2_s; p.b2_t = P + pk.b2_t; p.b2_a = P + pk.b2_a;
        p.out0cat = ws + L.out0cat; p.out1cat = ws + L.out1cat;
        dim3 grid((W2 + kStemTW - 1) / kStemTW, (H2 + kStemTH - 1) / kStemTH, B);
        {
            ProfScope _ps(h, "stem", st);
            if (a->in_fmt == 0) launch_k(h, kPdlStem, stem_kernel<0>, grid, 256, 0, st, p);
            else if (a->in_fmt == 1) launch_k(h, kPdlStem, stem_kernel<1>, grid, 256, 0, st, p);
            else launch_k(h, kPdlStem, stem_kernel<2>, grid, 256, 0, st, p);
        }
        LAUNCH_COUNT();
        CUDA_TRY(h, cudaPeekAtLastError());
        h->stages["b1"] = {ws + L.out0cat, (size_t)B * 19 * H2 * W2};
    }
    int rc;
    // one DownSamplerB / ESP block = reduce + branch stage, on CUDA cores (fp32) or tensor cores (fp16 operands)
    auto block_l2 = [&](const BlockW& bw, bool down, const float* in, float* out, float* out2, int c2_off) -> int {
        int r;
        if (tcm) {
            r = down ? (h->tc_reduce ? run_reduce3x3_tc<19, 16, 2>(h, in, bw, o1h, B, H2, W2, st) : run_reduce3x3_f16<19, 12, 2>(h, in, bw.c1, o1h, B, H2, W2, st))
                     : (h->tc_reduce ? run_reduce1x1_tc<64, 16, 2>(h, in, bw, o1h, B, H4, W4, st) : run_reduce1x1_f16<64, 12, 2>(h, in, bw.c1, o1h, B, H4, W4, st));
            if (r) return r;
            return run_branch_tc<2, 16, 16, 12>(h, bw, o1h, down ? nullptr : in, out, out2, 131, c2_off, pk.b2_s, pk.b2_t, pk.b2_a, B, H4, W4, st);
        }
        if (h->fp32_impl == 1) {   // fp32-equivalent on tensor cores: 3-term fp16 operand splits, fp32 accumulation
            r = down ? (h->tc_reduce ? run_reduce3x3_tc<19, 16, 2, true>(h, in, bw, o1h, B, H2, W2, st) : run_reduce3x3_f16<19, 12, 2, true>(h, in, bw.c1, o1h, B, H2, W2, st))
                     : run_reduce1x1_tc<64, 16, 2, true>(h, in, bw, o1h, B, H4, W4, st);
            if (r) return r;
            return run_branch_tc<2, 16, 16, 12, true>(h, bw, o1h, down ? nullptr : in, out, out2, 131, c2_off, pk.b2_s, pk.b2_t, pk.b2_a, B, H4, W4, st);
        }
        r = down ? run_reduce3x3<19, 12>(h, in, bw.c1, ws + L.o1, B, H2, W2, st) : run_reduce1x1<64, 12>(h, in, bw.c1, ws + L.o1, B, H4, W4, st);
        if (r) return r;
        return run_branch<12, 16, 12>(h, bw, ws + L.o1, down ? nullptr : in, out, out2, 131, c2_off, pk.b2_s, pk.b2_t, pk.b2_a, B, H4, W4, st);
    };
    auto block_l3 = [&](const BlockW& bw, bool down, const float* in, float* out, float* out2, int c2_off) -> int {
        int r;
        if (tcm) {
            r = down ? (h->tc_reduce ? run_reduce3x3_tc<131, 32, 4>(h, in, bw, o1h, B, H4, W4, st) : run_reduce3x3_f16<131, 25, 4>(h, in, bw.c1, o1h, B, H4, W4, st))
                     : (h->tc_reduce ? run_reduce1x1_tc<128, 32, 4>(h, in, bw, o1h, B, H8, W8, st) : run_reduce1x1_f16<128, 25, 4>(h, in, bw.c1, o1h, B, H8, W8, st));
            if (r) return r;
            return run_branch_tc<4, 32, 28, 25>(h, bw, o1h, down ? nullptr : in, out, out2, 256, c2_off, pk.b3_s, pk.b3_t, pk.b3_a, B, H8, W8, st);
        }
        if (h->fp32_impl == 1) {
            r = down ? (h->tc_reduce ? run_reduce3x3_tc<131, 32, 4, true>(h, in, bw, o1h, B, H4, W4, st) : run_reduce3x3_f16<131, 25, 4, true>(h, in, bw.c1, o1h, B, H4, W4, st))
                     : run_reduce1x1_tc<128, 32, 4, true>(h, in, bw, o1h, B, H8, W8, st);
            if (r) return r;
            return run_branch_tc<4, 32, 28, 25, true>(h, bw, o1h, down ? nullptr : in, out, out2, 256, c2_off, pk.b3_s, pk.b3_t, pk.b3_a, B, H8, W8, st);
        }
        r = down ? run_reduce3x3<131, 25>(h, in, bw.c1, ws + L.o1, B, H4, W4, st) : run_reduce1x1<128, 25>(h, in, bw.c1, ws + L.o1, B, H8, W8, st);
        if (r) return r;
        return run_branch<25, 28, 25>(h, bw, ws + L.o1, down ? nullptr : in, out, out2, 256, c2_off, pk.b3_s, pk.b3_t, pk.b3_a, B, H8, W8, st);
    };
    // ---- S2/S3 level 2 -----------------------------------------------------------------------------
    rc = block_l2(pk.l2_0, true, ws + L.out0cat, ws + L.l2a, ws + L.out1cat, 64);
    if (rc) return rc;
    {
        float* cur = ws + L.l2a;
        float* nxt = ws + L.l2b;
        std::string name_cur = "level2_0", name_nxt;
        for (int i = 0; i < h->p; ++i) {
            const bool last = i == h->p - 1;
            rc = block_l2(pk.l2[i], false, cur, last ? nullptr : nxt, last ? ws + L.out1cat : nullptr, 0);
            if (rc) return rc;
            if (!last) name_nxt = "level2." + std::to_string(i);
            float* t = cur; cur = nxt; nxt = t;
            std::swap(name_cur, name_nxt);
        }
        // block outputs that survived the ping-pong (parity taps)
        if (!name_cur.empty()) h->stages[name_cur] = {cur, (size_t)B * 64 * H4 * W4};
        if (!name_nxt.empty()) h->stages[name_nxt] = {nxt, (size_t)B * 64 * H4 * W4};
        h->stages["b2"] = {ws + L.out1cat, (size_t)B * 131 * H4 * W4};
    }
    // ---- S5/S6 level 3 -----------------------------------------------------------------------------
    rc = block_l3(pk.l3_0, true, ws + L.out1cat, ws + L.l3a, ws + L.out2cat, 0);
    if (rc) return rc;
    {
        float* cur = ws + L.l3a;
        float* nxt = ws + L.l3b;
        std::string name_cur = "level3_0", name_nxt;
        for (int i = 0; i < h->q; ++i) {
            const bool last = i == h->q - 1;
            rc = block_l3(pk.l3[i], false, cur, last ? nullptr : nxt, last ? ws + L.out2cat : nullptr, 128);
            if (rc) return rc;
            if (!last) name_nxt = "level3." + std::to_string(i);
            float* t = cur; cur = nxt; nxt = t;
            std::swap(name_cur, name_nxt);
        }
        if (!name_cur.empty()) h->stages[name_cur] = {cur, (size_t)B * 128 * H8 * W8};
        if (!name_nxt.empty()) h->stages[name_nxt] = {nxt, (size_t)B * 128 * H8 * W8};
        h->stages["b3"] = {ws + L.out2cat, (size_t)B * 256 * H8 * W8};
    }
    // ---- S7..S10 heads / decoder -------------------------------------------------------------------
    if (h->classes == 5 && !h->tail_impl) return run_tail<5>(h, a, L, ws, st);
    if (h->classes == 20 && !h->tail_impl) return run_tail<20>(h, a, L, ws, st);
    return run_tail_generic(h, a, L, ws, st);
}

// One forward with FIXED buffers recorded into a CUDA graph: ~30 kernel launches become one cudaGraphLaunch, which is what the
// reference's per-crop loop (VisualizeResults_iou.py:100-129, batch 1) needs -- at batch 1 the forward is launch-latency bound.
int espnet_graph_capture(espnet_t* h, const espnet_forward_args* a, int* graph_id) {
    if (!h || !a || !graph_id) return fail(h, ESPNET_EINVAL, "espnet_graph_capture: NULL argument");
    if (h->profiling) return fail(h, ESPNET_ESTATE, "espnet_graph_capture: switch profiling off first (events cannot be captured per kernel)");
    DeviceGuard g(h->device);
    espnet_forward_args b = *a;
    b.stream = h->own_stream;
    // an uncaptured pass first: raises the shared-memory limits (cudaFuncSetAttribute) and validates the arguments outside the capture
    int rc = espnet_forward(h, &b);
    if (rc) return rc;
    CUDA_TRY(h, cudaStreamSynchronize(h->own_stream));
    CUDA_TRY(h, cudaStreamBeginCapture(h->own_stream, cudaStreamCaptureModeRelaxed));
    rc = espnet_forward(h, &b);
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(h->own_stream, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess || !graph) return fail(h, ESPNET_ECUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
    espnet_handle::GraphRec rec;
    e = cudaGraphInstantiate(&rec.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return fail(h, ESPNET_ECUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
    h->graphs.push_back(rec);
    *graph_id = (int)h->graphs.size() - 1;
    return ESPNET_OK;
}

int espnet_graph_launch(espnet_t* h, int graph_id, void* stream) {
    if (!h || graph_id < 0 || graph_id >= (int)h->graphs.size() || !h->graphs[graph_id].exec)
        return fail(h, ESPNET_EINVAL, "espnet_graph_launch: unknown graph (a repack of the weights drops all captured graphs)");
    DeviceGuard g(h->device);
    CUDA_TRY(h, cudaGraphLaunch(h->graphs[graph_id].exec, (cudaStream_t)stream));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return ESPNET_OK;
}

int espnet_graph_destroy(espnet_t* h, int graph_id) {
    if (!h || graph_id < 0 || graph_id >= (int)h->graphs.size()) return ESPNET_EINVAL;
    DeviceGuard g(h->device);
    if (h->graphs[graph_id].exec) { cudaGraphExecDestroy(h->graphs[graph_id].exec); h->graphs[graph_id].exec = nullptr; }
    return ESPNET_OK;
}

// Hardware self-test of the tcgen05 operand convention (kernels_tc.cuh: tc_selftest_kernel): random fp16 A / B,
// one tap shifted by (dy, dx), compared with a double-precision host reference.  Returns the max abs error.
int espnet_tc_selftest(int device, int nkc, int nout, int dy, int dx, int use_tma, float* max_abs_err) {
    if (!max_abs_err || (nkc != 2 && nkc != 4) || (nout != 16 && nout != 32) || dy < -16 || dy > 16 || dx < -16 || dx > 16)
        return fail(nullptr, ESPNET_EINVAL, "espnet_tc_selftest: bad arguments");
    DeviceGuard g(device);
    // a small map (24 x 20) whose 48 x 48 box around the tile origin (4, 6) hangs over every edge
    const int Hm = 24, Wm = 20, ox = 4, oy = 6;
    std::vector<__half> amap((size_t)nkc * Hm * Wm * 8), apad((size_t)nkc * kTcBox * kTcBox * 8), bw((size_t)nkc * nout * 8);
    std::vector<float> amap_f(amap.size()), bw_f(bw.size());
    uint32_t seed = 12345u + (uint32_t)(nkc * 131 + nout * 17 + (dy + 16) * 37 + (dx + 16));
    auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return ((seed >> 8) & 0xFFFF) / 65536.0f - 0.5f; };
    for (size_t i = 0; i < amap.size(); ++i) { amap[i] = __float2half_rn(rnd()); amap_f[i] = __half2float(amap[i]); }
    for (size_t i = 0; i < bw.size(); ++i) { bw[i] = __float2half_rn(rnd()); bw_f[i] = __half2float(bw[i]); }
    auto at = [&](int kc, int y, int x, int j) -> float {   // zero outside the map
        if (y < 0 || y >= Hm || x < 0 || x >= Wm) return 0.f;
        return amap_f[(((size_t)kc * Hm + y) * Wm + x) * 8 + j];
    };
    for (int kc = 0; kc < nkc; ++kc)
        for (int r = 0; r < kTcBox; ++r)
            for (int c = 0; c < kTcBox; ++c)
                for (int j = 0; j < 8; ++j)
                    apad[(((size_t)kc * kTcBox + r) * kTcBox + c) * 8 + j] = __float2half_rn(at(kc, oy - kTcHalo + r, ox - kTcHalo + c, j));
    __half *d_map = nullptr, *d_pad = nullptr, *d_bw = nullptr;
    float* d_out = nullptr;
    auto cleanup = [&]() { cudaFree(d_map); cudaFree(d_pad); cudaFree(d_bw); cudaFree(d_out); };
#define ST_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { cleanup(); return fail(nullptr, ESPNET_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); } } while (0)
    ST_TRY(cudaMalloc(&d_map, amap.size() * 2));
    ST_TRY(cudaMalloc(&d_pad, apad.size() * 2));
    ST_TRY(cudaMalloc(&d_bw, bw.size() * 2));
    ST_TRY(cudaMalloc(&d_out, (size_t)128 * nout * 4));
    ST_TRY(cudaMemcpy(d_map, amap.data(), amap.size() * 2, cudaMemcpyHostToDevice));
    ST_TRY(cudaMemcpy(d_pad, apad.data(), apad.size() * 2, cudaMemcpyHostToDevice));
    ST_TRY(cudaMemcpy(d_bw, bw.data(), bw.size() * 2, cudaMemcpyHostToDevice));
    ST_TRY(cudaMemset(d_out, 0xFF, (size_t)128 * nout * 4));
    CUtensorMap map;
    int rc = make_o1h_map(nullptr, &map, d_map, 1, nkc, Hm, Wm, kTcBox, kTcBox, nkc);
    if (rc) { cleanup(); return rc; }
    const size_t smem = (size_t)4 * kTcPlaneBytes + 4 * 32 * 16 + 64;
    if (nout == 32) {
        ST_TRY(cudaFuncSetAttribute(tc_selftest_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc_selftest_kernel<32><<<1, 128, smem>>>(map, d_pad, d_bw, d_out, nkc, dy, dx, use_tma, ox, oy);
    } else {
        ST_TRY(cudaFuncSetAttribute(tc_selftest_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc_selftest_kernel<16><<<1, 128, smem>>>(map, d_pad, d_bw, d_out, nkc, dy, dx, use_tma, ox, oy);
    }
    LAUNCH_COUNT();
    ST_TRY(cudaGetLastError());
    ST_TRY(cudaDeviceSynchronize());
    std::vector<float> out((size_t)128 * nout);
    ST_TRY(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
#undef ST_TRY
    cleanup();
    double worst = 0.0;
    for (int l = 0; l < 128; ++l)
        for (int n = 0; n < nout; ++n) {
            double ref = 0.0;
            for (int kc = 0; kc < nkc; ++kc)
                for (int j = 0; j < 8; ++j)
                    ref += (double)at(kc, oy + dy + l / 8, ox + dx + l % 8, j) * (double)bw_f[((size_t)kc * nout + n) * 8 + j];
            const double got = (double)out[(size_t)l * nout + n];
            const double e = std::isfinite(got) ? std::fabs(got - ref) : 1e30;
            if (e > worst) worst = e;
        }
    *max_abs_err = (float)worst;
    return ESPNET_OK;
}

int espnet_set_profiling(espnet_t* h, int on) {
    if (!h) return ESPNET_EINVAL;
    DeviceGuard g(h->device);
    for (auto& r : h->prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    h->prof.clear();
    h->profiling = on != 0;
    return ESPNET_OK;
}

int espnet_get_profile(espnet_t* h, char* names, float* total_ms, int* launches, int max_entries, int* n_entries) {
    if (!h || !names || !total_ms || !launches || !n_entries || max_entries <= 0) return fail(h, ESPNET_EINVAL, "espnet_get_profile: bad arguments");
    DeviceGuard g(h->device);
    std::vector<std::string> order;
    std::map<std::string, std::pair<double, int>> agg;
    for (auto& r : h->prof) {
        CUDA_TRY(h, cudaEventSynchronize(r.e1));
        float ms = 0.f;
        CUDA_TRY(h, cudaEventElapsedTime(&ms, r.e0, r.e1));
        if (!agg.count(r.name)) order.push_back(r.name);
        agg[r.name].first += ms;
        agg[r.name].second += 1;
    }
    int n = 0;
    for (auto& k : order) {
        if (n >= max_entries) break;
        std::snprintf(names + 64 * n, 64, "%s", k.c_str());
        total_ms[n] = (float)agg[k].first;
        launches[n] = agg[k].second;
        ++n;
    }
    *n_entries = n;
    return ESPNET_OK;
}

int espnet_read_stage(espnet_t* h, const char* stage, float* dst, size_t dst_elems, size_t* count, void* stream) {
    if (!h || !stage) return fail(h, ESPNET_EINVAL, "espnet_read_stage: NULL argument");
    auto it = h->stages.find(stage);
    if (it == h->stages.end()) return fail(h, ESPNET_EINVAL, std::string("espnet_read_stage: unknown stage '") + stage + "'");
    if (count) *count = it->second.count;
    if (!dst) return ESPNET_OK;
    if (dst_elems < it->second.count) return fail(h, ESPNET_ESTATE, "espnet_read_stage: destination too small");
    DeviceGuard g(h->device);
    CUDA_TRY(h, cudaMemcpyAsync(dst, it->second.ptr, it->second.count * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return ESPNET_OK;
}

int espnet_segment_host(espnet_t* h, const uint8_t* crops_host, int B, int H, int W, const float mean[3], const float std_[3],
                        uint8_t* masks_host) {
    if (!h || !crops_host || !masks_host || !mean || !std_) return fail(h, ESPNET_EINVAL, "espnet_segment_host: NULL argument");
    if (B <= 0 || H <= 0 || W <= 0 || (H % 8) || (W % 8)) return fail(h, ESPNET_ESHAPE, "espnet_segment_host: bad shape");
    DeviceGuard g(h->device);
    const size_t in_sz = (size_t)B * H * W * 3, mask_sz = (size_t)B * H * W, ws_sz = espnet_workspace_bytes(h, B, H, W);
    auto grow = [&](void*& ptr, size_t& have, size_t need) -> cudaError_t {
        if (have >= need) return cudaSuccess;
        if (ptr) cudaFree(ptr);
        ptr = nullptr; have = 0;
        cudaError_t e = cudaMalloc(&ptr, need);
        if (e == cudaSuccess) have = need;
        return e;
    };
    CUDA_TRY(h, grow(h->hb_in, h->hb_in_sz, in_sz));
    CUDA_TRY(h, grow(h->hb_mask, h->hb_mask_sz, mask_sz));
    CUDA_TRY(h, grow(h->hb_ws, h->hb_ws_sz, ws_sz));
    CUDA_TRY(h, cudaMemcpyAsync(h->hb_in, crops_host, in_sz, cudaMemcpyHostToDevice, h->own_stream));
    espnet_forward_args a{};
    a.x = h->hb_in; a.in_fmt = ESPNET_IN_U8_BGR_HWC; a.B = B; a.H = H; a.W = W;
    for (int c = 0; c < 3; ++c) { a.mean[c] = mean[c]; a.std_[c] = std_[c]; }
    a.mask = (uint8_t*)h->hb_mask;
    a.workspace = h->hb_ws; a.workspace_bytes = h->hb_ws_sz; a.stream = h->own_stream;
    int rc = espnet_forward(h, &a);
    if (rc) return rc;
    CUDA_TRY(h, cudaMemcpyAsync(masks_host, h->hb_mask, mask_sz, cudaMemcpyDeviceToHost, h->own_stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->own_stream));
    return ESPNET_OK;
}

// ---------------------------------------------------------------------------------- stitching
int espnet_stitch_boxes(uint8_t* slide_mask, int slide_h, int slide_w, int y_limit, const int32_t* boxes,
                        const int64_t* mask_offsets, const uint8_t* masks, int n_boxes, void* stream) {
    PtrDeviceGuard _dg(slide_mask);
    if (!slide_mask || !boxes || !mask_offsets || !masks || slide_h <= 0 || slide_w <= 0 || n_boxes < 0) return ESPNET_EINVAL;
    if (((uintptr_t)slide_mask & 3) != 0) return ESPNET_EINVAL;   // 32-bit merge words
    if (n_boxes == 0) return ESPNET_OK;
    dim3 grid(n_boxes, 32);
    stitch_boxes_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(slide_mask, slide_h, slide_w, y_limit, boxes,
                                                               (const long long*)mask_offsets, masks, n_boxes);
    LAUNCH_COUNT();
    if (((size_t)slide_h * slide_w) & 3) {
        stitch_boxes_tail_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(slide_mask, slide_h, slide_w, y_limit, boxes, (const long long*)mask_offsets,
                                                                    masks, n_boxes);
        LAUNCH_COUNT();
    }
    return cudaPeekAtLastError() == cudaSuccess ? ESPNET_OK : ESPNET_ECUDA;
}

static int stitch_grid_launch(uint8_t* out, int out_y0, int out_rows, int slide_h, int slide_w, int y_limit, const uint8_t* tile_masks, int n_x,
                              int n_y, int win_x, int win_y, int stride_x, int stride_y, long long k0, long long k1, int overwrite, void* stream) {
    PtrDeviceGuard _dg(tile_masks);      // `out` may be a peer GPU's memory (P2P band placement): launch where the tiles live
    if (!out || !tile_masks || slide_h <= 0 || slide_w <= 0 || n_x <= 0 || n_y <= 0 || win_x <= 0 || win_y <= 0 || stride_x <= 0 ||
        stride_y <= 0 || k0 < 0 || k1 < k0 || k1 > (long long)n_x * n_y || out_y0 < 0 || out_rows < 0)
        return ESPNET_EINVAL;
    if (k1 == k0 || out_rows == 0) return ESPNET_OK;
    long long ylo = (k0 / n_x) * stride_y;
    long long yhi = ((k1 - 1) / n_x) * stride_y + win_y;
    if (yhi > slide_h) yhi = slide_h;
    if (yhi > y_limit) yhi = y_limit;
    if (ylo < out_y0) ylo = out_y0;
    if (yhi > (long long)out_y0 + out_rows) yhi = (long long)out_y0 + out_rows;
    if (yhi <= ylo) return ESPNET_OK;
    int gx = (slide_w / 4 + 255) / 256;  // four pixels per thread on the aligned path (the byte path strides)
    if (gx > 64) gx = 64;
    if (gx < 1) gx = 1;
    long long gy = yhi - ylo;
    if (gy > 16384) gy = 16384;          // the kernel strides over the rows: no 65535-row limit on the slide
    dim3 grid(gx, (unsigned)gy);
    stitch_grid_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(out, out_y0, out_rows, slide_h, slide_w, y_limit, tile_masks, n_x, n_y, win_x,
                                                              win_y, stride_x, stride_y, (int)k0, (int)k1, overwrite);
    LAUNCH_COUNT();
    return cudaPeekAtLastError() == cudaSuccess ? ESPNET_OK : ESPNET_ECUDA;
}

int espnet_stitch_grid(uint8_t* slide_mask, int slide_h, int slide_w, int y_limit, const uint8_t* tile_masks, int n_x, int n_y,
                       int win_x, int win_y, int stride_x, int stride_y, int tile_row0, int tile_rows, void* stream) {
    if (tile_row0 < 0 || tile_rows < 0 || tile_row0 + tile_rows > n_y) return ESPNET_EINVAL;
    return stitch_grid_launch(slide_mask, 0, slide_h, slide_h, slide_w, y_limit, tile_masks, n_x, n_y, win_x, win_y, stride_x, stride_y,
                              (long long)tile_row0 * n_x, (long long)(tile_row0 + tile_rows) * n_x, 0, stream);
}

int espnet_stitch_grid_band(uint8_t* band_mask, int band_y0, int band_rows, int slide_h, int slide_w, int y_limit, const uint8_t* tile_masks,
                            int n_x, int n_y, int win_x, int win_y, int stride_x, int stride_y, int tile_k0, int tile_k1, int overwrite,
                            void* stream) {
    if (band_y0 + (long long)band_rows > slide_h) return ESPNET_EINVAL;
    return stitch_grid_launch(band_mask, band_y0, band_rows, slide_h, slide_w, y_limit, tile_masks, n_x, n_y, win_x, win_y, stride_x,
                              stride_y, tile_k0, tile_k1, overwrite != 0, stream);
}

// ---- peer-visible buffers (CUDA IPC over NVLink): rank 0 of a multi-GPU run allocates the slide mask here and exports it;
// the other ranks (processes) map it and their stitch kernels write their bands straight into it (espnet_stitch_grid_band,
// overwrite form).  The only device memory besides the packed weights that this library ever allocates, and only on request.
int espnet_peer_alloc(size_t bytes, int device, void** dptr, uint8_t handle64[64]) {
    if (!dptr || !handle64 || bytes == 0) return ESPNET_EINVAL;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    DeviceGuard g(device);
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return ESPNET_ECUDA; }
    cudaIpcMemHandle_t hd;
    if (cudaMemset(p, 0, bytes) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess || cudaIpcGetMemHandle(&hd, p) != cudaSuccess) {
        cudaGetLastError(); cudaFree(p); return ESPNET_ECUDA;
    }
    std::memcpy(handle64, &hd, 64);
    *dptr = p;
    return ESPNET_OK;
}

// maps a buffer exported by ANOTHER process for kernels running on `device` (peer access is enabled lazily by the driver)
int espnet_peer_open(const uint8_t handle64[64], int device, void** dptr) {
    if (!dptr || !handle64) return ESPNET_EINVAL;
    DeviceGuard g(device);
    cudaIpcMemHandle_t hd;
    std::memcpy(&hd, handle64, 64);
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return ESPNET_ECUDA; }
    *dptr = p;
    return ESPNET_OK;
}

int espnet_peer_close(void* dptr, int device) {
    if (!dptr) return ESPNET_EINVAL;
    DeviceGuard g(device);
    if (cudaIpcCloseMemHandle(dptr) != cudaSuccess) { cudaGetLastError(); return ESPNET_ECUDA; }
    return ESPNET_OK;
}

int espnet_peer_free(void* dptr, int device) {
    if (!dptr) return ESPNET_EINVAL;
    DeviceGuard g(device);
    if (cudaFree(dptr) != cudaSuccess) { cudaGetLastError(); return ESPNET_ECUDA; }
    return ESPNET_OK;
}

int espnet_max_merge_u8(uint8_t* dst, const uint8_t* src, size_t n, void* stream) {
    PtrDeviceGuard _dg(dst);
    if (!dst || !src) return ESPNET_EINVAL;
    if (n == 0) return ESPNET_OK;
    size_t g = (n / 16 + 255) / 256;
    if (g > 1184) g = 1184;
    if (g < 1) g = 1;
    max_merge_u8_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(dst, src, n);
    LAUNCH_COUNT();
    return cudaPeekAtLastError() == cudaSuccess ? ESPNET_OK : ESPNET_ECUDA;
}

// Source index LUT of the reference's per-window /8 nearest resize + paste along one axis
// (eval_wsi_segmentation.py:180-195, 225-240): windows [k*ws, min((k+1)*ws, len)), each resized to
// int(w/8) samples with cv2's INTER_NEAREST rule sx = min(floor(dx * (w / int(w/8))), w-1) and pasted
// at [xmin//8, xmax//8).  `limit`: windows whose max exceeds it are skipped (the y-vs-width quirk, :194);
// pass slide_len for the x axis.  Entries nobody writes are -1 (stay 0 in the output).
int espnet_ds8_lut(int slide_len, int ws, int limit, int32_t* lut, int lut_len) {
    if (!lut || slide_len <= 0 || ws <= 0 || (ws % 8) != 0 || lut_len != (int)((double)slide_len / 8)) return ESPNET_EINVAL;
    for (int i = 0; i < lut_len; ++i) lut[i] = -1;
    for (int k = 0; k <= slide_len / ws; ++k) {
        const int lo = k * ws;
        const int hi = (k == slide_len / ws) ? slide_len : (k + 1) * ws;
        if (hi > limit) continue;
        const int w = hi - lo;
        if (w <= 0) continue;
        const int dw = (int)((double)w / 8);
        if (dw <= 0) continue;
        const int d0 = lo / 8, d1 = hi / 8;
        if (d1 - d0 != dw) return ESPNET_ESHAPE;   // numpy would refuse the paste in the reference too
        const double scale = (double)w / (double)dw;
        for (int i = 0; i < dw; ++i) {
            int s = (int)std::floor((double)i * scale);
            if (s > w - 1) s = w - 1;
            lut[d0 + i] = lo + s;
        }
    }
    return ESPNET_OK;
}

int espnet_downsample_lut(const uint8_t* level0, int slide_h, int slide_w, uint8_t* ds, int ds_h, int ds_w, const int32_t* ysrc_dev,
                          const int32_t* xsrc_dev, void* stream) {
    PtrDeviceGuard _dg(level0);
    if (!level0 || !ds || !ysrc_dev || !xsrc_dev || slide_h <= 0 || slide_w <= 0 || ds_h <= 0 || ds_w <= 0) return ESPNET_EINVAL;
    if (ds_h > 65535) return ESPNET_ESHAPE;
    int gx = (ds_w + 255) / 256;
    if (gx > 64) gx = 64;
    dim3 grid(gx, ds_h);
    downsample_lut_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(level0, slide_w, ds, ds_h, ds_w, ysrc_dev, xsrc_dev);
    LAUNCH_COUNT();
    return cudaPeekAtLastError() == cudaSuccess ? ESPNET_OK : ESPNET_ECUDA;
}

// ---------------------------------------------------------------------------------- front-end / render
// cv2.resize INTER_LINEAR source coordinates (OpenCV's generic resizeLinear): fx = (float)((d + 0.5) * scale - 0.5),
// sx = floor(fx), fx -= sx, clamped at both ends; scale = 1 / ((double)dst / src).
int espnet_bilinear_lut(int src_len, int dst_len, int32_t* idx, float* wgt) {
    if (!idx || !wgt || src_len <= 0 || dst_len <= 0) return ESPNET_EINVAL;
    const double scale = 1.0 / ((double)dst_len / (double)src_len);
    for (int d = 0; d < dst_len; ++d) {
        float fx = (float)(((double)d + 0.5) * scale - 0.5);
        int sx = (int)std::floor(fx);
        fx -= (float)sx;
        if (sx < 0) { fx = 0.f; sx = 0; }
        if (sx >= src_len - 1) { fx = 0.f; sx = src_len - 1; }
        idx[d] = sx;
        wgt[d] = fx;
    }
    return ESPNET_OK;
}

// cv2.resize INTER_NEAREST source index: min(floor(d * (src / dst)), src - 1)
int espnet_nearest_lut(int src_len, int dst_len, int32_t* idx) {
    if (!idx || src_len <= 0 || dst_len <= 0) return ESPNET_EINVAL;
    const double scale = (double)src_len / (double)dst_len;
    for (int d = 0; d < dst_len; ++d) {
        int s = (int)std::floor((double)d * scale);
        idx[d] = s > src_len - 1 ? src_len - 1 : s;
    }
    return ESPNET_OK;
}

int espnet_preprocess_resize(const uint8_t* crops, int B, int h, int w, const float mean[3], const float std_[3], const int32_t* xs_dev,
                             const float* xf_dev, const int32_t* ys_dev, const float* yf_dev, float* out, int H, int W, void* stream) {
    PtrDeviceGuard _dg(crops);
    if (!crops || !mean || !std_ || !xs_dev || !xf_dev || !ys_dev || !yf_dev || !out || B <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0 || B > 65535)
        return ESPNET_EINVAL;
    dim3 grid((W + 31) / 32, (H + 7) / 8, B);
    preprocess_resize_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(crops, B, h, w, mean[0], mean[1], mean[2], std_[0], std_[1], std_[2], xs_dev,
                                                                    xf_dev, ys_dev, yf_dev, out, H, W);
    LAUNCH_COUNT();
    return cudaPeekAtLastError() == cudaSuccess ? ESPNET_OK : ESPNET_ECUDA;
}

int espnet_preprocess_resize_boxes(const uint8_t* slide, int slide_h, int slide_w, const int32_t* boxes_dev, int B, const float mean[3],
                                   const float std_[3], const int32_t* xs_dev, const float* xf_dev, const int32_t* ys_dev, const float* yf_dev,
                                   float* out, int H, int W, void* stream) {
    PtrDeviceGuard _dg(slide);
    if (!slide || !boxes_dev || !mean || !std_ || !xs_dev || !xf_dev || !ys_dev || !yf_dev || !out || B <= 0 || slide_h <= 0 || slide_w <= 0 ||
        H <= 0 || W <= 0 || B > 65535)
        return ESPNET_EINVAL;
    dim3 grid((W + 31) / 32, (H + 7) / 8, B);
    preprocess_resize_boxes_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(slide, slide_h, slide_w, boxes_dev, mean[0], mean[1], mean[2], std_[0],
                                                                          std_[1], std_[2], xs_dev, xf_dev, ys_dev, yf_dev, out, H, W);
    LAUNCH_COUNT();
    return cudaPeekAtLastError() == cudaSuccess ? ESPNET_OK : ESPNET_ECUDA;
}

int espnet_resize_nearest_u8(const uint8_t* src, int B, int sh, int sw, uint8_t* dst, int dh, int dw, const int32_t* ysrc_dev,
                             const int32_t* xsrc_dev, void* stream) {
    PtrDeviceGuard _dg(src);
    if (!src || !dst || !ysrc_dev || !xsrc_dev || B <= 0 || sh <= 0 || sw <= 0 || dh <= 0 || dw <= 0 || B > 65535 || dh > 65535) return ESPNET_EINVAL;
    int gx = (dw + 255) / 256;
    if (gx > 64) gx = 64;
    dim3 grid(gx, dh, B);
    resize_nearest_u8_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, B, sh, sw, dst, dh, dw, ysrc_dev, xsrc_dev);
    LAUNCH_COUNT();
    return cudaPeekAtLastError() == cudaSuccess ? ESPNET_OK : ESPNET_ECUDA;
}

int espnet_palette_overlay(const uint8_t* img, const uint8_t* label, size_t npix, const uint8_t* palette_dev, int n_pal, uint8_t* color_out,
                           uint8_t* overlay_out, void* stream) {
    PtrDeviceGuard _dg(label);
    if (!label || !palette_dev || n_pal <= 0 || n_pal > 256 || (!color_out && !overlay_out) || (overlay_out && !img)) return ESPNET_EINVAL;
    if (npix == 0) return ESPNET_OK;
    size_t g = (npix + 256 * 8 - 1) / (256 * 8);
    if (g > 4736) g = 4736;
    palette_overlay_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(img, label, npix, palette_dev, n_pal, color_out, overlay_out);
    LAUNCH_COUNT();
    return cudaPeekAtLastError() == cudaSuccess ? ESPNET_OK : ESPNET_ECUDA;
}

int espnet_render_ds8(const uint8_t* slide, const uint8_t* label, int slide_h, int slide_w, const uint8_t* palette_dev, int n_pal, uint8_t* out,
                      int ds_h, int ds_w, const int32_t* ysrc_dev, const int32_t* xsrc_dev, void* stream) {
    PtrDeviceGuard _dg(slide);
    if (!slide || !label || !palette_dev || !out || !ysrc_dev || !xsrc_dev || slide_h <= 0 || slide_w <= 0 || ds_h <= 0 || ds_w <= 0 || n_pal <= 0 ||
        n_pal > 256 || ds_h > 65535)
        return ESPNET_EINVAL;
    int gx = (ds_w + 255) / 256;
    if (gx > 64) gx = 64;
    dim3 grid(gx, ds_h);
    render_ds8_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(slide, label, slide_w, palette_dev, n_pal, out, ds_h, ds_w, ysrc_dev, xsrc_dev);
    LAUNCH_COUNT();
    return cudaPeekAtLastError() == cudaSuccess ? ESPNET_OK : ESPNET_ECUDA;
}

int espnet_class_counts(const uint8_t* maps, int B, size_t pix_per_map, int n_classes, unsigned long long* counts_dev, void* stream) {
    PtrDeviceGuard _dg(maps);
    if (!maps || !counts_dev || B <= 0 || B > 65535 || n_classes <= 0 || n_classes > 32) return ESPNET_EINVAL;
    if (pix_per_map == 0) return ESPNET_OK;
    size_t g = (pix_per_map + 256 * 64 - 1) / (256 * 64);
    if (g > 148) g = 148;
    if (g < 1) g = 1;
    dim3 grid((unsigned)g, B);
    class_count_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(maps, pix_per_map, n_classes, counts_dev);
    LAUNCH_COUNT();
    return cudaPeekAtLastError() == cudaSuccess ? ESPNET_OK : ESPNET_ECUDA;
}

int espnet_confusion_hist(const uint8_t* pred, const uint8_t* gt, size_t count, int n_classes, unsigned long long* hist_dev, void* stream) {
    PtrDeviceGuard _dg(pred);
    if (!pred || !gt || !hist_dev || n_classes <= 0 || n_classes > 32) return ESPNET_EINVAL;
    if (count == 0) return ESPNET_OK;
    size_t g = (count + 256 * 64 - 1) / (256 * 64);
    if (g > 1184) g = 1184;
    if (g < 1) g = 1;
    confusion_hist_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(pred, gt, count, n_classes, hist_dev);
    LAUNCH_COUNT();
    return cudaPeekAtLastError() == cudaSuccess ? ESPNET_OK : ESPNET_ECUDA;
}

}  // extern "C"
