// redsec_b200/csrc/api.cu -- C-ABI (include/redsec_b200.h) over the sm_100a kernels.
// Replaces lib/GPU/gates.cu + libredcufhe for the bootstrap hot path (SURVEY.md 8b "B-inner").
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <unordered_map>
#include <string>
#include <vector>

#include "../../include/redsec_b200.h"
#include "blind_rotate.cuh"
#include "blind_rotate_ws.cuh"
#include "blind_rotate_tm.cuh"
#include "lwe_kernels.cuh"
#include "keyswitch_mma.cuh"
#include "params.h"

namespace {

std::string g_create_error;

struct ProfEvent {
    cudaEvent_t start, stop;
    int kind;
};

}  // namespace

// A lane = one CUDA stream plus the scratch a bootstrap needs (extracted samples, gate pre-combination), so that independent
// chains of launches (the channel blocks of a max-pool layer: sign -> OR level 1 -> OR level 2) can be in flight at once and
// the half-empty last wave of one launch is filled by the CTAs of another.  Lane 0 is the context's own stream.
struct Lane {
    cudaStream_t stream = nullptr;
    uint32_t* ext = nullptr; size_t ext_cap = 0;
    uint32_t* lin = nullptr; size_t lin_cap = 0;
    cudaEvent_t ev = nullptr;
};

struct rs_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    std::string err;
    // evaluation key, device resident
    double2* bsk_f = nullptr;     // [n][BK_ROWS][2][NH]
    uint32_t* ksk = nullptr;      // [N][t][base][LWE_STRIDE]  (un-tiled keyswitch, variant 1)
    uint32_t* ksk7 = nullptr;     // [N][t][7][LWE_STRIDE]     (tiled keyswitch, default)
    bool key_loaded = false;
    // tensor-core keyswitch (variant 2): pre-tiled u8 key, transposed a' scratch
    uint8_t* kskb = nullptr;
    bool kskb_ready = false;      // false until built from the CURRENT key (rs_load_eval_key resets it)
    uint32_t* abar_t = nullptr; size_t abar_cap = 0;
    uint32_t* bprime = nullptr; size_t bprime_cap = 0;
    // scratch (grow-only; no allocation on the steady-state hot path)
    uint32_t* ext = nullptr; size_t ext_cap = 0;        // extracted samples
    uint32_t* lin = nullptr; size_t lin_cap = 0;        // gate pre-combination
    uint32_t* wire = nullptr; size_t wire_cap = 0;      // wire-format staging (host variants)
    uint32_t* io0 = nullptr; uint32_t* io1 = nullptr; size_t io_cap = 0;
    // caching allocator behind rs_lwe_alloc / rs_dev_alloc: layer forward allocates and frees an activation batch per layer,
    // and cudaMalloc / cudaFree (a device-wide sync each) cost ~0.5 s per CIFAR image; freed blocks are kept and handed out
    // again by size class.  All work is ordered on ctx->stream, so reuse right after a free is stream-safe.
    std::multimap<size_t, void*> free_blocks;        // capacity -> block
    std::unordered_map<void*, size_t> live_blocks;   // block -> capacity
    // measurement
    bool profiling = false;
    std::vector<ProfEvent> events;
    std::vector<ProfEvent> pool;
    double prof_ms[RS_K_COUNT] = {0, 0, 0, 0};
    uint64_t prof_n[RS_K_COUNT] = {0, 0, 0, 0};
    uint64_t launches = 0;
    int br_variant = 0;
    int ks_variant = 0;           // 0 auto (tensor cores for large batches, shared-memory gather below), 1 un-tiled gather, 2 tensor cores, 3 gather
    int ks_mma_min = 2048;        // RS_KS_MMA_MIN: smallest batch the auto mode sends to the tensor cores
    int ws_split = 4;             // largest row-split factor of the warp-specialised kernel for batches below 2 ciphertexts per SM:
                                  // <= sm_count ciphertexts are spread over 4 slots of a CTA each, <= 2 sm_count over 2 (RS_WS_SPLIT=1
                                  // disables, =2 caps at 2).  With the producer warp 4 slots beat 2 (3.00 against 3.35 ms per launch);
                                  // with the round-1 claim protocol they lost (4.15 against 3.90 ms: four front warps claiming slabs)
    float l2_keep = 0.45f;        // fraction of the BSK stream hinted L2 evict_last (RS_L2_KEEP; measured optimum)
    bool ws_tail_split = true;    // RS_WS_TAIL_SPLIT=0: never cut the small last wave of a multi-wave batch into a row-split launch of its own
    int ws_gate = 0;              // RS_WS_GATE=n: CTAs of every n-th wave of a long un-split launch wait for the earlier waves (0 = off, the default:
                                  // n = 1 cuts the launch's DRAM reads from 70-220 GB to 7 GB and costs 1.2 % of time; HBM is 4 % busy either way)
    unsigned* wave_done = nullptr;   // [16] one gate counter per lane
    int ws_lookahead = 5;         // RS_WS_LOOKAHEAD (1..5): BSK slabs the producer warp requests ahead of the slowest consumer
    bool ws_producer = true;      // 16-warp build of the warp-specialised kernel with a dedicated BSK producer warp; RS_WS_PRODUCER=0: the 12-warp
                                  // build whose front warps claim the slabs (A/B knob)
    bool ws_stress = false;       // RS_WS_STRESS=1: row-split launches use the instantiation that delays one back-warp pair (tests)
    std::vector<Lane> lanes;      // parked state of the lanes that are not selected (stream / ext / lin below belong to lane `cur_lane`)
    int cur_lane = 0;
    cudaEvent_t fork_ev = nullptr;
    int refs = 0;                 // objects (layers, nets, communicators) that hold this context: rs_ctx_destroy refuses while > 0
    size_t pool_cached_bytes = 0; // bytes parked in free_blocks
    size_t pool_cap_bytes = (size_t)16 << 30;   // RS_POOL_CAP_MB: parked bytes above which rs_*_free returns blocks to the driver
};

namespace {

// Every entry point that allocates or launches makes the context's device current for its duration: a process may hold
// contexts on several devices (the drop-in with NUM_GPUS > 1), and torch may have changed the current device.
struct DeviceGuard {
    int prev = -1; bool switched = false;
    explicit DeviceGuard(const rs_ctx* ctx) {
        if (!ctx) return;
        if (cudaGetDevice(&prev) == cudaSuccess && prev != ctx->device) { cudaSetDevice(ctx->device); switched = true; }
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};

int fail(rs_ctx* ctx, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_create_error = buf;
    return code;
}

#define RS_CUDA(ctx, call)                                                                          \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return fail(ctx, RS_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

struct LaunchScope {   // brackets one kernel launch with events when profiling is on
    rs_ctx* ctx; int kind; ProfEvent ev{}; bool on;
    LaunchScope(rs_ctx* c, int k) : ctx(c), kind(k), on(c->profiling) {
        ctx->launches++;
        if (!on) return;
        if (!ctx->pool.empty()) { ev = ctx->pool.back(); ctx->pool.pop_back(); }
        else { cudaEventCreate(&ev.start); cudaEventCreate(&ev.stop); }
        ev.kind = kind;
        cudaEventRecord(ev.start, ctx->stream);
    }
    ~LaunchScope() {
        if (!on) return;
        cudaEventRecord(ev.stop, ctx->stream);
        ctx->events.push_back(ev);
    }
};

int pool_alloc(rs_ctx* ctx, size_t bytes, void** out) {
    const size_t want = ((bytes ? bytes : 1) + 255) & ~(size_t)255;
    auto it = ctx->free_blocks.lower_bound(want);
    if (it != ctx->free_blocks.end() && it->first <= want + want / 4 + 4096) {     // close enough in size: reuse
        *out = it->second;
        ctx->live_blocks[it->second] = it->first;
        ctx->pool_cached_bytes -= it->first;
        ctx->free_blocks.erase(it);
        return RS_OK;
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {      // out of memory: drop the cache and retry once
        cudaGetLastError();
        cudaStreamSynchronize(ctx->stream);
        for (auto& kv : ctx->free_blocks) cudaFree(kv.second);
        ctx->free_blocks.clear();
        ctx->pool_cached_bytes = 0;
        e = cudaMalloc(&p, want);
    }
    if (e != cudaSuccess) return fail(ctx, RS_ERR_CUDA, "cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e));
    ctx->live_blocks[p] = want;
    *out = p;
    return RS_OK;
}
int pool_free(rs_ctx* ctx, void* p) {
    if (!p) return RS_OK;
    auto it = ctx->live_blocks.find(p);
    if (it == ctx->live_blocks.end()) return fail(ctx, RS_ERR_ARG, "free of a pointer this context did not allocate");
    ctx->free_blocks.emplace(it->second, p);
    ctx->pool_cached_bytes += it->second;
    ctx->live_blocks.erase(it);
    // bound the cache: beyond the cap the largest parked blocks go back to the driver (cudaFree orders itself after pending work)
    while (ctx->pool_cached_bytes > ctx->pool_cap_bytes && !ctx->free_blocks.empty()) {
        auto last = std::prev(ctx->free_blocks.end());
        cudaFree(last->second);
        ctx->pool_cached_bytes -= last->first;
        ctx->free_blocks.erase(last);
    }
    return RS_OK;
}

int grow(rs_ctx* ctx, uint32_t** p, size_t* cap, size_t words) {
    if (*cap >= words) return RS_OK;
    if (*p) RS_CUDA(ctx, cudaFree(*p));
    *p = nullptr; *cap = 0;
    RS_CUDA(ctx, cudaMalloc(p, words * sizeof(uint32_t)));
    *cap = words;
    return RS_OK;
}

// Blind-rotation variants: (ciphertext groups per CTA, BSK ring stages).  4 groups = 8 warps = 2 per SM
// sub-partition (255 registers/thread).  More groups would put 3 warps on a sub-partition (168 registers/thread,
// which spills).  Variant 0 (default) = warp-specialised kernel (blind_rotate_ws.cuh: 12 warps, front/back roles,
// setmaxnreg); variant 1 = single-role kernel with a 7-stage BSK ring; variant 2 = same with a 4-stage ring;
// variant 3 = warp-specialised kernel with the BSK served from tensor memory (blind_rotate_tm.cuh: 16 warps,
// front/back/producer roles; measured slower than variant 0, kept for the record -- see DESIGN.md);
// variant 4 = variant 3 with a 4-stage smem ring.
struct BrVariant { int groups, stages, smem; void (*set_attr)(cudaError_t*); };
template <int G, int S>
void br_launch(rs_ctx* ctx, int grid, const uint32_t* in, int count, uint32_t mu, uint32_t* ext, const uint32_t* lut, int lut_mod) {
    rs::blind_rotate_kernel<G, S><<<grid, G * 64, rs::BrSmem<G, S>::kTotal, ctx->stream>>>(in, count, mu, ctx->bsk_f, ext, lut, lut_mod);
}
template <int G, int S>
cudaError_t br_prepare() {
    return cudaFuncSetAttribute(rs::blind_rotate_kernel<G, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, rs::BrSmem<G, S>::kTotal);
}
constexpr int kWsStages = 5, kWsSlots = 3;
constexpr int cmax(int a, int b) { return a > b ? a : b; }
constexpr int kMaxSmemNeeded = cmax(cmax(rs::BrSmem<4, 7>::kTotal, rs::WsSmem<kWsStages, kWsSlots>::kTotal),
                                    cmax(rs::TmSmem<5, 3>::kTotal, rs::TmSmem<5, 3>::kTotal));

int launch_blind_rotate_one(rs_ctx* ctx, uint32_t* ext, const uint32_t* in, size_t count, uint32_t mu, const uint32_t* lut, int lut_mod);

// A batch of several waves whose last wave would hold at most two ciphertexts per SM is cut in two launches: the whole waves
// un-split, the tail in a row-split mode -- an un-split wave costs 7.1 ms whatever it holds, the tail 3.0 ms (<= 1 per SM) or 4.4 ms
// (<= 2 per SM).  3 072 ciphertexts: 42.8 -> 38.9 ms.  (Test-vector batches only when the tail starts at a multiple of the table
// count, so that "ciphertext c uses table c % lut_mod" still holds inside the second launch.)
int launch_blind_rotate(rs_ctx* ctx, uint32_t* ext, const uint32_t* in, size_t count, uint32_t mu,
                        const uint32_t* lut = nullptr, int lut_mod = 1) {
    const size_t wave = 4 * (size_t)ctx->sm_count;
    if (ctx->br_variant == 0 && ctx->ws_split > 1 && ctx->ws_tail_split && count > wave) {
        const size_t tail = count % wave, head = count - tail;
        if (tail > 0 && tail <= 2 * (size_t)ctx->sm_count && (!lut || head % (size_t)lut_mod == 0)) {
            const int rc = launch_blind_rotate_one(ctx, ext, in, head, mu, lut, lut_mod);
            if (rc != RS_OK) return rc;
            return launch_blind_rotate_one(ctx, ext + head * rs::EXT_STRIDE, in + head * rs::LWE_STRIDE, tail, mu, lut, lut_mod);
        }
    }
    return launch_blind_rotate_one(ctx, ext, in, count, mu, lut, lut_mod);
}

int launch_blind_rotate_one(rs_ctx* ctx, uint32_t* ext, const uint32_t* in, size_t count, uint32_t mu, const uint32_t* lut, int lut_mod) {
    if (!ctx->key_loaded) return fail(ctx, RS_ERR_STATE, "rs_load_eval_key has not been called");
    if (count == 0) return RS_OK;
    if (lut && (ctx->br_variant == 3 || ctx->br_variant == 4))
        return fail(ctx, RS_ERR_STATE, "test-vector bootstraps are not implemented in the tensor-memory variants (rs_set_tuning 3/4)");
    constexpr int G = 4;
    const int grid = (int)((count + G - 1) / G);
    // warp-specialised kernel: a batch below two ciphertexts per SM is latency-bound (one ciphertext alone on an SM needs 6.1 ms),
    // so it is spread over all SMs and each ciphertext over 2 or 4 of the CTA's slots (row-split modes, blind_rotate_ws.cuh);
    // larger batches below one wave are spread at <= 4 ciphertexts per CTA; full waves use ceil(count/4) CTAs of 4
    const size_t sms = (size_t)ctx->sm_count;
    const int split = std::min(ctx->ws_split, count <= sms ? 4 : count <= 2 * sms ? 2 : 1);
    const int bgrid = count < (size_t)G * sms ? (int)std::min<size_t>(count, sms) : grid;
    {
        LaunchScope ls(ctx, RS_K_BLIND_ROTATE);
        if (ctx->br_variant == 3)
            rs::blind_rotate_tm_kernel<5, 3, 120, 184><<<grid, 512, rs::TmSmem<5, 3>::kTotal, ctx->stream>>>(
                in, (int)count, mu, ctx->bsk_f, ext);
        else if (ctx->br_variant == 4)
            rs::blind_rotate_tm_kernel<4, 3, 120, 184><<<grid, 512, rs::TmSmem<5, 3>::kTotal, ctx->stream>>>(in, (int)count, mu, ctx->bsk_f, ext);
        else if (ctx->br_variant == 0) {
            // warp-specialised kernel: 16-warp build with a BSK producer warp (default), or the 12-warp build whose front warps
            // claim the slabs themselves (RS_WS_PRODUCER=0, the round-1 shape, kept for A/B runs)
            const int smem = rs::WsSmem<kWsStages, kWsSlots>::kTotal;
            // wave gate of long un-split launches (blind_rotate_ws.cuh): one counter per lane, cleared on the launch stream
            unsigned* gate = nullptr;
            if (ctx->ws_producer && ctx->ws_gate > 0 && split == 1 && (size_t)bgrid > (size_t)ctx->ws_gate * sms && ctx->wave_done) {
                gate = ctx->wave_done + ctx->cur_lane;
                RS_CUDA(ctx, cudaMemsetAsync(gate, 0, sizeof(unsigned), ctx->stream));
            }
#define RS_WS_LAUNCH(SPLIT_, STRESS_)                                                                                              \
            do {                                                                                                                   \
                if (ctx->ws_producer)                                                                                              \
                    rs::blind_rotate_ws_kernel<kWsStages, kWsSlots, SPLIT_, STRESS_, true><<<bgrid, 512, smem, ctx->stream>>>(     \
                        in, (int)count, mu, ctx->bsk_f, ext, ctx->l2_keep, lut, lut_mod, ctx->ws_lookahead, gate, (int)sms, ctx->ws_gate); \
                else                                                                                                               \
                    rs::blind_rotate_ws_kernel<kWsStages, kWsSlots, SPLIT_, STRESS_, false><<<bgrid, 384, smem, ctx->stream>>>(    \
                        in, (int)count, mu, ctx->bsk_f, ext, ctx->l2_keep, lut, lut_mod, ctx->ws_lookahead, gate, (int)sms, ctx->ws_gate); \
            } while (0)
            if (split == 4) RS_WS_LAUNCH(4, false);
            else if (split == 2 && ctx->ws_stress) RS_WS_LAUNCH(2, true);
            else if (split == 2) RS_WS_LAUNCH(2, false);
            else RS_WS_LAUNCH(1, false);
#undef RS_WS_LAUNCH
        }
        else if (ctx->br_variant == 1) br_launch<4, 7>(ctx, grid, in, (int)count, mu, ext, lut, lut_mod);
        else br_launch<4, 4>(ctx, grid, in, (int)count, mu, ext, lut, lut_mod);
    }
    RS_CUDA(ctx, cudaGetLastError());
    return RS_OK;
}

int grid_for(rs_ctx* ctx, size_t total, int block) {
    size_t g = (total + block - 1) / block;
    size_t cap = (size_t)ctx->sm_count * 16;
    return (int)(g < cap ? (g ? g : 1) : cap);
}

constexpr int kKsTile = 4;
template <int TILE>
void ks_tiled_launch(rs_ctx* ctx, uint32_t* out, const uint32_t* ext, int count) {
    // split the sum over i until the grid covers the machine about twice (partial sums are combined with red.add)
    const int tiles = (count + TILE - 1) / TILE;
    int ir = rs::KS_IR_MAX;
    while (ir > 16 && tiles * (rs::N / ir) < 2 * ctx->sm_count) ir >>= 1;
    dim3 grid((unsigned)tiles, (unsigned)(rs::N / ir));
    rs::keyswitch_tiled_kernel<TILE><<<grid, 384, rs::KsSmem<TILE>::kTotal, ctx->stream>>>(ext, count, ir, ctx->ksk7, out);
}
int launch_keyswitch(rs_ctx* ctx, uint32_t* out, const uint32_t* ext, size_t count) {
    if (!ctx->key_loaded) return fail(ctx, RS_ERR_STATE, "rs_load_eval_key has not been called");
    if (count == 0) return RS_OK;
    // variant 0 (default) chooses by batch size: the tensor-core GEMM needs 256 ciphertexts x 64 words per CTA to fill the machine
    // (2^16: 5.3 ms against 24.9 ms for the gather kernel), the shared-memory gather kernel splits the sum over the 1024
    // coefficients across CTAs and wins on small batches (592: 0.27 ms against 0.5 ms)
    const bool use_mma = ctx->ks_variant == 2 || (ctx->ks_variant == 0 && count >= (size_t)ctx->ks_mma_min);
    if (use_mma) {
        // exact int8 GEMM on the tensor cores (keyswitch_mma.cuh): tiled key built on first use from the padded table
        if (!ctx->kskb) RS_CUDA(ctx, cudaMalloc(&ctx->kskb, rs::KSKB_BYTES));
        if (!ctx->kskb_ready) {          // first use, or the first use after another rs_load_eval_key
            LaunchScope ls(ctx, RS_K_OTHER);
            rs::kskb_build_kernel<<<ctx->sm_count * 16, 256, 0, ctx->stream>>>(ctx->ksk, ctx->kskb);
            ctx->kskb_ready = true;
        }
        const int stride = (int)((count + rs::KM_CTS - 1) / rs::KM_CTS) * rs::KM_CTS;
        if (int r = grow(ctx, &ctx->abar_t, &ctx->abar_cap, (size_t)rs::N * stride)) return r;
        if (int r = grow(ctx, &ctx->bprime, &ctx->bprime_cap, (size_t)stride)) return r;
        {
            LaunchScope ls(ctx, RS_K_KEYSWITCH);
            dim3 grid((unsigned)(stride / 32), rs::N / 32), block(32, 8);
            rs::ks_mma_prep_kernel<<<grid, block, 0, ctx->stream>>>(ext, (int)count, stride, ctx->abar_t, ctx->bprime);
        }
        RS_CUDA(ctx, cudaGetLastError());
        LaunchScope ls(ctx, RS_K_KEYSWITCH);
        dim3 grid((unsigned)(stride / rs::KM_CTS), rs::KM_NT);
        rs::keyswitch_mma_kernel<<<grid, 320, rs::KmSmem::kTotal, ctx->stream>>>(ctx->abar_t, ctx->bprime, (int)count, stride, ctx->kskb, out);
    } else if (ctx->ks_variant == 1) {
        const int grid = (int)((count + kKsTile - 1) / kKsTile);
        LaunchScope ls(ctx, RS_K_KEYSWITCH);
        rs::keyswitch_kernel<kKsTile><<<grid, rs::LWE_STRIDE, 0, ctx->stream>>>(ext, (int)count, ctx->ksk, out);
    } else {
        {
            LaunchScope ls(ctx, RS_K_KEYSWITCH);
            rs::keyswitch_init_kernel<<<grid_for(ctx, count * rs::LWE_STRIDE, 256), 256, 0, ctx->stream>>>(ext, (int)count, out);
        }
        RS_CUDA(ctx, cudaGetLastError());
        LaunchScope ls(ctx, RS_K_KEYSWITCH);
        if (count >= 16384) ks_tiled_launch<64>(ctx, out, ext, (int)count);
        else if (count >= 1024) ks_tiled_launch<32>(ctx, out, ext, (int)count);
        else ks_tiled_launch<16>(ctx, out, ext, (int)count);
    }
    RS_CUDA(ctx, cudaGetLastError());
    return RS_OK;
}

// DFMA throughput probe (the FP64 roofline denominator; MEASURED_PEAKS.json carries no FP64 figure).  Same loop as
// scripts/probes/dmma_probe.cu, which measured 36.27 TFLOP/s: 8 independent accumulator chains per thread, the multiplier an
// immediate (two register operands per DFMA, so the register file is not the limit), 32 DFMAs per loop trip so loop overhead
// is ~3 %, 12 or 16 resident warps per SM.  rs_fp64_peak reports the best of both shapes.
__global__ void __launch_bounds__(512) fp64_peak_kernel(double* out, const double* in, int iters) {
    double f[8], g[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { f[i] = in[(threadIdx.x + i) & 31]; g[i] = in[(threadIdx.x + 8 + i) & 31]; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int i = 0; i < 8; i++) f[i] = fma(g[i], 1.0000001, f[i]);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// same loop with three distinct register operands per FMA: the register file feeds the FP64 pipe one fresh 64-bit operand
// per cycle, so such an FMA takes 3 cycles instead of 2 (scripts/probes/fp64_probe2.cu) -- the practical ceiling of FMA-heavy code
__global__ void fp64_peak3_kernel(double* out, const double* in, int iters) {
    double a[8], b[8], c[8];
    for (int i = 0; i < 8; i++) { a[i] = in[i] + threadIdx.x; b[i] = in[8 + i] + threadIdx.x; c[i] = in[16 + i] + 1e-9 * threadIdx.x; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = fma(b[i], c[i], a[i]);
    }
    double s = 0;
    for (int i = 0; i < 8; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace

extern "C" {

int rs_ctx_create(rs_ctx** out, int device) {
    if (!out) return fail(nullptr, RS_ERR_ARG, "rs_ctx_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, RS_ERR_CUDA, "no CUDA device (%s); this engine has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, RS_ERR_ARG, "device %d out of range (%d devices)", device, ndev);
    RS_CUDA(nullptr, cudaSetDevice(device));
    cudaDeviceProp prop;
    RS_CUDA(nullptr, cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(nullptr, RS_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    if ((size_t)prop.sharedMemPerBlockOptin < (size_t)kMaxSmemNeeded)
        return fail(nullptr, RS_ERR_CUDA, "device offers %zu B opt-in shared memory, kernel needs %d", (size_t)prop.sharedMemPerBlockOptin, kMaxSmemNeeded);
    rs_ctx* ctx = new rs_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete ctx; return fail(nullptr, RS_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
    e = cudaSuccess;
    {
        const int smem = rs::WsSmem<kWsStages, kWsSlots>::kTotal;
        auto opt_in = [&](auto kernel) { if (e == cudaSuccess) e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); };
        opt_in(rs::blind_rotate_ws_kernel<kWsStages, kWsSlots, 1, false, true>);
        opt_in(rs::blind_rotate_ws_kernel<kWsStages, kWsSlots, 2, false, true>);
        opt_in(rs::blind_rotate_ws_kernel<kWsStages, kWsSlots, 2, true, true>);
        opt_in(rs::blind_rotate_ws_kernel<kWsStages, kWsSlots, 4, false, true>);
        opt_in(rs::blind_rotate_ws_kernel<kWsStages, kWsSlots, 1, false, false>);
        opt_in(rs::blind_rotate_ws_kernel<kWsStages, kWsSlots, 2, false, false>);
        opt_in(rs::blind_rotate_ws_kernel<kWsStages, kWsSlots, 2, true, false>);
        opt_in(rs::blind_rotate_ws_kernel<kWsStages, kWsSlots, 4, false, false>);
    }
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(rs::blind_rotate_tm_kernel<5, 3, 120, 184>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 rs::TmSmem<5, 3>::kTotal);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(rs::blind_rotate_tm_kernel<4, 3, 120, 184>, cudaFuncAttributeMaxDynamicSharedMemorySize, rs::TmSmem<5, 3>::kTotal);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rs::keyswitch_tiled_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, rs::KsSmem<64>::kTotal);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rs::keyswitch_tiled_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, rs::KsSmem<32>::kTotal);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rs::keyswitch_tiled_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, rs::KsSmem<16>::kTotal);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rs::keyswitch_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, rs::KmSmem::kTotal);
    if (e == cudaSuccess) e = br_prepare<4, 7>();
    if (e == cudaSuccess) e = br_prepare<4, 4>();
    if (e != cudaSuccess) { delete ctx; return fail(nullptr, RS_ERR_CUDA, "cudaFuncSetAttribute(max dynamic smem): %s", cudaGetErrorString(e)); }
    if (const char* env = getenv("RS_L2_KEEP")) { float v = (float)atof(env); if (v >= 0.f && v <= 1.f) ctx->l2_keep = v; }
    if (const char* env = getenv("RS_POOL_CAP_MB")) ctx->pool_cap_bytes = (size_t)atoll(env) << 20;
    if (const char* env = getenv("RS_WS_STRESS")) ctx->ws_stress = atoi(env) != 0;
    if (const char* env = getenv("RS_WS_PRODUCER")) ctx->ws_producer = atoi(env) != 0;
    if (const char* env = getenv("RS_WS_TAIL_SPLIT")) ctx->ws_tail_split = atoi(env) != 0;
    if (const char* env = getenv("RS_WS_GATE")) ctx->ws_gate = std::max(atoi(env), 0);
    if (const char* env = getenv("RS_WS_LOOKAHEAD")) ctx->ws_lookahead = std::min(std::max(atoi(env), 1), kWsStages);
    if (const char* env = getenv("RS_WS_SPLIT")) { const int v = atoi(env); ctx->ws_split = v >= 4 ? 4 : v >= 2 ? 2 : 1; }
    if (ctx->l2_keep > 0.f)   // the evict_last hint only holds lines inside the persisting carve-out (82.9 MB max on B200); best effort
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)prop.persistingL2CacheMaxSize);
    if (const char* env = getenv("RS_KS_VARIANT")) { int v = atoi(env); if (v >= 0 && v <= 3) ctx->ks_variant = v; }
    if (const char* env = getenv("RS_KS_MMA_MIN")) { int v = atoi(env); if (v >= 1) ctx->ks_mma_min = v; }
    if (const char* env = getenv("RS_BR_VARIANT")) { int v = atoi(env); if (v >= 0 && v <= 4) ctx->br_variant = v; }
    *out = ctx;
    return RS_OK;
}

static void lane_park(rs_ctx* ctx) {     // store the selected lane's live fields back into its slot
    Lane& l = ctx->lanes[ctx->cur_lane];
    l.stream = ctx->stream; l.ext = ctx->ext; l.ext_cap = ctx->ext_cap; l.lin = ctx->lin; l.lin_cap = ctx->lin_cap;
}
int rs_lanes(rs_ctx* ctx, int n) {
    if (!ctx || n < 1 || n > 16) return fail(ctx, RS_ERR_ARG, "rs_lanes: 1..16 lanes");
    DeviceGuard dg(ctx);
    if (ctx->lanes.empty()) ctx->lanes.resize(1);
    while ((int)ctx->lanes.size() < n) {
        Lane l;
        RS_CUDA(ctx, cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking));
        RS_CUDA(ctx, cudaEventCreateWithFlags(&l.ev, cudaEventDisableTiming));
        ctx->lanes.push_back(l);
    }
    if (!ctx->fork_ev) RS_CUDA(ctx, cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming));
    return RS_OK;
}
int rs_lane_count(const rs_ctx* ctx) { return ctx ? (ctx->lanes.empty() ? 1 : (int)ctx->lanes.size()) : 0; }
int rs_lane_select(rs_ctx* ctx, int lane) {
    if (!ctx) return RS_ERR_ARG;
    if (ctx->lanes.empty()) ctx->lanes.resize(1);
    if (lane < 0 || lane >= (int)ctx->lanes.size()) return fail(ctx, RS_ERR_ARG, "rs_lane_select: lane %d of %zu", lane, ctx->lanes.size());
    if (lane == ctx->cur_lane) return RS_OK;
    lane_park(ctx);
    const Lane& l = ctx->lanes[lane];
    ctx->stream = l.stream; ctx->ext = l.ext; ctx->ext_cap = l.ext_cap; ctx->lin = l.lin; ctx->lin_cap = l.lin_cap;
    ctx->cur_lane = lane;
    return RS_OK;
}
int rs_lane_fork(rs_ctx* ctx) {   // lanes 1.. wait for everything issued so far on lane 0
    if (!ctx) return RS_ERR_ARG;
    if (ctx->lanes.size() < 2) return RS_OK;
    if (ctx->cur_lane != 0) return fail(ctx, RS_ERR_STATE, "rs_lane_fork: select lane 0 first");
    DeviceGuard dg(ctx);
    RS_CUDA(ctx, cudaEventRecord(ctx->fork_ev, ctx->stream));
    for (size_t k = 1; k < ctx->lanes.size(); k++) RS_CUDA(ctx, cudaStreamWaitEvent(ctx->lanes[k].stream, ctx->fork_ev, 0));
    return RS_OK;
}
int rs_lane_join(rs_ctx* ctx) {   // lane 0 waits for everything issued so far on lanes 1..
    if (!ctx) return RS_ERR_ARG;
    if (ctx->lanes.size() < 2) return RS_OK;
    if (ctx->cur_lane != 0) return fail(ctx, RS_ERR_STATE, "rs_lane_join: select lane 0 first");
    DeviceGuard dg(ctx);
    for (size_t k = 1; k < ctx->lanes.size(); k++) {
        RS_CUDA(ctx, cudaEventRecord(ctx->lanes[k].ev, ctx->lanes[k].stream));
        RS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->lanes[k].ev, 0));
    }
    return RS_OK;
}

int rs_ctx_retain(rs_ctx* ctx) { if (!ctx) return RS_ERR_ARG; ctx->refs++; return RS_OK; }
int rs_ctx_release(rs_ctx* ctx) { if (!ctx || ctx->refs <= 0) return RS_ERR_ARG; ctx->refs--; return RS_OK; }

int rs_pool_trim(rs_ctx* ctx) {
    if (!ctx) return RS_ERR_ARG;
    DeviceGuard dg(ctx);
    RS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (auto& kv : ctx->free_blocks) cudaFree(kv.second);
    ctx->free_blocks.clear();
    ctx->pool_cached_bytes = 0;
    return RS_OK;
}

int rs_ctx_destroy(rs_ctx* ctx) {
    if (!ctx) return RS_OK;
    // layers, nets and communicators free their device tables through this context in their destructors: destroying it
    // first would leave them with a dangling pointer, so refuse (the caller destroys its nets first, then the context)
    if (ctx->refs > 0) return fail(ctx, RS_ERR_STATE, "rs_ctx_destroy: %d layer / net / communicator objects still use this context", ctx->refs);
    DeviceGuard dg(ctx);
    rs_lane_select(ctx, 0);
    cudaStreamSynchronize(ctx->stream);
    for (size_t k = 1; k < ctx->lanes.size(); k++) {
        Lane& l = ctx->lanes[k];
        cudaStreamSynchronize(l.stream);
        cudaFree(l.ext); cudaFree(l.lin); cudaEventDestroy(l.ev); cudaStreamDestroy(l.stream);
    }
    if (ctx->fork_ev) cudaEventDestroy(ctx->fork_ev);
    for (auto& ev : ctx->events) { cudaEventDestroy(ev.start); cudaEventDestroy(ev.stop); }
    for (auto& ev : ctx->pool) { cudaEventDestroy(ev.start); cudaEventDestroy(ev.stop); }
    cudaFree(ctx->bsk_f); cudaFree(ctx->ksk); cudaFree(ctx->ksk7); cudaFree(ctx->ext); cudaFree(ctx->lin); cudaFree(ctx->wire);
    cudaFree(ctx->io0); cudaFree(ctx->io1); cudaFree(ctx->kskb); cudaFree(ctx->abar_t); cudaFree(ctx->bprime); cudaFree(ctx->wave_done);
    for (auto& kv : ctx->free_blocks) cudaFree(kv.second);
    for (auto& kv : ctx->live_blocks) cudaFree(kv.first);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return RS_OK;
}

const char* rs_last_error(const rs_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int rs_set_stream(rs_ctx* ctx, void* cuda_stream) {
    DeviceGuard dg(ctx);
    if (!ctx) return RS_ERR_ARG;
    if (ctx->cur_lane != 0) return fail(ctx, RS_ERR_STATE, "rs_set_stream: select lane 0 first");
    RS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (cuda_stream) {
        if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
        ctx->stream = (cudaStream_t)cuda_stream;
        ctx->own_stream = false;
    } else if (!ctx->own_stream) {
        RS_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ctx->own_stream = true;
    }
    return RS_OK;
}

int rs_sync(rs_ctx* ctx) {
    DeviceGuard dg(ctx);
    if (!ctx) return RS_ERR_ARG;
    RS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (size_t k = 0; k < ctx->lanes.size(); k++)
        if ((int)k != ctx->cur_lane && ctx->lanes[k].stream) RS_CUDA(ctx, cudaStreamSynchronize(ctx->lanes[k].stream));
    RS_CUDA(ctx, cudaGetLastError());
    return RS_OK;
}

int rs_load_eval_key(rs_ctx* ctx, const uint32_t* bsk_host, const uint32_t* ksk_host) {
    DeviceGuard dg(ctx);
    if (!ctx || !bsk_host || !ksk_host) return fail(ctx, RS_ERR_ARG, "rs_load_eval_key: NULL argument");
    if (!ctx->bsk_f) RS_CUDA(ctx, cudaMalloc(&ctx->bsk_f, rs::BSK_F_BYTES));
    if (!ctx->ksk) RS_CUDA(ctx, cudaMalloc(&ctx->ksk, rs::KSK_DEV_WORDS * sizeof(uint32_t)));
    if (!ctx->ksk7) RS_CUDA(ctx, cudaMalloc(&ctx->ksk7, rs::KSK_TILED_WORDS * sizeof(uint32_t)));
    if (!ctx->wave_done) RS_CUDA(ctx, cudaMalloc(&ctx->wave_done, 16 * sizeof(unsigned)));
    // staging for the torus32 keys (freed after conversion)
    uint32_t* stage = nullptr;
    const size_t ksk_bytes = RS_KSK_WORDS * sizeof(uint32_t), bsk_bytes = RS_BSK_WORDS * sizeof(uint32_t);
    RS_CUDA(ctx, cudaMalloc(&stage, ksk_bytes > bsk_bytes ? ksk_bytes : bsk_bytes));
    RS_CUDA(ctx, cudaMemcpyAsync(stage, bsk_host, bsk_bytes, cudaMemcpyHostToDevice, ctx->stream));
    const int npolys = rs::LWE_N * rs::BK_ROWS * 2;
    {
        LaunchScope ls(ctx, RS_K_OTHER);
        rs::bsk_to_fourier_kernel<<<ctx->sm_count * 8, 64, 0, ctx->stream>>>(reinterpret_cast<const int32_t*>(stage), ctx->bsk_f, npolys);
    }
    RS_CUDA(ctx, cudaGetLastError());
    RS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    RS_CUDA(ctx, cudaMemcpyAsync(stage, ksk_host, ksk_bytes, cudaMemcpyHostToDevice, ctx->stream));
    const size_t rows = (size_t)rs::N * rs::KS_T * rs::KS_BASE;
    {
        LaunchScope ls(ctx, RS_K_OTHER);
        rs::ksk_pad_kernel<<<grid_for(ctx, rows * rs::LWE_STRIDE, 256), 256, 0, ctx->stream>>>(stage, ctx->ksk, rows);
    }
    RS_CUDA(ctx, cudaGetLastError());
    const size_t rows7 = (size_t)rs::N * rs::KS_T * rs::KS_DIGITS;
    {
        LaunchScope ls(ctx, RS_K_OTHER);
        rs::ksk_tile_kernel<<<grid_for(ctx, rows7 * rs::LWE_STRIDE, 256), 256, 0, ctx->stream>>>(stage, ctx->ksk7, rows7);
    }
    RS_CUDA(ctx, cudaGetLastError());
    RS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    RS_CUDA(ctx, cudaFree(stage));
    ctx->key_loaded = true;
    ctx->kskb_ready = false;      // the tensor-core keyswitch table is rebuilt from the new key on its next use
    return RS_OK;
}

int rs_lwe_alloc(rs_ctx* ctx, size_t count, uint32_t** dev_out) {
    DeviceGuard dg(ctx);
    if (!ctx || !dev_out) return fail(ctx, RS_ERR_ARG, "rs_lwe_alloc: NULL argument");
    return pool_alloc(ctx, (count ? count : 1) * rs::LWE_STRIDE * sizeof(uint32_t), reinterpret_cast<void**>(dev_out));
}
int rs_lwe_free(rs_ctx* ctx, uint32_t* dev) {
    DeviceGuard dg(ctx);
    if (!ctx) return RS_ERR_ARG;
    return pool_free(ctx, dev);
}
int rs_lwe_upload(rs_ctx* ctx, uint32_t* dev, const uint32_t* host_wire, size_t count) {
    DeviceGuard dg(ctx);
    if (!ctx || !dev || !host_wire) return fail(ctx, RS_ERR_ARG, "rs_lwe_upload: NULL argument");
    if (count == 0) return RS_OK;
    if (int r = grow(ctx, &ctx->wire, &ctx->wire_cap, count * rs::LWE_WORDS)) return r;
    RS_CUDA(ctx, cudaMemcpyAsync(ctx->wire, host_wire, count * rs::LWE_WORDS * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    {
        LaunchScope ls(ctx, RS_K_OTHER);
        rs::lwe_pad_kernel<<<grid_for(ctx, count * rs::LWE_STRIDE, 256), 256, 0, ctx->stream>>>(ctx->wire, dev, (int)count);
    }
    RS_CUDA(ctx, cudaGetLastError());
    return RS_OK;
}
int rs_lwe_download(rs_ctx* ctx, uint32_t* host_wire, const uint32_t* dev, size_t count) {
    DeviceGuard dg(ctx);
    if (!ctx || !dev || !host_wire) return fail(ctx, RS_ERR_ARG, "rs_lwe_download: NULL argument");
    if (count == 0) return RS_OK;
    if (int r = grow(ctx, &ctx->wire, &ctx->wire_cap, count * rs::LWE_WORDS)) return r;
    {
        LaunchScope ls(ctx, RS_K_OTHER);
        rs::lwe_unpad_kernel<<<grid_for(ctx, count * rs::LWE_WORDS, 256), 256, 0, ctx->stream>>>(dev, ctx->wire, (int)count);
    }
    RS_CUDA(ctx, cudaGetLastError());
    RS_CUDA(ctx, cudaMemcpyAsync(host_wire, ctx->wire, count * rs::LWE_WORDS * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    RS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RS_OK;
}
int rs_lwe_copy(rs_ctx* ctx, uint32_t* dst_dev, const uint32_t* src_dev, size_t count) {
    DeviceGuard dg(ctx);
    if (!ctx || !dst_dev || !src_dev) return fail(ctx, RS_ERR_ARG, "rs_lwe_copy: NULL argument");
    if (count == 0) return RS_OK;
    RS_CUDA(ctx, cudaMemcpyAsync(dst_dev, src_dev, count * rs::LWE_STRIDE * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
    return RS_OK;
}
int rs_host_alloc(void** out, size_t bytes) {
    if (!out) return RS_ERR_ARG;
    return cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault) == cudaSuccess ? RS_OK : RS_ERR_CUDA;
}
int rs_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? RS_OK : RS_ERR_CUDA; }

int rs_blind_rotate_batch(rs_ctx* ctx, uint32_t* ext_dev, const uint32_t* in_dev, size_t count, uint32_t mu) {
    DeviceGuard dg(ctx);
    if (!ctx || !ext_dev || !in_dev) return fail(ctx, RS_ERR_ARG, "rs_blind_rotate_batch: NULL argument");
    return launch_blind_rotate(ctx, ext_dev, in_dev, count, mu);
}
int rs_keyswitch_batch(rs_ctx* ctx, uint32_t* out_dev, const uint32_t* ext_dev, size_t count) {
    DeviceGuard dg(ctx);
    if (!ctx || !ext_dev || !out_dev) return fail(ctx, RS_ERR_ARG, "rs_keyswitch_batch: NULL argument");
    return launch_keyswitch(ctx, out_dev, ext_dev, count);
}
int rs_ext_alloc(rs_ctx* ctx, size_t count, uint32_t** dev_out) {
    DeviceGuard dg(ctx);
    if (!ctx || !dev_out) return fail(ctx, RS_ERR_ARG, "rs_ext_alloc: NULL argument");
    return pool_alloc(ctx, (count ? count : 1) * rs::EXT_STRIDE * sizeof(uint32_t), reinterpret_cast<void**>(dev_out));
}
int rs_ext_upload(rs_ctx* ctx, uint32_t* dev, const uint32_t* host, size_t count) {
    DeviceGuard dg(ctx);
    if (!ctx || !dev || !host) return fail(ctx, RS_ERR_ARG, "rs_ext_upload: NULL argument");
    RS_CUDA(ctx, cudaMemsetAsync(dev, 0, count * rs::EXT_STRIDE * sizeof(uint32_t), ctx->stream));
    RS_CUDA(ctx, cudaMemcpy2DAsync(dev, rs::EXT_STRIDE * 4, host, (rs::N + 1) * 4, (rs::N + 1) * 4, count, cudaMemcpyHostToDevice, ctx->stream));
    RS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RS_OK;
}
int rs_ext_download(rs_ctx* ctx, uint32_t* host, const uint32_t* dev, size_t count) {
    DeviceGuard dg(ctx);
    if (!ctx || !dev || !host) return fail(ctx, RS_ERR_ARG, "rs_ext_download: NULL argument");
    RS_CUDA(ctx, cudaMemcpy2DAsync(host, (rs::N + 1) * 4, dev, rs::EXT_STRIDE * 4, (rs::N + 1) * 4, count, cudaMemcpyDeviceToHost, ctx->stream));
    RS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RS_OK;
}

int rs_reserve_scratch(rs_ctx* ctx, size_t count) {
    DeviceGuard dg(ctx);
    if (!ctx) return RS_ERR_ARG;
    return grow(ctx, &ctx->ext, &ctx->ext_cap, count * rs::EXT_STRIDE);
}

int rs_pbs_batch(rs_ctx* ctx, uint32_t* out_dev, const uint32_t* in_dev, size_t count, uint32_t mu) {
    DeviceGuard dg(ctx);
    if (!ctx || !out_dev || !in_dev) return fail(ctx, RS_ERR_ARG, "rs_pbs_batch: NULL argument");
    if (int r = grow(ctx, &ctx->ext, &ctx->ext_cap, count * rs::EXT_STRIDE)) return r;
    if (int r = launch_blind_rotate(ctx, ctx->ext, in_dev, count, mu)) return r;
    return launch_keyswitch(ctx, out_dev, ctx->ext, count);
}

int rs_pbs_lut_batch(rs_ctx* ctx, uint32_t* out_dev, const uint32_t* in_dev, size_t count, const uint32_t* lut_dev, int lut_mod) {
    DeviceGuard dg(ctx);
    if (!ctx || !out_dev || !in_dev || !lut_dev) return fail(ctx, RS_ERR_ARG, "rs_pbs_lut_batch: NULL argument");
    if (lut_mod < 1) return fail(ctx, RS_ERR_ARG, "rs_pbs_lut_batch: lut_mod %d < 1", lut_mod);
    if (int r = grow(ctx, &ctx->ext, &ctx->ext_cap, count * rs::EXT_STRIDE)) return r;
    if (int r = launch_blind_rotate(ctx, ctx->ext, in_dev, count, 0u, lut_dev, lut_mod)) return r;
    return launch_keyswitch(ctx, out_dev, ctx->ext, count);
}

int rs_gate_batch(rs_ctx* ctx, int gate, uint32_t* out_dev, const uint32_t* in0_dev, const uint32_t* in1_dev, size_t count,
                  uint32_t mu) {
    DeviceGuard dg(ctx);
    if (!ctx || !out_dev || !in0_dev || !in1_dev) return fail(ctx, RS_ERR_ARG, "rs_gate_batch: NULL argument");
    if (gate < 0 || gate > RS_GATE_XNOR) return fail(ctx, RS_ERR_ARG, "rs_gate_batch: unknown gate %d", gate);
    if (count == 0) return RS_OK;
    // lib/GPU/gates.cu:246-286: NAND fix=+1/8, OR +1/8, AND -1/8, NOR -1/8, XOR +1/4 (x2), XNOR -1/4 (x2)
    static const uint32_t fix[6] = {0x20000000u, 0x20000000u, 0xE0000000u, 0xE0000000u, 0x40000000u, 0xC0000000u};
    static const uint32_t mul[6] = {0xFFFFFFFFu, 1u, 1u, 0xFFFFFFFFu, 2u, 0xFFFFFFFEu};
    if (int r = grow(ctx, &ctx->lin, &ctx->lin_cap, count * rs::LWE_STRIDE)) return r;
    {
        LaunchScope ls(ctx, RS_K_LINEAR);
        rs::gate_linear_kernel<<<grid_for(ctx, count * rs::LWE_STRIDE, 256), 256, 0, ctx->stream>>>(ctx->lin, in0_dev, in1_dev,
                                                                                                 (int)count, mul[gate], fix[gate]);
    }
    RS_CUDA(ctx, cudaGetLastError());
    return rs_pbs_batch(ctx, out_dev, ctx->lin, count, mu);
}

static int ensure_io(rs_ctx* ctx, size_t count) {
    DeviceGuard dg(ctx);
    if (ctx->io_cap >= count) return RS_OK;
    if (ctx->io0) cudaFree(ctx->io0);
    if (ctx->io1) cudaFree(ctx->io1);
    ctx->io0 = ctx->io1 = nullptr; ctx->io_cap = 0;
    RS_CUDA(ctx, cudaMalloc(&ctx->io0, count * rs::LWE_STRIDE * sizeof(uint32_t)));
    RS_CUDA(ctx, cudaMalloc(&ctx->io1, count * rs::LWE_STRIDE * sizeof(uint32_t)));
    ctx->io_cap = count;
    return RS_OK;
}

int rs_pbs_batch_host(rs_ctx* ctx, uint32_t* out_host, const uint32_t* in_host, size_t count, uint32_t mu) {
    DeviceGuard dg(ctx);
    if (!ctx || !out_host || !in_host) return fail(ctx, RS_ERR_ARG, "rs_pbs_batch_host: NULL argument");
    if (int r = ensure_io(ctx, count)) return r;
    if (int r = rs_lwe_upload(ctx, ctx->io0, in_host, count)) return r;
    if (int r = rs_pbs_batch(ctx, ctx->io0, ctx->io0, count, mu)) return r;
    return rs_lwe_download(ctx, out_host, ctx->io0, count);
}

int rs_gate_batch_host(rs_ctx* ctx, int gate, uint32_t* out_host, const uint32_t* in0_host, const uint32_t* in1_host,
                       size_t count, uint32_t mu) {
    DeviceGuard dg(ctx);
    if (!ctx || !out_host || !in0_host || !in1_host) return fail(ctx, RS_ERR_ARG, "rs_gate_batch_host: NULL argument");
    if (int r = ensure_io(ctx, count)) return r;
    if (int r = rs_lwe_upload(ctx, ctx->io0, in0_host, count)) return r;
    if (int r = rs_lwe_upload(ctx, ctx->io1, in1_host, count)) return r;
    if (int r = rs_gate_batch(ctx, gate, ctx->io0, ctx->io0, ctx->io1, count, mu)) return r;
    return rs_lwe_download(ctx, out_host, ctx->io0, count);
}

int rs_lwe_lincomb(rs_ctx* ctx, uint32_t* out_dev, size_t out_count, const uint32_t* in_dev, const int32_t* rowptr_dev,
                   const int32_t* col_dev, const int8_t* sign_dev, const uint32_t* bias_dev) {
    DeviceGuard dg(ctx);
    if (!ctx || !out_dev || !in_dev || !rowptr_dev) return fail(ctx, RS_ERR_ARG, "rs_lwe_lincomb: NULL argument");
    if (out_count == 0) return RS_OK;
    const size_t cap = (size_t)ctx->sm_count * 32;
    const int grid = (int)(out_count < cap ? out_count : cap);
    {
        LaunchScope ls(ctx, RS_K_LINEAR);
        rs::lwe_lincomb_kernel<<<grid, rs::LWE_STRIDE / 4, 0, ctx->stream>>>(out_dev, (int)out_count, in_dev, rowptr_dev, col_dev,
                                                                            sign_dev, bias_dev);
    }
    RS_CUDA(ctx, cudaGetLastError());
    return RS_OK;
}

int rs_lwe_conv(rs_ctx* ctx, uint32_t* out_dev, const uint32_t* in_dev, const int8_t* wpacked_dev, const uint32_t* bias_dev,
                const rs_conv_desc* desc) {
    DeviceGuard dg(ctx);
    if (!ctx || !out_dev || !in_dev || !wpacked_dev || !desc) return fail(ctx, RS_ERR_ARG, "rs_lwe_conv: NULL argument");
    if (desc->od_begin % rs::CONV_OD_TILE != 0 || desc->od_end <= desc->od_begin || desc->od_end > desc->out_dep)
        return fail(ctx, RS_ERR_ARG, "rs_lwe_conv: bad channel slice [%d,%d) of %d (begin must be a multiple of %d)", desc->od_begin,
                    desc->od_end, desc->out_dep, rs::CONV_OD_TILE);
    rs::ConvDesc d;
    d.in_h = desc->in_h; d.in_w = desc->in_w; d.in_dep = desc->in_dep;
    d.out_h = desc->out_h; d.out_w = desc->out_w; d.out_dep = desc->out_dep;
    d.win_h = desc->win_h; d.win_w = desc->win_w; d.stride_h = desc->stride_h; d.stride_w = desc->stride_w;
    d.ofs_h = desc->ofs_h; d.ofs_w = desc->ofs_w; d.od_begin = desc->od_begin; d.od_end = desc->od_end;
    d.unit = 1u << 20;   // modSwitchToTorus32(1, 4096)
    dim3 grid((unsigned)(d.out_h * d.out_w), (unsigned)((d.od_end - d.od_begin + rs::CONV_OD_TILE - 1) / rs::CONV_OD_TILE));
    {
        LaunchScope ls(ctx, RS_K_LINEAR);
        if (desc->int_mode) rs::lwe_conv_kernel<true><<<grid, rs::CONV_THREADS, 0, ctx->stream>>>(out_dev, in_dev, wpacked_dev, bias_dev, d);
        else rs::lwe_conv_kernel<false><<<grid, rs::CONV_THREADS, 0, ctx->stream>>>(out_dev, in_dev, wpacked_dev, bias_dev, d);
    }
    RS_CUDA(ctx, cudaGetLastError());
    return RS_OK;
}

int rs_lwe_interleave(rs_ctx* ctx, uint32_t* out_dev, const uint32_t* gathered_dev, size_t pixels, int c_local, int world) {
    DeviceGuard dg(ctx);
    if (!ctx || !out_dev || !gathered_dev || c_local <= 0 || world <= 0) return fail(ctx, RS_ERR_ARG, "rs_lwe_interleave: bad argument");
    const size_t rows = pixels * (size_t)c_local * world;
    if (rows == 0) return RS_OK;
    const size_t cap = (size_t)ctx->sm_count * 32;
    {
        LaunchScope ls(ctx, RS_K_LINEAR);
        rs::lwe_interleave_kernel<<<(unsigned)(rows < cap ? rows : cap), rs::LWE_STRIDE / 4, 0, ctx->stream>>>(
            reinterpret_cast<uint4*>(out_dev), reinterpret_cast<const uint4*>(gathered_dev), pixels, c_local, world);
    }
    RS_CUDA(ctx, cudaGetLastError());
    return RS_OK;
}

int rs_lwe_add_const(rs_ctx* ctx, uint32_t* dev, size_t count, uint32_t value) {
    DeviceGuard dg(ctx);
    if (!ctx || !dev) return fail(ctx, RS_ERR_ARG, "rs_lwe_add_const: NULL argument");
    if (count == 0) return RS_OK;
    {
        LaunchScope ls(ctx, RS_K_LINEAR);
        rs::lwe_add_const_kernel<<<grid_for(ctx, count, 256), 256, 0, ctx->stream>>>(dev, count, value);
    }
    RS_CUDA(ctx, cudaGetLastError());
    return RS_OK;
}

int rs_lwe_add_bias(rs_ctx* ctx, uint32_t* dev, size_t count, const uint32_t* bias_dev, int mod) {
    DeviceGuard dg(ctx);
    if (!ctx || !dev || !bias_dev || mod < 1) return fail(ctx, RS_ERR_ARG, "rs_lwe_add_bias: bad argument");
    if (count == 0) return RS_OK;
    {
        LaunchScope ls(ctx, RS_K_LINEAR);
        rs::lwe_add_bias_kernel<<<grid_for(ctx, count, 256), 256, 0, ctx->stream>>>(dev, count, bias_dev, mod);
    }
    RS_CUDA(ctx, cudaGetLastError());
    return RS_OK;
}

int rs_lwe_axpby(rs_ctx* ctx, uint32_t* out_dev, const uint32_t* in0_dev, const uint32_t* in1_dev, size_t count, uint32_t m0,
                 uint32_t m1, uint32_t fix) {
    DeviceGuard dg(ctx);
    if (!ctx || !out_dev || !in0_dev) return fail(ctx, RS_ERR_ARG, "rs_lwe_axpby: NULL argument");
    if (count == 0) return RS_OK;
    {
        LaunchScope ls(ctx, RS_K_LINEAR);
        rs::lwe_axpby_kernel<<<grid_for(ctx, count * rs::LWE_STRIDE, 256), 256, 0, ctx->stream>>>(out_dev, in0_dev, in1_dev, count, m0, m1, fix);
    }
    RS_CUDA(ctx, cudaGetLastError());
    return RS_OK;
}

int rs_ctx_device(const rs_ctx* ctx) { return ctx ? ctx->device : -1; }
int rs_get_stream(rs_ctx* ctx, void** cuda_stream) {
    if (!ctx || !cuda_stream) return RS_ERR_ARG;
    *cuda_stream = (void*)ctx->stream;
    return RS_OK;
}

int rs_dev_alloc(rs_ctx* ctx, size_t bytes, void** dev_out) {
    DeviceGuard dg(ctx);
    if (!ctx || !dev_out) return fail(ctx, RS_ERR_ARG, "rs_dev_alloc: NULL argument");
    return pool_alloc(ctx, bytes, dev_out);
}
int rs_dev_free(rs_ctx* ctx, void* dev) {
    DeviceGuard dg(ctx);
    if (!ctx) return RS_ERR_ARG;
    return pool_free(ctx, dev);
}
int rs_dev_upload(rs_ctx* ctx, void* dev, const void* host, size_t bytes) {
    DeviceGuard dg(ctx);
    if (!ctx || !dev || !host) return fail(ctx, RS_ERR_ARG, "rs_dev_upload: NULL argument");
    RS_CUDA(ctx, cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    RS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RS_OK;
}
int rs_dev_download(rs_ctx* ctx, void* host, const void* dev, size_t bytes) {
    DeviceGuard dg(ctx);
    if (!ctx || !dev || !host) return fail(ctx, RS_ERR_ARG, "rs_dev_download: NULL argument");
    RS_CUDA(ctx, cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    RS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RS_OK;
}

int rs_profile_enable(rs_ctx* ctx, int on) {
    if (!ctx) return RS_ERR_ARG;
    ctx->profiling = on != 0;
    return RS_OK;
}
static int profile_drain(rs_ctx* ctx) {
    RS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (auto& ev : ctx->events) {
        float ms = 0.f;
        RS_CUDA(ctx, cudaEventElapsedTime(&ms, ev.start, ev.stop));
        ctx->prof_ms[ev.kind] += ms;
        ctx->prof_n[ev.kind] += 1;
        ctx->pool.push_back(ev);
    }
    ctx->events.clear();
    return RS_OK;
}
int rs_profile_get(rs_ctx* ctx, int kind, double* total_ms, uint64_t* launches) {
    DeviceGuard dg(ctx);
    if (!ctx || kind < 0 || kind >= RS_K_COUNT) return fail(ctx, RS_ERR_ARG, "rs_profile_get: bad argument");
    if (int r = profile_drain(ctx)) return r;
    if (total_ms) *total_ms = ctx->prof_ms[kind];
    if (launches) *launches = ctx->prof_n[kind];
    return RS_OK;
}
int rs_profile_reset(rs_ctx* ctx) {
    DeviceGuard dg(ctx);
    if (!ctx) return RS_ERR_ARG;
    if (int r = profile_drain(ctx)) return r;
    for (int k = 0; k < RS_K_COUNT; k++) { ctx->prof_ms[k] = 0; ctx->prof_n[k] = 0; }
    return RS_OK;
}
uint64_t rs_launch_count(const rs_ctx* ctx) { return ctx ? ctx->launches : 0; }

int rs_fp64_peak(rs_ctx* ctx, double* tflops) {
    if (!ctx || !tflops) return fail(ctx, RS_ERR_ARG, "rs_fp64_peak: NULL argument");
    DeviceGuard dg(ctx);
    const int iters = 4000;
    double *out = nullptr, *in = nullptr;
    RS_CUDA(ctx, cudaMalloc(&out, (size_t)ctx->sm_count * 2 * 512 * sizeof(double)));
    RS_CUDA(ctx, cudaMalloc(&in, 32 * sizeof(double)));
    RS_CUDA(ctx, cudaMemsetAsync(in, 0, 32 * sizeof(double), ctx->stream));
    cudaEvent_t e0, e1;
    RS_CUDA(ctx, cudaEventCreate(&e0));
    RS_CUDA(ctx, cudaEventCreate(&e1));
    double best_tf = 0.0;
    const int shapes[3][2] = {{ctx->sm_count, 384}, {ctx->sm_count, 512}, {ctx->sm_count * 2, 512}};   // 12, 16, 32 warps per SM
    for (auto& sh : shapes) {
        for (int rep = 0; rep < 4; rep++) {
            RS_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
            fp64_peak_kernel<<<sh[0], sh[1], 0, ctx->stream>>>(out, in, iters);
            ctx->launches++;
            RS_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
            RS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            float ms = 0.f;
            RS_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
            const double tf = 2.0 * 32.0 * (double)iters * sh[0] * sh[1] / (ms * 1e-3) / 1e12;
            if (rep > 0 && tf > best_tf) best_tf = tf;
        }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out); cudaFree(in);
    *tflops = best_tf;
    return RS_OK;
}

int rs_fp64_peak_three_operand(rs_ctx* ctx, double* tflops) {
    DeviceGuard dg(ctx);
    if (!ctx || !tflops) return fail(ctx, RS_ERR_ARG, "rs_fp64_peak_three_operand: NULL argument");
    const int block = 256, grid = ctx->sm_count * 8, iters = 20000;
    double *out = nullptr, *in = nullptr;
    RS_CUDA(ctx, cudaMalloc(&out, (size_t)grid * block * sizeof(double)));
    RS_CUDA(ctx, cudaMalloc(&in, 32 * sizeof(double)));
    RS_CUDA(ctx, cudaMemsetAsync(in, 0, 32 * sizeof(double), ctx->stream));
    cudaEvent_t e0, e1;
    RS_CUDA(ctx, cudaEventCreate(&e0));
    RS_CUDA(ctx, cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        RS_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
        fp64_peak3_kernel<<<grid, block, 0, ctx->stream>>>(out, in, iters);
        ctx->launches++;
        RS_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
        RS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        float ms = 0.f;
        RS_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out); cudaFree(in);
    *tflops = 2.0 * 8.0 * (double)iters * grid * block / (best * 1e-3) / 1e12;
    return RS_OK;
}

#ifdef RS_ERR_STATS
int rs_debug_err_stats(unsigned long long* hist8, double* max_err) {   // debug builds only
    cudaDeviceSynchronize();
    unsigned long long bits = 0;
    if (cudaMemcpyFromSymbol(hist8, rs::g_err_hist, sizeof(unsigned long long) * 8) != cudaSuccess) return 1;
    if (cudaMemcpyFromSymbol(&bits, rs::g_err_max_bits, sizeof(bits)) != cudaSuccess) return 1;
    memcpy(max_err, &bits, sizeof(double));
    return 0;
}
#endif

#ifdef RS_WS_PROF
int rs_debug_ws_prof(long long* out96) {   // debug builds only: phase timers of CTA 0, [12 warps][8 phases]
    cudaDeviceSynchronize();
    return cudaMemcpyFromSymbol(out96, rs::g_ws_prof, sizeof(long long) * 96) == cudaSuccess ? 0 : 1;
}
int rs_debug_ws_rowwait(long long* out60) {   // [3][20], see blind_rotate_ws.cuh
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out60, rs::g_ws_rowwait, sizeof(long long) * 60) != cudaSuccess) return 1;
    static const long long zero[60] = {0};
    return cudaMemcpyToSymbol(rs::g_ws_rowwait, zero, sizeof(zero)) == cudaSuccess ? 0 : 1;   // read and reset
}
#endif

#ifdef RS_BR_STATS
int* rs_debug_progress() {   // host-mapped progress words of block 0 (debug builds only)
    static int* host = nullptr;
    if (!host) {
        cudaHostAlloc(&host, 64 * sizeof(int), cudaHostAllocMapped);
        memset(host, 0xff, 64 * sizeof(int));
        int* dev = nullptr;
        cudaHostGetDevicePointer(&dev, host, 0);
        cudaMemcpyToSymbol(rs::g_br_progress, &dev, sizeof(dev));
    }
    return host;
}
int rs_debug_stats(unsigned long long* out8, int reset) {
    cudaDeviceSynchronize();
    if (out8) cudaMemcpyFromSymbol(out8, rs::g_br_stats, sizeof(unsigned long long) * 8);
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(rs::g_br_stats, z, sizeof(z)); }
    return 0;
}
#endif

int rs_set_ks_variant(rs_ctx* ctx, int ks_variant) {
    if (!ctx || ks_variant < 0 || ks_variant > 3) return fail(ctx, RS_ERR_ARG, "rs_set_ks_variant: 0 (auto), 1 (un-tiled gather), 2 (tensor cores) or 3 (shared-memory gather)");
    ctx->ks_variant = ks_variant;
    return RS_OK;
}

int rs_set_tuning(rs_ctx* ctx, int br_variant) {
    if (!ctx || br_variant < 0 || br_variant > 4) return fail(ctx, RS_ERR_ARG, "rs_set_tuning: br_variant must be 0..4");
    ctx->br_variant = br_variant;
    return RS_OK;
}

int rs_device_info(rs_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, size_t* smem_optin) {
    DeviceGuard dg(ctx);
    if (!ctx) return RS_ERR_ARG;
    cudaDeviceProp prop;
    RS_CUDA(ctx, cudaGetDeviceProperties(&prop, ctx->device));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (smem_optin) *smem_optin = prop.sharedMemPerBlockOptin;
    return RS_OK;
}

}  // extern "C"
