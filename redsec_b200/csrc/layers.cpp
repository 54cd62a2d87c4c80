// redsec_b200/csrc/layers.cpp -- host-side Layer forward over the C-ABI.
//
// What the reference does per neuron / per gate (lib/BinFunc.cpp:1044-1075, lib/IntFunc.cpp:860-889,
// lib/BinFunc.cpp:880-925, GPU twins lib/GPU/BinFunc_gpu.cu:465-531,591-629) this file does per layer:
//   linear part (conv / FC / sum-pool / bias) -> ONE rs_pbs_batch over every neuron of the layer
//   -> max-pool as an OR tree, one batched launch per tree level.
// Host C++ only: every device operation is an rs_* call from include/redsec_b200.h.
//
// Encodings: bit b <-> torus +-1/4096; bias b <-> trivial LWE modSwitchToTorus32(b,4096)
// (lib/BinOps_enc.cpp:292-295); sign activation = bootstrap with mu=1/4096 (lib/BinOps_enc.cpp:182-186).
// Max-pool (reference defect R3): the reference ORs +-1/4096 bits with a gate that expects +-1/8,
// starting from an uninitialised accumulator (lib/BinFunc.cpp:891,917).  Here the sign bootstrap in front of a
// max-pool emits +-1/8, the OR tree runs at +-1/8 and its last level emits +-1/4096 again.
// DoReFa ReLU (lib/IntFunc.cpp:934-973): the reference's encrypted branch (multiply_pc_ints on a
// never-cleared scratch + binarize + bootsMUX on non-gate encodings, defect R6) cannot work; the plaintext branch defines
// out = clamp((slope*x + bias) >> slope_bits, 0, 2^shift_bits - 1).  Here that staircase is ONE test-vector bootstrap per
// neuron (rs_pbs_lut_batch) with one table per output channel; see relu_tables().
//
// Launch structure of a max-pool layer: the three dependent bootstrap launches (sign, OR level 1, OR level 2) each end in a
// partly filled wave of CTAs.  The layer is therefore cut into blocks of output rows -- a block's sign neurons, OR gates and
// pooled outputs depend on nothing outside it -- and the blocks' chains are issued round-robin on the context's lanes
// (rs_lanes), so the tail wave of one launch is filled by CTAs of another block's launch.
//
// Multi-GPU: execute_sharded() = this rank's output-channel block -> rs_allgather (NCCL on the engine stream) ->
// rs_lwe_interleave; Net::run_sharded chains the layers without a single host synchronisation.
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <utility>

#include "../host/redsec_layers.hpp"

namespace redsec {
namespace {

constexpr uint32_t kUnit = 1u << 20;        // 1/4096 on the torus
constexpr uint32_t kEighth = 1u << 29;      // 1/8
constexpr int kTile = 16;                   // CONV_OD_TILE of the conv kernel
constexpr int kMaxLanes = 2;            // see run_maxpool_sign: three or more concurrent chains were measured to corrupt rows
constexpr size_t kWaveCts = 4 * 148;        // ciphertexts of one full wave of blind-rotate CTAs (4 per SM)

#define RS_TRY(call) do { int rc_ = (call); if (rc_ != RS_OK) return rc_; } while (0)

struct Csr {   // host-side sparse rows: out[o] = (0,bias[o]) + sum w*in[col]
    std::vector<int32_t> rowptr{0}, col;
    std::vector<int8_t> w;
    std::vector<uint32_t> bias;
    void entry(int32_t c, int8_t weight) { col.push_back(c); w.push_back(weight); }
    void end_row(uint32_t b) { rowptr.push_back((int32_t)col.size()); bias.push_back(b); }
    size_t rows() const { return rowptr.size() - 1; }
};

struct DevCsr {
    rs_ctx* ctx = nullptr;
    void *rowptr = nullptr, *col = nullptr, *w = nullptr, *bias = nullptr;
    size_t rows = 0;
    DevCsr() = default;
    DevCsr(const DevCsr&) = delete;
    DevCsr& operator=(const DevCsr&) = delete;
    ~DevCsr() { release(); }
    void release() {
        if (!ctx) return;
        rs_dev_free(ctx, rowptr); rs_dev_free(ctx, col); rs_dev_free(ctx, w); rs_dev_free(ctx, bias);
        rowptr = col = w = bias = nullptr;
    }
    int upload(rs_ctx* c, const Csr& h) {
        release();
        ctx = c; rows = h.rows();
        auto up = [&](void** dst, const void* src, size_t bytes) -> int {
            RS_TRY(rs_dev_alloc(ctx, bytes ? bytes : 4, dst));
            if (bytes) RS_TRY(rs_dev_upload(ctx, *dst, src, bytes));
            return RS_OK;
        };
        RS_TRY(up(&rowptr, h.rowptr.data(), h.rowptr.size() * 4));
        RS_TRY(up(&col, h.col.data(), h.col.size() * 4));
        RS_TRY(up(&w, h.w.data(), h.w.size()));
        RS_TRY(up(&bias, h.bias.data(), h.bias.size() * 4));
        return RS_OK;
    }
    int apply(uint32_t* out, const uint32_t* in) const { return apply_range(out, in, 0, rows); }
    // rows [r0,r1) only; `out` receives them starting at its row 0 (column indices stay absolute into `in`)
    int apply_range(uint32_t* out, const uint32_t* in, size_t r0, size_t r1) const {
        if (r1 <= r0) return RS_OK;
        return rs_lwe_lincomb(ctx, out, r1 - r0, in, (const int32_t*)rowptr + r0, (const int32_t*)col, (const int8_t*)w,
                              (const uint32_t*)bias + r0);
    }
};

// one max-pool tree level: rows reading the tree buffer, bootstrapped with mu
struct PoolStep {
    DevCsr csr;
    size_t dst_offset = 0;            // row offset in the tree buffer where the bootstrapped results land
    uint32_t mu = kEighth;
    std::vector<uint32_t> first_row;  // [n_out+1]: rows of pooled outputs < o come before first_row[o] (rows are emitted in output order)
};

struct PoolPlan {   // OR tree for one channel count
    std::vector<std::unique_ptr<PoolStep>> steps;
    size_t in_count = 0, buf_count = 0, out_count = 0;
    DevCsr final_gather;      // tree buffer -> canonical (oph,opw,c) order
    size_t gates = 0;
    bool blockable = false;   // windows tile the input (stride == window, valid padding): blocks of output rows are independent
};

struct ConvWeights { void* packed = nullptr; void* bias = nullptr; };

int out_extent_same(int in, int stride) { return (in - 1) / stride + 1; }

// device rows owned until released: every early return of forward() gives its temporaries back to the pool
struct Rows {
    rs_ctx* ctx; uint32_t* p = nullptr;
    explicit Rows(rs_ctx* c) : ctx(c) {}
    Rows(const Rows&) = delete;
    Rows& operator=(const Rows&) = delete;
    ~Rows() { reset(); }
    int alloc(size_t count) { reset(); return rs_lwe_alloc(ctx, count, &p); }
    void reset(uint32_t* q = nullptr) { if (p) rs_lwe_free(ctx, p); p = q; }
    uint32_t* release() { uint32_t* q = p; p = nullptr; return q; }
};

}  // namespace

class LayerImpl {
public:
    rs_ctx* ctx = nullptr;
    bool int_inputs = false;
    eConvType ec = E_NO_CONV; uint16_t depth = 1; ePoolType ep = E_NO_POOL; eQuantType eq = E_ACTIVATION_NONE;
    tNetParams np{};
    bool prepared = false;
    // conv geometry (after the FC flatten)
    bool has_conv = false;
    int cin_h = 0, cin_w = 0, cin_dep = 0, cout_h = 0, cout_w = 0, cout_dep = 0, ofs_h = 0, ofs_w = 0;
    std::vector<int8_t> weights;        // [K][cout_dep] in {-1,0,1}, K = (fh*win_w+fw)*cin_dep+di
    // sum-pool geometry
    bool has_sumpool = false;
    int sp_in_h = 0, sp_in_w = 0, sp_out_h = 0, sp_out_w = 0, sp_ofs_h = 0, sp_ofs_w = 0;
    // quantize
    int q_h = 0, q_w = 0, q_dep = 0;    // dims at the activation
    std::vector<uint32_t> bias_torus;   // [q_dep]
    void* bias_dev = nullptr;           // same on the device (stage use: rs_lwe_add_bias)
    // DoReFa ReLU (IntLayer only)
    std::vector<int32_t> bias_int, slope;   // [q_dep]; slope empty = 1
    int shift_bits = 0, slope_bits = 0;
    // IntFunc conv conventions: false = the reference's ENCRYPTED branch (zero weight / padding contribute -1/4096,
    // lib/IntFunc.cpp:268,277); true = its plaintext twin (weight -1 contributes ~x = -x-1, IntOps::invert
    // lib/IntOps.cpp:72-82; zero weight 0), which is what the shipped ReLU weights were converted for
    bool twin_conv = false;
    std::vector<int32_t> neg_count;     // [cout_dep] number of -1 weights per output channel (twin_conv)
    // max-pool geometry
    bool has_maxpool = false;
    int mp_out_h = 0, mp_out_w = 0;
    // caches keyed by channel slice
    std::map<std::pair<std::pair<int, int>, bool>, ConvWeights> conv_cache;   // (channel range, bias fused)
    typedef std::pair<std::pair<int, int>, std::pair<int, int>> SliceKey;     // (channel range, output-pixel range)
    std::map<std::pair<SliceKey, bool>, std::unique_ptr<DevCsr>> sumpool_cache, identity_cache;
    std::map<int, std::unique_ptr<PoolPlan>> maxpool_cache;
    std::map<std::pair<int, int>, void*> relu_cache;

    explicit LayerImpl(rs_ctx* c) : ctx(c) { if (ctx) rs_ctx_retain(ctx); }
    ~LayerImpl() {
        if (!ctx) return;
        for (auto& kv : conv_cache) { rs_dev_free(ctx, kv.second.packed); rs_dev_free(ctx, kv.second.bias); }
        for (auto& kv : relu_cache) rs_dev_free(ctx, kv.second);
        rs_dev_free(ctx, bias_dev);
        sumpool_cache.clear(); identity_cache.clear(); maxpool_cache.clear();   // their DevCsr members free through ctx
        rs_ctx_release(ctx);
    }
    bool is_relu() const { return eq == E_ACTIVATION_RELU; }
    bool conv_is_twin() const { return twin_conv && int_inputs && has_conv; }

    int channels() const { return q_dep; }
    size_t final_h() const { return has_maxpool ? mp_out_h : q_h; }
    size_t final_w() const { return has_maxpool ? mp_out_w : q_w; }

    // ---- weight file blocks (readers lib/BinOps_enc.cpp:247-297)
    static int read_ternary(FILE* fd, std::vector<int8_t>& out, size_t len) {
        uint8_t tag = 0;
        if (fread(&tag, 1, 1, fd) != 1 || (tag != 1 && tag != 2)) return RS_ERR_ARG;
        const int nbits = tag == 1 ? 1 : 2;
        std::vector<uint8_t> packed((len * nbits + 7) / 8);
        if (fread(packed.data(), 1, packed.size(), fd) != packed.size()) return RS_ERR_ARG;
        out.resize(len);
        for (size_t i = 0; i < len; i++) {
            const size_t bit = i * nbits;
            const int sign = (packed[bit >> 3] >> (7 - (bit & 7))) & 1;               // MSB first: 1 -> +1, 0 -> -1
            int zero = 0;
            if (nbits == 2) zero = (packed[(bit + 1) >> 3] >> (7 - ((bit + 1) & 7))) & 1;
            out[i] = zero ? 0 : (sign ? 1 : -1);
        }
        return RS_OK;
    }
    static int read_ints(FILE* fd, std::vector<int32_t>& out, size_t len) {
        uint8_t tag = 0;
        if (fread(&tag, 1, 1, fd) != 1 || (tag != 3 && tag != 4)) return RS_ERR_ARG;
        out.resize(len);
        return fread(out.data(), 4, len, fd) == len ? RS_OK : RS_ERR_ARG;
    }

    // ---- prep, one piece per Func object of the reference ({Bin,Int}Layer::run(E_PREP): lib/BinLayer.cpp:150-241,
    // lib/IntLayer.cpp:153-235).  Each piece updates *dim exactly as the corresponding Func::prep does.
    int prep_conv(FILE* fd, tDimensions* dim) {   // Convolution::prep, lib/BinFunc.cpp:76-133
        if (!fd || !dim) return RS_ERR_ARG;
        has_conv = true;
        const tConvParams& cv = np.conv;
        cin_h = dim->hw.h; cin_w = dim->hw.w; cin_dep = (int)dim->in_dep; cout_dep = depth;
        if (cv.window.h < 1 || cv.window.w < 1 || cv.stride.h < 1 || cv.stride.w < 1) return RS_ERR_ARG;
        if (cv.same_pad) {   // lib/BinFunc.cpp:87-95
            cout_h = out_extent_same(cin_h, cv.stride.h); cout_w = out_extent_same(cin_w, cv.stride.w);
            ofs_h = cv.stride.h == 1 ? (cv.window.h - 1) / 2 : (cout_h * cv.stride.h - cin_h) / 2;
            ofs_w = cv.stride.w == 1 ? (cv.window.w - 1) / 2 : (cout_w * cv.stride.w - cin_w) / 2;
        } else {             // lib/BinFunc.cpp:96-104
            ofs_h = ofs_w = 0;
            cout_h = (cin_h - 2 * ((cv.window.h - 1) / 2)) / cv.stride.h;
            cout_w = (cin_w - 2 * ((cv.window.w - 1) / 2)) / cv.stride.w;
        }
        const size_t K = (size_t)cv.window.h * cv.window.w * cin_dep;
        RS_TRY(read_ternary(fd, weights, K * cout_dep));
        neg_count.assign(cout_dep, 0);
        for (size_t k = 0; k < K; k++)
            for (int od = 0; od < cout_dep; od++) neg_count[od] += weights[k * cout_dep + od] < 0;
        dim->up_bound *= (uint32_t)(dim->filter_bits * cv.window.w * cv.window.h) * dim->in_dep;
        for (dim->in_bits = dim->in_bits; (dim->up_bound >> dim->in_bits) > 0; dim->in_bits++) {}
        dim->hw.h = (int16_t)cout_h; dim->hw.w = (int16_t)cout_w; dim->in_dep = cout_dep;
        q_h = cout_h; q_w = cout_w; q_dep = cout_dep;
        return RS_OK;
    }
    int prep_sumpool(tDimensions* dim) {          // SumPooling::prep, lib/IntFunc.cpp:598-634
        if (!dim) return RS_ERR_ARG;
        has_sumpool = true;
        const tPoolParams& pl = np.pool;
        if (pl.window.h < 1 || pl.window.w < 1 || pl.stride.h < 1 || pl.stride.w < 1) return RS_ERR_ARG;
        sp_in_h = dim->hw.h; sp_in_w = dim->hw.w;
        if (pl.same_pad) {
            sp_out_h = out_extent_same(sp_in_h, pl.stride.h); sp_out_w = out_extent_same(sp_in_w, pl.stride.w);
            sp_ofs_h = pl.stride.h == 1 ? (pl.window.h - 1) / 2 : (sp_out_h * pl.stride.h - sp_in_h) / 2;
            sp_ofs_w = pl.stride.w == 1 ? (pl.window.w - 1) / 2 : (sp_out_w * pl.stride.w - sp_in_w) / 2;
        } else {
            sp_ofs_h = sp_ofs_w = 0;
            sp_out_h = (sp_in_h - pl.window.h / 2 - 1) / pl.stride.h + 1;
            sp_out_w = (sp_in_w - pl.window.w / 2 - 1) / pl.stride.w + 1;
        }
        dim->up_bound *= (uint32_t)(pl.window.w * pl.window.h);
        dim->scale *= (float)(pl.window.w * pl.window.h);
        dim->hw.h = (int16_t)sp_out_h; dim->hw.w = (int16_t)sp_out_w;
        q_h = sp_out_h; q_w = sp_out_w; q_dep = (int)dim->in_dep;
        return RS_OK;
    }
    // Quantize::prep: bias block of length in_dep (lib/BinFunc.cpp:1001-1003), slope block for a batch-normed ReLU
    // (lib/IntFunc.cpp:800-840)
    int prep_quant(FILE* fd, tDimensions* dim, bool read_slope) {
        if (!fd || !dim) return RS_ERR_ARG;
        q_h = dim->hw.h; q_w = dim->hw.w; q_dep = (int)dim->in_dep;
        RS_TRY(read_ints(fd, bias_int, (size_t)q_dep));
        bias_torus.resize(q_dep);
        for (int i = 0; i < q_dep; i++) bias_torus[i] = (uint32_t)bias_int[i] * kUnit;   // modSwitchToTorus32(b, 4096)
        if (is_relu()) {
            if (!int_inputs) return RS_ERR_STATE;   // BinFunc::Quantize::relu_shift (no shipped net uses it) is not built
            shift_bits = np.quant.shift_bits;
            if (shift_bits < 2 || shift_bits > 8) return RS_ERR_ARG;
            if (read_slope) RS_TRY(read_ints(fd, slope, (size_t)q_dep));                 // slope block (IntLayer.cpp:96-100)
            int sc_b = 0;
            while ((float)(1 << sc_b) < dim->scale) sc_b++;                              // log2(scale), IntFunc.cpp:813-814
            slope_bits = 8 + sc_b - shift_bits;                                          // SLOPE_BITS = 8 (IntFunc.cpp:45,815)
            if (slope_bits < 0) return RS_ERR_ARG;
            dim->in_bits = (uint8_t)shift_bits;
            dim->scale = (float)((1 << shift_bits) - 1);
            dim->up_bound = 1u << (shift_bits - 1);
        }
        if (eq == E_ACTIVATION_SIGN) { dim->in_bits = 1; dim->up_bound = 1; dim->scale = 1.0f; }
        dim->out_bits = SINGLE_BIT;
        return RS_OK;
    }
    int prep_maxpool(tDimensions* dim) {          // MaxPooling::prep, lib/BinFunc.cpp:836-872
        if (!dim) return RS_ERR_ARG;
        has_maxpool = true;
        const tPoolParams& pl = np.pool;
        if (pl.window.h < 1 || pl.window.w < 1 || pl.stride.h < 1 || pl.stride.w < 1) return RS_ERR_ARG;
        q_h = dim->hw.h; q_w = dim->hw.w; q_dep = (int)dim->in_dep;
        if (pl.same_pad) { mp_out_h = out_extent_same(q_h, pl.stride.h); mp_out_w = out_extent_same(q_w, pl.stride.w); }
        else { mp_out_h = q_h / pl.window.h; mp_out_w = q_w / pl.window.w; }
        dim->hw.h = (int16_t)mp_out_h; dim->hw.w = (int16_t)mp_out_w;
        return RS_OK;
    }
    int prep(FILE* fd, tDimensions* dim) {
        if (prepared || !fd || !dim) return RS_ERR_ARG;
        if (ec == E_FC || ec == E_FC_FINAL) {   // flatten (lib/BinLayer.cpp:157-167)
            dim->in_dep *= (uint32_t)dim->hw.h * dim->hw.w;
            dim->hw.h = 1; dim->hw.w = 1;
        }
        if (ec != E_NO_CONV) RS_TRY(prep_conv(fd, dim));
        if (ep == E_SUMPOOL) RS_TRY(prep_sumpool(dim));
        RS_TRY(prep_quant(fd, dim, np.e_bias == E_BNORM));
        if (ep == E_MAXPOOL && eq == E_ACTIVATION_SIGN && ec != E_FC_FINAL) RS_TRY(prep_maxpool(dim));   // lib/BinFunc.cpp:855-866
        prepared = true;
        return RS_OK;
    }

    // ---- device tables, built lazily per output-channel slice [c0,c1)
    // fuse_bias: the layer bias rides on the conv (layer forward, when nothing linear follows the conv and the activation is
    // not a ReLU, whose bias lives inside the test vector); stage use passes false (Quantize adds the bias)
    int conv_tables(int c0, int c1, bool fuse_bias, ConvWeights** out) {
        auto key = std::make_pair(std::make_pair(c0, c1), fuse_bias);
        auto it = conv_cache.find(key);
        if (it == conv_cache.end()) {
            const size_t K = (size_t)np.conv.window.h * np.conv.window.w * cin_dep;
            const int cl = c1 - c0, tiles = (cl + kTile - 1) / kTile;
            std::vector<int8_t> packed((size_t)tiles * K * kTile, 0);
            for (int t = 0; t < tiles; t++)
                for (size_t k = 0; k < K; k++)
                    for (int o = 0; o < kTile; o++) {
                        const int od = c0 + t * kTile + o;
                        if (od < c1) packed[((size_t)t * K + k) * kTile + o] = weights[k * cout_dep + od];
                    }
            ConvWeights cw;
            RS_TRY(rs_dev_alloc(ctx, packed.size(), &cw.packed));
            if (int rc = rs_dev_upload(ctx, cw.packed, packed.data(), packed.size())) { rs_dev_free(ctx, cw.packed); return rc; }
            // the plaintext-twin convention adds -1 per negative weight
            if (fuse_bias || conv_is_twin()) {
                std::vector<uint32_t> cb(cl, 0u);
                for (int c = 0; c < cl; c++) {
                    if (fuse_bias) cb[c] = bias_torus[c0 + c];
                    if (conv_is_twin()) cb[c] -= (uint32_t)neg_count[c0 + c] * kUnit;
                }
                int rc = rs_dev_alloc(ctx, (size_t)cl * 4, &cw.bias);
                if (rc == RS_OK) rc = rs_dev_upload(ctx, cw.bias, cb.data(), (size_t)cl * 4);
                if (rc != RS_OK) { rs_dev_free(ctx, cw.packed); rs_dev_free(ctx, cw.bias); return rc; }
            }
            it = conv_cache.emplace(key, cw).first;
        }
        *out = &it->second;
        return RS_OK;
    }

    // sum-pool over a local channel slice; rows (oph,opw,cl), inputs (ih,iw,cl) (lib/IntFunc.cpp:665-697)
    int sumpool_table(int c0, int c1, bool with_bias, DevCsr** out, int p0 = 0, int p1 = -1) {
        if (p1 < 0) p1 = sp_out_h * sp_out_w;
        auto key = std::make_pair(std::make_pair(std::make_pair(c0, c1), std::make_pair(p0, p1)), with_bias);
        auto it = sumpool_cache.find(key);
        if (it == sumpool_cache.end()) {
            const tPoolParams& pl = np.pool;
            const int cl = c1 - c0;
            Csr csr;
            for (int oph = 0; oph < sp_out_h; oph++)
                for (int opw = 0; opw < sp_out_w; opw++)
                    for (int c = 0; c < cl; c++) {
                        if (oph * sp_out_w + opw < p0 || oph * sp_out_w + opw >= p1) continue;   // pixel-sharded input layers
                        const int ih0 = oph * pl.stride.h - sp_ofs_h, iw0 = opw * pl.stride.w - sp_ofs_w;
                        for (int fh = 0; fh < pl.window.h && ih0 + fh < sp_in_h; fh++) {
                            if (ih0 + fh < 0) continue;
                            for (int fw = 0; fw < pl.window.w && iw0 + fw < sp_in_w; fw++) {
                                if (iw0 + fw < 0) continue;
                                csr.entry(((ih0 + fh) * sp_in_w + iw0 + fw) * cl + c, 1);
                            }
                        }
                        csr.end_row(with_bias ? bias_torus[c0 + c] : 0u);
                    }
            auto dev = std::make_unique<DevCsr>();
            RS_TRY(dev->upload(ctx, csr));
            it = sumpool_cache.emplace(key, std::move(dev)).first;
        }
        *out = it->second.get();
        return RS_OK;
    }

    // bias add only (E_NO_CONV without pooling, e.g. CIFAR layer 0); input (h,w,C), output slice (h,w,cl)
    int identity_table(int c0, int c1, bool with_bias, DevCsr** out, int p0 = 0, int p1 = -1) {
        if (p1 < 0) p1 = q_h * q_w;
        auto key = std::make_pair(std::make_pair(std::make_pair(c0, c1), std::make_pair(p0, p1)), with_bias);
        auto it = identity_cache.find(key);
        if (it == identity_cache.end()) {
            Csr csr;
            for (int p = p0; p < p1; p++)
                for (int c = c0; c < c1; c++) { csr.entry(p * q_dep + c, 1); csr.end_row(with_bias ? bias_torus[c] : 0u); }
            auto dev = std::make_unique<DevCsr>();
            RS_TRY(dev->upload(ctx, csr));
            it = identity_cache.emplace(key, std::move(dev)).first;
        }
        *out = it->second.get();
        return RS_OK;
    }

    // Test vectors of the encrypted DoReFa ReLU for channels [c0,c1): uint32[cl][1024] on the device.  The neuron value x
    // sits on the torus in units of 1/4096 and a bootstrap resolves 2N = 2048 phase slots, so slot j <-> x = 2j.  The
    // staircase f(x) = clamp((slope*x + bias) >> slope_bits, 0, 2^shift_bits - 1) saturates on both sides, so
    // g = f - (2^shift_bits - 1)/2 extends negacyclically: slots [0,512) hold g(2j) (x in [0,1024); x in [-2048,-1024) reads
    // -g there, the saturated value), slots [512,1024) hold -g(2j-2048) (x in [-1024,0)).  forward() adds the constant back.
    uint32_t relu_half() const { return (uint32_t)((1 << shift_bits) - 1) * (kUnit / 2); }
    int relu_tables(int c0, int c1, void** out) {
        auto key = std::make_pair(c0, c1);
        auto it = relu_cache.find(key);
        if (it == relu_cache.end()) {
            const int cl = c1 - c0;
            std::vector<uint32_t> tv((size_t)cl * RS_TLWE_N);
            const int64_t top = (1 << shift_bits) - 1;
            for (int c = 0; c < cl; c++) {
                const int64_t sl = slope.empty() ? 1 : slope[c0 + c], b = bias_int[c0 + c];
                for (int j = 0; j < RS_TLWE_N; j++) {
                    const int64_t x = j < RS_TLWE_N / 2 ? 2 * j : 2 * j - 2 * RS_TLWE_N;
                    int64_t v = (sl * x + b) >> slope_bits;        // IntOps::shift: arithmetic shift (lib/IntOps.cpp:194)
                    v = v < 0 ? 0 : (v > top ? top : v);           // IntOps::relu (lib/IntOps.cpp:152-171)
                    const uint32_t g = (uint32_t)v * kUnit - relu_half();
                    tv[(size_t)c * RS_TLWE_N + j] = j < RS_TLWE_N / 2 ? g : 0u - g;
                }
            }
            void* dev = nullptr;
            RS_TRY(rs_dev_alloc(ctx, tv.size() * 4, &dev));
            if (int rc = rs_dev_upload(ctx, dev, tv.data(), tv.size() * 4)) { rs_dev_free(ctx, dev); return rc; }
            it = relu_cache.emplace(key, dev).first;
        }
        *out = it->second;
        return RS_OK;
    }

    // OR tree over each pooling window for cl local channels (replaces the depth-4 chain of lib/BinFunc.cpp:896-919)
    int maxpool_plan(int cl, PoolPlan** out) {
        auto it = maxpool_cache.find(cl);
        if (it == maxpool_cache.end()) {
            const tPoolParams& pl = np.pool;
            auto plan = std::make_unique<PoolPlan>();
            const size_t n_out = (size_t)mp_out_h * mp_out_w * cl;
            plan->in_count = (size_t)q_h * q_w * cl;
            plan->out_count = n_out;
            plan->blockable = !pl.same_pad && pl.stride.h == pl.window.h && pl.stride.w == pl.window.w;
            std::vector<std::vector<int32_t>> items(n_out);   // indices into the tree buffer
            for (int oph = 0; oph < mp_out_h; oph++)
                for (int opw = 0; opw < mp_out_w; opw++)
                    for (int c = 0; c < cl; c++) {
                        auto& v = items[((size_t)oph * mp_out_w + opw) * cl + c];
                        const int ih0 = oph * pl.stride.h, iw0 = opw * pl.stride.w;   // offset_window is 0 (reference R4)
                        for (int fh = 0; fh < pl.window.h && ih0 + fh < q_h; fh++)
                            for (int fw = 0; fw < pl.window.w && iw0 + fw < q_w; fw++)
                                v.push_back(((ih0 + fh) * q_w + iw0 + fw) * cl + c);
                    }
            size_t next = plan->in_count;
            std::vector<int32_t> final_pos(n_out, -1);
            bool more = true;
            while (more) {
                more = false;
                Csr mid, fin;                 // OR gates whose result is reduced further / is the pooled bit
                std::vector<size_t> fin_owner;
                std::vector<std::pair<size_t, size_t>> mid_slots;   // (output, slot in its new item list)
                std::vector<std::vector<int32_t>> nxt(n_out);
                std::vector<uint32_t> mid_first(n_out + 1, 0), fin_first(n_out + 1, 0);
                for (size_t o = 0; o < n_out; o++) {
                    mid_first[o] = (uint32_t)mid.rows(); fin_first[o] = (uint32_t)fin.rows();
                    auto& v = items[o];
                    if (final_pos[o] >= 0 || v.empty()) continue;
                    if (v.size() == 1) {      // single element window: re-encode +-1/8 -> +-1/4096 with one bootstrap
                        fin.entry(v[0], 1); fin.end_row(0); fin_owner.push_back(o);
                        continue;
                    }
                    const bool last = v.size() == 2;
                    for (size_t k = 0; k + 1 < v.size(); k += 2) {
                        Csr& dst = last ? fin : mid;
                        dst.entry(v[k], 1); dst.entry(v[k + 1], 1); dst.end_row(kEighth);   // OR: (0,1/8)+a+b
                        if (last) fin_owner.push_back(o);
                        else { mid_slots.push_back({o, nxt[o].size()}); nxt[o].push_back(-1); }
                    }
                    if (v.size() & 1) nxt[o].push_back(v.back());
                }
                mid_first[n_out] = (uint32_t)mid.rows(); fin_first[n_out] = (uint32_t)fin.rows();
                plan->gates += mid.rows() + fin.rows();
                if (mid.rows()) {
                    auto st = std::make_unique<PoolStep>();
                    RS_TRY(st->csr.upload(ctx, mid));
                    st->dst_offset = next; st->mu = kEighth; st->first_row = std::move(mid_first);
                    for (size_t r = 0; r < mid_slots.size(); r++) nxt[mid_slots[r].first][mid_slots[r].second] = (int32_t)(next + r);
                    next += mid.rows();
                    plan->steps.push_back(std::move(st));
                    more = true;
                }
                if (fin.rows()) {
                    auto st = std::make_unique<PoolStep>();
                    RS_TRY(st->csr.upload(ctx, fin));
                    st->dst_offset = next; st->mu = kUnit; st->first_row = std::move(fin_first);
                    for (size_t r = 0; r < fin_owner.size(); r++) final_pos[fin_owner[r]] = (int32_t)(next + r);
                    next += fin.rows();
                    plan->steps.push_back(std::move(st));
                }
                for (size_t o = 0; o < n_out; o++) if (final_pos[o] < 0) items[o] = std::move(nxt[o]);
            }
            plan->buf_count = next;
            Csr gather;
            for (size_t o = 0; o < n_out; o++) { gather.entry(final_pos[o], 1); gather.end_row(0); }
            RS_TRY(plan->final_gather.upload(ctx, gather));
            it = maxpool_cache.emplace(cl, std::move(plan)).first;
        }
        *out = it->second.get();
        return RS_OK;
    }

    // builds (and caches) every device table forward() needs for the channel slice [c0,c1): packed conv weights, pooling /
    // bias CSR rows, the max-pool OR tree, the ReLU test vectors.  forward() builds them lazily; calling this from prep keeps
    // the host-side packing out of the first inference, like the reference's prep() (weights are read before "Inference Time").
    int build_tables(int c0, int c1, int p0 = 0, int p1 = -1) {
        if (!prepared) return RS_ERR_STATE;
        if (!ctx) return RS_ERR_STATE;
        if (has_conv) { ConvWeights* cw = nullptr; RS_TRY(conv_tables(c0, c1, conv_fuses_bias(), &cw)); }
        if (has_sumpool) {
            if (!has_conv && (c0 != 0 || c1 != q_dep)) return RS_ERR_ARG;
            DevCsr* sp = nullptr; RS_TRY(sumpool_table(c0, c1, !is_relu(), &sp, has_conv ? 0 : p0, has_conv ? -1 : p1));
        }
        if (!has_conv && !has_sumpool) { DevCsr* id = nullptr; RS_TRY(identity_table(c0, c1, !is_relu(), &id, p0, p1)); }
        if (is_relu()) { void* lut = nullptr; RS_TRY(relu_tables(c0, c1, &lut)); }
        if (has_maxpool && eq == E_ACTIVATION_SIGN) { PoolPlan* plan = nullptr; RS_TRY(maxpool_plan(c1 - c0, &plan)); }
        return RS_OK;
    }
    bool conv_fuses_bias() const { return !has_sumpool && !is_relu(); }

    // Layers without a conv stage (the input layers: bias / sum-pool + activation on the client's ciphertexts) have too few
    // channels to shard by channel; they shard by OUTPUT PIXEL instead: rank r computes pixels [r*P/world, (r+1)*P/world) for all
    // channels, and the all-gather of those row blocks is already in canonical (h,w,c) order.
    bool pixel_shardable(int world) const {
        return world > 1 && !has_conv && !has_maxpool && eq != E_ACTIVATION_NONE && (q_h * q_w) % world == 0;
    }
    int shard_mode(int world) const {
        if (world <= 1) return 0;
        if (pixel_shardable(world)) return 2;
        return (has_conv && q_dep % world == 0) ? 1 : 0;
    }

    // ---- the stages (each allocates its output; none frees its input)
    int run_conv(const uint32_t* in, size_t in_count, int c0, int c1, bool fuse_bias, Rows* out, size_t* out_count) {
        if (in_count != (size_t)cin_h * cin_w * cin_dep) return RS_ERR_ARG;
        const int cl = c1 - c0;
        ConvWeights* cw = nullptr;
        RS_TRY(conv_tables(c0, c1, fuse_bias, &cw));
        rs_conv_desc d{};
        d.in_h = cin_h; d.in_w = cin_w; d.in_dep = cin_dep; d.out_h = cout_h; d.out_w = cout_w; d.out_dep = cl;
        d.win_h = np.conv.window.h; d.win_w = np.conv.window.w; d.stride_h = np.conv.stride.h; d.stride_w = np.conv.stride.w;
        d.ofs_h = ofs_h; d.ofs_w = ofs_w; d.int_mode = (int_inputs && !twin_conv) ? 1 : 0; d.od_begin = 0; d.od_end = cl;
        *out_count = (size_t)cout_h * cout_w * cl;
        RS_TRY(out->alloc(*out_count));
        return rs_lwe_conv(ctx, out->p, in, (const int8_t*)cw->packed, (const uint32_t*)cw->bias, &d);
    }

    // sign -> OR tree.  `cur` holds the pre-activations (rows (q_h,q_w,cl)) and is reused as the gate pre-combination buffer.
    int run_maxpool_sign(uint32_t* cur, size_t cur_count, int cl, Rows* pooled_out, size_t* out_count) {
        PoolPlan* plan = nullptr;
        RS_TRY(maxpool_plan(cl, &plan));
        if (plan->in_count != cur_count) return RS_ERR_ARG;
        Rows tree(ctx), pooled(ctx);
        RS_TRY(tree.alloc(plan->buf_count));
        RS_TRY(pooled.alloc(plan->out_count));
        const size_t S = RS_LWE_STRIDE;
        // Default: ONE launch per tree level (sign, then each OR level) -- with every launch tens of waves long the partly filled last
        // wave costs ~1 %.  RS_POOL_BLOCKS=1 selects the block-pipelined form instead: blocks of output rows issued round-robin on two
        // lanes so that one block's tail wave overlaps the other's launch.  Measured at the sizes one rank sees when cifar/binarynet
        // runs on 1 / 2 / 8 GPUs it is never faster and 4 % slower on the largest 8-GPU shard (the blocks' own launches are only 1-3
        // waves long; scripts/pool_block_ab.py, profiles/r2_pool_block_ab.log), and it is the one place where kernels of this
        // library run concurrently, so it is off by default.
        int nb = 1;
        if (plan->blockable && mp_out_h > 1 && getenv("RS_POOL_BLOCKS") && !getenv("RS_NO_LANES")) {
            const size_t waves = cur_count / kWaveCts;
            nb = (int)std::min<size_t>((size_t)mp_out_h, std::min<size_t>(16, waves / 3));
            if (nb < 2) nb = 1;
        }
        if (nb == 1) {
            RS_TRY(rs_pbs_batch(ctx, tree.p, cur, cur_count, kEighth));      // sign bits at +-1/8 for the OR gates
            for (auto& st : plan->steps) {
                if (st->csr.rows > cur_count) return RS_ERR_STATE;
                RS_TRY(st->csr.apply(cur, tree.p));
                RS_TRY(rs_pbs_batch(ctx, tree.p + st->dst_offset * S, cur, st->csr.rows, st->mu));
            }
            RS_TRY(plan->final_gather.apply(pooled.p, tree.p));
        } else {
            int lanes = std::min(nb, kMaxLanes);
            if (const char* e = getenv("RS_LANES_MAX")) lanes = std::max(1, std::min(lanes, atoi(e)));     // diagnostics
            const bool serial = getenv("RS_LANE_SERIAL") != nullptr;                                       // diagnostics: never two blocks in flight
            RS_TRY(rs_lanes(ctx, lanes));
            // every lane's bootstrap scratch is sized before the fork: no device allocation while other lanes' launches are in flight
            const size_t per_block = ((size_t)(mp_out_h + nb - 1) / nb) * (size_t)np.pool.stride.h * q_w * cl + (size_t)q_w * cl * np.pool.stride.h;
            for (int l = lanes - 1; l >= 0; l--) {
                RS_TRY(rs_lane_select(ctx, l));
                RS_TRY(rs_reserve_scratch(ctx, std::min(per_block, cur_count)));
            }
            RS_TRY(rs_lane_fork(ctx));
            const size_t win_rows = (size_t)mp_out_w * cl;                       // pooled outputs per output row
            const size_t in_rows = (size_t)np.pool.stride.h * q_w * cl;          // neurons per output row of windows
            int rc = RS_OK;
            for (int b = 0; b < nb && rc == RS_OK; b++) {
                const int a0 = (int)((long long)b * mp_out_h / nb), a1 = (int)((long long)(b + 1) * mp_out_h / nb);
                const size_t o0 = a0 * win_rows, o1 = a1 * win_rows;
                const size_t i0 = a0 * in_rows, i1 = (b == nb - 1) ? cur_count : a1 * in_rows;   // the last block takes ragged rows
                rc = rs_lane_select(ctx, b % lanes);
                if (rc == RS_OK) rc = rs_pbs_batch(ctx, tree.p + i0 * S, cur + i0 * S, i1 - i0, kEighth);
                for (auto& st : plan->steps) {
                    if (rc != RS_OK) break;
                    const size_t r0 = st->first_row[o0], r1 = st->first_row[o1];
                    if (r1 - r0 > i1 - i0) { rc = RS_ERR_STATE; break; }
                    // the block's own (already consumed) pre-activation rows serve as its pre-combination scratch
                    rc = st->csr.apply_range(cur + i0 * S, tree.p, r0, r1);
                    if (rc == RS_OK) rc = rs_pbs_batch(ctx, tree.p + (st->dst_offset + r0) * S, cur + i0 * S, r1 - r0, st->mu);
                }
                if (rc == RS_OK) rc = plan->final_gather.apply_range(pooled.p + o0 * S, tree.p, o0, o1);
                if (serial && rc == RS_OK) { rs_lane_select(ctx, 0); rc = rs_lane_join(ctx); if (rc == RS_OK) rc = rs_lane_fork(ctx); }
            }
            rs_lane_select(ctx, 0);
            const int rj = rs_lane_join(ctx);       // always re-join: the buffers above go back to the pool after this point
            if (rc != RS_OK) return rc;
            RS_TRY(rj);
        }
        *out_count = plan->out_count;
        pooled_out->reset(pooled.release());
        return RS_OK;
    }

    // ---- forward for the channel slice [c0,c1) (and, for conv-less layers, the output-pixel range [p0,p1)); does not free `in`
    int forward(const Batch& in, int c0, int c1, Batch* out, int p0 = 0, int p1 = -1) {
        if (!prepared || !ctx) return RS_ERR_STATE;
        if (has_conv && (p0 != 0 || p1 >= 0)) return RS_ERR_ARG;
        const int cl = c1 - c0;
        Rows cur(ctx);                 // linear-part result, rows (h,w,cl)
        size_t cur_count = 0;

        if (has_conv) RS_TRY(run_conv(in.dev, in.count, c0, c1, conv_fuses_bias(), &cur, &cur_count));
        if (has_sumpool) {
            DevCsr* sp = nullptr;
            const uint32_t* src = cur.p;
            if (!has_conv) {   // pooling the raw input: only the full channel range makes sense (in_dep is 1 or 3)
                if (c0 != 0 || c1 != q_dep) return RS_ERR_ARG;
                if (in.count != (size_t)sp_in_h * sp_in_w * q_dep) return RS_ERR_ARG;
                src = in.dev;
            }
            RS_TRY(sumpool_table(c0, c1, !is_relu(), &sp, has_conv ? 0 : p0, has_conv ? -1 : p1));
            Rows pooled(ctx);
            RS_TRY(pooled.alloc(sp->rows));
            RS_TRY(sp->apply(pooled.p, src));
            cur.reset(pooled.release()); cur_count = sp->rows;
        }
        if (!has_conv && !has_sumpool) {
            if (in.count != (size_t)q_h * q_w * q_dep) return RS_ERR_ARG;
            DevCsr* id = nullptr;
            RS_TRY(identity_table(c0, c1, !is_relu(), &id, p0, p1));
            cur_count = id->rows;
            RS_TRY(cur.alloc(cur_count));
            RS_TRY(id->apply(cur.p, in.dev));
        }
        if (eq == E_ACTIVATION_NONE) { out->dev = cur.release(); out->count = cur_count; return RS_OK; }   // Quantize::add_bias
        if (is_relu()) {   // ONE test-vector bootstrap per neuron (rows are channel-fastest: row % cl = local channel)
            void* lut = nullptr;
            RS_TRY(relu_tables(c0, c1, &lut));
            RS_TRY(rs_pbs_lut_batch(ctx, cur.p, cur.p, cur_count, (const uint32_t*)lut, cl));
            RS_TRY(rs_lwe_add_const(ctx, cur.p, cur_count, relu_half()));
            out->dev = cur.release(); out->count = cur_count;
            return RS_OK;
        }
        // ---- sign activation: ONE batched bootstrap for every neuron of the (sliced) layer
        if (!has_maxpool) {
            RS_TRY(rs_pbs_batch(ctx, cur.p, cur.p, cur_count, kUnit));
            out->dev = cur.release(); out->count = cur_count;
            return RS_OK;
        }
        Rows pooled(ctx);
        size_t pooled_count = 0;
        RS_TRY(run_maxpool_sign(cur.p, cur_count, cl, &pooled, &pooled_count));
        out->dev = pooled.release(); out->count = pooled_count;
        return RS_OK;
    }

    size_t bootstraps(int cl) {
        if (is_relu()) return (size_t)q_h * q_w * cl;
        if (eq != E_ACTIVATION_SIGN) return 0;
        size_t n = (size_t)q_h * q_w * cl;
        if (has_maxpool) {
            const tPoolParams& pl = np.pool;      // OR gates of the tree: (window elements - 1) per pooled output
            for (int oph = 0; oph < mp_out_h; oph++)
                for (int opw = 0; opw < mp_out_w; opw++) {
                    const int eh = std::min<int>(pl.window.h, q_h - oph * pl.stride.h), ew = std::min<int>(pl.window.w, q_w - opw * pl.stride.w);
                    n += (size_t)std::max(1, eh * ew - 1) * cl;
                }
        }
        return n;
    }
};

// ---------------------------------------------------------------------------------------------------- Layer
Layer::Layer(rs_ctx* ctx, bool int_inputs, eConvType ec, uint16_t depth, ePoolType ep, eQuantType eq, tNetParams* np)
    : impl_(new LayerImpl(ctx)) {
    assert(np != nullptr && ec < NUM_CONVS && ep < NUM_POOLS && eq < NUM_ACTIVATIONS);
    impl_->int_inputs = int_inputs; impl_->ec = ec; impl_->depth = depth; impl_->ep = ep; impl_->eq = eq;
    impl_->np = *np;
    tNetParams& p = impl_->np;
    if (p.version < 1) {   // set_version back-compat (lib/BinLayer.cpp:252-261)
        p.conv.stride.h = p.conv.stride.w = 1;
        p.pool.stride.h = p.pool.stride.w = 0;
    }
    if (ec == E_FC || ec == E_FC_FINAL) { p.conv.window.h = p.conv.window.w = 1; p.conv.same_pad = true; p.conv.stride.h = p.conv.stride.w = 1; }
    if (p.pool.stride.h == 0) p.pool.stride.h = p.pool.window.h;   // MaxPooling ctor (lib/BinFunc.cpp:816-817)
    if (p.pool.stride.w == 0) p.pool.stride.w = p.pool.window.w;
    if (ep == E_MAXPOOL) assert(eq == E_ACTIVATION_SIGN);
}
Layer::~Layer() = default;

tDimensions* Layer::prep(FILE* fd, tDimensions* dim) {
    in_dim = *dim;
    if (impl_->prep(fd, dim) != RS_OK) { printf("Bad Weights File. Exiting...\r\n"); return nullptr; }
    out_dim = *dim;
    return dim;
}

Batch Layer::execute(Batch in) {
    Batch out;
    int rc = impl_->forward(in, 0, impl_->channels(), &out);
    if (impl_->ctx) rs_lwe_free(impl_->ctx, in.dev);   // callee frees its input, like every Func::execute of the reference
    if (rc != RS_OK) { out.dev = nullptr; out.count = 0; }
    return out;
}

Batch Layer::execute_shard(const Batch& in, ShardSpec shard, int* ch_begin, int* ch_end) {
    int c0 = 0, c1 = impl_->channels();
    rs_shard_range(impl_->channels(), impl_->has_conv ? 1 : 0, shard.rank, shard.world, &c0, &c1);
    Batch out;
    int p0 = 0, p1 = -1;
    if (impl_->pixel_shardable(shard.world)) {
        const int per = impl_->q_h * impl_->q_w / shard.world;
        p0 = shard.rank * per; p1 = p0 + per;
    }
    if (impl_->forward(in, c0, c1, &out, p0, p1) != RS_OK) { out.dev = nullptr; out.count = 0; }
    if (ch_begin) *ch_begin = c0;
    if (ch_end) *ch_end = c1;
    return out;
}

int Layer::shard_mode(int world) const { return impl_->shard_mode(world); }

Batch Layer::execute_sharded(Batch in, rs_comm* comm) {
    Batch out = forward_sharded(in, comm);
    if (impl_->ctx) rs_lwe_free(impl_->ctx, in.dev);
    return out;
}

Batch Layer::forward_sharded(const Batch& in, rs_comm* comm) {
    const int world = rs_comm_world(comm), rank = rs_comm_rank(comm);
    rs_ctx* ctx = impl_->ctx;
    ShardSpec sh; sh.rank = rank; sh.world = comm ? world : 1;
    int c0 = 0, c1 = 0;
    Batch part = execute_shard(in, sh, &c0, &c1);
    Batch fail;
    if (!comm || world <= 1) return part;
    if (!part.dev) return fail;
    const int mode = impl_->shard_mode(world);
    if (mode == 0) return part;                          // computed whole on every rank
    Rows mine(ctx), gathered(ctx);
    mine.reset(part.dev);
    if (gathered.alloc(part.count * world) != RS_OK) return fail;
    if (rs_allgather(comm, gathered.p, mine.p, part.count) != RS_OK) return fail;   // the exchange step between layers
    Batch out;
    out.count = part.count * world;
    if (mode == 2) { out.dev = gathered.release(); return out; }   // blocks of pixel rows: already canonical
    Rows full(ctx);
    if (full.alloc(out.count) != RS_OK) return fail;
    const int cl = c1 - c0;
    if (rs_lwe_interleave(ctx, full.p, gathered.p, part.count / cl, cl, world) != RS_OK) return fail;
    out.dev = full.release();
    return out;
}

size_t Layer::out_count() const { return impl_->final_h() * impl_->final_w() * (size_t)impl_->channels(); }
int Layer::out_channels() const { return impl_->channels(); }
size_t Layer::bootstraps() const { return impl_->bootstraps(impl_->channels()); }
int Layer::build_tables(ShardSpec shard) {
    int c0 = 0, c1 = impl_->channels();
    rs_shard_range(impl_->channels(), impl_->has_conv ? 1 : 0, shard.rank, shard.world, &c0, &c1);
    int p0 = 0, p1 = -1;
    if (impl_->pixel_shardable(shard.world)) {
        const int per = impl_->q_h * impl_->q_w / shard.world;
        p0 = shard.rank * per; p1 = p0 + per;
    }
    return impl_->build_tables(c0, c1, p0, p1);
}
void Layer::set_int_conv_twin(bool on) { impl_->twin_conv = on; }
bool Layer::is_relu() const { return impl_->is_relu(); }

// ---------------------------------------------------------------------------------------------------- Net
Layer* Net::add(bool int_layer, eConvType ec, uint16_t depth, ePoolType ep, eQuantType eq, tNetParams* np) {
    if (int_layer) layers_.emplace_back(new IntLayer(ctx_, ec, depth, ep, eq, np));
    else layers_.emplace_back(new BinLayer(ctx_, ec, depth, ep, eq, np));
    return layers_.back().get();
}
int Net::prep(FILE* weights, tDimensions* dim) {
    // a net with DoReFa-ReLU layers follows the plaintext twin's IntFunc conv convention throughout (its weights were
    // converted for it; the reference's encrypted branch of those nets is not functional)
    bool any_relu = false;
    for (auto& l : layers_) any_relu = any_relu || l->is_relu();
    if (any_relu) for (auto& l : layers_) l->set_int_conv_twin(true);
    for (auto& l : layers_) if (!l->prep(weights, dim)) return RS_ERR_ARG;
    return RS_OK;
}
Batch Net::run(Batch in) {
    for (auto& l : layers_) { in = l->execute(in); if (!in.dev) break; }
    return in;
}
Batch Net::run_sharded(Batch in, rs_comm* comm) {
    for (auto& l : layers_) { in = l->execute_sharded(in, comm); if (!in.dev) break; }
    return in;
}
size_t Net::bootstraps() const { size_t n = 0; for (auto& l : layers_) n += l->bootstraps(); return n; }

// ---------------------------------------------------------------------------------------------------- Func-level stages
namespace {
tNetParams stage_params() {
    tNetParams p{};
    p.conv.window = {1, 1}; p.conv.stride = {1, 1}; p.conv.same_pad = true; p.conv.tern_thresh = 0.05f;
    p.pool.window = {2, 2}; p.pool.stride = {2, 2}; p.pool.same_pad = false;
    p.bnorm = {false, 0.001f}; p.quant.shift_bits = 1; p.e_bias = E_NO_BIAS; p.version = 2;
    return p;
}
Batch consume(rs_ctx* ctx, Batch in, int rc, Rows& out, size_t count) {   // callee-frees-input + result hand-over
    if (ctx) rs_lwe_free(ctx, in.dev);
    Batch b;
    if (rc == RS_OK) { b.dev = out.release(); b.count = count; }
    return b;
}
}  // namespace

ConvStage::ConvStage(rs_ctx* ctx, bool int_inputs, uint32_t out_depth, const tConvParams& conv) : impl_(new LayerImpl(ctx)) {
    impl_->np = stage_params(); impl_->np.conv = conv;
    impl_->int_inputs = int_inputs; impl_->ec = E_CONV; impl_->depth = (uint16_t)out_depth;
}
ConvStage::~ConvStage() = default;
int ConvStage::out_depth() const { return impl_->cout_dep; }
tDimensions* ConvStage::prep(FILE* fd, tDimensions* dim) {
    if (impl_->prep_conv(fd, dim) != RS_OK) { printf("Bad Weights File. Exiting...\r\n"); return nullptr; }
    impl_->prepared = true;
    return dim;
}
Batch ConvStage::execute(Batch in, int c0, int c1) {
    Rows out(impl_->ctx);
    size_t n = 0;
    if (c1 < 0) c1 = impl_->cout_dep;
    int rc = (impl_->prepared && c0 >= 0 && c0 < c1 && c1 <= impl_->cout_dep) ? RS_OK : RS_ERR_STATE;
    if (rc == RS_OK) rc = impl_->run_conv(in.dev, in.count, c0, c1, false, &out, &n);
    return consume(impl_->ctx, in, rc, out, n);
}

SumPoolStage::SumPoolStage(rs_ctx* ctx, const tPoolParams& pool) : impl_(new LayerImpl(ctx)) {
    impl_->np = stage_params(); impl_->np.pool = pool; impl_->ep = E_SUMPOOL;
    if (impl_->np.pool.stride.h == 0) impl_->np.pool.stride.h = pool.window.h;
    if (impl_->np.pool.stride.w == 0) impl_->np.pool.stride.w = pool.window.w;
}
SumPoolStage::~SumPoolStage() = default;
tDimensions* SumPoolStage::prep(tDimensions* dim) {
    if (impl_->prep_sumpool(dim) != RS_OK) return nullptr;
    impl_->bias_torus.assign(impl_->q_dep, 0u);
    impl_->prepared = true;
    return dim;
}
Batch SumPoolStage::execute(Batch in) {
    Rows out(impl_->ctx);
    DevCsr* sp = nullptr;
    int rc = impl_->prepared ? RS_OK : RS_ERR_STATE;
    const size_t pixels = (size_t)impl_->sp_in_h * impl_->sp_in_w;
    const int cl = pixels ? (int)(in.count / pixels) : 0;           // a channel slice pools like the full layer
    if (rc == RS_OK && (cl < 1 || cl > impl_->q_dep || (size_t)cl * pixels != in.count)) rc = RS_ERR_ARG;
    if (rc == RS_OK) rc = impl_->sumpool_table(0, cl, false, &sp);
    if (rc == RS_OK) rc = out.alloc(sp->rows);
    if (rc == RS_OK) rc = sp->apply(out.p, in.dev);
    return consume(impl_->ctx, in, rc, out, sp ? sp->rows : 0);
}

QuantizeStage::QuantizeStage(rs_ctx* ctx, bool int_inputs, const tQParams& q) : impl_(new LayerImpl(ctx)) {
    impl_->np = stage_params(); impl_->np.quant = q; impl_->int_inputs = int_inputs;
    // The reference's Quantize only stores shift_bits; which activation runs is decided by the method the layer calls
    // (execute / add_bias / relu_shift, lib/GPU/BinLayer.cu:160-175).  shift_bits still shapes prep(): 0 none, 1 sign, 2..8 a
    // DoReFa ReLU (lib/IntFunc.cpp:812-840).  The generated drivers leave tNetParams::quant uninitialised for layers without a
    // ReLU (nets/cifar/binarynet/net.cu:78-93 never sets it), so any other value is treated as "no ReLU tables" rather than an error.
    const bool relu_ok = int_inputs && q.shift_bits >= 2 && q.shift_bits <= 8;
    impl_->eq = q.shift_bits == 1 ? E_ACTIVATION_SIGN : relu_ok ? E_ACTIVATION_RELU : E_ACTIVATION_NONE;
}
QuantizeStage::~QuantizeStage() = default;
const std::vector<int32_t>& QuantizeStage::bias() const { return impl_->bias_int; }
int QuantizeStage::channels() const { return impl_->q_dep; }
tDimensions* QuantizeStage::prep(FILE* fd, tDimensions* dim, bool read_slope) {
    if (impl_->prep_quant(fd, dim, read_slope && impl_->is_relu()) != RS_OK) { printf("Bad Weights File. Exiting...\r\n"); return nullptr; }
    if (impl_->ctx) {
        if (rs_dev_alloc(impl_->ctx, impl_->bias_torus.size() * 4, &impl_->bias_dev) != RS_OK) return nullptr;
        if (rs_dev_upload(impl_->ctx, impl_->bias_dev, impl_->bias_torus.data(), impl_->bias_torus.size() * 4) != RS_OK) return nullptr;
    }
    impl_->prepared = true;
    return dim;
}
Batch QuantizeStage::add_bias(Batch in, int c0) {
    Batch fail;
    const size_t pixels = (size_t)impl_->q_h * impl_->q_w;
    const int cl = pixels ? (int)(in.count / pixels) : 0;
    if (!impl_->prepared || !impl_->ctx || cl < 1 || c0 < 0 || c0 + cl > impl_->q_dep || (size_t)cl * pixels != in.count ||
        rs_lwe_add_bias(impl_->ctx, in.dev, in.count, (const uint32_t*)impl_->bias_dev + c0, cl) != RS_OK) {
        if (impl_->ctx) rs_lwe_free(impl_->ctx, in.dev);
        return fail;
    }
    return in;      // in place: the "fresh array" of the reference is the same rows
}
Batch QuantizeStage::pre_sign(Batch in, int c0) { return add_bias(in, c0); }
int QuantizeStage::sign_bootstrap(rs_ctx* ctx, Batch& pre, uint32_t mu) {
    return rs_pbs_batch(ctx, pre.dev, pre.dev, pre.count, mu);
}
Batch QuantizeStage::relu_shift(Batch in, int c0) {
    Batch fail;
    void* lut = nullptr;
    const size_t pixels = (size_t)impl_->q_h * impl_->q_w;
    const int cl = pixels ? (int)(in.count / pixels) : 0;
    int rc = (impl_->prepared && impl_->is_relu() && cl >= 1 && c0 >= 0 && c0 + cl <= impl_->q_dep && (size_t)cl * pixels == in.count) ? RS_OK : RS_ERR_STATE;
    if (rc == RS_OK) rc = impl_->relu_tables(c0, c0 + cl, &lut);
    if (rc == RS_OK) rc = rs_pbs_lut_batch(impl_->ctx, in.dev, in.dev, in.count, (const uint32_t*)lut, cl);
    if (rc == RS_OK) rc = rs_lwe_add_const(impl_->ctx, in.dev, in.count, impl_->relu_half());
    if (rc != RS_OK) { if (impl_->ctx) rs_lwe_free(impl_->ctx, in.dev); return fail; }
    return in;
}

MaxPoolStage::MaxPoolStage(rs_ctx* ctx, const tPoolParams& pool) : impl_(new LayerImpl(ctx)) {
    impl_->np = stage_params(); impl_->np.pool = pool; impl_->ep = E_MAXPOOL; impl_->eq = E_ACTIVATION_SIGN;
    if (impl_->np.pool.stride.h == 0) impl_->np.pool.stride.h = pool.window.h;   // MaxPooling ctor (lib/BinFunc.cpp:816-817)
    if (impl_->np.pool.stride.w == 0) impl_->np.pool.stride.w = pool.window.w;
}
MaxPoolStage::~MaxPoolStage() = default;
tDimensions* MaxPoolStage::prep(tDimensions* dim) {
    if (impl_->prep_maxpool(dim) != RS_OK) return nullptr;
    impl_->prepared = true;
    return dim;
}
Batch MaxPoolStage::execute_from_preact(Batch pre) {
    rs_ctx* ctx = impl_->ctx;
    Rows pooled(ctx);
    size_t n = 0;
    const size_t pixels = (size_t)impl_->q_h * impl_->q_w;
    const int cl = pixels ? (int)(pre.count / pixels) : 0;
    int rc = (impl_->prepared && ctx && cl >= 1 && cl <= impl_->q_dep && (size_t)cl * pixels == pre.count) ? RS_OK : RS_ERR_STATE;
    if (rc == RS_OK) rc = impl_->run_maxpool_sign(pre.dev, pre.count, cl, &pooled, &n);
    return consume(ctx, pre, rc, pooled, n);
}

Batch gather_channels(rs_ctx* ctx, rs_comm* comm, Batch part, int c_local) {
    Batch fail;
    const int world = rs_comm_world(comm);
    if (!comm || world <= 1) return part;
    Rows mine(ctx), gathered(ctx), full(ctx);
    mine.reset(part.dev);
    if (c_local < 1 || part.count % (size_t)c_local) return fail;
    if (gathered.alloc(part.count * world) != RS_OK) return fail;
    if (rs_allgather(comm, gathered.p, mine.p, part.count) != RS_OK) return fail;
    if (full.alloc(part.count * world) != RS_OK) return fail;
    if (rs_lwe_interleave(ctx, full.p, gathered.p, part.count / c_local, c_local, world) != RS_OK) return fail;
    Batch out;
    out.dev = full.release(); out.count = part.count * world;
    return out;
}

}  // namespace redsec

// ---------------------------------------------------------------------------------------------------- flat C view
// (used by the Python harness through ctypes; a C++ caller such as nets/*/net.cu uses the classes directly)
struct rs_net { redsec::Net net; rs_ctx* ctx; explicit rs_net(rs_ctx* c) : net(c), ctx(c) {} };

extern "C" {

// Neuron partition: contiguous blocks of output channels; a layer is sharded only when it has a
// conv/FC stage and its channel count divides evenly, otherwise every rank computes it whole (replicated).
int rs_shard_range(int channels, int has_conv, int rank, int world, int* ch_begin, int* ch_end) {
    if (!ch_begin || !ch_end || channels <= 0 || world <= 0 || rank < 0 || rank >= world) return RS_ERR_ARG;
    *ch_begin = 0; *ch_end = channels;
    if (world > 1 && has_conv && channels % world == 0) {
        const int per = channels / world;
        *ch_begin = rank * per; *ch_end = *ch_begin + per;
    }
    return RS_OK;
}

// ctx may be NULL: a "dry" net that can be prepared and asked for its layer shapes and shard plan (host logic only; used by
// the CPU tests), but not run
rs_net* rs_net_create(rs_ctx* ctx) { return new rs_net(ctx); }
void rs_net_destroy(rs_net* n) { delete n; }

int rs_net_add_layer(rs_net* n, int int_layer, int conv_type, int out_depth, int pool_type, int quant_type, const rs_layer_params* p) {
    if (!n || !p) return RS_ERR_ARG;
    tNetParams np{};
    np.conv.window = {(int16_t)p->conv_win_h, (int16_t)p->conv_win_w};
    np.conv.stride = {(int16_t)p->conv_stride_h, (int16_t)p->conv_stride_w};
    np.conv.same_pad = p->conv_same_pad != 0; np.conv.tern_thresh = 0.05f;
    np.pool.window = {(int16_t)p->pool_win_h, (int16_t)p->pool_win_w};
    np.pool.stride = {(int16_t)p->pool_stride_h, (int16_t)p->pool_stride_w};
    np.pool.same_pad = p->pool_same_pad != 0;
    np.bnorm = {false, 0.001f};
    np.quant.shift_bits = (uint8_t)p->shift_bits;
    np.e_bias = (eBiasType)p->e_bias;
    np.version = (uint16_t)p->version;
    n->net.add(int_layer != 0, (eConvType)conv_type, (uint16_t)out_depth, (ePoolType)pool_type, (eQuantType)quant_type, &np);
    return RS_OK;
}

int rs_net_prep(rs_net* n, const char* weights_path, int in_h, int in_w, int in_dep) {
    return rs_net_prep_ex(n, weights_path, in_h, in_w, in_dep, 9, 255 * 2, 255.0f);   // nets/mnist/sign1024x1/net.cpp:100-105
}

int rs_net_prep_ex(rs_net* n, const char* weights_path, int in_h, int in_w, int in_dep, int in_bits, int up_bound, float scale) {
    if (!n || !weights_path) return RS_ERR_ARG;
    FILE* fd = fopen(weights_path, "rb");
    if (!fd) return RS_ERR_ARG;
    tDimensions dim{};
    dim.hw = {(int16_t)in_h, (int16_t)in_w}; dim.in_dep = (uint32_t)in_dep;
    dim.in_bits = (uint8_t)in_bits; dim.out_bits = SINGLE_BIT; dim.filter_bits = SINGLE_BIT; dim.bias_bits = SINGLE_BIT;
    dim.up_bound = (uint32_t)up_bound; dim.scale = scale;
    int rc = n->net.prep(fd, &dim);
    // the file must be consumed exactly (format check of SURVEY.md 5.4)
    if (rc == RS_OK) { int c = fgetc(fd); if (c != EOF) rc = RS_ERR_ARG; }
    fclose(fd);
    return rc;
}

int rs_net_build_tables(rs_net* n, int rank, int world) {
    if (!n || world <= 0 || rank < 0 || rank >= world) return RS_ERR_ARG;
    redsec::ShardSpec sh; sh.rank = rank; sh.world = world;
    for (size_t i = 0; i < n->net.num_layers(); i++)
        if (int rc = n->net.layer(i)->build_tables(sh)) return rc;
    return RS_OK;
}

int rs_net_num_layers(const rs_net* n) { return n ? (int)const_cast<rs_net*>(n)->net.num_layers() : 0; }
int rs_net_layer_info(rs_net* n, int i, size_t* out_count, int* channels, size_t* bootstraps, int* out_h, int* out_w) {
    if (!n || i < 0 || (size_t)i >= n->net.num_layers()) return RS_ERR_ARG;
    redsec::Layer* l = n->net.layer(i);
    if (out_count) *out_count = l->out_count();
    if (channels) *channels = l->out_channels();
    if (bootstraps) *bootstraps = l->bootstraps();
    if (out_h) *out_h = l->out_dim.hw.h;
    if (out_w) *out_w = l->out_dim.hw.w;
    return RS_OK;
}

// how layer i is split over `world` ranks: mode 0 = whole on every rank, 1 = output-channel blocks (all-gather + interleave),
// 2 = output-pixel blocks (all-gather only); rows_per_rank = ciphertexts each rank contributes to the all-gather
int rs_net_shard_plan(rs_net* n, int i, int world, int* mode, size_t* rows_per_rank, int* c_local) {
    if (!n || i < 0 || (size_t)i >= n->net.num_layers() || world < 1) return RS_ERR_ARG;
    redsec::Layer* l = n->net.layer(i);
    const int m = l->shard_mode(world);
    if (mode) *mode = m;
    if (rows_per_rank) *rows_per_rank = m == 0 ? l->out_count() : l->out_count() / world;
    if (c_local) *c_local = m == 1 ? l->out_channels() / world : l->out_channels();
    return RS_OK;
}

// runs layer i on in_dev WITHOUT consuming it; caller frees *out_dev with rs_lwe_free
int rs_net_layer_forward(rs_net* n, int i, const uint32_t* in_dev, size_t in_count, int rank, int world, uint32_t** out_dev,
                         size_t* out_count, int* ch_begin, int* ch_end) {
    if (!n || !n->ctx || i < 0 || (size_t)i >= n->net.num_layers() || !in_dev || !out_dev || !out_count) return RS_ERR_ARG;
    redsec::Batch in; in.dev = const_cast<uint32_t*>(in_dev); in.count = in_count;
    redsec::ShardSpec sh; sh.rank = rank; sh.world = world;
    redsec::Batch out = n->net.layer(i)->execute_shard(in, sh, ch_begin, ch_end);
    if (!out.dev) return RS_ERR_STATE;
    *out_dev = out.dev; *out_count = out.count;
    return RS_OK;
}

// one layer, neuron-sharded over the communicator (comm == NULL: whole layer on this GPU); does NOT consume in_dev; every
// rank receives the full layer output
int rs_net_layer_forward_sharded(rs_net* n, int i, rs_comm* comm, const uint32_t* in_dev, size_t in_count, uint32_t** out_dev,
                                 size_t* out_count) {
    if (!n || !n->ctx || i < 0 || (size_t)i >= n->net.num_layers() || !in_dev || !out_dev || !out_count) return RS_ERR_ARG;
    if (comm && rs_comm_ctx(comm) != n->ctx) return RS_ERR_ARG;
    redsec::Batch in; in.dev = const_cast<uint32_t*>(in_dev); in.count = in_count;
    redsec::Batch out = n->net.layer(i)->forward_sharded(in, comm);
    if (!out.dev) return RS_ERR_STATE;
    *out_dev = out.dev; *out_count = out.count;
    return RS_OK;
}

// HeBNN::run for the whole network on the engine stream: comm == NULL runs on one GPU, otherwise every layer is neuron-sharded
// over the communicator's ranks with an NCCL all-gather between layers -- no host synchronisation anywhere.  Does NOT consume
// in_dev (it is copied once); the caller frees *out_dev with rs_lwe_free.  Every rank receives the full output.
int rs_net_run(rs_net* n, rs_comm* comm, const uint32_t* in_dev, size_t in_count, uint32_t** out_dev, size_t* out_count) {
    if (!n || !n->ctx || !in_dev || !out_dev || !out_count) return RS_ERR_ARG;
    if (comm && rs_comm_ctx(comm) != n->ctx) return RS_ERR_ARG;
    redsec::Batch in;
    in.count = in_count;
    if (int rc = rs_lwe_alloc(n->ctx, in_count, &in.dev)) return rc;
    if (int rc = rs_lwe_copy(n->ctx, in.dev, in_dev, in_count)) { rs_lwe_free(n->ctx, in.dev); return rc; }
    redsec::Batch out = comm ? n->net.run_sharded(in, comm) : n->net.run(in);
    if (!out.dev) return RS_ERR_STATE;
    *out_dev = out.dev; *out_count = out.count;
    return RS_OK;
}

}  // extern "C"
