// UNet handle: architecture walk, weight registry/repacking, workspace planning and the forward launcher.
//
// Mirrors the constructor logic of the reference UNets (model/sr3_modules/unet.py:162-233,
// model/ddpm_modules/unet.py:148-219) to derive (a) the state_dict keys and shapes it must be fed and (b) a
// flat list of kernel launches ("plan") for a given (B, H, W, precision).  The forward pass
// (unet.py:235-259) then is a loop over that list on the caller's stream: no allocation, no sync.
#include <string.h>
#include <map>
#include <string>
#include <vector>

#include "common.cuh"
#include "tc.cuh"

namespace ds {

// ------------------------------------------------------------------------------------------ weights
enum WKind { WK_CONV, WK_VEC, WK_LINEAR, WK_BUF };

struct WeightSpec {
    std::string name;
    int ndim;
    int64_t shape[4];
    WKind kind;
    size_t off;        // float offset of the packed fp32 copy in the arena
    size_t elems;      // packed element count
    size_t off_bf16;   // byte offset of the bf16 pack in the tensor-core arena (convs only)
    int npad;          // conv: padded Cout ; vec: padded length
    int tc_kc, tc_up;  // tensor-core pack variant (tc_kc == 0: layer has no bf16 pack)
    bool halo;         // additionally packed for the fused GroupNorm+Swish->conv kernel (tc_halo.cu)
    size_t off_halo;   // byte offset of that pack in the bf16 arena
    bool chain;        // additionally packed for the per-sample persistent chain kernel (tc_chain.cu)
    size_t off_chain;
    bool entry32;      // the net's first conv: a TF32 TMA-fed image with the input channels padded to 8 (K chunk 8) ...
    size_t off_entry32;   // ... at this byte offset of the tensor-core arena
    int tc_kc32;       // TF32 images (ds_unet_desc.tf32_weights): K chunk of the TMA-fed pack (0 = none) ...
    size_t off_tf32;   // ... its byte offset in the tensor-core arena ...
    bool halo32;       // ... and the fused kernel's TF32 pack
    size_t off_halo32;
    bool loaded;
};

struct ConvW {
    int w = -1, b = -1;
    int cin = 0, cout = 0, ks = 0, npad = 0;
};
struct GNW {
    int w = -1, b = -1, C = 0;
};
struct ResW {
    GNW gn1, gn2, agn;
    ConvW conv1, conv2, res, qkv, aout;
    bool has_res = false, attn = false;
    int cin = 0, cout = 0;
    int temb_off = -1;
    int film_w = -1, film_b = -1;
};
enum LKind { L_CONV, L_RES, L_DOWN, L_UP };
struct Layer {
    LKind kind;
    std::string name;
    ConvW conv;
    ResW res;
    int section;   // 0 downs, 1 mid, 2 ups
};

// ------------------------------------------------------------------------------------------ plan
constexpr int ENTRY_CPAD = 16;   // channels of the TF32 NHWC repack of the network input (64-byte rows: the persistent conv takes it)
enum OpKind { OP_TEMB, OP_CONV, OP_GN, OP_ATTN, OP_GN_STATS, OP_CH_SUMS, OP_PACK_IN };
constexpr int64_t EXT_XA = -1, EXT_XB = -2, EXT_OUT = -3, NONE = -100;

struct Op {
    OpKind kind;
    // sources: GroupNorm reads the fp32 copies, tensor-core convs the bf16 copies, fp32 convs fp32
    int64_t src_a = NONE, src_b = NONE;
    int ca = 0, cb = 0, src_nchw = 0, Hs = 0, Ws = 0, up = 0, stride = 1;
    int Ho = 0, Wo = 0;
    const ConvW* cw = nullptr;
    int temb_off = -1;
    int64_t residual = NONE;             // fp32 NHWC
    int64_t dst = NONE;                  // fp32 NHWC (or EXT_OUT), may be NONE in bf16 mode
    int64_t dst_b16 = NONE;              // bf16 NHWC copy (bf16 mode only)
    int out_nchw = 0;
    // gn
    const GNW* gw = nullptr;
    int swish = 0, HW = 0;
    // conv with the GroupNorm(+Swish) of its input fused into the operand staging (sources are the fp32 copies)
    int halo = 0;
    const GNW* fgn = nullptr;
    int fswish = 0;
    // tensor-core conv whose sources are the fp32 copies, read as TF32 operands (DS_PREC_TF32)
    int tf32 = 0;
    // conv executed inside a per-sample persistent chain launch (tc_chain.cu): consecutive chain ops form one launch
    int chain = 0, chain_src_b16 = 0;
    // chain conv2 with the block's 1x1 res_conv folded in: raw second operand (fp32 copies) and its weights
    int64_t xsrc_a = NONE, xsrc_b = NONE;
    int xca = 0, xcb = 0;
    const ConvW* xw = nullptr;
    // the network's first conv on the dedicated small-Cin kernel (conv_entry.cu), which also emits the statistics
    int entry = 0;
    // ... or, in the throughput regime, on the TMA-fed TF32 kernel reading the 8-channel NHWC repack of the input (OP_PACK_IN)
    int entry32 = 0;
    // bf16 mode GroupNorm statistics: per-channel fp64 (sum, sumsq) slots in the plan's statistics arena
    int64_t sums_out = NONE;             // slot the producer's epilogue accumulates into
    int64_t sums_a = NONE, sums_b = NONE;   // slots of the (two) sources a fused / apply-only GroupNorm reads
    // fork / join: `side` ops run on the handle's side stream concurrently with the main chain (the 1x1 res_conv of a
    // block is independent of conv1); the op flagged `join` waits for it
    int side = 0, join = 0;
    // attn
    int N = 0, C = 0;
};

struct Tap {
    int64_t off;
    int C, H, W, bf16;
};

struct Plan {
    int B, H, W, prec;
    std::vector<Op> ops;
    std::vector<TcConvPlan> tc;          // per op (bf16 mode): TMA descriptors + launch geometry
    std::vector<AttnTcPlan> attn;        // per op (bf16 mode, attention ops whose channel count is a multiple of 64)
    std::vector<ChainPlan> chains;       // per op: the launch of the chain that STARTS at this op (chain_len[i] ops)
    std::vector<int> chain_len;          // > 0 at the first op of a chain, 0 elsewhere
    int chain_time_len = -1;
    const void* tc_ws = nullptr;         // workspace base the descriptors were encoded for
    size_t bytes = 0;
    int64_t temb_buf = NONE, gn_scratch = NONE;
    size_t stats_base = 0, stats_bytes = 0;   // per-channel fp64 statistics slots, zeroed at the start of every forward
    size_t counters_off = 0;                  // offset (inside that region) of the fp32 GroupNorm's per-sample counters
    std::map<std::string, Tap> taps;
    int launches = 0;
};

class Arena {   // first-fit offset allocator used only while planning
  public:
    explicit Arena(bool reuse) : reuse_(reuse) {}
    int64_t alloc(size_t bytes) {
        bytes = align_up(bytes, 256);
        if (reuse_) {
            for (size_t i = 0; i < free_.size(); ++i) {
                if (free_[i].second >= bytes) {
                    int64_t off = free_[i].first;
                    if (free_[i].second == bytes) free_.erase(free_.begin() + i);
                    else { free_[i].first += bytes; free_[i].second -= bytes; }
                    size_[off] = bytes;
                    return off;
                }
            }
        }
        int64_t off = (int64_t)top_;
        top_ += bytes;
        size_[off] = bytes;
        return off;
    }
    void release(int64_t off) {
        if (off < 0 || !reuse_) return;
        size_t bytes = size_[off];
        // insert sorted + coalesce
        size_t i = 0;
        while (i < free_.size() && free_[i].first < off) ++i;
        free_.insert(free_.begin() + i, std::make_pair(off, bytes));
        if (i + 1 < free_.size() && free_[i].first + (int64_t)free_[i].second == free_[i + 1].first) {
            free_[i].second += free_[i + 1].second;
            free_.erase(free_.begin() + i + 1);
        }
        if (i > 0 && free_[i - 1].first + (int64_t)free_[i - 1].second == free_[i].first) {
            free_[i - 1].second += free_[i].second;
            free_.erase(free_.begin() + i);
        }
    }
    size_t top() const { return top_; }

  private:
    bool reuse_;
    size_t top_ = 0;
    std::vector<std::pair<int64_t, size_t>> free_;
    std::map<int64_t, size_t> size_;
};

struct Act {
    int64_t f32 = NONE, b16 = NONE;      // workspace offsets of the fp32 / bf16 copies
    int64_t sums = NONE;                 // offset of the tensor's per-channel statistics slot (bf16 mode)
    int C = 0, H = 0, W = 0;
    int* rc = nullptr;                   // shared refcount
};

}  // namespace ds

using namespace ds;

struct ds_unet {
    ds_unet_desc d;
    std::vector<WeightSpec> specs;
    std::map<std::string, int> by_name;
    std::vector<Layer> layers;
    GNW final_gn;
    ConvW final_conv;
    int temb_total = 0;
    int w_inv_freq = -1, w_l1 = -1, b_l1 = -1, w_l2 = -1, b_l2 = -1;
    size_t film_off = 0, film_bias_off = 0;     // stacked per-block projection [total, dim] / [total]
    size_t arena_floats = 0;
    size_t arena_bf16_bytes = 0;
    float* d_arena = nullptr;
    uint8_t* d_arena_bf16 = nullptr;
    int arena_device = -1;               // device the weight arenas (and streams / events) were created on: the handle is bound to it
    bool weights_ready = false;
    std::vector<Plan*> plans;
    bool keep_taps = false;
    cudaStream_t side_stream = nullptr;  // fork / join partner of the caller's stream
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_temb = nullptr;
    int skip_mask = 0;                   // timing experiments only (DIFFSPLIT_B200_SKIP): bit OpKind = do not launch

    const float* wp(int i) const { return d_arena + specs[i].off; }
};

namespace ds {

static int add_spec(ds_unet* n, const std::string& name, WKind kind, std::initializer_list<int64_t> shape, int npad = 0) {
    WeightSpec s;
    s.name = name;
    s.kind = kind;
    s.ndim = (int)shape.size();
    int i = 0;
    for (int d = 0; d < 4; ++d) s.shape[d] = 1;
    for (auto v : shape) s.shape[i++] = v;
    s.npad = npad;
    s.loaded = false;
    s.off_bf16 = 0;
    s.tc_kc = 0;
    s.tc_up = 0;
    s.halo = false;
    s.off_halo = 0;
    s.chain = false;
    s.off_chain = 0;
    s.tc_kc32 = 0;
    s.off_tf32 = 0;
    s.entry32 = false;
    s.off_entry32 = 0;
    s.halo32 = false;
    s.off_halo32 = 0;
    size_t elems = 1;
    if (kind == WK_CONV) elems = (size_t)s.shape[1] * s.shape[2] * s.shape[3] * npad;
    else if (kind == WK_VEC) elems = (size_t)(npad ? npad : s.shape[0]);
    else for (int d = 0; d < s.ndim; ++d) elems *= (size_t)s.shape[d];
    s.elems = elems;
    s.off = n->arena_floats;
    n->arena_floats += align_up(elems, 64);
    if (kind == WK_CONV) {
        s.off_bf16 = n->arena_bf16_bytes;
        n->arena_bf16_bytes += align_up(tc_packed_weight_bytes((int)s.shape[0], (int)s.shape[1], (int)s.shape[2]), 1024);
    }
    n->by_name[name] = (int)n->specs.size();
    n->specs.push_back(s);
    return (int)n->specs.size() - 1;
}

static int npad_of(int cout) {
    int p = (cout + 15) / 16 * 16;
    if (p > 32) p = (cout + 63) / 64 * 64;
    return p;
}

// ca/cb: channel split of the input as the kernels will see it (two-source concat), up: nearest-x2 folded in,
// tc = false: the layer never runs on the tensor-core path (entry conv reads fp32 NCHW)
static ConvW add_conv(ds_unet* n, const std::string& p, int cin, int cout, int ks, bool bias = true, int ca = -1, int cb = 0,
                      int up = 0, bool tc = true, bool halo = false) {
    ConvW c;
    c.cin = cin; c.cout = cout; c.ks = ks; c.npad = npad_of(cout);
    c.w = add_spec(n, p + ".weight", WK_CONV, {cout, cin, ks, ks}, c.npad);
    if (ca < 0) ca = cin;
    if (tc) {
        n->specs[c.w].tc_kc = tc_pick_kc(ca, cb);
        n->specs[c.w].tc_up = up;
    }
    if (halo && cin % 16 == 0 && cin <= 224) {
        n->specs[c.w].halo = true;
        n->specs[c.w].off_halo = n->arena_bf16_bytes;
        n->arena_bf16_bytes += align_up(halo_packed_weight_bytes(cout, cin, ks), 1024);
    }
    if (n->d.tf32_weights) {
        if (tc && tc_pick_kc(ca, cb, 1)) {
            n->specs[c.w].tc_kc32 = tc_pick_kc(ca, cb, 1);
            n->specs[c.w].off_tf32 = n->arena_bf16_bytes;
            n->arena_bf16_bytes += align_up(tc_packed_weight_bytes(cout, cin, ks, 1), 1024);
        }
        if (halo && cin % 16 == 0 && cin <= 224) {
            n->specs[c.w].halo32 = true;
            n->specs[c.w].off_halo32 = n->arena_bf16_bytes;
            n->arena_bf16_bytes += align_up(halo_packed_weight_bytes(cout, cin, ks, 1), 1024);
        }
    }
    if (tc && !up && cin % 16 == 0 && cin <= 256 && (cout + 15) / 16 * 16 <= 256) {
        n->specs[c.w].chain = true;
        n->specs[c.w].off_chain = n->arena_bf16_bytes;
        n->arena_bf16_bytes += align_up(chain_packed_weight_bytes(cout, cin, ks), 1024);
    }
    if (bias) c.b = add_spec(n, p + ".bias", WK_VEC, {cout}, c.npad);
    return c;
}

static GNW add_gn(ds_unet* n, const std::string& p, int C) {
    GNW g;
    g.C = C;
    g.w = add_spec(n, p + ".weight", WK_VEC, {C});
    g.b = add_spec(n, p + ".bias", WK_VEC, {C});
    return g;
}

static bool in_attn_res(const ds_unet_desc& d, int res) {
    for (int i = 0; i < d.n_attn_res; ++i) if (d.attn_res[i] == res) return true;
    return false;
}

static ResW add_res(ds_unet* n, const std::string& p, int cin, int cout, bool attn, int skip = 0) {
    const ds_unet_desc& d = n->d;
    ResW r;
    r.cin = cin; r.cout = cout; r.attn = attn; r.has_res = cin != cout;
    const std::string rb = p + ".res_block";
    // registration order follows the reference module's state_dict order
    if (d.with_time_emb) {
        // sr3: FeatureWiseAffine.noise_func[0]; ddpm: mlp = [Swish, Linear] -> index 1
        const std::string f = d.variant == DS_UNET_SR3 ? rb + ".noise_func.noise_func.0" : rb + ".mlp.1";
        r.film_w = add_spec(n, f + ".weight", WK_LINEAR, {cout, d.inner_channel});
        r.film_b = add_spec(n, f + ".bias", WK_VEC, {cout});
        r.temb_off = n->temb_total;
        n->temb_total += cout;
    }
    r.gn1 = add_gn(n, rb + ".block1.block.0", cin);
    r.conv1 = add_conv(n, rb + ".block1.block.3", cin, cout, 3, true, -1, 0, 0, true, true);
    r.gn2 = add_gn(n, rb + ".block2.block.0", cout);
    r.conv2 = add_conv(n, rb + ".block2.block.3", cout, cout, 3, true, -1, 0, 0, true, true);
    if (r.has_res) r.res = add_conv(n, rb + ".res_conv", cin, cout, 1, true, cin - skip, skip);
    if (attn) {
        r.agn = add_gn(n, p + ".attn.norm", cout);
        r.qkv = add_conv(n, p + ".attn.qkv", cout, 3 * cout, 1, false, -1, 0, 0, true, true);
        r.aout = add_conv(n, p + ".attn.out", cout, cout, 1);
    }
    return r;
}

static int build_arch(ds_unet* n) {
    const ds_unet_desc& d = n->d;
    DS_REQUIRE(d.variant == DS_UNET_SR3 || d.variant == DS_UNET_DDPM, "unet: unknown variant %d", d.variant);
    DS_REQUIRE(d.n_mults >= 1 && d.n_mults <= DS_MAX_LEVELS, "unet: n_mults %d outside [1,%d]", d.n_mults, DS_MAX_LEVELS);
    DS_REQUIRE(d.n_attn_res >= 0 && d.n_attn_res <= DS_MAX_LEVELS, "unet: n_attn_res %d", d.n_attn_res);
    DS_REQUIRE(d.in_channel > 0 && d.out_channel > 0 && d.inner_channel > 0 && d.res_blocks >= 0, "unet: bad channel counts");
    DS_REQUIRE(d.norm_groups > 0 && d.norm_groups <= 64, "unet: norm_groups %d outside [1,64]", d.norm_groups);
    DS_REQUIRE(d.inner_channel % 2 == 0 && d.inner_channel <= 256, "unet: inner_channel %d must be even and <= 256", d.inner_channel);
    const int inner = d.inner_channel;
    if (d.with_time_emb) {
        const std::string key = d.variant == DS_UNET_SR3 ? "noise_level_mlp" : "time_mlp";
        if (d.variant == DS_UNET_DDPM) n->w_inv_freq = add_spec(n, key + ".0.inv_freq", WK_BUF, {inner / 2});
        n->w_l1 = add_spec(n, key + ".1.weight", WK_LINEAR, {4 * inner, inner});
        n->b_l1 = add_spec(n, key + ".1.bias", WK_VEC, {4 * inner});
        n->w_l2 = add_spec(n, key + ".3.weight", WK_LINEAR, {inner, 4 * inner});
        n->b_l2 = add_spec(n, key + ".3.bias", WK_VEC, {inner});
    }
    std::vector<int> skip_ch;
    int ch = inner, res = d.image_size, idx = 0;
    {
        Layer L; L.kind = L_CONV; L.name = "downs.0"; L.section = 0;
        L.conv = add_conv(n, "downs.0", d.in_channel, inner, 3, true, -1, 0, 0, false);
        if (d.in_channel <= 8) {
            n->specs[L.conv.w].entry32 = true;
            n->specs[L.conv.w].off_entry32 = n->arena_bf16_bytes;
            n->arena_bf16_bytes += align_up(tc_packed_weight_bytes(inner, ENTRY_CPAD, 3, 1), 1024);
        }
        n->layers.push_back(L);
        skip_ch.push_back(inner);
        idx = 1;
    }
    for (int lvl = 0; lvl < d.n_mults; ++lvl) {
        const int cout = inner * d.channel_mults[lvl];
        DS_REQUIRE(d.channel_mults[lvl] > 0, "unet: channel multiplier %d", d.channel_mults[lvl]);
        for (int r = 0; r < d.res_blocks; ++r) {
            Layer L; L.kind = L_RES; L.name = "downs." + std::to_string(idx++); L.section = 0;
            L.res = add_res(n, L.name, ch, cout, in_attn_res(d, res));
            n->layers.push_back(L);
            ch = cout;
            skip_ch.push_back(ch);
        }
        if (lvl != d.n_mults - 1) {
            Layer L; L.kind = L_DOWN; L.name = "downs." + std::to_string(idx++); L.section = 0;
            L.conv = add_conv(n, L.name + ".conv", ch, ch, 3);
            n->layers.push_back(L);
            skip_ch.push_back(ch);
            res /= 2;
        }
    }
    for (int m = 0; m < 2; ++m) {
        Layer L; L.kind = L_RES; L.name = "mid." + std::to_string(m); L.section = 1;
        L.res = add_res(n, L.name, ch, ch, m == 0);
        n->layers.push_back(L);
    }
    idx = 0;
    for (int lvl = d.n_mults - 1; lvl >= 0; --lvl) {
        const int cout = inner * d.channel_mults[lvl];
        for (int r = 0; r < d.res_blocks + 1; ++r) {
            DS_REQUIRE(!skip_ch.empty(), "unet: skip stack underflow");
            Layer L; L.kind = L_RES; L.name = "ups." + std::to_string(idx++); L.section = 2;
            L.res = add_res(n, L.name, ch + skip_ch.back(), cout, in_attn_res(d, res), skip_ch.back());
            skip_ch.pop_back();
            n->layers.push_back(L);
            ch = cout;
        }
        if (lvl >= 1) {
            Layer L; L.kind = L_UP; L.name = "ups." + std::to_string(idx++); L.section = 2;
            L.conv = add_conv(n, L.name + ".conv", ch, ch, 3, true, -1, 0, 1);
            n->layers.push_back(L);
            res *= 2;
        }
    }
    n->final_gn = add_gn(n, "final_conv.block.0", ch);
    n->final_conv = add_conv(n, "final_conv.block.3", ch, d.out_channel, 3, true, -1, 0, 0, true, true);
    // every GroupNorm must divide evenly
    for (auto& s : n->specs)
        if (s.kind == WK_VEC && s.name.find(".block.0.weight") != std::string::npos)
            DS_REQUIRE(s.shape[0] % d.norm_groups == 0, "unet: %s has %lld channels, not divisible by %d groups", s.name.c_str(),
                       (long long)s.shape[0], d.norm_groups);
    // stacked conditioning projection
    n->film_off = n->arena_floats;
    n->arena_floats += align_up((size_t)n->temb_total * inner, 64);
    n->film_bias_off = n->arena_floats;
    n->arena_floats += align_up((size_t)n->temb_total, 64);
    return DS_OK;
}

// ------------------------------------------------------------------------------------------ planning
enum Fmt { F32 = 1, B16 = 2 };

#define DS_ASSERT_CHAIN(op) do { if (!((op).kind == OP_CONV && (op).chain)) { fprintf(stderr, "diffsplit_b200: internal error, res fold on a non-chain op\n"); abort(); } } while (0)

// environment switches read once per process
static bool env_set(const char* name) {
    return getenv(name) != nullptr;
}
static bool no_side_stream() { static const bool v = env_set("DIFFSPLIT_B200_NO_SIDE"); return v; }
static bool attn_cuda_cores() { static const bool v = env_set("DIFFSPLIT_B200_ATTN_CUDA_CORES"); return v; }

struct Planner {
    ds_unet* n;
    Plan* p;
    Arena arena;
    int B;
    bool tc;          // bf16 / tf32 mode: tensor-core convs, fp32 residual stream
    bool tf32 = false;   // tf32 mode: the convs read the fp32 tensors themselves as TF32 operands - no bf16 copies except around
                         // the attention kernel (bf16 q / k / v)
    bool tf32_strict = false;
    size_t stats_top = 0;
    Planner(ds_unet* n_, Plan* p_, bool reuse) : n(n_), p(p_), arena(reuse) {}

    // what a tensor is stored as.  fp32 mode: always fp32.  bf16 mode: block-level tensors keep both copies (GroupNorm
    // statistics / residual adds read fp32, TMA-fed convs read bf16), conv operands are bf16 only, hidden = fp32 only.
    Act make(int C, int H, int W, int fmt, bool want_sums = true) {
        Act a;
        a.C = C; a.H = H; a.W = W;
        if (!tc) fmt = F32;
        if (tf32 && (fmt & F32)) fmt = F32;
        if (tc && (fmt & F32) && want_sums) {       // a GroupNorm will read this tensor: its producer emits the statistics
            a.sums = (int64_t)stats_top;
            stats_top += align_up((size_t)TC_SUM_COPIES * B * C * 2 * sizeof(double), 256);
        }
        if (fmt & F32) a.f32 = arena.alloc((size_t)B * H * W * C * 4);
        if (fmt & B16) a.b16 = arena.alloc((size_t)B * H * W * C * 2);
        a.rc = new int(1);
        return a;
    }
    void retain(Act& a) { if (a.rc) ++*a.rc; }
    void release(Act& a) {
        if (!a.rc) return;
        if (--*a.rc == 0) { arena.release(a.f32); arena.release(a.b16); delete a.rc; }
        a.rc = nullptr;
    }
    void tap(const std::string& name, const Act& a) {
        p->taps[name] = a.f32 != NONE ? Tap{a.f32, a.C, a.H, a.W, 0} : Tap{a.b16, a.C, a.H, a.W, 1};
    }
    bool src_tf32(const Act& a) const { return tf32 && a.f32 != NONE; }
    int64_t conv_src(const Act& a) const { return (tc && !src_tf32(a)) ? a.b16 : a.f32; }

    void gn(const Act& a, const Act* b, const GNW& g, int swish, const Act& out) {
        Op o; o.kind = OP_GN;
        o.src_a = a.f32; o.ca = a.C;
        if (b) { o.src_b = b->f32; o.cb = b->C; }
        o.gw = &g; o.swish = swish; o.HW = a.H * a.W;
        o.dst = out.f32; o.dst_b16 = out.b16;
        o.sums_a = a.sums;
        if (b) o.sums_b = b->sums;
        p->ops.push_back(o);
    }
    // per-sample persistent chain (tc_chain.cu): every stride-1 conv of a level whose padded sample fits `chain_mtiles`
    // 128-row tiles
    int chain_mtiles = 1;
    bool chain_ok(const Act& a, const Act* b, const ConvW& w, int stride, int up) const {
        // tf32 mode: no bf16-operand chains unless DIFFSPLIT_B200_TF32_HYBRID is set (then the 8 x 8 level - 11 of the ~37
        // convolutions of a splitting net - and the wide concat layers keep bf16 operands: hagen 64^2 x 16 0.54 instead of 0.69 ms per
        // step, whole-network max-rel 4-6.5e-3 instead of 1-2e-3, PSNR delta of a T = 5 JointIndi chain 0.08 instead of 0.03 dB)
        if (!tc || (tf32 && tf32_strict) || chain_mtiles <= 0 || stride != 1 || up || !n->specs[w.w].chain) return false;
        if ((a.H + 2) * (a.W + 2) > 128 * chain_mtiles) return false;
        return chain_conv_supported(a.C, b ? b->C : 0, w.cout, w.ks, a.H, a.W);
    }
    void chain_conv(const Act& a, const Act* b, const GNW* g, int swish, const ConvW& w, int temb_off, const Act* residual,
                    const Act& out) {
        Op o; o.kind = OP_CONV;
        o.chain = 1;
        o.fgn = g; o.fswish = swish;
        if (g) { o.sums_a = a.sums; if (b) o.sums_b = b->sums; }
        o.chain_src_b16 = a.f32 == NONE ? 1 : 0;
        o.src_a = o.chain_src_b16 ? a.b16 : a.f32; o.ca = a.C; o.Hs = a.H; o.Ws = a.W;
        if (b) { o.src_b = o.chain_src_b16 ? b->b16 : b->f32; o.cb = b->C; }
        o.cw = &w; o.temb_off = temb_off; o.Ho = out.H; o.Wo = out.W;
        if (residual) o.residual = residual->f32;
        o.dst = out.f32; o.dst_b16 = out.b16;
        o.sums_out = out.sums;
        p->ops.push_back(o);
    }
    void conv(const Act& a, const Act* b, const ConvW& w, int stride, int up, int temb_off, const Act* residual, const Act& out) {
        if (chain_ok(a, b, w, stride, up) && ((a.f32 != NONE && (!b || b->f32 != NONE)) || (a.f32 == NONE && a.b16 != NONE && !b))) {
            chain_conv(a, b, nullptr, 0, w, temb_off, residual, out);
            return;
        }
        Op o; o.kind = OP_CONV;
        o.tf32 = src_tf32(a) ? 1 : 0;
        o.src_a = conv_src(a); o.ca = a.C; o.Hs = a.H; o.Ws = a.W;
        if (b) { o.src_b = conv_src(*b); o.cb = b->C; }
        o.cw = &w; o.stride = stride; o.up = up; o.temb_off = temb_off;
        o.Ho = out.H; o.Wo = out.W;
        if (residual) o.residual = residual->f32;
        o.dst = out.f32; o.dst_b16 = out.b16;
        o.sums_out = out.sums;
        p->ops.push_back(o);
    }

    // out = conv(swish?(GroupNorm(cat[a, b])))  -  one statistics pass + one fused tensor-core kernel when the shape
    // allows it, else GroupNorm-apply into a bf16 operand tensor followed by the TMA-fed conv
    void gn_conv(const Act& a, const Act* b, const GNW& g, int swish, const ConvW& w, int temb_off, const Act* residual,
                 const Act& out) {
        const int cb = b ? b->C : 0;
        if (chain_ok(a, b, w, 1, 0) && a.f32 != NONE && (!b || b->f32 != NONE) && a.sums != NONE && (!b || b->sums != NONE)) {
            chain_conv(a, b, &g, swish, w, temb_off, residual, out);
            return;
        }
        bool use32 = tf32 && n->specs[w.w].halo32 && halo_conv_preferred(a.C, cb, w.cout, w.ks, B, a.H, a.W, 1);
        // wide concat inputs whose TF32 operand image does not fit shared memory: the fused kernel with bf16 operands rather than
        // GroupNorm-apply + a deep-K TMA-fed TF32 conv (50 us for 256 -> 128 at 8 x 8 x 16)
        const bool use16 = tc && !use32 && (!tf32 || !tf32_strict) && n->specs[w.w].halo && halo_conv_preferred(a.C, cb, w.cout, w.ks, B, a.H, a.W, 0);
        if (use32 || use16) {
            Op o; o.kind = OP_CONV;
            o.tf32 = use32 ? 1 : 0;
            o.sums_a = a.sums;
            if (b) o.sums_b = b->sums;
            o.sums_out = out.sums;
            o.halo = 1; o.fgn = &g; o.fswish = swish;
            o.src_a = a.f32; o.ca = a.C; o.Hs = a.H; o.Ws = a.W;
            if (b) { o.src_b = b->f32; o.cb = cb; }
            o.cw = &w; o.temb_off = temb_off; o.Ho = out.H; o.Wo = out.W;
            if (residual) o.residual = residual->f32;
            o.dst = out.f32; o.dst_b16 = out.b16;
            p->ops.push_back(o);
            return;
        }
        Act act = make(a.C + cb, a.H, a.W, tf32 ? F32 : B16, false);
        gn(a, b, g, swish, act);
        conv(act, nullptr, w, 1, 0, temb_off, residual, out);
        release(act);
    }

    Act resblock(const Layer& L, Act x, const Act* skip) {
        const ResW& r = L.res;
        const int H = x.H, W = x.W;
        Act resid = x;
        Act rbuf;
        bool res_side = false;
        // per-sample chain: the 1x1 res_conv becomes an extra K range of conv2 (one op less in the chain)
        Act hprobe; hprobe.C = r.cout; hprobe.H = H; hprobe.W = W;
        const bool fold_res = r.has_res && getenv("DIFFSPLIT_B200_NO_RES_FOLD") == nullptr && x.f32 != NONE &&
                              (!skip || skip->f32 != NONE) && chain_ok(x, skip, r.res, 1, 0) && chain_ok(hprobe, nullptr, r.conv2, 1, 0) &&
                              chain_ok(x, skip, r.conv1, 1, 0);
        // throughput regime: conv2 on the persistent TMA-fed kernel takes the 1x1 res_conv as extra K chunks (no res_conv launch, no
        // fp32 round trip of its output through the residual operand)
        bool fold_tcs = false;
        if (r.has_res && !fold_res && tc && getenv("DIFFSPLIT_B200_NO_RES_FOLD") == nullptr) {
            const int t32 = tf32 ? 1 : 0;                       // conv2 reads the GroupNorm-apply output: fp32 (tf32 mode) or bf16
            const int kc2 = t32 ? n->specs[r.conv2.w].tc_kc32 : n->specs[r.conv2.w].tc_kc;
            const int kcr = t32 ? n->specs[r.res.w].tc_kc32 : n->specs[r.res.w].tc_kc;
            const bool srcs_ok = t32 ? (x.f32 != NONE && (!skip || skip->f32 != NONE)) : (x.b16 != NONE && (!skip || skip->b16 != NONE));
            const bool unfused = !(n->specs[r.conv2.w].halo32 && t32 && halo_conv_preferred(r.cout, 0, r.cout, 3, B, H, W, 1)) &&
                                 !(!t32 && n->specs[r.conv2.w].halo && halo_conv_preferred(r.cout, 0, r.cout, 3, B, H, W, 0)) &&
                                 !(t32 && !tf32_strict && n->specs[r.conv2.w].halo && halo_conv_preferred(r.cout, 0, r.cout, 3, B, H, W, 0));
            fold_tcs = kc2 > 0 && kc2 == kcr && srcs_ok && unfused && !chain_ok(hprobe, nullptr, r.conv2, 1, 0) &&
                       tc_conv_persistent(r.cout, 0, H, W, B, r.cout, 3, 1, 0, t32);
        }
        if (r.has_res && !fold_res && !fold_tcs) {   // res_conv(x) first: it runs on the side stream while conv1 runs on the main one
            rbuf = make(r.cout, H, W, F32, false);
            conv(x, skip, r.res, 1, 0, -1, nullptr, rbuf);
            p->ops.back().side = (tc && !p->ops.back().chain) ? 1 : 0;
            res_side = p->ops.back().side != 0;
            resid = rbuf;
        }
        Act h = make(r.cout, H, W, F32);
        gn_conv(x, skip, r.gn1, 1, r.conv1, r.temb_off, nullptr, h);
        Act out = make(r.cout, H, W, F32 | B16);
        gn_conv(h, nullptr, r.gn2, 1, r.conv2, -1, (fold_res || fold_tcs) ? nullptr : &resid, out);
        if (fold_tcs) {
            Op& o2 = p->ops.back();
            if (!(o2.kind == OP_CONV && !o2.chain && !o2.halo && o2.cw == &r.conv2)) { fprintf(stderr, "diffsplit_b200: internal error, res fold on an unexpected op\n"); abort(); }
            o2.xsrc_a = conv_src(x); o2.xca = x.C;
            if (skip) { o2.xsrc_b = conv_src(*skip); o2.xcb = skip->C; }
            o2.xw = &r.res;
        }
        if (fold_res) {
            Op& o2 = p->ops.back();
            DS_ASSERT_CHAIN(o2);
            o2.xsrc_a = x.f32; o2.xca = x.C;
            if (skip) { o2.xsrc_b = skip->f32; o2.xcb = skip->C; }
            o2.xw = &r.res;
        }
        if (r.has_res && tc && res_side) {
            for (size_t i = p->ops.size(); i-- > 0;)
                if (p->ops[i].kind == OP_CONV && p->ops[i].cw == &r.conv2) { p->ops[i].join = 1; break; }
        }
        release(h);
        if (r.has_res && !fold_res && !fold_tcs) release(rbuf);
        if (r.attn) {
            Act qkv = make(3 * r.cout, H, W, B16);
            gn_conv(out, nullptr, r.agn, 0, r.qkv, -1, nullptr, qkv);
            Act att = make(r.cout, H, W, B16);
            Op o; o.kind = OP_ATTN;
            o.src_a = conv_src(qkv); o.dst = att.f32; o.dst_b16 = att.b16; o.N = H * W; o.C = r.cout;
            p->ops.push_back(o);
            release(qkv);
            Act out2 = make(r.cout, H, W, F32 | B16);
            conv(att, nullptr, r.aout, 1, 0, -1, &out, out2);
            release(att);
            release(out);
            out = out2;
        }
        return out;
    }
};

static int build_plan(ds_unet* n, int B, int H, int W, int prec, Plan** out) {
    const ds_unet_desc& d = n->d;
    const int down = 1 << (d.n_mults - 1);
    DS_REQUIRE(B > 0 && H > 0 && W > 0, "unet: bad shape B=%d H=%d W=%d", B, H, W);
    DS_REQUIRE(B <= GN_MAX_BATCH, "unet: batch %d > %d", B, GN_MAX_BATCH);
    DS_REQUIRE(H % down == 0 && W % down == 0, "unet: H=%d W=%d must be divisible by %d", H, W, down);
    DS_REQUIRE(prec == DS_PREC_FP32 || prec == DS_PREC_BF16 || prec == DS_PREC_TF32, "unet: unknown precision %d", prec);
    DS_REQUIRE(prec != DS_PREC_TF32 || d.tf32_weights, "unet: DS_PREC_TF32 needs a handle created with ds_unet_desc.tf32_weights = 1");
    Plan* p = new Plan();
    p->B = B; p->H = H; p->W = W; p->prec = prec;
    Planner P(n, p, !n->keep_taps);
    P.B = B;
    P.tc = prec != DS_PREC_FP32;
    P.tf32 = prec == DS_PREC_TF32;
    P.tf32_strict = getenv("DIFFSPLIT_B200_TF32_HYBRID") == nullptr;
    {
        const char* e = getenv("DIFFSPLIT_B200_CHAIN_MTILES");      // 0 disables the per-sample persistent chains
        P.chain_mtiles = e ? atoi(e) : 1;
        if (P.chain_mtiles > CHAIN_MAX_MTILES) P.chain_mtiles = CHAIN_MAX_MTILES;
    }
    if (d.with_time_emb) p->temb_buf = P.arena.alloc((size_t)B * n->temb_total * sizeof(float));
    p->gn_scratch = P.arena.alloc(gn_scratch_bytes(B, d.norm_groups));

    std::vector<Act> skips;
    Act x;
    int h = H, w = W;
    for (const Layer& L : n->layers) {
        if (L.section == 0) {
            if (L.kind == L_CONV) {
                Act o = P.make(L.conv.cout, h, w, F32 | B16);
                Op op; op.kind = OP_CONV;
                op.src_a = EXT_XA; op.src_b = EXT_XB; op.src_nchw = 1; op.Hs = h; op.Ws = w;
                op.cw = &L.conv; op.Ho = h; op.Wo = w; op.dst = o.f32; op.dst_b16 = o.b16;
                // throughput regime (>= 4 tiles of 128 pixels per SM): repack to 8-channel TF32 NHWC + the TMA-fed tensor-core conv
                // (wide first layers only: at 16 output channels the small-Cin CUDA-core kernel is faster)
                const bool entry_tc = P.tc && n->specs[L.conv.w].entry32 && (int64_t)B * h * w >= (int64_t)4 * 148 * 128 &&
                                      L.conv.cout >= 64 && getenv("DIFFSPLIT_B200_NO_ENTRY_TC") == nullptr &&
                                      tc_conv_shape_supported(ENTRY_CPAD, 0, L.conv.ks, 1, 0, h, w, 1);
                if (entry_tc) {
                    Act in8 = P.make(ENTRY_CPAD, h, w, F32, false);
                    Op pk; pk.kind = OP_PACK_IN;
                    pk.src_a = EXT_XA; pk.src_b = EXT_XB; pk.dst = in8.f32; pk.Hs = h; pk.Ws = w; pk.HW = h * w;
                    p->ops.push_back(pk);
                    Op c8; c8.kind = OP_CONV;
                    c8.tf32 = 1; c8.entry32 = 1;
                    c8.src_a = in8.f32; c8.ca = ENTRY_CPAD; c8.Hs = h; c8.Ws = w;
                    c8.cw = &L.conv; c8.Ho = h; c8.Wo = w; c8.dst = o.f32; c8.dst_b16 = o.b16; c8.sums_out = o.sums;
                    p->ops.push_back(c8);
                    P.release(in8);
                    x = o;
                    P.retain(x);
                    Act s0 = x;
                    skips.push_back(s0);
                    continue;
                }
                const bool entry_kernel = P.tc && entry_conv_supported(d.in_channel, 0, L.conv.cout, L.conv.ks) &&
                                          getenv("DIFFSPLIT_B200_NO_ENTRY_KERNEL") == nullptr;
                if (entry_kernel) { op.entry = 1; op.sums_out = o.sums; }
                p->ops.push_back(op);
                if (P.tc && !entry_kernel) {   // the generic CUDA-core conv does not emit statistics: one pass over its output
                    Op cs; cs.kind = OP_CH_SUMS;
                    cs.src_a = o.f32; cs.ca = o.C; cs.HW = h * w; cs.sums_out = o.sums;
                    p->ops.push_back(cs);
                }
                x = o;
            } else if (L.kind == L_RES) {
                Act o = P.resblock(L, x, nullptr);
                P.release(x);
                x = o;
            } else {
                h /= 2; w /= 2;
                Act o = P.make(L.conv.cout, h, w, F32 | B16);
                P.conv(x, nullptr, L.conv, 2, 0, -1, nullptr, o);
                P.release(x);
                x = o;
            }
            P.retain(x);
            Act s = x;
            skips.push_back(s);
        } else if (L.section == 1) {
            Act o = P.resblock(L, x, nullptr);
            P.release(x);
            x = o;
        } else {
            if (L.kind == L_RES) {
                Act s = skips.back();
                skips.pop_back();
                DS_REQUIRE(s.H == x.H && s.W == x.W, "unet: skip shape mismatch at %s", L.name.c_str());
                Act o = P.resblock(L, x, &s);
                P.release(x);
                P.release(s);
                x = o;
            } else {
                Act o = P.make(L.conv.cout, 2 * x.H, 2 * x.W, F32 | B16);
                P.conv(x, nullptr, L.conv, 1, 1, -1, nullptr, o);
                P.release(x);
                x = o;
            }
        }
        P.tap(L.name, x);
    }
    {
        Act ext;                       // the network output: external fp32 NCHW
        ext.f32 = EXT_OUT; ext.C = n->final_conv.cout; ext.H = x.H; ext.W = x.W;
        const size_t first_new = p->ops.size();
        P.chain_mtiles = 0;                  // the final conv writes the caller's NCHW tensor: never part of a chain
        P.gn_conv(x, nullptr, n->final_gn, 1, n->final_conv, -1, nullptr, ext);
        for (size_t i = first_new; i < p->ops.size(); ++i)
            if (p->ops[i].kind == OP_CONV) p->ops[i].out_nchw = 1;
    }
    P.release(x);
    p->stats_base = align_up(P.arena.top(), 256);
    // "last CTA of the sample" counters of the fp32-mode GroupNorm statistics kernel: per plan (i.e. per workspace), inside
    // the region zeroed at the start of every forward - two UNets running concurrently (the two branches of JointIndi) must
    // not share them
    p->counters_off = P.stats_top;
    P.stats_top += align_up((size_t)B * sizeof(unsigned), 256);
    p->stats_bytes = P.stats_top;
    p->bytes = p->stats_base + p->stats_bytes;
    int launches = d.with_time_emb ? 1 : 0;
    int run = 0;                                                                 // ops of the current persistent chain
    for (size_t i = 0; i < p->ops.size(); ++i) {
        const Op& o = p->ops[i];
        if (o.kind == OP_CONV && o.chain) {
            const bool cont = run > 0 && run < CHAIN_MAX_OPS && p->ops[i - 1].Hs == o.Hs && p->ops[i - 1].Ws == o.Ws &&
                              chain_cluster_size(p->ops[i - 1].cw->cout) == chain_cluster_size(o.cw->cout) &&
                              chain_slice_rows(p->ops[i - 1].cw->cout) == chain_slice_rows(o.cw->cout);
            if (!cont) { ++launches; run = 0; }
            ++run;
            continue;
        }
        run = 0;
        launches += o.kind == OP_GN ? 2 : 1;                                      // GroupNorm = statistics (or their fold) + apply
    }
    if (p->stats_bytes) ++launches;                                              // the statistics-arena memset
    p->launches = launches;
    *out = p;
    return DS_OK;
}


static int get_plan(ds_unet* n, int B, int H, int W, int prec, Plan** out) {
    for (Plan* p : n->plans)
        if (p->B == B && p->H == H && p->W == W && p->prec == prec) { *out = p; return DS_OK; }
    Plan* p = nullptr;
    int rc = build_plan(n, B, H, W, prec, &p);
    if (rc != DS_OK) return rc;
    if (n->plans.size() >= 16) { delete n->plans.front(); n->plans.erase(n->plans.begin()); }
    n->plans.push_back(p);
    *out = p;
    return DS_OK;
}

}  // namespace ds

// ============================================================================================ C ABI
extern "C" int ds_unet_create(const ds_unet_desc* desc, ds_unet** out) {
    DS_REQUIRE(desc && out, "unet_create: null argument");
    ds_unet* n = new ds_unet();
    n->d = *desc;
    const char* e = getenv("DIFFSPLIT_B200_TAPS");
    n->keep_taps = e && e[0] == '1';
    const char* sk = getenv("DIFFSPLIT_B200_SKIP");
    if (sk) n->skip_mask = atoi(sk);
    int rc = build_arch(n);
    if (rc != DS_OK) { delete n; return rc; }
    *out = n;
    return DS_OK;
}

// the handle's device becomes current for the scope (weights, streams and events live there)
struct DeviceScope {
    int prev = -1;
    explicit DeviceScope(int dev) {
        if (dev >= 0 && cudaGetDevice(&prev) == cudaSuccess && prev != dev) cudaSetDevice(dev); else prev = -1;
    }
    ~DeviceScope() { if (prev >= 0) cudaSetDevice(prev); }
};

extern "C" void ds_unet_destroy(ds_unet* n) {
    if (!n) return;
    DeviceScope scope(n->arena_device);
    if (n->d_arena) cudaFree(n->d_arena);
    if (n->d_arena_bf16) cudaFree(n->d_arena_bf16);
    if (n->side_stream) cudaStreamDestroy(n->side_stream);
    if (n->ev_fork) cudaEventDestroy(n->ev_fork);
    if (n->ev_join) cudaEventDestroy(n->ev_join);
    if (n->ev_temb) cudaEventDestroy(n->ev_temb);
    for (Plan* p : n->plans) delete p;
    delete n;
}

extern "C" int ds_unet_num_weights(const ds_unet* n) { return n ? (int)n->specs.size() : 0; }

extern "C" const char* ds_unet_weight_name(const ds_unet* n, int i) {
    if (!n || i < 0 || i >= (int)n->specs.size()) return nullptr;
    return n->specs[i].name.c_str();
}

extern "C" int ds_unet_weight_shape(const ds_unet* n, int i, int32_t* ndim, int64_t shape[4]) {
    DS_REQUIRE(n && i >= 0 && i < (int)n->specs.size() && ndim && shape, "unet_weight_shape: bad index %d", i);
    *ndim = n->specs[i].ndim;
    for (int d = 0; d < 4; ++d) shape[d] = n->specs[i].shape[d];
    return DS_OK;
}

extern "C" int ds_unet_load_weights(ds_unet* n, const ds_tensor_view* ws, int cnt, void* stream) {
    DS_REQUIRE(n && ws && cnt > 0, "unet_load_weights: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    int cur_dev = -1;
    DS_CHECK_CUDA(cudaGetDevice(&cur_dev));
    DS_REQUIRE(n->arena_device < 0 || n->arena_device == cur_dev,
               "unet_load_weights: this handle's weights live on device %d, the current device is %d (create a new handle per device)",
               n->arena_device, cur_dev);
    if (!n->d_arena) {
        n->arena_device = cur_dev;
        DS_CHECK_CUDA(cudaMalloc(&n->d_arena, n->arena_floats * sizeof(float)));
        DS_CHECK_CUDA(cudaMemsetAsync(n->d_arena, 0, n->arena_floats * sizeof(float), st));
        DS_CHECK_CUDA(cudaMalloc(&n->d_arena_bf16, n->arena_bf16_bytes ? n->arena_bf16_bytes : 256));
        DS_CHECK_CUDA(cudaMemsetAsync(n->d_arena_bf16, 0, n->arena_bf16_bytes ? n->arena_bf16_bytes : 256, st));
    }
    for (int i = 0; i < cnt; ++i) {
        const ds_tensor_view& v = ws[i];
        DS_REQUIRE(v.name && v.d_data, "unet_load_weights: entry %d has a null name/pointer", i);
        auto it = n->by_name.find(v.name);
        if (it == n->by_name.end()) continue;      // foreign keys (e.g. sampler buffers) are ignored
        WeightSpec& s = n->specs[it->second];
        bool ok = v.ndim == s.ndim;
        for (int d = 0; ok && d < s.ndim; ++d) ok = v.shape[d] == s.shape[d];
        if (!ok) {
            set_error("unet_load_weights: %s has shape [%lld,%lld,%lld,%lld] (ndim %d), expected [%lld,%lld,%lld,%lld] (ndim %d)",
                      v.name, (long long)v.shape[0], (long long)v.shape[1], (long long)v.shape[2], (long long)v.shape[3], v.ndim,
                      (long long)s.shape[0], (long long)s.shape[1], (long long)s.shape[2], (long long)s.shape[3], s.ndim);
            return DS_ERR_WEIGHT;
        }
        float* dst = n->d_arena + s.off;
        const float* src = (const float*)v.d_data;
        if (s.kind == WK_CONV) {
            int rc = launch_pack_conv_weight_f32(src, dst, (int)s.shape[0], (int)s.shape[1], (int)s.shape[2], s.npad, st);
            if (rc != DS_OK) return rc;
            if (s.tc_kc) {
                rc = tc_pack_conv_weight(src, n->d_arena_bf16 + s.off_bf16, (int)s.shape[0], (int)s.shape[1], (int)s.shape[2],
                                         s.tc_up, s.tc_kc, 0, st);
                if (rc != DS_OK) return rc;
            }
            if (s.halo) {
                rc = halo_pack_conv_weight(src, n->d_arena_bf16 + s.off_halo, (int)s.shape[0], (int)s.shape[1], (int)s.shape[2], 0, st);
                if (rc != DS_OK) return rc;
            }
            if (s.entry32) {
                rc = tc_pack_conv_weight(src, n->d_arena_bf16 + s.off_entry32, (int)s.shape[0], ENTRY_CPAD, (int)s.shape[2], 0, ENTRY_CPAD, 1, st, (int)s.shape[1]);
                if (rc != DS_OK) return rc;
            }
            if (s.tc_kc32) {
                rc = tc_pack_conv_weight(src, n->d_arena_bf16 + s.off_tf32, (int)s.shape[0], (int)s.shape[1], (int)s.shape[2],
                                         s.tc_up, s.tc_kc32, 1, st);
                if (rc != DS_OK) return rc;
            }
            if (s.halo32) {
                rc = halo_pack_conv_weight(src, n->d_arena_bf16 + s.off_halo32, (int)s.shape[0], (int)s.shape[1], (int)s.shape[2], 1, st);
                if (rc != DS_OK) return rc;
            }
            if (s.chain) {
                rc = chain_pack_conv_weight(src, n->d_arena_bf16 + s.off_chain, (int)s.shape[0], (int)s.shape[1], (int)s.shape[2], st);
                if (rc != DS_OK) return rc;
            }
        } else {
            size_t cnt_el = 1;
            for (int d = 0; d < s.ndim; ++d) cnt_el *= (size_t)s.shape[d];
            DS_CHECK_CUDA(cudaMemcpyAsync(dst, src, cnt_el * sizeof(float), cudaMemcpyDeviceToDevice, st));
        }
        s.loaded = true;
    }
    for (auto& s : n->specs) {
        if (!s.loaded) {
            if (s.kind == WK_BUF) continue;       // inv_freq is recomputed below if it was not supplied
            set_error("unet_load_weights: missing weight %s", s.name.c_str());
            return DS_ERR_WEIGHT;
        }
    }
    if (n->w_inv_freq >= 0 && !n->specs[n->w_inv_freq].loaded) {
        const int half = n->d.inner_channel / 2;
        std::vector<float> f(half);
        for (int j = 0; j < half; ++j) f[j] = expf((float)(2 * j) * (float)(-log(10000.0) / (double)n->d.inner_channel));
        DS_CHECK_CUDA(cudaMemcpyAsync(n->d_arena + n->specs[n->w_inv_freq].off, f.data(), half * sizeof(float),
                                      cudaMemcpyHostToDevice, st));
        DS_CHECK_CUDA(cudaStreamSynchronize(st));
    }
    // stack the per-block conditioning projections
    if (n->d.with_time_emb) {
        const int dim = n->d.inner_channel;
        for (const Layer& L : n->layers) {
            if (L.kind != L_RES) continue;
            const ResW& r = L.res;
            DS_CHECK_CUDA(cudaMemcpyAsync(n->d_arena + n->film_off + (size_t)r.temb_off * dim, n->wp(r.film_w),
                                          (size_t)r.cout * dim * sizeof(float), cudaMemcpyDeviceToDevice, st));
            DS_CHECK_CUDA(cudaMemcpyAsync(n->d_arena + n->film_bias_off + r.temb_off, n->wp(r.film_b),
                                          (size_t)r.cout * sizeof(float), cudaMemcpyDeviceToDevice, st));
        }
    }
    if (!n->side_stream) {
        DS_CHECK_CUDA(cudaStreamCreateWithFlags(&n->side_stream, cudaStreamNonBlocking));
        DS_CHECK_CUDA(cudaEventCreateWithFlags(&n->ev_fork, cudaEventDisableTiming));
        DS_CHECK_CUDA(cudaEventCreateWithFlags(&n->ev_join, cudaEventDisableTiming));
        DS_CHECK_CUDA(cudaEventCreateWithFlags(&n->ev_temb, cudaEventDisableTiming));
    }
    n->weights_ready = true;
    return DS_OK;
}

extern "C" size_t ds_unet_workspace_bytes(ds_unet* n, int B, int H, int W, int precision) {
    if (!n) return 0;
    Plan* p = nullptr;
    if (get_plan(n, B, H, W, precision, &p) != DS_OK) return 0;
    return p->bytes;
}

extern "C" int ds_unet_launches(ds_unet* n, int B, int H, int W, int precision) {
    if (!n) return 0;
    Plan* p = nullptr;
    if (get_plan(n, B, H, W, precision, &p) != DS_OK) return 0;
    return p->launches;
}

extern "C" double ds_unet_flops(const ds_unet* n, int H, int W) {
    if (!n) return 0.0;
    double fl = 0.0;
    int h = H, w = W;
    auto res = [&](const ResW& r) {
        double f = 2.0 * h * w * 9.0 * ((double)r.cin * r.cout + (double)r.cout * r.cout);
        if (r.has_res) f += 2.0 * h * w * (double)r.cin * r.cout;
        if (r.attn) {
            const double N = (double)h * w, C = r.cout;
            f += 2.0 * N * C * 3.0 * C + 2.0 * N * C * C + 4.0 * N * N * C;
        }
        return f;
    };
    for (const Layer& L : n->layers) {
        if (L.kind == L_RES) fl += res(L.res);
        else {
            if (L.kind == L_DOWN) { h /= 2; w /= 2; }
            if (L.kind == L_UP) { h *= 2; w *= 2; }
            fl += 2.0 * h * w * 9.0 * (double)L.conv.cin * L.conv.cout;
        }
    }
    fl += 2.0 * h * w * 9.0 * (double)n->final_conv.cin * n->final_conv.cout;
    return fl;
}

namespace ds {
struct Prof {
    ds_op_profile* out;
    int max_ops, n;
    int reps;                 // each operator is launched `reps` times back to back between the two events (warm L2)
    cudaEvent_t e0, e1;
};
}  // namespace ds

static int run_forward(ds_unet* n, const float* d_xa, int ca, const float* d_xb, int cb, const float* d_time,
                       int time_len, float* d_out, int B, int H, int W, int precision, void* d_ws, size_t ws_bytes,
                       void* stream, Prof* prof) {
    DS_REQUIRE(n && d_xa && d_out && d_ws, "unet_forward: null argument");
    DS_REQUIRE(n->weights_ready, "unet_forward: weights not loaded");
    DS_REQUIRE(ca + cb == n->d.in_channel && ca > 0 && cb >= 0 && (cb == 0 || d_xb),
               "unet_forward: input channels %d+%d != in_channel %d", ca, cb, n->d.in_channel);
    DS_REQUIRE(!n->d.with_time_emb || (d_time && (time_len == 1 || time_len == B)),
               "unet_forward: time must have 1 or B=%d entries (got %d)", B, time_len);
    cudaStream_t st = (cudaStream_t)stream;
    {
        int cur_dev = -1;
        DS_CHECK_CUDA(cudaGetDevice(&cur_dev));
        DS_REQUIRE(cur_dev == n->arena_device, "unet_forward: the handle's weights live on device %d, the current device is %d", n->arena_device,
                   cur_dev);
    }
    Plan* p = nullptr;
    int rc = get_plan(n, B, H, W, precision, &p);
    if (rc != DS_OK) return rc;
    if (ws_bytes < p->bytes) {
        set_error("unet_forward: workspace %zu < required %zu bytes", ws_bytes, p->bytes);
        return DS_ERR_WORKSPACE;
    }
    DS_REQUIRE((reinterpret_cast<uintptr_t>(d_ws) & 255) == 0, "unet_forward: workspace must be 256-byte aligned");
    uint8_t* base = (uint8_t*)d_ws;
    auto ptr = [&](int64_t off) -> float* {
        if (off == EXT_XA) return const_cast<float*>(d_xa);
        if (off == EXT_XB) return const_cast<float*>(d_xb);
        if (off == EXT_OUT) return d_out;
        if (off == NONE) return nullptr;
        return reinterpret_cast<float*>(base + off);
    };
    float* temb = nullptr;
    bool temb_pending = false;               // the conditioning kernel is in flight on the side stream
    if (n->d.with_time_emb) {
        TembParams tp;
        tp.variant = n->d.variant;
        tp.dim = n->d.inner_channel;
        tp.inv_freq = n->w_inv_freq >= 0 ? n->wp(n->w_inv_freq) : nullptr;
        tp.w1 = n->wp(n->w_l1); tp.b1 = n->wp(n->b_l1);
        tp.w2 = n->wp(n->w_l2); tp.b2 = n->wp(n->b_l2);
        tp.wf = n->d_arena + n->film_off; tp.bf = n->d_arena + n->film_bias_off;
        tp.total = n->temb_total;
        temb = ptr(p->temb_buf);
        if (prof) {
            rc = launch_temb_f32(tp, d_time, time_len, temb, st);      // warm-up
            if (rc != DS_OK) return rc;
            cudaEventRecord(prof->e0, st);
        }
        // the conditioning vectors depend only on the noise level: their kernel runs on the forked stream beside the entry
        // conv; the first consumer waits for it
        const bool temb_side = !prof && n->side_stream && !no_side_stream();
        if (temb_side) {
            DS_CHECK_CUDA(cudaEventRecord(n->ev_fork, st));
            DS_CHECK_CUDA(cudaStreamWaitEvent(n->side_stream, n->ev_fork, 0));
        }
        for (int rep_i = 0; rep_i < (prof ? prof->reps : 1); ++rep_i) {
            rc = launch_temb_f32(tp, d_time, time_len, temb, temb_side ? n->side_stream : st);
            if (rc != DS_OK) return rc;
        }
        if (temb_side) {
            DS_CHECK_CUDA(cudaEventRecord(n->ev_temb, n->side_stream));
            temb_pending = true;
        }
        if (prof && prof->n < prof->max_ops) {
            cudaEventRecord(prof->e1, st);
            cudaEventSynchronize(prof->e1);
            ds_op_profile& r = prof->out[prof->n++];
            memset(&r, 0, sizeof(r));
            r.kind = 0;
            cudaEventElapsedTime(&r.ms, prof->e0, prof->e1);
            r.ms /= prof->reps;
            r.launches = 1;
        }
    }
    void* gn_scratch = ptr(p->gn_scratch);
    unsigned* counters = reinterpret_cast<unsigned*>(base + p->stats_base + p->counters_off);
    const bool tc = precision != DS_PREC_FP32;
    if (tc && p->tc_ws != d_ws) {
        // (re)encode the TMA descriptors for this workspace address
        p->tc.assign(p->ops.size(), TcConvPlan());
        p->attn.assign(p->ops.size(), AttnTcPlan());
        for (size_t i = 0; i < p->ops.size(); ++i) {
            const Op& o = p->ops[i];
            if (o.kind == OP_ATTN && attn_tc_supported(o.N, o.C)) {
                rc = attn_tc_build(&p->attn[i], ptr(o.src_a), ptr(o.dst_b16), B, o.N, o.C);
                if (rc != DS_OK) return rc;
            }
            if (o.kind != OP_CONV || o.src_nchw || o.halo || o.chain) continue;
            const int want_kc = o.entry32 ? ENTRY_CPAD : (o.tf32 ? n->specs[o.cw->w].tc_kc32 : n->specs[o.cw->w].tc_kc);
            if (!want_kc || !tc_conv_shape_supported(o.ca, o.cb, o.cw->ks, o.stride, o.up, o.Hs, o.Ws, o.tf32)) {
                set_error("unet_forward: the tensor-core modes need channel counts that are multiples of 16 (layer %s: %d+%d -> %d); use fp32",
                          n->specs[o.cw->w].name.c_str(), o.ca, o.cb, o.cw->cout);
                return DS_ERR_INVALID;
            }
            rc = tc_build_conv(&p->tc[i], ptr(o.src_a), o.ca, ptr(o.src_b), o.cb, o.Hs, o.Ws, B, o.cw->cout, o.cw->ks, o.stride, o.up, o.tf32,
                               o.xw ? ptr(o.xsrc_a) : nullptr, o.xw ? o.xca : 0, o.xw ? ptr(o.xsrc_b) : nullptr, o.xw ? o.xcb : 0);
            if (rc != DS_OK) return rc;
            if (p->tc[i].kc != want_kc) {
                set_error("unet_forward: internal error, K-chunk mismatch for %s", n->specs[o.cw->w].name.c_str());
                return DS_ERR_INVALID;
            }
        }
    }
    auto sums = [&](int64_t off) -> double* {
        return off == NONE ? nullptr : reinterpret_cast<double*>(base + p->stats_base + off);
    };
    if (tc && (p->tc_ws != d_ws || p->chain_time_len != time_len)) {
        // per-sample persistent chains: maximal runs of consecutive chain ops at one resolution -> one launch each
        p->chains.assign(p->ops.size(), ChainPlan());
        p->chain_len.assign(p->ops.size(), 0);
        size_t i = 0;
        while (i < p->ops.size()) {
            if (!(p->ops[i].kind == OP_CONV && p->ops[i].chain)) { ++i; continue; }
            size_t j = i;
            std::vector<ChainOpDesc> descs;
            while (j < p->ops.size() && p->ops[j].kind == OP_CONV && p->ops[j].chain && p->ops[j].Hs == p->ops[i].Hs &&
                   p->ops[j].Ws == p->ops[i].Ws && (int)descs.size() < CHAIN_MAX_OPS &&
                   chain_cluster_size(p->ops[j].cw->cout) == chain_cluster_size(p->ops[i].cw->cout) &&
                   chain_slice_rows(p->ops[j].cw->cout) == chain_slice_rows(p->ops[i].cw->cout)) {
                const Op& o = p->ops[j];
                ChainOpDesc d;
                memset(&d, 0, sizeof(d));
                d.src_a = ptr(o.src_a); d.src_b = ptr(o.src_b); d.ca = o.ca; d.cb = o.cb; d.src_b16 = o.chain_src_b16;
                d.norm = o.fgn ? 1 : 0; d.swish = o.fswish; d.G = n->d.norm_groups;
                d.sums_a = sums(o.sums_a); d.sums_b = sums(o.sums_b);
                d.gamma = o.fgn ? n->wp(o.fgn->w) : nullptr; d.beta = o.fgn ? n->wp(o.fgn->b) : nullptr;
                d.w = n->d_arena_bf16 + n->specs[o.cw->w].off_chain;
                d.cout = o.cw->cout; d.ks = o.cw->ks;
                if (o.xw) {
                    d.xsrc_a = ptr(o.xsrc_a); d.xsrc_b = ptr(o.xsrc_b); d.xca = o.xca; d.xcb = o.xcb;
                    d.wx = n->d_arena_bf16 + n->specs[o.xw->w].off_chain;
                    d.bias2 = o.xw->b >= 0 ? n->wp(o.xw->b) : nullptr;
                }
                d.epi.bias = o.cw->b >= 0 ? n->wp(o.cw->b) : nullptr;
                d.epi.temb = o.temb_off >= 0 ? temb : nullptr;
                d.epi.temb_off = o.temb_off >= 0 ? o.temb_off : 0;
                d.epi.temb_stride = n->temb_total;
                d.epi.temb_bcast = (time_len == 1);
                d.epi.residual = ptr(o.residual);
                d.out_f32 = ptr(o.dst); d.out_b16 = ptr(o.dst_b16);
                d.sums_out = sums(o.sums_out);
                descs.push_back(d);
                ++j;
            }
            rc = chain_build(&p->chains[i], descs.data(), (int)descs.size(), B, p->ops[i].Hs, p->ops[i].Ws);
            if (rc != DS_OK) return rc;
            p->chain_len[i] = (int)descs.size();
            i = j;
        }
        p->chain_time_len = time_len;
    }
    if (tc) p->tc_ws = d_ws;
    if (p->stats_bytes) DS_CHECK_CUDA(cudaMemsetAsync(base + p->stats_base, 0, p->stats_bytes, st));
    size_t op_index = 0;
    for (const Op& o : p->ops) {
        bool used_tc = false;
        const size_t oi = op_index++;
        if (n->skip_mask & (1 << (int)o.kind)) continue;
        const int nrep = prof ? prof->reps + 1 : 1;          // first repetition = warm-up, outside the events
        for (int rep_i = 0; rep_i < nrep; ++rep_i) {
        if (prof && rep_i == 1) cudaEventRecord(prof->e0, st);
        switch (o.kind) {
            case OP_GN:
                if (tc)      // the normalised conv operand: bf16, or fp32 when the consumer reads it as TF32
                    rc = launch_gn_apply_sums(ptr(o.src_a), o.ca, ptr(o.src_b), o.cb, sums(o.sums_a), sums(o.sums_b), n->wp(o.gw->w),
                                              n->wp(o.gw->b), o.dst_b16 != NONE ? (void*)ptr(o.dst_b16) : (void*)ptr(o.dst), B, o.HW,
                                              n->d.norm_groups, o.swish, o.dst_b16 != NONE ? 1 : 0, gn_scratch, st);
                else
                    rc = launch_groupnorm(ptr(o.src_a), o.ca, ptr(o.src_b), o.cb, n->wp(o.gw->w), n->wp(o.gw->b), ptr(o.dst), B, o.HW,
                                          n->d.norm_groups, o.swish, gn_scratch, counters, 0, st);
                break;
            case OP_PACK_IN:
                rc = tc_pack_input(d_xa, ca, d_xb, cb, ptr(o.dst), ENTRY_CPAD, B, o.Hs, o.Ws, st);
                break;
            case OP_CH_SUMS:
                rc = launch_ch_sums(ptr(o.src_a), o.ca, B, o.HW, sums(o.sums_out), st);
                break;
            case OP_GN_STATS:
                rc = launch_gn_stats(ptr(o.src_a), o.ca, ptr(o.src_b), o.cb, B, o.HW, n->d.norm_groups, gn_scratch, counters, st);
                break;
            case OP_CONV: {
                cudaStream_t st_main = st;
                if (temb_pending && (o.temb_off >= 0 || o.chain)) {
                    DS_CHECK_CUDA(cudaStreamWaitEvent(st_main, n->ev_temb, 0));
                    temb_pending = false;
                }
                const bool on_side = o.side && !prof && n->side_stream && !no_side_stream();
                if (on_side) {
                    DS_CHECK_CUDA(cudaEventRecord(n->ev_fork, st_main));
                    DS_CHECK_CUDA(cudaStreamWaitEvent(n->side_stream, n->ev_fork, 0));
                }
                if (o.join && !prof && n->side_stream && !no_side_stream())
                    DS_CHECK_CUDA(cudaStreamWaitEvent(st_main, n->ev_join, 0));
                cudaStream_t st = on_side ? n->side_stream : st_main;
                ConvSrc s;
                s.a = ptr(o.src_a); s.b = ptr(o.src_b);
                s.ca = o.src_nchw ? ca : o.ca; s.cb = o.src_nchw ? cb : o.cb;
                s.nchw = o.src_nchw; s.Hs = o.Hs; s.Ws = o.Ws; s.up = o.up;
                ConvEpi e;
                e.bias = o.cw->b >= 0 ? n->wp(o.cw->b) : nullptr;
                e.temb = o.temb_off >= 0 ? temb : nullptr;
                e.temb_off = o.temb_off >= 0 ? o.temb_off : 0;
                e.temb_stride = n->temb_total;
                e.temb_bcast = (time_len == 1);
                e.residual = ptr(o.residual);
                e.out_nchw = o.out_nchw;
                e.out2_bf16 = ptr(o.dst_b16);
                if (o.entry) {
                    rc = launch_conv_entry(d_xa, ca, d_xb, cb, n->wp(o.cw->w), o.cw->npad, e.bias, o.cw->cout, B, o.Ho, o.Wo, ptr(o.dst),
                                           ptr(o.dst_b16), sums(o.sums_out), st);
                } else if (o.chain) {
                    used_tc = true;
                    rc = p->chain_len[oi] > 0 ? chain_launch(&p->chains[oi], st) : DS_OK;   // later ops of a chain: nothing to do
                } else if (o.halo) {
                    used_tc = true;
                    HaloNorm nm;
                    nm.stats = nullptr; nm.sums_a = sums(o.sums_a); nm.sums_b = sums(o.sums_b);
                    nm.gamma = n->wp(o.fgn->w); nm.beta = n->wp(o.fgn->b); nm.G = n->d.norm_groups; nm.swish = o.fswish;
                    rc = halo_launch_conv(ptr(o.src_a), o.ca, ptr(o.src_b), o.cb, nm,
                                          n->d_arena_bf16 + (o.tf32 ? n->specs[o.cw->w].off_halo32 : n->specs[o.cw->w].off_halo),
                                          o.cw->cout, o.cw->ks, B, o.Hs, o.Ws, e, o.out_nchw ? nullptr : ptr(o.dst), ptr(o.dst_b16),
                                          o.out_nchw ? ptr(o.dst) : nullptr, sums(o.sums_out), o.tf32, st);
                } else if (tc && !o.src_nchw) {
                    used_tc = true;
                    rc = tc_launch_conv(&p->tc[oi],
                                        n->d_arena_bf16 + (o.entry32 ? n->specs[o.cw->w].off_entry32
                                                                     : (o.tf32 ? n->specs[o.cw->w].off_tf32 : n->specs[o.cw->w].off_bf16)), e,
                                        o.out_nchw ? nullptr : ptr(o.dst),
                                        ptr(o.dst_b16), o.out_nchw ? ptr(o.dst) : nullptr, sums(o.sums_out), st,
                                        o.xw ? n->d_arena_bf16 + (o.tf32 ? n->specs[o.xw->w].off_tf32 : n->specs[o.xw->w].off_bf16) : nullptr,
                                        (o.xw && o.xw->b >= 0) ? n->wp(o.xw->b) : nullptr);
                } else {
                    rc = launch_conv_f32(s, n->wp(o.cw->w), o.cw->npad, o.cw->cout, o.cw->ks, o.stride, B, o.Ho, o.Wo, e,
                                         ptr(o.dst), st);
                }
                if (on_side && rc == DS_OK) DS_CHECK_CUDA(cudaEventRecord(n->ev_join, n->side_stream));
                break;
            }
            case OP_ATTN:
                if (tc && attn_tc_supported(o.N, o.C) && !attn_cuda_cores())
                    rc = attn_tc_launch(&p->attn[oi], st);
                else
                    rc = launch_attention(ptr(o.src_a), tc ? (void*)ptr(o.dst_b16) : (void*)ptr(o.dst), B, o.N, o.C, tc ? 1 : 0, st);
                break;
            default:
                rc = DS_OK;
        }
        if (rc != DS_OK) return rc;
        }
        if (prof && prof->n < prof->max_ops) {
            cudaEventRecord(prof->e1, st);
            cudaEventSynchronize(prof->e1);
            ds_op_profile& r = prof->out[prof->n++];
            memset(&r, 0, sizeof(r));
            cudaEventElapsedTime(&r.ms, prof->e0, prof->e1);
            r.ms /= prof->reps;
            r.launches = 1;
            if (o.kind == OP_CONV) {
                const int cin = o.src_nchw ? ca + cb : o.ca + o.cb;
                r.kind = o.chain ? 7 : (o.halo ? 6 : (used_tc ? 4 : 1));
                if (o.chain && p->chain_len[oi] == 0) r.launches = 0;
                r.cin = cin; r.cout = o.cw->cout; r.ksize = o.cw->ks; r.h = o.Ho; r.w = o.Wo;
                r.flops = 2.0 * B * o.Ho * o.Wo * (double)o.cw->ks * o.cw->ks * cin * o.cw->cout;
                r.bytes = 4.0 * B * ((double)o.Hs * o.Ws * cin + (double)o.Ho * o.Wo * o.cw->cout *
                                     (o.residual != NONE ? 2.0 : 1.0)) + 4.0 * o.cw->ks * o.cw->ks * cin * o.cw->cout;
                if (o.xw) {   // folded 1x1 res_conv: its flops and its input
                    r.flops += 2.0 * B * o.Ho * o.Wo * (double)(o.xca + o.xcb) * o.cw->cout;
                    r.bytes += 4.0 * B * (double)o.Hs * o.Ws * (o.xca + o.xcb) + 4.0 * (o.xca + o.xcb) * o.cw->cout;
                }
            } else if (o.kind == OP_PACK_IN) {
                r.kind = 1;
                r.cin = ca + cb; r.cout = ENTRY_CPAD; r.h = o.Hs; r.w = o.Ws;
                r.bytes = 4.0 * B * (double)o.HW * (ca + cb + ENTRY_CPAD);
            } else if (o.kind == OP_GN_STATS || o.kind == OP_CH_SUMS) {
                r.kind = 5;
                r.cin = r.cout = o.ca + o.cb; r.h = o.HW; r.w = 1;
                r.bytes = 4.0 * B * (double)o.HW * (o.ca + o.cb);        // one read
            } else if (o.kind == OP_GN) {
                r.kind = 2;
                r.cin = r.cout = o.ca + o.cb; r.h = o.HW; r.w = 1;
                r.launches = 2;
                r.bytes = 4.0 * B * (double)o.HW * (o.ca + o.cb) * 3.0;   // standalone two-pass: 2 reads + 1 write
            } else if (o.kind == OP_ATTN) {
                r.kind = 3;
                r.cin = r.cout = o.C; r.h = o.N; r.w = 1;
                r.flops = 4.0 * B * (double)o.N * o.N * o.C;
                r.bytes = 4.0 * B * (double)o.N * o.C * 4.0;
            }
        }
    }
    if (temb_pending) DS_CHECK_CUDA(cudaStreamWaitEvent(st, n->ev_temb, 0));      // no consumer: still join the fork
    return DS_OK;
}

extern "C" int ds_unet_forward(ds_unet* n, const float* d_xa, int ca, const float* d_xb, int cb, const float* d_time,
                               int time_len, float* d_out, int B, int H, int W, int precision, void* d_ws, size_t ws_bytes,
                               void* stream) {
    return run_forward(n, d_xa, ca, d_xb, cb, d_time, time_len, d_out, B, H, W, precision, d_ws, ws_bytes, stream, nullptr);
}

extern "C" int ds_unet_forward_profiled(ds_unet* n, const float* d_xa, int ca, const float* d_xb, int cb,
                                        const float* d_time, int time_len, float* d_out, int B, int H, int W, int precision,
                                        void* d_ws, size_t ws_bytes, void* stream, ds_op_profile* ops, int max_ops,
                                        int* n_ops) {
    DS_REQUIRE(ops && n_ops && max_ops > 0, "unet_forward_profiled: null argument");
    Prof pr;
    pr.out = ops; pr.max_ops = max_ops; pr.n = 0;
    pr.reps = 20;
    DS_CHECK_CUDA(cudaEventCreate(&pr.e0));
    DS_CHECK_CUDA(cudaEventCreate(&pr.e1));
    int rc = run_forward(n, d_xa, ca, d_xb, cb, d_time, time_len, d_out, B, H, W, precision, d_ws, ws_bytes, stream, &pr);
    cudaEventDestroy(pr.e0);
    cudaEventDestroy(pr.e1);
    *n_ops = pr.n;
    return rc;
}

namespace ds {
__global__ void nhwc_to_nchw_kernel(const void* __restrict__ src, float* __restrict__ dst, int C, int HW, int64_t total, int bf16) {
    for (int64_t i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int64_t b = i / ((int64_t)C * HW);
        const int64_t r = i - b * C * HW;
        const int c = (int)(r / HW), p = (int)(r - (int64_t)c * HW);
        const int64_t j = (b * HW + p) * C + c;
        dst[i] = bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(src)[j]) : reinterpret_cast<const float*>(src)[j];
    }
}
}  // namespace ds

extern "C" int ds_unet_read_tap(ds_unet* n, const char* name, float* d_out, size_t out_elems, void* d_ws, void* stream) {
    DS_REQUIRE(n && name && d_out && d_ws, "unet_read_tap: null argument");
    DS_REQUIRE(n->keep_taps, "unet_read_tap: create the net with DIFFSPLIT_B200_TAPS=1 (buffers are recycled otherwise)");
    DS_REQUIRE(!n->plans.empty(), "unet_read_tap: no forward has been planned");
    Plan* p = n->plans.back();
    auto it = p->taps.find(name);
    DS_REQUIRE(it != p->taps.end(), "unet_read_tap: unknown tap %s", name);
    const Tap& t = it->second;
    const int64_t total = (int64_t)p->B * t.C * t.H * t.W;
    DS_REQUIRE((size_t)total <= out_elems, "unet_read_tap: output too small (%zu < %lld)", out_elems, (long long)total);
    const void* src = (uint8_t*)d_ws + t.off;
    int blocks = (int)((total + 255) / 256 > 4096 ? 4096 : (total + 255) / 256);
    nhwc_to_nchw_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, d_out, t.C, t.H * t.W, total, t.bf16);
    DS_CHECK_LAUNCH("nhwc_to_nchw");
    return DS_OK;
}
