// Miscellaneous C-ABI entry points: error text, device info, and the standalone operator wrappers that the
// parity tests call one by one.
#include <mutex>
#include <vector>

#include "common.cuh"
#include "tc.cuh"

namespace ds {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

}  // namespace ds

namespace ds {
bool pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("DIFFSPLIT_B200_PDL");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}
}  // namespace ds

namespace ds {
constexpr int TRACE_MAX = 8192;
static unsigned long long* g_trace = nullptr;
static int g_trace_n = 0;
static int g_trace_kind[TRACE_MAX];
TraceSlot trace_next(int kind) {
    static int on = -1;
    if (on < 0) on = getenv("DIFFSPLIT_B200_TRACE") ? 1 : 0;
    TraceSlot t = {nullptr, 0};
    if (!on) return t;
    if (!g_trace) {
        if (cudaMalloc(&g_trace, TRACE_MAX * 2 * sizeof(unsigned long long)) != cudaSuccess) return t;
        cudaMemset(g_trace, 0, TRACE_MAX * 2 * sizeof(unsigned long long));
    }
    if (g_trace_n >= TRACE_MAX) return t;
    g_trace_kind[g_trace_n] = kind;
    t.buf = g_trace;
    t.id = g_trace_n++;
    return t;
}
}  // namespace ds

// timeline debugging: forget all ids / (re)arm the slots recorded so far / read them back
extern "C" int ds_debug_trace_reset(int forget_ids) {
    using namespace ds;
    if (!g_trace) return DS_OK;
    if (forget_ids) g_trace_n = 0;
    std::vector<unsigned long long> init(TRACE_MAX * 2);
    for (int i = 0; i < TRACE_MAX; ++i) { init[2 * i] = ~0ull; init[2 * i + 1] = 0ull; }
    DS_CHECK_CUDA(cudaDeviceSynchronize());
    DS_CHECK_CUDA(cudaMemcpy(g_trace, init.data(), init.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice));
    return DS_OK;
}
extern "C" int ds_debug_trace_read(unsigned long long* h_start_end, int* h_kind, int max_n) {
    using namespace ds;
    if (!g_trace) return 0;
    cudaDeviceSynchronize();
    const int n = g_trace_n < max_n ? g_trace_n : max_n;
    cudaMemcpy(h_start_end, g_trace, (size_t)n * 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    for (int i = 0; i < n; ++i) h_kind[i] = g_trace_kind[i];
    return n;
}

using namespace ds;

extern "C" const char* ds_last_error(void) { return g_err; }

extern "C" int ds_version(void) { return 100; }

extern "C" int ds_device_info(int* sm_count, int* max_threads_per_sm, int* cc_major, int* cc_minor) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        set_error("no CUDA device: %s", cudaGetErrorString(e));
        return DS_ERR_NO_DEVICE;
    }
    cudaDeviceProp p;
    DS_CHECK_CUDA(cudaGetDeviceProperties(&p, dev));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (max_threads_per_sm) *max_threads_per_sm = p.maxThreadsPerMultiProcessor;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    return DS_OK;
}

namespace ds {
int gn_counters(unsigned** out) {
    static unsigned* bufs[64] = {nullptr};
    int dev = 0;
    DS_CHECK_CUDA(cudaGetDevice(&dev));
    DS_REQUIRE(dev >= 0 && dev < 64, "device index %d out of range", dev);
    if (!bufs[dev]) {
        DS_CHECK_CUDA(cudaMalloc(&bufs[dev], GN_MAX_BATCH * sizeof(unsigned)));
        DS_CHECK_CUDA(cudaMemset(bufs[dev], 0, GN_MAX_BATCH * sizeof(unsigned)));
    }
    *out = bufs[dev];
    return DS_OK;
}
}  // namespace ds

extern "C" size_t ds_groupnorm_scratch_bytes(int B, int groups) { return gn_scratch_bytes(B, groups); }

extern "C" int ds_groupnorm_swish_f32(const float* d_a, int ca, const float* d_b, int cb, const float* d_gamma,
                                      const float* d_beta, float* d_out, int B, int H, int W, int groups, int apply_swish,
                                      void* d_scratch, size_t scratch_bytes, void* stream) {
    DS_REQUIRE(d_a && d_gamma && d_beta && d_out && d_scratch, "groupnorm: null argument");
    DS_REQUIRE(ca > 0 && cb >= 0 && (cb == 0 || d_b), "groupnorm: bad channel split %d+%d", ca, cb);
    DS_REQUIRE(scratch_bytes >= gn_scratch_bytes(B, groups), "groupnorm: scratch too small");
    unsigned* counters = nullptr;
    int rc = gn_counters(&counters);
    if (rc != DS_OK) return rc;
    return launch_groupnorm(d_a, ca, d_b, cb, d_gamma, d_beta, d_out, B, H * W, groups, apply_swish, d_scratch, counters, 0,
                            (cudaStream_t)stream);
}

static int conv_npad(int cout) {
    int p = (cout + 15) / 16 * 16;
    if (p > 32) p = (cout + 63) / 64 * 64;
    return p;
}

extern "C" size_t ds_conv2d_scratch_bytes(int cin, int cout, int ksize) {
    return (size_t)ksize * ksize * cin * conv_npad(cout) * sizeof(float);
}

extern "C" int ds_conv2d_f32(const float* d_x, const float* d_w_oihw, const float* d_bias, float* d_out, int B, int H, int W,
                             int cin, int cout, int ksize, int stride, int upsample2x, void* d_scratch, size_t scratch_bytes,
                             void* stream) {
    DS_REQUIRE(d_x && d_w_oihw && d_out && d_scratch, "conv2d: null argument");
    DS_REQUIRE(scratch_bytes >= ds_conv2d_scratch_bytes(cin, cout, ksize), "conv2d: scratch too small");
    DS_REQUIRE(!(upsample2x && stride != 1), "conv2d: upsample with stride != 1");
    cudaStream_t st = (cudaStream_t)stream;
    const int npad = conv_npad(cout);
    int rc = launch_pack_conv_weight_f32(d_w_oihw, (float*)d_scratch, cout, cin, ksize, npad, st);
    if (rc != DS_OK) return rc;
    ConvSrc s;
    s.a = d_x; s.b = nullptr; s.ca = cin; s.cb = 0; s.nchw = 0; s.Hs = H; s.Ws = W; s.up = upsample2x;
    const int Hin = upsample2x ? 2 * H : H, Win = upsample2x ? 2 * W : W;
    const int pad = ksize / 2;
    const int Ho = (Hin + 2 * pad - ksize) / stride + 1, Wo = (Win + 2 * pad - ksize) / stride + 1;
    ConvEpi e;
    e.bias = d_bias; e.temb = nullptr; e.temb_off = 0; e.temb_stride = 0; e.temb_bcast = 0; e.residual = nullptr; e.out_nchw = 0;
    e.out2_bf16 = nullptr;
    return launch_conv_f32(s, (const float*)d_scratch, npad, cout, ksize, stride, B, Ho, Wo, e, d_out, st);
}

extern "C" int ds_attention_f32(const float* d_qkv, float* d_out, int B, int N, int C, void* stream) {
    DS_REQUIRE(d_qkv && d_out, "attention: null argument");
    return launch_attention(d_qkv, d_out, B, N, C, 0, (cudaStream_t)stream);
}

// Two chained fused ops in ONE per-sample persistent kernel (tc_chain.cu), shaped like a ResnetBlock:
//   h = conv1(swish(GN1(cat[xa, xb]))) + b1 ;  y = conv2(swish(GN2(h))) + b2 [+ residual]
// GN2 reads the statistics op 1's epilogue emitted inside the same launch.
static size_t chain2_layout(int B, int H, int W, int cin, int cmid, int cout, int ks, size_t* off_w2, size_t* off_sa, size_t* off_sb,
                            size_t* off_sh, size_t* off_h, int ca, int cb) {
    size_t o = 0;
    o += align_up(chain_packed_weight_bytes(cmid, cin, ks), 1024);
    *off_w2 = o; o += align_up(chain_packed_weight_bytes(cout, cmid, ks), 1024);
    *off_sa = o; o += align_up((size_t)TC_SUM_COPIES * B * ca * 16, 256);
    *off_sb = o; o += align_up((size_t)TC_SUM_COPIES * B * (cb > 0 ? cb : 1) * 16, 256);
    *off_sh = o; o += align_up((size_t)TC_SUM_COPIES * B * cmid * 16, 256);
    *off_h = o; o += align_up((size_t)B * H * W * cmid * 4, 256);
    return o;
}

extern "C" size_t ds_chain2_bf16_scratch_bytes(int B, int H, int W, int ca, int cb, int cmid, int cout, int ksize) {
    size_t a, b, c, d, e;
    return chain2_layout(B, H, W, ca + cb, cmid, cout, ksize, &a, &b, &c, &d, &e, ca, cb);
}

extern "C" int ds_chain2_bf16(const float* d_xa, int ca, const float* d_xb, int cb, const float* d_gamma1, const float* d_beta1,
                              const float* d_w1_oihw, const float* d_b1, const float* d_gamma2, const float* d_beta2,
                              const float* d_w2_oihw, const float* d_b2, const float* d_residual, int groups, float* d_out_f32,
                              void* d_out_b16, int B, int H, int W, int cmid, int cout, int ksize, void* d_scratch,
                              size_t scratch_bytes, void* stream) {
    DS_REQUIRE(d_xa && d_w1_oihw && d_w2_oihw && d_scratch && (d_out_f32 || d_out_b16) && groups > 0, "chain2_bf16: null argument");
    size_t off_w2, off_sa, off_sb, off_sh, off_h;
    const size_t need = chain2_layout(B, H, W, ca + cb, cmid, cout, ksize, &off_w2, &off_sa, &off_sb, &off_sh, &off_h, ca, cb);
    DS_REQUIRE(scratch_bytes >= need, "chain2_bf16: scratch too small");
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t* sc = (uint8_t*)d_scratch;
    int rc = chain_pack_conv_weight(d_w1_oihw, sc, cmid, ca + cb, ksize, st);
    if (rc != DS_OK) return rc;
    rc = chain_pack_conv_weight(d_w2_oihw, sc + off_w2, cout, cmid, ksize, st);
    if (rc != DS_OK) return rc;
    DS_CHECK_CUDA(cudaMemsetAsync(sc + off_sa, 0, off_h - off_sa, st));
    rc = launch_ch_sums(d_xa, ca, B, H * W, (double*)(sc + off_sa), st);
    if (rc != DS_OK) return rc;
    if (cb > 0) {
        rc = launch_ch_sums(d_xb, cb, B, H * W, (double*)(sc + off_sb), st);
        if (rc != DS_OK) return rc;
    }
    ChainOpDesc ops[2];
    memset(ops, 0, sizeof(ops));
    ops[0].src_a = d_xa; ops[0].src_b = cb > 0 ? d_xb : nullptr; ops[0].ca = ca; ops[0].cb = cb;
    ops[0].norm = 1; ops[0].swish = 1; ops[0].G = groups;
    ops[0].sums_a = (const double*)(sc + off_sa); ops[0].sums_b = cb > 0 ? (const double*)(sc + off_sb) : nullptr;
    ops[0].gamma = d_gamma1; ops[0].beta = d_beta1; ops[0].w = sc; ops[0].cout = cmid; ops[0].ks = ksize;
    ops[0].epi.bias = d_b1; ops[0].out_f32 = (float*)(sc + off_h); ops[0].sums_out = (double*)(sc + off_sh);
    ops[1].src_a = sc + off_h; ops[1].ca = cmid;
    ops[1].norm = 1; ops[1].swish = 1; ops[1].G = groups;
    ops[1].sums_a = (const double*)(sc + off_sh);
    ops[1].gamma = d_gamma2; ops[1].beta = d_beta2; ops[1].w = sc + off_w2; ops[1].cout = cout; ops[1].ks = ksize;
    ops[1].epi.bias = d_b2; ops[1].epi.residual = d_residual; ops[1].out_f32 = d_out_f32; ops[1].out_b16 = d_out_b16;
    ChainPlan plan;
    rc = chain_build(&plan, ops, 2, B, H, W);
    if (rc != DS_OK) return rc;
    return chain_launch(&plan, st);
}

extern "C" int ds_attention_bf16(const void* d_qkv, void* d_out, int B, int N, int C, void* stream) {
    DS_REQUIRE(d_qkv && d_out, "attention (bf16): null argument");
    AttnTcPlan plan;
    int rc = attn_tc_build(&plan, d_qkv, d_out, B, N, C);
    if (rc != DS_OK) return rc;
    return attn_tc_launch(&plan, (cudaStream_t)stream);
}

extern "C" size_t ds_conv2d_bf16_scratch_bytes(int cin, int cout, int ksize) {
    return align_up(tc_packed_weight_bytes(cout, cin, ksize), 1024);
}

extern "C" int ds_conv2d_bf16(const void* d_xa, int ca, const void* d_xb, int cb, const float* d_w_oihw, const float* d_bias,
                              const float* d_residual, void* d_out, float* d_out_f32, int out_f32_nchw, int B, int H, int W,
                              int cout, int ksize, int stride, int upsample2x, void* d_scratch, size_t scratch_bytes,
                              void* stream) {
    DS_REQUIRE(d_xa && d_w_oihw && d_out && d_scratch, "conv2d_bf16: null argument");
    DS_REQUIRE(ca > 0 && cb >= 0 && (cb == 0 || d_xb), "conv2d_bf16: bad channel split %d+%d", ca, cb);
    DS_REQUIRE(scratch_bytes >= ds_conv2d_bf16_scratch_bytes(ca + cb, cout, ksize), "conv2d_bf16: scratch too small");
    DS_REQUIRE(tc_conv_shape_supported(ca, cb, ksize, stride, upsample2x, H, W),
               "conv2d_bf16: unsupported shape (channels must be multiples of 16; stride 2 needs even H, W)");
    cudaStream_t st = (cudaStream_t)stream;
    const int kc = tc_pick_kc(ca, cb);
    int rc = tc_pack_conv_weight(d_w_oihw, (uint8_t*)d_scratch, cout, ca + cb, ksize, upsample2x, kc, 0, st);
    if (rc != DS_OK) return rc;
    TcConvPlan plan;
    rc = tc_build_conv(&plan, d_xa, ca, d_xb, cb, H, W, B, cout, ksize, stride, upsample2x);
    if (rc != DS_OK) return rc;
    ConvEpi e;
    e.bias = d_bias; e.temb = nullptr; e.temb_off = 0; e.temb_stride = 0; e.temb_bcast = 0;
    e.residual = (const float*)d_residual;
    e.out_nchw = out_f32_nchw; e.out2_bf16 = nullptr;
    return tc_launch_conv(&plan, (const uint8_t*)d_scratch, e, d_out_f32, out_f32_nchw ? nullptr : d_out,
                          out_f32_nchw ? (float*)d_out : nullptr, nullptr, st);
}

extern "C" size_t ds_gnconv_bf16_scratch_bytes(int B, int groups, int cin, int cout, int ksize) {
    return align_up(halo_packed_weight_bytes(cout, cin, ksize), 1024) + align_up(gn_scratch_bytes(B, groups), 256);
}

extern "C" int ds_gnconv_bf16(const float* d_xa, int ca, const float* d_xb, int cb, const float* d_gamma, const float* d_beta,
                              int groups, int apply_swish, const float* d_w_oihw, const float* d_bias, const float* d_residual,
                              void* d_out_b16, float* d_out_f32, int B, int H, int W, int cout, int ksize, void* d_scratch,
                              size_t scratch_bytes, void* stream) {
    DS_REQUIRE(d_xa && d_w_oihw && d_scratch && (d_out_b16 || d_out_f32), "gnconv_bf16: null argument");
    DS_REQUIRE(scratch_bytes >= ds_gnconv_bf16_scratch_bytes(B, groups > 0 ? groups : 1, ca + cb, cout, ksize),
               "gnconv_bf16: scratch too small");
    DS_REQUIRE(halo_conv_supported(ca, cb, cout, ksize, B, H, W),
               "gnconv_bf16: unsupported shape (channel counts multiples of 8, total a multiple of 16 and <= 224)");
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t* wp = (uint8_t*)d_scratch;
    void* gscratch = wp + align_up(halo_packed_weight_bytes(cout, ca + cb, ksize), 1024);
    int rc = halo_pack_conv_weight(d_w_oihw, wp, cout, ca + cb, ksize, 0, st);
    if (rc != DS_OK) return rc;
    const float2* stats = nullptr;
    if (groups > 0) {
        unsigned* counters = nullptr;
        rc = gn_counters(&counters);
        if (rc != DS_OK) return rc;
        rc = launch_gn_stats(d_xa, ca, d_xb, cb, B, H * W, groups, gscratch, counters, st);
        if (rc != DS_OK) return rc;
        stats = gn_stats_ptr(gscratch, B, groups);
    }
    ConvEpi e;
    e.bias = d_bias; e.temb = nullptr; e.temb_off = 0; e.temb_stride = 0; e.temb_bcast = 0; e.residual = d_residual;
    e.out_nchw = 0; e.out2_bf16 = nullptr;
    HaloNorm nm;
    nm.stats = stats; nm.sums_a = nullptr; nm.sums_b = nullptr; nm.gamma = d_gamma; nm.beta = d_beta; nm.G = groups;
    nm.swish = apply_swish;
    return halo_launch_conv(d_xa, ca, d_xb, cb, nm, wp, cout, ksize, B, H, W, e, d_out_f32, d_out_b16, nullptr, nullptr, 0, st);
}

// ---- the same two operators with fp32 sources read as TF32 operands (precision mode DS_PREC_TF32)
extern "C" size_t ds_conv2d_tf32_scratch_bytes(int cin, int cout, int ksize) {
    return align_up(tc_packed_weight_bytes(cout, cin, ksize, 1), 1024);
}

extern "C" int ds_conv2d_tf32(const float* d_xa, int ca, const float* d_xb, int cb, const float* d_w_oihw, const float* d_bias,
                              const float* d_residual, float* d_out_f32, int out_nchw, int B, int H, int W, int cout, int ksize,
                              int stride, int upsample2x, void* d_scratch, size_t scratch_bytes, void* stream) {
    DS_REQUIRE(d_xa && d_w_oihw && d_out_f32 && d_scratch, "conv2d_tf32: null argument");
    DS_REQUIRE(ca > 0 && cb >= 0 && (cb == 0 || d_xb), "conv2d_tf32: bad channel split %d+%d", ca, cb);
    DS_REQUIRE(scratch_bytes >= ds_conv2d_tf32_scratch_bytes(ca + cb, cout, ksize), "conv2d_tf32: scratch too small");
    DS_REQUIRE(tc_conv_shape_supported(ca, cb, ksize, stride, upsample2x, H, W, 1),
               "conv2d_tf32: unsupported shape (channels must be multiples of 8; stride 2 needs even H, W)");
    cudaStream_t st = (cudaStream_t)stream;
    const int kc = tc_pick_kc(ca, cb, 1);
    int rc = tc_pack_conv_weight(d_w_oihw, (uint8_t*)d_scratch, cout, ca + cb, ksize, upsample2x, kc, 1, st);
    if (rc != DS_OK) return rc;
    TcConvPlan plan;
    rc = tc_build_conv(&plan, d_xa, ca, d_xb, cb, H, W, B, cout, ksize, stride, upsample2x, 1);
    if (rc != DS_OK) return rc;
    ConvEpi e;
    e.bias = d_bias; e.temb = nullptr; e.temb_off = 0; e.temb_stride = 0; e.temb_bcast = 0;
    e.residual = d_residual;
    e.out_nchw = out_nchw; e.out2_bf16 = nullptr;
    return tc_launch_conv(&plan, (const uint8_t*)d_scratch, e, out_nchw ? nullptr : d_out_f32, nullptr, out_nchw ? d_out_f32 : nullptr,
                          nullptr, st);
}

extern "C" size_t ds_gnconv_tf32_scratch_bytes(int B, int groups, int cin, int cout, int ksize) {
    return align_up(halo_packed_weight_bytes(cout, cin, ksize, 1), 1024) + align_up(gn_scratch_bytes(B, groups), 256);
}

extern "C" int ds_gnconv_tf32(const float* d_xa, int ca, const float* d_xb, int cb, const float* d_gamma, const float* d_beta,
                              int groups, int apply_swish, const float* d_w_oihw, const float* d_bias, const float* d_residual,
                              float* d_out_f32, int B, int H, int W, int cout, int ksize, void* d_scratch, size_t scratch_bytes,
                              void* stream) {
    DS_REQUIRE(d_xa && d_w_oihw && d_scratch && d_out_f32, "gnconv_tf32: null argument");
    DS_REQUIRE(scratch_bytes >= ds_gnconv_tf32_scratch_bytes(B, groups > 0 ? groups : 1, ca + cb, cout, ksize),
               "gnconv_tf32: scratch too small");
    DS_REQUIRE(halo_conv_supported(ca, cb, cout, ksize, B, H, W, 1),
               "gnconv_tf32: unsupported shape (channel counts multiples of 8, total a multiple of 16 and <= 224)");
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t* wp = (uint8_t*)d_scratch;
    void* gscratch = wp + align_up(halo_packed_weight_bytes(cout, ca + cb, ksize, 1), 1024);
    int rc = halo_pack_conv_weight(d_w_oihw, wp, cout, ca + cb, ksize, 1, st);
    if (rc != DS_OK) return rc;
    const float2* stats = nullptr;
    if (groups > 0) {
        unsigned* counters = nullptr;
        rc = gn_counters(&counters);
        if (rc != DS_OK) return rc;
        rc = launch_gn_stats(d_xa, ca, d_xb, cb, B, H * W, groups, gscratch, counters, st);
        if (rc != DS_OK) return rc;
        stats = gn_stats_ptr(gscratch, B, groups);
    }
    ConvEpi e;
    e.bias = d_bias; e.temb = nullptr; e.temb_off = 0; e.temb_stride = 0; e.temb_bcast = 0; e.residual = d_residual;
    e.out_nchw = 0; e.out2_bf16 = nullptr;
    HaloNorm nm;
    nm.stats = stats; nm.sums_a = nullptr; nm.sums_b = nullptr; nm.gamma = d_gamma; nm.beta = d_beta; nm.G = groups;
    nm.swish = apply_swish;
    return halo_launch_conv(d_xa, ca, d_xb, cb, nm, wp, cout, ksize, B, H, W, e, d_out_f32, nullptr, nullptr, nullptr, 1, st);
}
