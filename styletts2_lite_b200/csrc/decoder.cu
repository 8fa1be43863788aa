// Decoder handle, load-time weight packing and the forward "program" (the sequence of kernel
// launches that replaces Decoder.forward / Generator.forward, Modules/hifigan.py:446-475,
// :321-347 and Modules/istftnet.py:692-721, :542-573), plus the C ABI of include/st2_b200.h.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <new>
#include <string>
#include <vector>

#include "common.cuh"

namespace st2 {

thread_local char g_err[1024] = "";
thread_local int64_t g_launch_count = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

struct RawTensor {
    const float* ptr;
    std::vector<int64_t> shape;
    int64_t numel() const {
        int64_t n = 1;
        for (auto s : shape) n *= s;
        return n;
    }
};

struct ConvW {
    float* w32 = nullptr;       // [k][Cin][Cout]
    void* w16[3] = {nullptr, nullptr, nullptr};   // index by OutDtype: [k][CoutPad][CinPad]
    float* bias = nullptr;
    int Cin = 0, Cout = 0, k = 0;
    int cin_pad = 0, cout_pad = 0;
    bool transposed = false;
};

struct AdaINRef { int h_off = 0; int C = 0; };

struct ResBlock1W {           // AdaINResBlock1 (hifigan.py:26-74)
    ConvW c1[3], c2[3];
    AdaINRef n1[3], n2[3];
    float* alpha1[3] = {nullptr, nullptr, nullptr};
    float* alpha2[3] = {nullptr, nullptr, nullptr};
    int k = 0, C = 0;
    int dil[3] = {1, 3, 5};
    std::string name;
};

struct ResBlk1dW {            // AdainResBlk1d (hifigan.py:359-403)
    ConvW conv1, conv2, conv1x1;
    bool has_sc = false, upsample = false;
    AdaINRef norm1, norm2;
    float* pool_w = nullptr;  // [3][ld_in]
    float* pool_b = nullptr;  // [ld_in]
    int Cin = 0, Cout = 0;
    std::string name;
};

struct LstmW {                // bidirectional nn.LSTM(I, H): input half as two 1x1 convs, W_hh^T / b_hh for the recurrence kernel
    ConvW ih[2];
    float* whh = nullptr;     // [2][H][4H]
    float* bhh = nullptr;     // [2][4H]
};

struct Tap { float* dst; int64_t cap; };

}  // namespace st2

using namespace st2;

struct st2_decoder {
    st2_config cfg;
    std::map<std::string, RawTensor> raw;
    bool finalized = false;
    std::vector<void*> allocs;
    int64_t num_params = 0;
    int64_t last_launches = 0;
    const uint64_t* seed_dev = nullptr;     // optional device-resident Philox seed (st2_decoder_set_seed_buffer)
    bool tc_ok = false;

    ResBlk1dW encode, decode[4];
    float *f0_w = nullptr, *f0_b = nullptr, *n_w = nullptr, *n_b = nullptr;
    ConvW asr_res;
    float *lin_w = nullptr, *lin_b = nullptr;
    float* gen_alpha[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    ConvW noise_convs[4], ups[4], conv_post;
    ResBlock1W noise_res[4], resblocks[12];
    float *fc_w = nullptr, *fc_b = nullptr;
    int fc_rows = 0;
    float *stft_fr = nullptr, *stft_fi = nullptr, *stft_br = nullptr, *stft_bi = nullptr;
    std::map<std::string, Tap> taps;

    // variant 2: the F0 / energy predictor ProsodyPredictor.F0Ntrain (models.py:407-419, :448-461); cfg.dim_in = d_hid
    LstmW shared;                           // models.py:407
    // duration half (row N2; packed only when the caller handed its weights over): DurationEncoder (models.py:468-483),
    // `lstm` (models.py:404), `duration_proj` (models.py:405)
    bool has_duration = false;
    int dur_layers = 0;
    LstmW enc_lstm[4];
    AdaINRef enc_norm[4];                   // AdaLayerNorm fc rows (gamma(C) | beta(C)) in the shared style matrix
    LstmW dur_lstm;
    float *dur_w = nullptr, *dur_b = nullptr;   // duration_proj.linear_layer [max_dur][d_hid], [max_dur]
    int max_dur = 0;

    // variant 3: TextEncoder (models.py:238-285); cfg.dim_in = channels
    int te_depth = 0, te_kernel = 0, te_symbols = 0;
    float* te_embedding = nullptr;          // [n_symbols][channels]
    ConvW te_conv[8];
    float* te_gamma[8] = {};                // gamma[c] followed by beta[c]
    LstmW te_lstm;
    ResBlk1dW pred_blk[2][3];               // F0.{0,1,2}, N.{0,1,2}
    ConvW pred_proj[2];                     // F0_proj, N_proj

    // per-launch event profile (one boundary event after every launch of a profiled forward)
    struct ProfRec { int cat; double flops; double bytes; };
    bool profiling = false;
    std::vector<cudaEvent_t> prof_events;
    std::vector<ProfRec> prof_recs;

    int spf() const {    // samples per asr frame
        int p = 2;
        for (int i = 0; i < cfg.n_stages; ++i) p *= cfg.upsample_rates[i];
        return p * (cfg.variant == 1 ? cfg.gen_istft_hop_size : 1);
    }
    int stage_channels(int i) const { return cfg.upsample_initial_channel >> (i + 1); }
};

namespace st2 {

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

// ------------------------------------------------------------------------------------------
// load time
// ------------------------------------------------------------------------------------------
struct Packer {
    st2_decoder* d;
    cudaStream_t st;
    int err = ST2_OK;
    std::vector<std::pair<std::string, int>> adain_list;   // (prefix, C) in h_off order

    void* dalloc(size_t bytes) {
        void* p = nullptr;
        if (cudaMalloc(&p, bytes ? bytes : 4) != cudaSuccess) {
            set_error("cudaMalloc(%zu) failed while packing weights", bytes);
            err = ST2_ERR_CUDA;
            return nullptr;
        }
        d->allocs.push_back(p);
        return p;
    }
    const RawTensor* get(const std::string& n, bool required = true) {
        auto it = d->raw.find(n);
        if (it == d->raw.end()) {
            if (required && err == ST2_OK) {
                set_error("missing weight '%s'", n.c_str());
                err = ST2_ERR_INVALID;
            }
            return nullptr;
        }
        return &it->second;
    }
    float* copy(const std::string& n, int64_t expect_numel) {
        const RawTensor* t = get(n);
        if (!t) return nullptr;
        if (t->numel() != expect_numel) {
            set_error("weight '%s' has %lld elements, expected %lld", n.c_str(), (long long)t->numel(),
                      (long long)expect_numel);
            err = ST2_ERR_INVALID;
            return nullptr;
        }
        float* p = (float*)dalloc(expect_numel * sizeof(float));
        if (p && cudaMemcpyAsync(p, t->ptr, expect_numel * sizeof(float), cudaMemcpyDeviceToDevice, st) != cudaSuccess)
            err = ST2_ERR_CUDA;
        return p;
    }
    void conv(ConvW& c, const std::string& n, int Cin, int Cout, int k, bool transposed, bool bias, bool want16) {
        c.Cin = Cin; c.Cout = Cout; c.k = k; c.transposed = transposed;
        const RawTensor* g = get(n + ".weight_g", false);
        const RawTensor* v = g ? get(n + ".weight_v") : get(n + ".weight");
        if (!v) return;
        const int d0 = transposed ? Cin : Cout, d1 = transposed ? Cout : Cin;
        if (v->shape.size() != 3 || v->shape[0] != d0 || v->shape[1] != d1 || v->shape[2] != k ||
            (g && g->numel() != d0)) {
            set_error("weight '%s' has the wrong shape (expected [%d,%d,%d])", n.c_str(), d0, d1, k);
            err = ST2_ERR_INVALID;
            return;
        }
        c.w32 = (float*)dalloc((size_t)k * Cin * Cout * sizeof(float));
        if (!c.w32) return;
        if (launch_fold_pack(g ? g->ptr : nullptr, v->ptr, c.w32, d0, d1, k, transposed ? 1 : 0, st) != ST2_OK)
            err = ST2_ERR_CUDA;
        if (bias) c.bias = copy(n + ".bias", Cout);
        if (want16 && d->tc_ok) {
            c.cin_pad = round_up(Cin, 64);
            c.cout_pad = round_up(Cout, 16);
            for (int dt = DT_BF16; dt <= DT_F16; ++dt) {
                c.w16[dt] = dalloc((size_t)k * c.cin_pad * c.cout_pad * 2);
                if (c.w16[dt] &&
                    launch_pack_w16(c.w32, c.w16[dt], k, Cin, Cout, c.cin_pad, c.cout_pad, dt, st) != ST2_OK)
                    err = ST2_ERR_CUDA;
            }
        }
    }
    // nn.Linear / nn.LSTM input matrix [Cout, Cin] as a 1x1 conv (packed [1][Cin][Cout] + 16-bit copies)
    void linear(ConvW& c, const std::string& wname, const std::string& bname, int Cin, int Cout) {
        c.Cin = Cin; c.Cout = Cout; c.k = 1; c.transposed = false;
        const RawTensor* v = get(wname);
        if (!v) return;
        if (v->numel() != (int64_t)Cin * Cout || v->shape.empty() || v->shape[0] != Cout) {
            set_error("weight '%s' has the wrong shape (expected [%d,%d])", wname.c_str(), Cout, Cin);
            err = ST2_ERR_INVALID;
            return;
        }
        c.w32 = (float*)dalloc((size_t)Cin * Cout * sizeof(float));
        if (!c.w32) return;
        if (launch_fold_pack(nullptr, v->ptr, c.w32, Cout, Cin, 1, 0, st) != ST2_OK) err = ST2_ERR_CUDA;
        if (!bname.empty()) c.bias = copy(bname, Cout);
        if (d->tc_ok && Cin % 64 == 0 && Cout % 16 == 0) {
            c.cin_pad = Cin; c.cout_pad = Cout;
            for (int dt = DT_BF16; dt <= DT_F16; ++dt) {
                c.w16[dt] = dalloc((size_t)Cin * Cout * 2);
                if (c.w16[dt] && launch_pack_w16(c.w32, c.w16[dt], 1, Cin, Cout, Cin, Cout, dt, st) != ST2_OK) err = ST2_ERR_CUDA;
            }
        }
    }
    void adain(AdaINRef& a, const std::string& prefix, int C) {
        a.C = C;
        a.h_off = d->fc_rows;
        d->fc_rows += 2 * C;
        adain_list.push_back({prefix, C});
    }
    void resblk1d(ResBlk1dW& b, const std::string& n, int Cin, int Cout, bool upsample) {
        b.name = n; b.Cin = Cin; b.Cout = Cout; b.upsample = upsample; b.has_sc = (Cin != Cout);
        conv(b.conv1, n + ".conv1", Cin, Cout, 3, false, true, true);
        conv(b.conv2, n + ".conv2", Cout, Cout, 3, false, true, true);
        if (b.has_sc) conv(b.conv1x1, n + ".conv1x1", Cin, Cout, 1, false, false, true);
        adain(b.norm1, n + ".norm1", Cin);
        adain(b.norm2, n + ".norm2", Cout);
        if (upsample) {
            // depthwise ConvTranspose1d weight [C,1,3] -> [3][ld] (zero padded), bias [ld]
            const int ld = round_up(Cin, 64);
            ConvW tmp;
            conv(tmp, n + ".pool", Cin, 1, 3, true, false, false);
            tmp.bias = copy(n + ".pool.bias", Cin);
            b.pool_w = (float*)dalloc((size_t)3 * ld * sizeof(float));
            b.pool_b = (float*)dalloc((size_t)ld * sizeof(float));
            if (err != ST2_OK || !b.pool_w || !b.pool_b) return;
            cudaMemsetAsync(b.pool_w, 0, (size_t)3 * ld * sizeof(float), st);
            cudaMemsetAsync(b.pool_b, 0, (size_t)ld * sizeof(float), st);
            cudaMemcpy2DAsync(b.pool_w, (size_t)ld * sizeof(float), tmp.w32, (size_t)Cin * sizeof(float),
                              (size_t)Cin * sizeof(float), 3, cudaMemcpyDeviceToDevice, st);
            cudaMemcpyAsync(b.pool_b, tmp.bias, (size_t)Cin * sizeof(float), cudaMemcpyDeviceToDevice, st);
        }
    }
    void resblock1(ResBlock1W& b, const std::string& n, int C, int k, const int* dil) {
        b.name = n; b.C = C; b.k = k;
        for (int j = 0; j < 3; ++j) {
            b.dil[j] = dil[j];
            const std::string js = std::to_string(j);
            conv(b.c1[j], n + ".convs1." + js, C, C, k, false, true, true);
            conv(b.c2[j], n + ".convs2." + js, C, C, k, false, true, true);
            adain(b.n1[j], n + ".adain1." + js, C);
            adain(b.n2[j], n + ".adain2." + js, C);
            b.alpha1[j] = copy(n + ".alpha1." + js, C);
            b.alpha2[j] = copy(n + ".alpha2." + js, C);
        }
    }
};

// one bidirectional nn.LSTM: W_ih as 1x1 convs (+ b_ih), W_hh transposed for the recurrence kernel
static void pack_lstm(st2_decoder* d, Packer& P, LstmW& w, const std::string& name, int I, int H) {
    w.whh = (float*)P.dalloc((size_t)2 * H * 4 * H * sizeof(float));
    w.bhh = (float*)P.dalloc((size_t)2 * 4 * H * sizeof(float));
    for (int dir = 0; dir < 2; ++dir) {
        const std::string sfx = dir ? "_reverse" : "";
        P.linear(w.ih[dir], name + ".weight_ih_l0" + sfx, name + ".bias_ih_l0" + sfx, I, 4 * H);
        const RawTensor* whh = P.get(name + ".weight_hh_l0" + sfx);
        const RawTensor* bhh = P.get(name + ".bias_hh_l0" + sfx);
        if (!whh || !bhh || !w.whh || !w.bhh) return;
        if (whh->numel() != (int64_t)4 * H * H || bhh->numel() != 4 * H) {
            set_error("%s.weight_hh_l0%s / bias_hh_l0%s have the wrong shape", name.c_str(), sfx.c_str(), sfx.c_str());
            P.err = ST2_ERR_INVALID;
            return;
        }
        // [4H][H] -> [H][4H]: the recurrence kernel reads gate columns contiguously
        if (launch_fold_pack(nullptr, whh->ptr, w.whh + (size_t)dir * H * 4 * H, 4 * H, H, 1, 0, P.st) != ST2_OK)
            P.err = ST2_ERR_CUDA;
        if (cudaMemcpyAsync(w.bhh + (size_t)dir * 4 * H, bhh->ptr, (size_t)4 * H * sizeof(float), cudaMemcpyDeviceToDevice,
                            P.st) != cudaSuccess)
            P.err = ST2_ERR_CUDA;
    }
}

// ProsodyPredictor weights: F0Ntrain (models.py:407-419) and, when present, the duration half (models.py:399-405)
static void pack_predictor(st2_decoder* d, Packer& P) {
    const int dh = d->cfg.dim_in, H = dh / 2, I = dh + d->cfg.style_dim;
    pack_lstm(d, P, d->shared, "shared", I, H);
    d->has_duration = d->raw.count("duration_proj.linear_layer.weight") != 0;
    if (d->has_duration) {
        d->dur_layers = 0;
        while (d->dur_layers < 4 && d->raw.count("text_encoder.lstms." + std::to_string(2 * d->dur_layers) + ".weight_ih_l0"))
            ++d->dur_layers;
        for (int i = 0; i < d->dur_layers; ++i) {
            pack_lstm(d, P, d->enc_lstm[i], "text_encoder.lstms." + std::to_string(2 * i), I, H);
            P.adain(d->enc_norm[i], "text_encoder.lstms." + std::to_string(2 * i + 1), dh);
        }
        pack_lstm(d, P, d->dur_lstm, "lstm", I, H);
        const RawTensor* w = P.get("duration_proj.linear_layer.weight");
        if (w && w->shape.size() == 2 && w->shape[1] == dh) {
            d->max_dur = (int)w->shape[0];
            d->dur_w = P.copy("duration_proj.linear_layer.weight", (int64_t)d->max_dur * dh);
            d->dur_b = P.copy("duration_proj.linear_layer.bias", d->max_dur);
        } else if (P.err == ST2_OK) {
            set_error("duration_proj.linear_layer.weight must be [max_dur, %d]", dh);
            P.err = ST2_ERR_INVALID;
        }
    }
    const char* br[2] = {"F0", "N"};
    for (int i = 0; i < 2; ++i) {
        P.resblk1d(d->pred_blk[i][0], std::string(br[i]) + ".0", dh, dh, false);
        P.resblk1d(d->pred_blk[i][1], std::string(br[i]) + ".1", dh, H, true);
        P.resblk1d(d->pred_blk[i][2], std::string(br[i]) + ".2", H, H, false);
        P.conv(d->pred_proj[i], std::string(br[i]) + "_proj", H, 1, 1, false, true, false);
    }
}

// TextEncoder weights (models.py:241-256)
static void pack_text_encoder(st2_decoder* d, Packer& P) {
    const int C = d->cfg.dim_in;
    d->te_embedding = P.copy("embedding.weight", (int64_t)d->te_symbols * C);
    for (int i = 0; i < d->te_depth; ++i) {
        const std::string n = "cnn." + std::to_string(i);
        P.conv(d->te_conv[i], n + ".0", C, C, d->te_kernel, false, true, true);
        d->te_gamma[i] = (float*)P.dalloc((size_t)2 * C * sizeof(float));
        const RawTensor* g = P.get(n + ".1.gamma");
        const RawTensor* b = P.get(n + ".1.beta");
        if (!g || !b || !d->te_gamma[i]) return;
        if (g->numel() != C || b->numel() != C) {
            set_error("%s.1.gamma / beta must have %d elements", n.c_str(), C);
            P.err = ST2_ERR_INVALID;
            return;
        }
        if (cudaMemcpyAsync(d->te_gamma[i], g->ptr, (size_t)C * sizeof(float), cudaMemcpyDeviceToDevice, P.st) != cudaSuccess ||
            cudaMemcpyAsync(d->te_gamma[i] + C, b->ptr, (size_t)C * sizeof(float), cudaMemcpyDeviceToDevice, P.st) != cudaSuccess)
            P.err = ST2_ERR_CUDA;
    }
    pack_lstm(d, P, d->te_lstm, "lstm", C, C / 2);
}

// Decoder weights (hifigan.py:416-443, istftnet.py:660-690)
static void pack_decoder(st2_decoder* d, Packer& P) {
    const st2_config& c = d->cfg;
    const int dim_in = c.dim_in;
    P.resblk1d(d->encode, "encode", dim_in + 2, 1024, false);
    for (int i = 0; i < 3; ++i) P.resblk1d(d->decode[i], "decode." + std::to_string(i), 1024 + 2 + 64, 1024, false);
    P.resblk1d(d->decode[3], "decode.3", 1024 + 2 + 64, c.upsample_initial_channel, true);
    {
        ConvW t;
        P.conv(t, "F0_conv", 1, 1, 3, false, true, false);
        d->f0_w = t.w32; d->f0_b = t.bias;
        ConvW u;
        P.conv(u, "N_conv", 1, 1, 3, false, true, false);
        d->n_w = u.w32; d->n_b = u.bias;
    }
    P.conv(d->asr_res, "asr_res.0", dim_in, 64, 1, false, true, true);
    d->lin_w = P.copy("generator.m_source.l_linear.weight", 9);
    d->lin_b = P.copy("generator.m_source.l_linear.bias", 1);
    const bool istft = c.variant == 1;
    const int c0 = c.upsample_initial_channel;
    if (!istft) d->gen_alpha[0] = P.copy("generator.alphas.0", c0);
    const int dil135[3] = {1, 3, 5};
    for (int i = 0; i < c.n_stages; ++i) {
        const int C = d->stage_channels(i);
        const std::string is = std::to_string(i);
        int sf = 1;
        for (int j = i + 1; j < c.n_stages; ++j) sf *= c.upsample_rates[j];
        const bool last = (i + 1 == c.n_stages);
        const int nc_cin = istft ? c.gen_istft_n_fft + 2 : 1;
        P.conv(d->noise_convs[i], "generator.noise_convs." + is, nc_cin, C, last ? 1 : 2 * sf, false, true, false);
        P.resblock1(d->noise_res[i], "generator.noise_res." + is, C, last ? 11 : 7, dil135);
        P.conv(d->ups[i], "generator.ups." + is, 2 * C, C, c.upsample_kernel_sizes[i], true, true, true);
        if (!istft) d->gen_alpha[i + 1] = P.copy("generator.alphas." + std::to_string(i + 1), C);
        for (int j = 0; j < c.n_kernels; ++j)
            P.resblock1(d->resblocks[i * c.n_kernels + j], "generator.resblocks." + std::to_string(i * c.n_kernels + j),
                        C, c.resblock_kernel_sizes[j], c.resblock_dilations[j]);
    }
    {
        const int Cl = d->stage_channels(c.n_stages - 1);
        P.conv(d->conv_post, "generator.conv_post", Cl, istft ? c.gen_istft_n_fft + 2 : 1, 7, false, true, istft);
    }
    if (istft) {
        const int n = c.gen_istft_n_fft, bins = n / 2 + 1;
        d->stft_fr = P.copy("generator.stft.weight_forward_real", (int64_t)bins * n);
        d->stft_fi = P.copy("generator.stft.weight_forward_imag", (int64_t)bins * n);
        d->stft_br = P.copy("generator.stft.weight_backward_real", (int64_t)bins * n);
        d->stft_bi = P.copy("generator.stft.weight_backward_imag", (int64_t)bins * n);
    }
}

static int finalize_impl(st2_decoder* d, cudaStream_t st) {
    const st2_config& c = d->cfg;
    Packer P{d, st};
    d->fc_rows = 0;
    if (c.variant == 2) pack_predictor(d, P);
    else if (c.variant == 3) pack_text_encoder(d, P);
    else pack_decoder(d, P);
    if (P.err != ST2_OK) return P.err;
    // all AdaIN fc layers -> one [R,style] matrix (rows: gamma(C) | beta(C) per instance)
    d->fc_w = (float*)P.dalloc((size_t)d->fc_rows * c.style_dim * sizeof(float));
    d->fc_b = (float*)P.dalloc((size_t)d->fc_rows * sizeof(float));
    if (!d->fc_w || !d->fc_b) return P.err;
    int row = 0;
    for (auto& pr : P.adain_list) {
        const RawTensor* w = P.get(pr.first + ".fc.weight");
        const RawTensor* b = P.get(pr.first + ".fc.bias");
        if (!w || !b) return P.err;
        const int rows = 2 * pr.second;
        if (w->numel() != (int64_t)rows * c.style_dim || b->numel() != rows) {
            set_error("AdaIN fc '%s' has the wrong shape", pr.first.c_str());
            return ST2_ERR_INVALID;
        }
        ST2_CUDA_CHECK(cudaMemcpyAsync(d->fc_w + (size_t)row * c.style_dim, w->ptr,
                                       (size_t)rows * c.style_dim * sizeof(float), cudaMemcpyDeviceToDevice, st));
        ST2_CUDA_CHECK(cudaMemcpyAsync(d->fc_b + row, b->ptr, (size_t)rows * sizeof(float), cudaMemcpyDeviceToDevice, st));
        row += rows;
    }
    ST2_CUDA_CHECK(cudaStreamSynchronize(st));
    ST2_CUDA_CHECK(cudaGetLastError());
    d->num_params = 0;
    for (auto& kv : d->raw)
        if (kv.first.find("generator.stft.") == std::string::npos) d->num_params += kv.second.numel();
    d->finalized = true;
    return ST2_OK;
}

// ------------------------------------------------------------------------------------------
// forward program
// ------------------------------------------------------------------------------------------
struct Exec {
    st2_decoder* d;
    cudaStream_t st;
    bool dry;                 // plan only: compute the workspace high-water mark
    int prec;                 // st2_precision
    int B;
    char* base;
    int64_t cap, off = 0, peak = 0;
    int err = ST2_OK;
    const float* H = nullptr; // style rows [B][fc_rows]
    float* coef = nullptr;    // [B][2][2048]
    int coef_ld = 0;          // stride the last coefficient kernel wrote with

    void* alloc(int64_t bytes) {
        off = (off + 255) / 256 * 256;
        void* p = dry ? nullptr : base + off;
        off += bytes;
        if (off > peak) peak = off;
        if (!dry && off > cap && err == ST2_OK) {
            set_error("workspace too small: need at least %lld bytes, have %lld", (long long)off, (long long)cap);
            err = ST2_ERR_WORKSPACE;
        }
        return p;
    }
    float* allocf(int64_t n) { return (float*)alloc(n * (int64_t)sizeof(float)); }
    bool live() const { return !dry && err == ST2_OK; }
    void chk(int e) { if (e != ST2_OK && err == ST2_OK) err = e; }

    // boundary event after the launch(es) just issued; flops / bytes are the ALGORITHMIC figures
    void prof(int cat, double flops, double bytes) {
        if (!d->profiling || !live()) return;
        const size_t idx = d->prof_recs.size() + 1;      // event 0 = start of the forward
        while (d->prof_events.size() <= idx) {
            cudaEvent_t ev;
            if (cudaEventCreate(&ev) != cudaSuccess) { chk(ST2_ERR_CUDA); return; }
            d->prof_events.push_back(ev);
        }
        cudaEventRecord(d->prof_events[idx], st);
        d->prof_recs.push_back({cat, flops, bytes});
    }

    int fmt_for(const std::string& name) const {
        if (prec == ST2_PREC_FP32) return DT_F32;
        if (prec == ST2_PREC_FP16) return DT_F16;
        if (d->cfg.variant >= 2) return DT_F16;   // predictor / text encoder: 0.4 % of the decoder's FLOPs, its outputs steer the SineGen phase
        // bf16 for the generator resblocks / ups (96 % of the FLOPs).  fp16 operands (same tensor
        // throughput) for generator.noise_res -- bf16 there alone costs ~9 dB of SNR -- and for the
        // front half (encode / decode / asr_res, K up to 3270), whose five chained blocks otherwise
        // reach 1.2e-2 per-layer relative L2 (precision study).
        if (name.find("noise_res") != std::string::npos || name.find("encode") != std::string::npos ||
            name.find("decode") != std::string::npos || name.find("asr_res") != std::string::npos)
            return DT_F16;
        return DT_BF16;
    }

    void tap(const std::string& name, const float* src, int ld, int64_t rows, int C) {
        if (!live()) return;
        auto it = d->taps.find(name);
        if (it == d->taps.end() || it->second.dst == nullptr) return;
        if (rows * C > it->second.cap) {
            set_error("tap '%s' needs %lld floats, buffer has %lld", name.c_str(), (long long)(rows * C),
                      (long long)it->second.cap);
            err = ST2_ERR_INVALID;
            return;
        }
        chk(launch_copy_dense(src, ld, it->second.dst, rows, C, st));
        prof(PC_MISC, 0, 8.0 * rows * C);
    }

    // tap of a dense fp16 tensor (the intra-block tensor of the fused resblocks)
    void tap16(const std::string& name, const void* src, int64_t n) {
        if (!live()) return;
        auto it = d->taps.find(name);
        if (it == d->taps.end() || it->second.dst == nullptr) return;
        if (n > it->second.cap) {
            set_error("tap '%s' needs %lld floats, buffer has %lld", name.c_str(), (long long)n, (long long)it->second.cap);
            err = ST2_ERR_INVALID;
            return;
        }
        chk(launch_half_to_float(src, it->second.dst, n, st));
        prof(PC_MISC, 0, 6.0 * n);
    }

    // y = act(AdaIN(x)) or act(x) when `n` is null.  x fp32 [B,T,ld_x]; y dtype dt, pitch ld_y.
    void norm_act(const float* x, int ld_x, int T, int C, const AdaINRef* n, int act, float slope, const float* alpha,
                  void* y, int ld_y, int dt) {
        const int Cpad = ld_y < ld_x ? ld_y : ld_x;
        void* scratch = nullptr;
        const int64_t mark = off;
        if (n) scratch = alloc(adain_scratch_bytes(B, T, C));
        if (live()) {
            const double numel = (double)B * T * C;
            if (n) {
                chk(launch_in_stats(x, ld_x, B, T, C, scratch, st));
                prof(PC_NORM_STATS, 0, numel * 4);
            }
            if (err == ST2_OK) chk(launch_adain_coef(scratch, n ? H : nullptr, d->fc_rows, n ? n->h_off : 0, coef, B, T, C, Cpad, st));
            prof(PC_NORM_COEF, 0, 0);
            if (err == ST2_OK) chk(launch_affine_act(x, ld_x, coef, alpha, act, slope, y, ld_y, dt, B, T, Cpad, st));
            prof(PC_AFFINE_ACT, 0, numel * (4 + (dt == DT_F32 ? 4 : 2)));
        }
        off = mark;
    }

    bool use_tc(const ConvW& w, int dt) const { return dt != DT_F32 && w.w16[dt] != nullptr; }

    // geometry of Conv1d / (polyphase) ConvTranspose1d in the common ConvArgs contract
    bool fill_args(ConvArgs& a, const ConvW& w, int Tin, int Tout, int stride, int padding, int dilation, int out_row_shift) {
        memset(&a, 0, sizeof(a));
        a.B = B; a.Cin = w.Cin; a.Cout = w.Cout;
        a.Tin = Tin; a.Tout = Tout;
        a.w = w.w32; a.bias = w.bias;
        if (!w.transposed) {
            a.M = Tout; a.ntaps = w.k; a.tap_step = dilation; a.in_off = -padding; a.in_stride = stride;
            a.phases = 1; a.w_step = 1; a.out_stride = 1; a.out_pad = -out_row_shift;
        } else {
            if (w.k % stride != 0) {
                set_error("ConvTranspose1d k=%d must be a multiple of stride=%d", w.k, stride);
                err = ST2_ERR_UNSUPPORTED;
                return false;
            }
            a.ntaps = w.k / stride; a.tap_step = -1; a.in_off = 0; a.in_stride = 1;
            a.phases = stride; a.w_step = stride; a.out_stride = stride; a.out_pad = padding - out_row_shift;
            a.M = (Tout - 1 - out_row_shift + padding) / stride + 1;
        }
        return true;
    }

    // Conv1d (stride 1 here except noise_convs) / ConvTranspose1d through the common ConvArgs contract
    void conv(const ConvW& w, const void* x, int ld_x, int Tin, int dt, float* y, int ld_y, int Tout, int stride,
              int padding, int dilation, const float* res, int ld_res, int res_shift, float scale, int accumulate,
              int out_row_shift = 0, int mirror = 0) {
        if (!live()) return;
        ConvArgs a;
        if (!fill_args(a, w, Tin, Tout, stride, padding, dilation, out_row_shift)) return;
        a.res = res; a.ld_res = ld_res; a.res_shift = res_shift;
        a.y = y; a.ld_y = ld_y;
        a.scale = scale; a.accumulate = accumulate; a.mirror = mirror;
        // algorithmic work: Conv1d 2*B*Tout*Cout*Cin*k ; ConvTranspose1d 2*B*Tin*Cin*Cout*k
        const double flops = 2.0 * B * (w.transposed ? (double)Tin : (double)(Tout - out_row_shift)) * w.Cin * w.Cout * w.k;
        const bool tc = use_tc(w, dt);
        const double bytes = (double)B * ((double)w.Cin * Tin * (tc ? 2 : 4) +
                                          (double)w.Cout * Tout * 4 * (1 + (res ? 1 : 0) + (accumulate ? 1 : 0))) +
                             (double)w.k * w.Cin * w.Cout * (tc ? 2 : 4);
        if (tc) {
            a.x16 = x; a.ld_x16 = ld_x; a.w16 = w.w16[dt]; a.w16_cin_pad = w.cin_pad; a.w16_cout_pad = w.cout_pad;
            a.fmt16 = dt;
            chk(launch_conv_tc(a, st));
            prof(PC_CONV_TC, flops, bytes);
        } else {
            a.x = (const float*)x; a.ld_x = ld_x;
            chk(launch_conv_simt(a, st));
            prof(PC_CONV_SIMT, flops, bytes);
        }
    }

    // ---- fused path: statistics references, coefficient kernels, fused conv ---------------------
    struct StatRef { const void* ptr; int nparts; bool f2; };   // f2: float2 tile partials; else double2 slab partials

    // statistics of a tensor no fused epilogue produced (noise_convs output): standalone pass
    StatRef stats_standalone(const float* x, int ld_x, int T, int C) {
        void* scratch = alloc(adain_scratch_bytes(B, T, C));
        if (live()) {
            chk(launch_in_stats(x, ld_x, B, T, C, scratch, st));
            prof(PC_NORM_STATS, 0, (double)B * T * C * 4);
        }
        return StatRef{scratch, 0, false};
    }
    // coef <- (1+gamma)*rstd, beta - mean*(1+gamma)*rstd  (n == nullptr: identity)
    void coef_from(const StatRef& sr, const AdaINRef* n, int T, int C, int Cpad) {
        if (!live()) return;
        if (n == nullptr || !sr.f2)
            chk(launch_adain_coef(n ? sr.ptr : nullptr, n ? H : nullptr, d->fc_rows, n ? n->h_off : 0, coef, B, T, C, Cpad, st));
        else
            chk(launch_adain_coef_f2(sr.ptr, sr.nparts, H, d->fc_rows, n->h_off, coef, B, T, C, Cpad, st));
        prof(PC_NORM_COEF, 0, 0);
        coef_ld = Cpad;
    }
    bool can_fuse(const ConvW& w, int dt, int ld_x, int ld_y, int stride, int dilation) {
        if (!use_tc(w, dt) || getenv("ST2_NO_FUSED") != nullptr) return false;
        ConvArgs a;
        memset(&a, 0, sizeof(a));
        a.in_stride = 1; a.w16 = w.w16[dt]; a.w16_cin_pad = w.cin_pad; a.w16_cout_pad = w.cout_pad;
        a.Cin = w.Cin; a.Cout = w.Cout; a.ld_x = ld_x; a.ld_y = ld_y; a.ld_res = 4;
        a.ntaps = w.transposed ? w.k / stride : w.k;
        a.tap_step = w.transposed ? -1 : dilation;
        return conv_fused_supported(a);
    }
    int fused_parts(const ConvW& w, int Tout, int stride, int padding, int out_row_shift) {
        ConvArgs a;
        memset(&a, 0, sizeof(a));
        a.w16_cout_pad = w.cout_pad;
        a.phases = w.transposed ? stride : 1;
        a.M = w.transposed ? (Tout - 1 - out_row_shift + padding) / stride + 1 : Tout;
        return fused_stats_parts(a);
    }
    // would launch_conv_fused run this stride-1 conv on the TMA pipeline kernel (conv_pipe.cu) with these storage types?
    bool pipe_ok(const ConvW& w, int ld_x, int ld_y, int T, int padding, int dilation, bool has_res, int accumulate, int dt,
                 int x16in, int y16out, int res16 = 0, int acc16 = 0) {
        ConvArgs a;
        if (!fill_args(a, w, T, T, 1, padding, dilation, 0)) return false;
        a.accumulate = accumulate;
        a.ld_x = ld_x; a.ld_y = ld_y; a.ld_res = ld_y; a.res = has_res ? (const float*)this : nullptr;   // only null-ness matters
        a.w16 = w.w16[dt]; a.w16_cin_pad = w.cin_pad; a.w16_cout_pad = w.cout_pad; a.fmt16 = dt;
        a.x16in = x16in; a.y16out = y16out; a.res16 = res16; a.scale = 1.f;
        if (accumulate && acc16) { a.acc_src = this; a.acc16 = 1; }                                      // only null-ness matters
        return conv_pipe_supported(a);
    }
    // would launch_conv_fused run this upsampling conv on the TMA pipeline kernel and write a 16-bit output?
    bool pipe_ok_ups16(const ConvW& w, int ld_x, int ld_y, int Tin, int Tout, int stride, int padding, int shift, int dt) {
        ConvArgs a;
        if (shift != 0 || !fill_args(a, w, Tin, Tout, stride, padding, 1, 0)) return false;
        a.ld_x = ld_x; a.ld_y = ld_y; a.ld_res = ld_y; a.res = (const float*)this;                        // only null-ness matters
        a.w16 = w.w16[dt]; a.w16_cin_pad = w.cin_pad; a.w16_cout_pad = w.cout_pad; a.fmt16 = dt;
        a.y16out = 1; a.scale = 1.f;
        return conv_pipe_supported(a);
    }
    // y = epilogue( conv( act(coef.a * x + coef.b) ) ), statistics of y -> stats_out (float2 partials)
    void conv_fused(const ConvW& w, const float* x, int ld_x, int Tin, int dt, int act, float slope, const float* alpha,
                    float* y, int ld_y, int Tout, int stride, int padding, int dilation, const float* res, int ld_res,
                    int res_shift, float scale, int accumulate, void* stats_out, int out_row_shift = 0, int mirror = 0,
                    int x16in = 0, int y16out = 0, int res16 = 0, const void* acc_src = nullptr, int acc16 = 0) {
        if (!live()) return;
        ConvArgs a;
        if (!fill_args(a, w, Tin, Tout, stride, padding, dilation, out_row_shift)) return;
        a.res = res; a.ld_res = ld_res; a.res_shift = res_shift;
        a.y = y; a.ld_y = ld_y;
        a.scale = scale; a.accumulate = accumulate; a.mirror = mirror;
        a.x = x; a.ld_x = ld_x; a.w16 = w.w16[dt]; a.w16_cin_pad = w.cin_pad; a.w16_cout_pad = w.cout_pad; a.fmt16 = dt;
        a.x16in = x16in; a.y16out = y16out; a.res16 = res16; a.acc_src = acc_src; a.acc16 = acc16;
        chk(launch_conv_fused(a, coef, coef_ld, act, slope, alpha, stats_out, st));
        const double flops = 2.0 * B * (w.transposed ? (double)Tin : (double)(Tout - out_row_shift)) * w.Cin * w.Cout * w.k;
        const double bytes = (double)B * ((double)w.Cin * Tin * (x16in ? 2 : 4) +
                                          (double)w.Cout * Tout * ((y16out ? 2 : 4) + (res ? (res16 ? 2 : 4) : 0) + (accumulate ? (acc16 ? 2 : 4) : 0))) +
                             (double)w.k * w.Cin * w.Cout * 2;
        prof(conv_pipe_supported(a) ? PC_CONV_PIPE : PC_CONV_FUSED, flops, bytes);
    }

    // AdainResBlk1d.forward (hifigan.py:400-403).  x [B,T,ld_x] (Cin real channels) -> y [B,T or 2T,ld_y]
    void resblk1d(const ResBlk1dW& w, const float* x, int ld_x, int T, float* y, int ld_y) {
        const int64_t mark = off;
        const int dt = fmt_for(w.name);
        const bool tc1 = use_tc(w.conv1, dt);
        const int dt1 = tc1 ? dt : DT_F32;
        const int dta = (tc1 && !w.upsample) ? dt : DT_F32;     // the depthwise pool reads fp32, writes the 16-bit operand
        const int es1 = dta == DT_F32 ? 4 : 2;
        const int Tc = w.upsample ? 2 * T : T;
        void* xa = alloc((int64_t)B * T * ld_x * es1);
        norm_act(x, ld_x, T, w.Cin, &w.norm1, ACT_LRELU, 0.2f, nullptr, xa, ld_x, dta);
        const void* cin = xa;
        if (w.upsample) {
            void* xp = alloc((int64_t)B * Tc * ld_x * (tc1 ? 2 : 4));
            if (live()) {
                if (tc1) chk(launch_pool_dw16((const float*)xa, ld_x, w.pool_w, w.pool_b, xp, ld_x, dt, B, T, ld_x, st));
                else chk(launch_pool_dw((const float*)xa, ld_x, w.pool_w, w.pool_b, (float*)xp, ld_x, B, T, w.Cin, ld_x, st));
            }
            prof(PC_MISC, 0, 4.0 * B * w.Cin * 3.0 * T);
            cin = xp;
        }
        float* h1 = allocf((int64_t)B * Tc * w.Cout);
        conv(w.conv1, cin, ld_x, Tc, dt1, h1, w.Cout, Tc, 1, 1, 1, nullptr, 0, 0, 1.f, 0);
        tap(w.name + ".conv1", h1, w.Cout, (int64_t)B * Tc, w.Cout);
        const bool tc2 = use_tc(w.conv2, dt);
        const int dt2 = tc2 ? dt : DT_F32;
        void* xa2 = alloc((int64_t)B * Tc * w.Cout * (dt2 == DT_F32 ? 4 : 2));
        norm_act(h1, w.Cout, Tc, w.Cout, &w.norm2, ACT_LRELU, 0.2f, nullptr, xa2, w.Cout, dt2);
        const float* res = x;
        int ld_res = ld_x;
        if (w.has_sc) {
            float* sc = allocf((int64_t)B * T * w.Cout);
            const bool tcs = use_tc(w.conv1x1, dt);
            const void* xin = x;
            if (tcs) {   // 16-bit copy of the raw block input for the tensor-core 1x1
                void* x16 = alloc((int64_t)B * T * ld_x * 2);
                norm_act(x, ld_x, T, w.Cin, nullptr, ACT_NONE, 0.f, nullptr, x16, ld_x, dt);
                xin = x16;
            }
            conv(w.conv1x1, xin, ld_x, T, tcs ? dt : DT_F32, sc, w.Cout, T, 1, 0, 1, nullptr, 0, 0, 1.f, 0);
            res = sc;
            ld_res = w.Cout;
        }
        conv(w.conv2, xa2, w.Cout, Tc, dt2, y, ld_y, Tc, 1, 1, 1, res, ld_res, w.upsample ? 1 : 0,
             0.70710678118654752f, 0);
        tap(w.name, y, ld_y, (int64_t)B * Tc, w.Cout);
        off = mark;
    }

    // AdaINResBlock1.forward (hifigan.py:65-74) on x_in [B,T,C]; the running tensor lives in `run`
    // (may alias x_in for an in-place block); the last iteration writes
    // dest = (dest_old*accumulate + conv2 + run) * scale.
    // would resblock1 take its input tensor as fp16 (every conv of the block on the TMA pipeline kernel)?
    // would the last conv of this block write / accumulate the fp16 partial sum of the stage?
    //   mode 1: first block of a stage, writes the fp16 sum;  2: middle, fp16 sum in place;  3: last, fp16 sum -> fp32 stage output
    bool resblock1_sum16_ok(const ResBlock1W& w, int T, int mode, int x16) {
        const int C = w.C, dt = fmt_for(w.name);
        if (!can_fuse(w.c1[0], dt, C, C, 1, 5) || getenv("ST2_NO_XT16") || getenv("ST2_NO_RUN16") || getenv("ST2_NO_SUM16")) return false;
        for (int j = 0; j < 3; ++j) {        // every conv of the block on the pipeline kernel with the fp16 tensors it will see
            const int dil = w.dil[j], in16 = (j > 0 || x16) ? 1 : 0;
            if (!pipe_ok(w.c1[j], C, C, T, (w.k * dil - dil) / 2, dil, false, 0, dt, in16, 1) ||
                !pipe_ok(w.c2[j], C, C, T, (w.k - 1) / 2, 1, true, (j == 2 && mode > 1) ? 1 : 0, dt, 1, (j < 2 || mode < 3) ? 1 : 0, in16,
                         (j == 2 && mode > 1) ? 1 : 0))
                return false;
        }
        return true;
    }
    bool resblock1_x16_ok(const ResBlock1W& w, int T, int accumulate) {
        const int C = w.C, dt = fmt_for(w.name);
        if (!can_fuse(w.c1[0], dt, C, C, 1, 5) || getenv("ST2_NO_XT16") || getenv("ST2_NO_RUN16")) return false;
        for (int j = 0; j < 3; ++j) {
            const int dil = w.dil[j];
            if (!pipe_ok(w.c1[j], C, C, T, (w.k * dil - dil) / 2, dil, false, 0, dt, 1, 1) ||
                !pipe_ok(w.c2[j], C, C, T, (w.k - 1) / 2, 1, true, j == 2 ? accumulate : 0, dt, 1, j < 2, 1))
                return false;
        }
        return true;
    }
    // sum16 (modes above) with sum16buf: the stage's partial sum lives in fp16 until the last block writes `dest`
    void resblock1(const ResBlock1W& w, const float* x_in, float* run, int T, float* dest, float scale, int accumulate,
                   const StatRef* in_stats = nullptr, int x16 = 0, int sum16 = 0, void* sum16buf = nullptr) {
        const int64_t mark = off;
        const int C = w.C;
        const int dt = fmt_for(w.name);
        if (can_fuse(w.c1[0], dt, C, C, 1, 5)) {
            // fused: 2 tiny coefficient kernels + 2 fused convs per iteration; AdaIN statistics come from the
            // producing conv's epilogue (or from in_stats for the block input)
            const int nparts = fused_parts(w.c1[0], T, 1, 0, 0);
            void* st_xt = alloc((int64_t)B * nparts * C * 8);
            void* st_run = alloc((int64_t)B * nparts * C * 8);
            // the intra-block tensor xt (conv1 output, only consumed by conv2's transform) is stored as fp16 when both convs
            // run on the TMA pipeline kernel: 20 % fewer HBM bytes per iteration for -0.1 dB of SNR (its statistics still come
            // from the fp32 values in the epilogue).  ST2_NO_XT16=1 keeps it fp32.
            int xt16 = getenv("ST2_NO_XT16") == nullptr ? 1 : 0;
            for (int j = 0; j < 3 && xt16; ++j) {
                const int dil = w.dil[j];
                if (!pipe_ok(w.c1[j], C, C, T, (w.k * dil - dil) / 2, dil, false, 0, dt, 0, 1) ||
                    !pipe_ok(w.c2[j], C, C, T, (w.k - 1) / 2, 1, true, j == 2 ? accumulate : 0, dt, 1, 0))
                    xt16 = 0;
            }
            float* xt = (float*)alloc((int64_t)B * T * C * 4);      // sized for fp32 (the dry run must not depend on the device)
            // The running tensor between the three iterations (x + conv2 output of iterations 0 and 1; read as conv1's
            // input and conv2's residual by the next iteration) is private to the block: stored as fp16 as well when every
            // conv of the block takes it (25 % fewer HBM bytes per block; AdaIN statistics still come from the fp32 values in
            // the epilogue, the stage output the last iteration writes stays fp32).  ST2_NO_RUN16=1 keeps it fp32.
            int run16 = (xt16 && getenv("ST2_NO_RUN16") == nullptr) ? 1 : 0;
            for (int j = 0; j < 3 && run16; ++j) {
                const int dil = w.dil[j];
                if (!pipe_ok(w.c1[j], C, C, T, (w.k * dil - dil) / 2, dil, false, 0, dt, j > 0 || x16, 1) ||
                    !pipe_ok(w.c2[j], C, C, T, (w.k - 1) / 2, 1, true, j == 2 ? accumulate : 0, dt, 1, j < 2, j > 0 || x16))
                    run16 = 0;
            }
            void* r16buf = alloc((int64_t)B * T * C * 2);           // allocated either way: same workspace on every device
            if (sum16 && !run16 && err == ST2_OK) {                  // the caller asks resblock1_sum16_ok first
                set_error("resblock1: the fp16 stage sum needs the fp16 running-tensor path");
                err = ST2_ERR_STATE;
            }
            if (x16 && !(run16 && in_stats) && err == ST2_OK) {      // the caller asks resblock1_x16_ok first
                set_error("resblock1: fp16 block input needs the fp16 running-tensor path and producer statistics");
                err = ST2_ERR_STATE;
            }
            StatRef cur_st = in_stats ? *in_stats : stats_standalone(x_in, C, T, C);
            const float* cur = x_in;
            int cur16 = x16;
            for (int j = 0; j < 3; ++j) {
                const int dil = w.dil[j];
                coef_from(cur_st, &w.n1[j], T, C, C);
                conv_fused(w.c1[j], cur, C, T, dt, ACT_SNAKE, 0.f, w.alpha1[j], xt, C, T, 1, (w.k * dil - dil) / 2, dil, nullptr,
                           0, 0, 1.f, 0, st_xt, 0, 0, cur16, xt16);
                if (!xt16) tap(w.name + ".convs1." + std::to_string(j), xt, C, (int64_t)B * T, C);
                else tap16(w.name + ".convs1." + std::to_string(j), xt, (int64_t)B * T * C);
                coef_from(StatRef{st_xt, nparts, true}, &w.n2[j], T, C, C);
                const bool last = (j == 2);
                const bool s16 = last && sum16 != 0 && run16;                     // the caller asked resblock1_sum16_ok first
                const int out16 = ((run16 && !last) || (s16 && sum16 < 3)) ? 1 : 0;
                float* out = last ? ((s16 && sum16 < 3) ? (float*)sum16buf : dest) : (out16 ? (float*)r16buf : run);
                conv_fused(w.c2[j], xt, C, T, dt, ACT_SNAKE, 0.f, w.alpha2[j], out, C, T, 1, (w.k - 1) / 2, 1, cur, C, 0,
                           last ? scale : 1.f, last ? accumulate : 0, last ? nullptr : st_run, 0, 0, xt16, out16, cur16,
                           (s16 && sum16 > 1) ? sum16buf : nullptr, (s16 && sum16 > 1) ? 1 : 0);
                if (out16) tap16(w.name + ".iter" + std::to_string(j), out, (int64_t)B * T * C);
                else if (!last || (!accumulate && scale == 1.f)) tap(w.name + ".iter" + std::to_string(j), out, C, (int64_t)B * T, C);
                cur = out;
                cur16 = out16;
                cur_st = StatRef{st_run, nparts, true};
            }
            off = mark;
            return;
        }
        const bool tc = use_tc(w.c1[0], dt);
        const int dta = tc ? dt : DT_F32;
        void* xa = alloc((int64_t)B * T * C * (dta == DT_F32 ? 4 : 2));
        float* xt = allocf((int64_t)B * T * C);
        const float* cur = x_in;
        for (int j = 0; j < 3; ++j) {
            const int dil = w.dil[j];
            norm_act(cur, C, T, C, &w.n1[j], ACT_SNAKE, 0.f, w.alpha1[j], xa, C, dta);
            conv(w.c1[j], xa, C, T, dta, xt, C, T, 1, (w.k * dil - dil) / 2, dil, nullptr, 0, 0, 1.f, 0);
            tap(w.name + ".convs1." + std::to_string(j), xt, C, (int64_t)B * T, C);
            norm_act(xt, C, T, C, &w.n2[j], ACT_SNAKE, 0.f, w.alpha2[j], xa, C, dta);
            const bool last = (j == 2);
            float* out = last ? dest : run;
            conv(w.c2[j], xa, C, T, dta, out, C, T, 1, (w.k - 1) / 2, 1, cur, C, 0, last ? scale : 1.f,
                 last ? accumulate : 0);
            if (!last || (!accumulate && scale == 1.f)) tap(w.name + ".iter" + std::to_string(j), out, C, (int64_t)B * T, C);
            cur = out;
        }
        off = mark;
    }
};

static int forward_impl(st2_decoder* d, const float* asr, const float* f0, const float* nn, const float* s,
                        const float* noise, uint64_t seed, float* out, int B, int T, int prec, void* ws,
                        int64_t ws_bytes, cudaStream_t st, bool dry, int64_t* peak_out) {
    const st2_config& c = d->cfg;
    const bool istft = c.variant == 1;
    Exec E{d, st, dry, prec, B, (char*)ws, ws_bytes};
    const int spf = d->spf();
    const int S = spf * T, L2 = 2 * T;
    const int up_scale = spf / 2;
    const int C514 = c.dim_in + 2, LD514 = round_up(C514, 64);
    const int C1090 = 1024 + 2 + 64, LD1090 = round_up(C1090, 64);

    float* H = E.allocf((int64_t)B * d->fc_rows);
    E.H = H;
    E.coef = E.allocf((int64_t)B * 2 * 2048);
    float* frames = E.allocf((int64_t)B * L2 * 9);
    float* har = E.allocf((int64_t)B * S);
    float* x514 = E.allocf((int64_t)B * T * LD514);
    float* x1090 = E.allocf((int64_t)B * T * LD1090);
    const int C0 = c.upsample_initial_channel;
    float* xg = E.allocf((int64_t)B * 2 * T * C0);
    // per-stage outputs (the mean over the three resblocks), allocated up front
    float* stage_out[4];
    int stage_T[4];
    {
        int Tcur = 2 * T;
        for (int i = 0; i < c.n_stages; ++i) {
            Tcur *= c.upsample_rates[i];
            stage_T[i] = Tcur + ((istft && i + 1 == c.n_stages) ? 1 : 0);
            stage_out[i] = E.allocf((int64_t)B * stage_T[i] * d->stage_channels(i));
        }
    }
    const int har_frames = S / (istft ? c.gen_istft_hop_size : 1) + 1;
    const int HLD = 24;
    float* har22 = istft ? E.allocf((int64_t)B * har_frames * HLD) : nullptr;

    if (E.live()) {
        if (d->profiling) {
            d->prof_recs.clear();
            if (d->prof_events.empty()) {
                cudaEvent_t ev;
                if (cudaEventCreate(&ev) == cudaSuccess) d->prof_events.push_back(ev);
            }
            if (!d->prof_events.empty()) cudaEventRecord(d->prof_events[0], st);
        }
        E.chk(launch_style_fc(s, d->fc_w, d->fc_b, H, B, d->fc_rows, c.style_dim, st));
        E.chk(launch_cf_to_cl(asr, x514, LD514, B, c.dim_in, T, st));
        E.chk(launch_f0n_conv(f0, nn, d->f0_w, d->f0_b, d->n_w, d->n_b, x514, LD514, c.dim_in, C514, x1090, LD1090,
                              1024 + 64, C1090, B, T, st));
        E.prof(PC_MISC, 2.0 * B * d->fc_rows * c.style_dim,
               4.0 * ((double)d->fc_rows * c.style_dim + 2.0 * B * c.dim_in * T + 4.0 * B * L2));
        E.chk(launch_sinegen_frames(f0, frames, B, L2, up_scale, st));
        E.chk(launch_har_source(f0, frames, noise, seed, d->seed_dev, d->lin_w, d->lin_b, har, B, L2, up_scale, st));
        // SineGen algorithmic bytes: read 4*B*2T (+ 36*B*S of noise when taped), write 4*B*S
        E.prof(PC_SOURCE, 0, 4.0 * B * L2 + (noise ? 36.0 * B * S : 0.0) + 4.0 * B * S);
        if (istft) {
            E.chk(launch_stft_transform(har, d->stft_fr, d->stft_fi, har22, HLD, B, S, c.gen_istft_n_fft,
                                        c.gen_istft_hop_size, st));
            E.prof(PC_SOURCE, 0, 4.0 * B * S + 4.0 * B * har_frames * (c.gen_istft_n_fft + 2));
        }
    }
    E.tap("har_source", har, 1, (int64_t)B * S, 1);
    if (istft) E.tap("har", har22, HLD, (int64_t)B * har_frames, c.gen_istft_n_fft + 2);

    // ---- front half: encode, asr_res, decode[0..3] (hifigan.py:461-472)
    E.resblk1d(d->encode, x514, LD514, T, x1090, LD1090);
    {
        const int dt = E.fmt_for("asr_res");
        const bool tc = E.use_tc(d->asr_res, dt);
        const int64_t mark = E.off;
        const void* xin = x514;
        if (tc) {
            void* x16 = E.alloc((int64_t)B * T * LD514 * 2);
            E.norm_act(x514, LD514, T, C514, nullptr, ACT_NONE, 0.f, nullptr, x16, LD514, dt);
            xin = x16;
        }
        E.conv(d->asr_res, xin, LD514, T, tc ? dt : DT_F32, x1090 + 1024, LD1090, T, 1, 0, 1, nullptr, 0, 0, 1.f, 0);
        E.off = mark;
    }
    for (int i = 0; i < 3; ++i) E.resblk1d(d->decode[i], x1090, LD1090, T, x1090, LD1090);
    E.resblk1d(d->decode[3], x1090, LD1090, T, xg, C0);
    E.tap("decode.out", xg, C0, (int64_t)B * 2 * T, C0);

    // ---- generator (hifigan.py:328-345 / istftnet.py:552-573)
    const float* x = xg;
    int Tin = 2 * T;
    for (int i = 0; i < c.n_stages; ++i) {
        const int64_t mark = E.off;
        const int C = d->stage_channels(i), Cin = 2 * C;
        const int u = c.upsample_rates[i], ku = c.upsample_kernel_sizes[i];
        const int Tout = stage_T[i];
        const bool last = (i + 1 == c.n_stages);
        const int shift = (istft && last) ? 1 : 0;          // ReflectionPad1d((1,0))
        const std::string is = std::to_string(i);
        // x_source = noise_res[i](noise_convs[i](har_source), s)
        float* nc = E.allocf((int64_t)B * Tout * C);
        Exec::StatRef nc_stats{nullptr, 0, true};
        bool nc_have_stats = false;
        {
            const ConvW& w = d->noise_convs[i];
            int sf = 1;
            for (int j = i + 1; j < c.n_stages; ++j) sf *= c.upsample_rates[j];
            const int stride = last ? 1 : sf, pad = last ? 0 : (sf + 1) / 2;
            if (istft) {
                E.conv(w, har22, HLD, har_frames, DT_F32, nc, C, Tout, stride, pad, 1, nullptr, 0, 0, 1.f, 0);
            } else {
                // Conv1d(1 -> C) of the harmonic source: dedicated HBM-bound kernel that also emits the InstanceNorm
                // partials of its output (consumed by noise_res[i].adain1[0])
                nc_stats.nparts = noise_conv_parts(Tout);
                void* stp = E.alloc((int64_t)B * nc_stats.nparts * C * 8);
                nc_stats.ptr = stp;
                nc_have_stats = true;
                if (E.live()) {
                    E.chk(launch_noise_conv(har, w.w32, w.bias, nc, stp, B, S, Tout, C, w.k, stride, pad, st));
                    E.prof(PC_SOURCE, 2.0 * B * Tout * C * w.k, 4.0 * B * ((double)S + (double)Tout * C));
                }
            }
        }
        E.resblock1(d->noise_res[i], nc, nc, Tout, nc, 1.f, 0, nc_have_stats ? &nc_stats : nullptr);
        // x = ups[i](act(x)) + x_source
        const int dtu = E.fmt_for("generator.ups");
        const bool tcu = E.use_tc(d->ups[i], dtu);
        const int pu = istft ? (ku - u) / 2 : (u / 2 + u % 2);
        float* xu = E.allocf((int64_t)B * Tout * C);
        Exec::StatRef xu_stats{nullptr, 0, true};
        const bool fuse_u = E.can_fuse(d->ups[i], dtu, Cin, C, u, 1);
        // The stage input xu = ups(x) + x_source is read six times (input and residual of the first iteration of the three
        // resblocks): stored as fp16 when the ups conv and all six convs run on the TMA pipeline kernel (ST2_NO_XU16=1: fp32)
        int xu16 = 0;
        if (fuse_u && getenv("ST2_NO_XU16") == nullptr && E.pipe_ok_ups16(d->ups[i], Cin, C, Tin, Tout, u, pu, shift, dtu)) {
            xu16 = 1;
            for (int j = 0; j < c.n_kernels; ++j)
                if (!E.resblock1_x16_ok(d->resblocks[i * c.n_kernels + j], Tout, j > 0 ? 1 : 0)) xu16 = 0;
        }
        if (fuse_u) {
            // Snake / LeakyReLU applied on the A-operand path; InstanceNorm statistics of xu from the epilogue
            xu_stats.nparts = E.fused_parts(d->ups[i], Tout, u, pu, shift);
            void* st_xu = E.alloc((int64_t)B * xu_stats.nparts * C * 8);
            xu_stats.ptr = st_xu;
            E.coef_from(Exec::StatRef{nullptr, 0, false}, nullptr, Tin, Cin, Cin);
            E.conv_fused(d->ups[i], x, Cin, Tin, dtu, istft ? ACT_LRELU : ACT_SNAKE, 0.1f, istft ? nullptr : d->gen_alpha[i], xu, C,
                         Tout, u, pu, 1, nc, C, 0, 1.f, 0, st_xu, shift, shift, 0, xu16);
        } else {
            void* xs = E.alloc((int64_t)B * Tin * Cin * (tcu ? 2 : 4));
            if (istft) E.norm_act(x, Cin, Tin, Cin, nullptr, ACT_LRELU, 0.1f, nullptr, xs, Cin, tcu ? dtu : DT_F32);
            else E.norm_act(x, Cin, Tin, Cin, nullptr, ACT_SNAKE, 0.f, d->gen_alpha[i], xs, Cin, tcu ? dtu : DT_F32);
            // istftnet: ReflectionPad1d((1,0)) after the last ups = write at row t+1 and mirror row 2 into row 0
            E.conv(d->ups[i], xs, Cin, Tin, tcu ? dtu : DT_F32, xu, C, Tout, u, pu, 1, nc, C, 0, 1.f, 0, shift, shift);
        }
        if (xu16) E.tap16("generator.stage" + is + ".in", xu, (int64_t)B * Tout * C);
        else E.tap("generator.stage" + is + ".in", xu, C, (int64_t)B * Tout, C);
        float* run = E.allocf((int64_t)B * Tout * C);
        // the partial sum over the resblocks of the stage (written by the first, read + written by the middle ones, read by the
        // last, which writes the fp32 stage output) is kept in fp16 when all their last convs take it (ST2_NO_SUM16=1: fp32)
        void* sum16buf = E.alloc((int64_t)B * Tout * C * 2);
        int sum16 = c.n_kernels >= 2 ? 1 : 0;
        for (int j = 0; j < c.n_kernels && sum16; ++j)
            if (!E.resblock1_sum16_ok(d->resblocks[i * c.n_kernels + j], Tout, j == 0 ? 1 : (j + 1 == c.n_kernels ? 3 : 2), xu16)) sum16 = 0;
        for (int j = 0; j < c.n_kernels; ++j) {
            const bool lastk = (j + 1 == c.n_kernels);
            E.resblock1(d->resblocks[i * c.n_kernels + j], xu, run, Tout, stage_out[i], lastk ? 1.f / (float)c.n_kernels : 1.f,
                        j > 0 ? 1 : 0, fuse_u ? &xu_stats : nullptr, xu16, sum16 ? (j == 0 ? 1 : (lastk ? 3 : 2)) : 0, sum16buf);
        }
        E.tap("generator.stage" + is + ".out", stage_out[i], C, (int64_t)B * Tout, C);
        x = stage_out[i];
        Tin = Tout;
        E.off = mark;
    }
    const int Cl = d->stage_channels(c.n_stages - 1);
    if (!istft) {
        if (E.live())
            E.chk(launch_post_hifigan(x, Cl, d->gen_alpha[c.n_stages], d->conv_post.w32, d->conv_post.bias, out, B, S, Cl,
                                       prec != ST2_PREC_FP32 ? 1 : 0, st));
        E.prof(PC_POST, 2.0 * B * S * Cl * 7, 4.0 * B * S * (Cl + 1));
    } else {
        const int dt = E.fmt_for("generator.conv_post");
        const bool tc = E.use_tc(d->conv_post, dt);
        void* xa = E.alloc((int64_t)B * Tin * Cl * (tc ? 2 : 4));
        E.norm_act(x, Cl, Tin, Cl, nullptr, ACT_LRELU, 0.01f, nullptr, xa, Cl, tc ? dt : DT_F32);
        float* y22 = E.allocf((int64_t)B * Tin * HLD);
        E.conv(d->conv_post, xa, Cl, Tin, tc ? dt : DT_F32, y22, HLD, Tin, 1, 3, 1, nullptr, 0, 0, 1.f, 0);
        E.tap("generator.conv_post", y22, HLD, (int64_t)B * Tin, c.gen_istft_n_fft + 2);
        if (E.live())
            E.chk(launch_istft_head(y22, HLD, d->stft_br, d->stft_bi, out, B, Tin, S, c.gen_istft_n_fft,
                                    c.gen_istft_hop_size, st));
        E.prof(PC_POST, 0, 4.0 * B * Tin * (c.gen_istft_n_fft + 2) + 4.0 * B * S);
    }
    if (peak_out) *peak_out = E.peak;
    return E.err;
}

// bidirectional LSTM over channels-last x [B][T][I] -> y [B][T][2H]; G [B][T][8H] scratch for the input half of the gates
static void bilstm(Exec& E, const LstmW& w, const char* name, const float* x, float* G, float* y, int T, int I, int H) {
    const int B = E.B;
    const int64_t mark = E.off;
    const int dt = E.fmt_for(name);
    const bool tc = E.use_tc(w.ih[0], dt) && E.use_tc(w.ih[1], dt);
    const void* xin = x;
    if (tc) {
        void* x16 = E.alloc((int64_t)B * T * I * 2);
        E.norm_act(x, I, T, I, nullptr, ACT_NONE, 0.f, nullptr, x16, I, dt);
        xin = x16;
    }
    for (int dir = 0; dir < 2; ++dir)
        E.conv(w.ih[dir], xin, I, T, tc ? dt : DT_F32, G + (size_t)dir * 4 * H, 8 * H, T, 1, 0, 1, nullptr, 0, 0, 1.f, 0);
    if (E.live()) E.chk(launch_lstm_bidir(G, w.whh, w.bhh, y, B, T, H, E.st));
    E.prof(PC_LSTM, 2.0 * B * T * 2 * 4 * H * H, 4.0 * B * T * (8 * H + 2 * H) + 4.0 * 2 * 4 * H * H);
    E.off = mark;
}

// ProsodyPredictor.F0Ntrain(x, s) (models.py:448-461): en [B, d_hid+style, T], s [B, style] -> F0 [B,2T], N [B,2T]
static int f0n_forward_impl(st2_decoder* d, const float* en, const float* s, float* f0_out, float* n_out, int B, int T,
                            int prec, void* ws, int64_t ws_bytes, cudaStream_t st, bool dry, int64_t* peak_out) {
    const st2_config& c = d->cfg;
    const int dh = c.dim_in, H = dh / 2, I = dh + c.style_dim;
    Exec E{d, st, dry, prec, B, (char*)ws, ws_bytes};
    float* Hs = E.allocf((int64_t)B * d->fc_rows);
    E.H = Hs;
    E.coef = E.allocf((int64_t)B * 2 * 2048);
    float* x = E.allocf((int64_t)B * T * I);            // en, channels-last
    float* G = E.allocf((int64_t)B * T * 8 * H);        // input half of the gates, fwd 4H | rev 4H per row
    float* y = E.allocf((int64_t)B * T * dh);           // LSTM output, fwd H | rev H
    if (E.live()) {
        if (d->profiling) {
            d->prof_recs.clear();
            if (d->prof_events.empty()) {
                cudaEvent_t ev;
                if (cudaEventCreate(&ev) == cudaSuccess) d->prof_events.push_back(ev);
            }
            if (!d->prof_events.empty()) cudaEventRecord(d->prof_events[0], st);
        }
        E.chk(launch_style_fc(s, d->fc_w, d->fc_b, Hs, B, d->fc_rows, c.style_dim, st));
        E.chk(launch_cf_to_cl(en, x, I, B, I, T, st));
        E.prof(PC_MISC, 2.0 * B * d->fc_rows * c.style_dim, 4.0 * ((double)d->fc_rows * c.style_dim + 2.0 * B * I * T));
    }
    // x, _ = self.shared(x.transpose(-1, -2))   (models.py:449)
    bilstm(E, d->shared, "shared", x, G, y, T, I, H);
    E.tap("shared", y, dh, (int64_t)B * T, dh);
    for (int br = 0; br < 2; ++br) {                     // models.py:451-454 (F0) and :456-459 (N)
        const int64_t mark = E.off;
        float* a0 = E.allocf((int64_t)B * T * dh);
        E.resblk1d(d->pred_blk[br][0], y, dh, T, a0, dh);
        float* a1 = E.allocf((int64_t)B * 2 * T * H);
        E.resblk1d(d->pred_blk[br][1], a0, dh, T, a1, H);
        float* a2 = E.allocf((int64_t)B * 2 * T * H);
        E.resblk1d(d->pred_blk[br][2], a1, H, 2 * T, a2, H);
        E.conv(d->pred_proj[br], a2, H, 2 * T, DT_F32, br == 0 ? f0_out : n_out, 1, 2 * T, 1, 0, 1, nullptr, 0, 0, 1.f, 0);
        E.off = mark;
    }
    if (peak_out) *peak_out = E.peak;
    return E.err;
}

// inference.py:242-245 for an equal-length batch: d = predictor.text_encoder(t_en, s, lengths, mask) (DurationEncoder.forward,
// models.py:485-520), x = predictor.lstm(d), duration = sigmoid(duration_proj(x)).sum(-1).
// t_en [B, d_hid, L], s [B, style] -> d_out [B, L, d_hid+style] (the reference's layout of `d`), duration [B, L]
static int dur_forward_impl(st2_decoder* d, const float* t_en, const float* s, float* d_out, float* duration, int B, int L,
                            int prec, void* ws, int64_t ws_bytes, cudaStream_t st, bool dry, int64_t* peak_out) {
    const st2_config& c = d->cfg;
    const int dh = c.dim_in, H = dh / 2, I = dh + c.style_dim;
    Exec E{d, st, dry, prec, B, (char*)ws, ws_bytes};
    float* Hs = E.allocf((int64_t)B * d->fc_rows);
    E.H = Hs;
    E.coef = E.allocf((int64_t)B * 2 * 2048);
    float* xa = E.allocf((int64_t)B * L * I);           // layer input [B][L][d_hid | style]
    float* G = E.allocf((int64_t)B * L * 8 * H);
    float* y = E.allocf((int64_t)B * L * dh);
    if (E.live()) {
        if (d->profiling) {
            d->prof_recs.clear();
            if (d->prof_events.empty()) {
                cudaEvent_t ev;
                if (cudaEventCreate(&ev) == cudaSuccess) d->prof_events.push_back(ev);
            }
            if (!d->prof_events.empty()) cudaEventRecord(d->prof_events[0], st);
        }
        E.chk(launch_style_fc(s, d->fc_w, d->fc_b, Hs, B, d->fc_rows, c.style_dim, st));
        E.chk(launch_cf_to_cl(t_en, xa, I, B, dh, L, st));                       // x.permute / cat([x, s]) (models.py:488-490)
        E.chk(launch_concat_style(xa, I, dh, s, c.style_dim, B, L, st));
        E.prof(PC_MISC, 2.0 * B * d->fc_rows * c.style_dim, 4.0 * ((double)d->fc_rows * c.style_dim + 2.0 * B * I * L));
    }
    for (int i = 0; i < d->dur_layers; ++i) {
        const std::string nm = "text_encoder.lstms." + std::to_string(2 * i);
        bilstm(E, d->enc_lstm[i], nm.c_str(), xa, G, y, L, I, H);                 // models.py:503-509
        E.tap(nm, y, dh, (int64_t)B * L, dh);
        float* dst = (i + 1 == d->dur_layers) ? d_out : xa;                        // the last layer's output is `d`
        if (E.live()) {
            E.chk(launch_ada_layer_norm(y, Hs, d->fc_rows, d->enc_norm[i].h_off, dst, I, B, L, dh, st));   // models.py:498
            E.chk(launch_concat_style(dst, I, dh, s, c.style_dim, B, L, st));      // models.py:499
        }
        E.prof(PC_AFFINE_ACT, 0, 8.0 * B * L * dh);
        E.tap("text_encoder.lstms." + std::to_string(2 * i + 1), dst, I, (int64_t)B * L, dh);
    }
    bilstm(E, d->dur_lstm, "lstm", d_out, G, y, L, I, H);                          // inference.py:243
    E.tap("lstm", y, dh, (int64_t)B * L, dh);
    if (E.live()) E.chk(launch_duration_head(y, d->dur_w, d->dur_b, duration, B, L, dh, d->max_dur, st));   // inference.py:244-245
    E.prof(PC_MISC, 2.0 * B * L * dh * d->max_dur, 4.0 * B * L * (dh + 1));
    if (peak_out) *peak_out = E.peak;
    return E.err;
}

// TextEncoder.forward(x, input_lengths, m) (models.py:258-285) for an equal-length batch (mask all False):
// tokens [B, L] int64 -> out [B, channels, L]
static int text_forward_impl(st2_decoder* d, const int64_t* tokens, float* out, int B, int L, int prec, void* ws, int64_t ws_bytes,
                             cudaStream_t st, bool dry, int64_t* peak_out) {
    const int C = d->cfg.dim_in, H = C / 2;
    Exec E{d, st, dry, prec, B, (char*)ws, ws_bytes};
    E.coef = E.allocf((int64_t)B * 2 * 2048);
    float* xa = E.allocf((int64_t)B * L * C);
    float* xb = E.allocf((int64_t)B * L * C);
    float* G = E.allocf((int64_t)B * L * 8 * H);
    if (E.live()) {
        if (d->profiling) {
            d->prof_recs.clear();
            if (d->prof_events.empty()) {
                cudaEvent_t ev;
                if (cudaEventCreate(&ev) == cudaSuccess) d->prof_events.push_back(ev);
            }
            if (!d->prof_events.empty()) cudaEventRecord(d->prof_events[0], st);
        }
        E.chk(launch_embedding(tokens, d->te_embedding, xa, B, L, C, d->te_symbols, st));       // models.py:259-260
        E.prof(PC_MISC, 0, 8.0 * B * L * C);
    }
    const int dt = E.fmt_for("cnn");
    for (int i = 0; i < d->te_depth; ++i) {                                                       // models.py:264-266
        const int64_t mark = E.off;
        const bool tc = E.use_tc(d->te_conv[i], dt);
        const void* xin = xa;
        if (tc) {
            void* x16 = E.alloc((int64_t)B * L * C * 2);
            E.norm_act(xa, C, L, C, nullptr, ACT_NONE, 0.f, nullptr, x16, C, dt);
            xin = x16;
        }
        E.conv(d->te_conv[i], xin, C, L, tc ? dt : DT_F32, xb, C, L, 1, (d->te_kernel - 1) / 2, 1, nullptr, 0, 0, 1.f, 0);
        if (E.live()) E.chk(launch_layer_norm_lrelu(xb, d->te_gamma[i], d->te_gamma[i] + C, 0.2f, xa, B, L, C, st));
        E.prof(PC_AFFINE_ACT, 0, 8.0 * B * L * C);
        E.tap("cnn." + std::to_string(i), xa, C, (int64_t)B * L, C);
        E.off = mark;
    }
    bilstm(E, d->te_lstm, "lstm", xa, G, xb, L, C, H);                                           // models.py:268-277
    if (E.live()) E.chk(launch_cl_to_cf(xb, out, B, L, C, st));                                   // models.py:279
    E.prof(PC_MISC, 0, 8.0 * B * L * C);
    if (peak_out) *peak_out = E.peak;
    return E.err;
}

}  // namespace st2

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" {

int st2_abi_version(void) { return ST2_ABI_VERSION; }
const char* st2_last_error(void) { return st2::g_err; }

int st2_decoder_create(const st2_config* cfg, st2_decoder** out) {
    ST2_REQUIRE(cfg != nullptr && out != nullptr, "create: null argument");
    ST2_REQUIRE(cfg->variant == 0 || cfg->variant == 1, "create: variant must be 0 (hifigan) or 1 (istftnet)");
    ST2_REQUIRE(cfg->n_stages >= 1 && cfg->n_stages <= 4 && cfg->n_kernels == 3, "create: unsupported stage/kernel count");
    ST2_REQUIRE(cfg->dim_in == 512 && cfg->style_dim >= 4 && cfg->style_dim <= 1024, "create: dim_in must be 512");
    ST2_REQUIRE(cfg->upsample_initial_channel % 64 == 0 && (cfg->upsample_initial_channel >> cfg->n_stages) >= 4 &&
                    ((cfg->upsample_initial_channel >> cfg->n_stages) % 4) == 0,
                "create: unsupported upsample_initial_channel");
    for (int i = 0; i < cfg->n_stages; ++i)
        ST2_REQUIRE(cfg->upsample_rates[i] >= 1 && cfg->upsample_kernel_sizes[i] % cfg->upsample_rates[i] == 0,
                    "create: upsample kernel must be a multiple of its rate");
    if (cfg->variant == 1)
        ST2_REQUIRE(cfg->gen_istft_n_fft == 20 && cfg->gen_istft_hop_size == 5, "create: iSTFT head supports n_fft=20, hop=5");
    st2_decoder* d = new (std::nothrow) st2_decoder();
    ST2_REQUIRE(d != nullptr, "create: out of memory");
    d->cfg = *cfg;
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaGetDeviceProperties(&prop, dev) == cudaSuccess)
        d->tc_ok = (prop.major == 10);
    *out = d;
    return ST2_OK;
}

void st2_decoder_destroy(st2_decoder* d) {
    if (!d) return;
    for (void* p : d->allocs) cudaFree(p);
    for (cudaEvent_t ev : d->prof_events) cudaEventDestroy(ev);
    delete d;
}

int st2_decoder_set_weight(st2_decoder* d, const char* name, const float* dev_ptr, const int64_t* shape, int32_t ndim) {
    ST2_REQUIRE(d && name && dev_ptr && (shape || ndim == 0) && ndim >= 0 && ndim <= 4, "set_weight: bad argument");
    st2::RawTensor t;
    t.ptr = dev_ptr;
    for (int i = 0; i < ndim; ++i) t.shape.push_back(shape[i]);
    d->raw[name] = t;
    return ST2_OK;
}

int st2_decoder_finalize(st2_decoder* d, void* stream) {
    ST2_REQUIRE(d != nullptr, "finalize: null handle");
    for (void* p : d->allocs) cudaFree(p);
    d->allocs.clear();
    d->finalized = false;
    int e = st2::finalize_impl(d, (cudaStream_t)stream);
    if (e != ST2_OK) {
        for (void* p : d->allocs) cudaFree(p);
        d->allocs.clear();
    }
    return e;
}

int64_t st2_decoder_num_params(const st2_decoder* d) { return d ? d->num_params : 0; }

int64_t st2_decoder_workspace_bytes(const st2_decoder* d, int32_t B, int32_t T, int32_t precision) {
    if (!d || !d->finalized || d->cfg.variant >= 2 || B <= 0 || T <= 0) {
        st2::set_error("workspace_bytes: handle not finalized (or a predictor handle) or bad shape");
        return ST2_ERR_STATE;
    }
    int64_t peak = 0;
    int e = st2::forward_impl(const_cast<st2_decoder*>(d), nullptr, nullptr, nullptr, nullptr, nullptr, 0, nullptr, B, T,
                              precision, nullptr, 0, nullptr, true, &peak);
    if (e != ST2_OK) return e;
    return peak + 256;
}

int st2_decoder_forward(st2_decoder* d, const float* asr, const float* f0, const float* n, const float* s,
                        const float* noise, uint64_t seed, float* out, int32_t B, int32_t T, int32_t precision,
                        void* workspace, int64_t workspace_bytes, void* stream) {
    ST2_REQUIRE(d != nullptr, "forward: null handle");
    if (!d->finalized || d->cfg.variant >= 2) {
        st2::set_error("forward: st2_decoder_finalize has not been called (or this is a predictor / text-encoder handle)");
        return ST2_ERR_STATE;
    }
    ST2_REQUIRE(asr && f0 && n && s && out && workspace, "forward: null tensor");
    ST2_REQUIRE(B > 0 && T >= 2, "forward: need B>0 and T>=2 (got B=%d T=%d)", B, T);
    ST2_REQUIRE(precision >= ST2_PREC_FP32 && precision <= ST2_PREC_FP16, "forward: bad precision %d", precision);
    if (precision != ST2_PREC_FP32 && !d->tc_ok) {
        st2::set_error("forward: tensor-core precision requires an sm_100 device");
        return ST2_ERR_UNSUPPORTED;
    }
    ST2_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "forward: workspace must be 256-byte aligned");
    st2::g_launch_count = 0;
    int e = st2::forward_impl(d, asr, f0, n, s, noise, seed, out, B, T, precision, workspace, workspace_bytes,
                              (cudaStream_t)stream, false, nullptr);
    d->last_launches = st2::g_launch_count;
    return e;
}

/* ---- F0 / energy predictor: replaces ProsodyPredictor.F0Ntrain, models.py:448-461 ---- */
int st2_f0n_create(int32_t d_hid, int32_t style_dim, st2_decoder** out) {
    ST2_REQUIRE(out != nullptr, "f0n_create: null argument");
    ST2_REQUIRE(d_hid == 512, "f0n_create: d_hid must be 512 (got %d)", d_hid);
    ST2_REQUIRE(style_dim >= 4 && style_dim <= 1024 && (d_hid + style_dim) % 64 == 0,
                "f0n_create: d_hid + style_dim must be a multiple of 64 (style_dim=%d)", style_dim);
    st2_decoder* d = new (std::nothrow) st2_decoder();
    ST2_REQUIRE(d != nullptr, "f0n_create: out of memory");
    memset(&d->cfg, 0, sizeof(d->cfg));
    d->cfg.variant = 2;
    d->cfg.dim_in = d_hid;
    d->cfg.style_dim = style_dim;
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaGetDeviceProperties(&prop, dev) == cudaSuccess)
        d->tc_ok = (prop.major == 10);
    *out = d;
    return ST2_OK;
}

int64_t st2_f0n_workspace_bytes(const st2_decoder* d, int32_t B, int32_t T, int32_t precision) {
    if (!d || !d->finalized || d->cfg.variant != 2 || B <= 0 || T <= 0) {
        st2::set_error("f0n_workspace_bytes: not a finalized predictor handle, or bad shape");
        return ST2_ERR_STATE;
    }
    int64_t peak = 0;
    int e = st2::f0n_forward_impl(const_cast<st2_decoder*>(d), nullptr, nullptr, nullptr, nullptr, B, T, precision, nullptr, 0,
                                  nullptr, true, &peak);
    if (e != ST2_OK) return e;
    return peak + 256;
}

int st2_f0n_forward(st2_decoder* d, const float* en, const float* s, float* f0, float* n, int32_t B, int32_t T,
                    int32_t precision, void* workspace, int64_t workspace_bytes, void* stream) {
    ST2_REQUIRE(d != nullptr, "f0n_forward: null handle");
    if (!d->finalized || d->cfg.variant != 2) {
        st2::set_error("f0n_forward: not a finalized predictor handle (st2_f0n_create + st2_decoder_finalize)");
        return ST2_ERR_STATE;
    }
    ST2_REQUIRE(en && s && f0 && n && workspace, "f0n_forward: null tensor");
    ST2_REQUIRE(B > 0 && T >= 1, "f0n_forward: need B>0 and T>=1 (got B=%d T=%d)", B, T);
    ST2_REQUIRE(precision >= ST2_PREC_FP32 && precision <= ST2_PREC_FP16, "f0n_forward: bad precision %d", precision);
    if (precision != ST2_PREC_FP32 && !d->tc_ok) {
        st2::set_error("f0n_forward: tensor-core precision requires an sm_100 device");
        return ST2_ERR_UNSUPPORTED;
    }
    ST2_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "f0n_forward: workspace must be 256-byte aligned");
    st2::g_launch_count = 0;
    int e = st2::f0n_forward_impl(d, en, s, f0, n, B, T, precision, workspace, workspace_bytes, (cudaStream_t)stream, false, nullptr);
    d->last_launches = st2::g_launch_count;
    return e;
}

/* duration half: inference.py:242-245 */
int64_t st2_dur_workspace_bytes(const st2_decoder* d, int32_t B, int32_t L, int32_t precision) {
    if (!d || !d->finalized || d->cfg.variant != 2 || !d->has_duration || B <= 0 || L <= 0) {
        st2::set_error("dur_workspace_bytes: not a finalized predictor handle with the duration weights, or bad shape");
        return ST2_ERR_STATE;
    }
    int64_t peak = 0;
    int e = st2::dur_forward_impl(const_cast<st2_decoder*>(d), nullptr, nullptr, nullptr, nullptr, B, L, precision, nullptr, 0,
                                  nullptr, true, &peak);
    if (e != ST2_OK) return e;
    return peak + 256;
}

int st2_dur_forward(st2_decoder* d, const float* t_en, const float* s, float* d_out, float* duration, int32_t B, int32_t L,
                    int32_t precision, void* workspace, int64_t workspace_bytes, void* stream) {
    ST2_REQUIRE(d != nullptr, "dur_forward: null handle");
    if (!d->finalized || d->cfg.variant != 2 || !d->has_duration) {
        st2::set_error("dur_forward: needs a finalized predictor handle that was given text_encoder.* / lstm.* / duration_proj.*");
        return ST2_ERR_STATE;
    }
    ST2_REQUIRE(t_en && s && d_out && duration && workspace, "dur_forward: null tensor");
    ST2_REQUIRE(B > 0 && L >= 1, "dur_forward: need B>0 and L>=1 (got B=%d L=%d)", B, L);
    ST2_REQUIRE(precision >= ST2_PREC_FP32 && precision <= ST2_PREC_FP16, "dur_forward: bad precision %d", precision);
    if (precision != ST2_PREC_FP32 && !d->tc_ok) {
        st2::set_error("dur_forward: tensor-core precision requires an sm_100 device");
        return ST2_ERR_UNSUPPORTED;
    }
    ST2_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "dur_forward: workspace must be 256-byte aligned");
    st2::g_launch_count = 0;
    int e = st2::dur_forward_impl(d, t_en, s, d_out, duration, B, L, precision, workspace, workspace_bytes, (cudaStream_t)stream,
                                  false, nullptr);
    d->last_launches = st2::g_launch_count;
    return e;
}

/* ---- TextEncoder: replaces models.py:238-285 ---- */
int st2_text_create(int32_t channels, int32_t kernel_size, int32_t depth, int32_t n_symbols, st2_decoder** out) {
    ST2_REQUIRE(out != nullptr, "text_create: null argument");
    ST2_REQUIRE(channels == 512, "text_create: channels must be 512 (got %d)", channels);
    ST2_REQUIRE(kernel_size >= 1 && kernel_size <= 15 && (kernel_size & 1) && depth >= 1 && depth <= 8 && n_symbols >= 1,
                "text_create: unsupported kernel_size / depth / n_symbols (%d, %d, %d)", kernel_size, depth, n_symbols);
    st2_decoder* d = new (std::nothrow) st2_decoder();
    ST2_REQUIRE(d != nullptr, "text_create: out of memory");
    memset(&d->cfg, 0, sizeof(d->cfg));
    d->cfg.variant = 3;
    d->cfg.dim_in = channels;
    d->cfg.style_dim = 4;
    d->te_depth = depth; d->te_kernel = kernel_size; d->te_symbols = n_symbols;
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaGetDeviceProperties(&prop, dev) == cudaSuccess)
        d->tc_ok = (prop.major == 10);
    *out = d;
    return ST2_OK;
}

int64_t st2_text_workspace_bytes(const st2_decoder* d, int32_t B, int32_t L, int32_t precision) {
    if (!d || !d->finalized || d->cfg.variant != 3 || B <= 0 || L <= 0) {
        st2::set_error("text_workspace_bytes: not a finalized text-encoder handle, or bad shape");
        return ST2_ERR_STATE;
    }
    int64_t peak = 0;
    int e = st2::text_forward_impl(const_cast<st2_decoder*>(d), nullptr, nullptr, B, L, precision, nullptr, 0, nullptr, true, &peak);
    if (e != ST2_OK) return e;
    return peak + 256;
}

int st2_text_forward(st2_decoder* d, const int64_t* tokens, float* out, int32_t B, int32_t L, int32_t precision, void* workspace,
                     int64_t workspace_bytes, void* stream) {
    ST2_REQUIRE(d != nullptr, "text_forward: null handle");
    if (!d->finalized || d->cfg.variant != 3) {
        st2::set_error("text_forward: not a finalized text-encoder handle (st2_text_create + st2_decoder_finalize)");
        return ST2_ERR_STATE;
    }
    ST2_REQUIRE(tokens && out && workspace, "text_forward: null tensor");
    ST2_REQUIRE(B > 0 && L >= 1, "text_forward: need B>0 and L>=1 (got B=%d L=%d)", B, L);
    ST2_REQUIRE(precision >= ST2_PREC_FP32 && precision <= ST2_PREC_FP16, "text_forward: bad precision %d", precision);
    if (precision != ST2_PREC_FP32 && !d->tc_ok) {
        st2::set_error("text_forward: tensor-core precision requires an sm_100 device");
        return ST2_ERR_UNSUPPORTED;
    }
    ST2_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "text_forward: workspace must be 256-byte aligned");
    st2::g_launch_count = 0;
    int e = st2::text_forward_impl(d, tokens, out, B, L, precision, workspace, workspace_bytes, (cudaStream_t)stream, false, nullptr);
    d->last_launches = st2::g_launch_count;
    return e;
}

int st2_decoder_set_tap(st2_decoder* d, const char* name, float* dst, int64_t capacity) {
    ST2_REQUIRE(d && name, "set_tap: bad argument");
    if (dst == nullptr) d->taps.erase(name);
    else d->taps[name] = st2::Tap{dst, capacity};
    return ST2_OK;
}

int st2_decoder_set_seed_buffer(st2_decoder* d, const uint64_t* dev_seed) {
    ST2_REQUIRE(d != nullptr, "set_seed_buffer: null handle");
    d->seed_dev = dev_seed;
    return ST2_OK;
}

int64_t st2_decoder_last_launch_count(const st2_decoder* d) { return d ? d->last_launches : 0; }

int st2_decoder_set_profiling(st2_decoder* d, int32_t enable) {
    ST2_REQUIRE(d != nullptr, "set_profiling: null handle");
    d->profiling = enable != 0;
    d->prof_recs.clear();
    return ST2_OK;
}

int st2_profile_num_categories(void) { return st2::PC_COUNT; }

const char* st2_profile_category_name(int32_t cat) {
    static const char* names[st2::PC_COUNT] = {"conv_tc", "conv_simt", "norm_stats", "norm_coef", "affine_act",
                                               "source", "post", "misc", "conv_fused", "conv_pipe", "lstm"};
    return (cat >= 0 && cat < st2::PC_COUNT) ? names[cat] : "";
}

int64_t st2_decoder_get_profile_launches(st2_decoder* d, int64_t max_n, int32_t* cat, float* ms, double* flops,
                                         double* bytes) {
    if (!d || !cat || !ms || !flops || !bytes) { st2::set_error("get_profile_launches: null argument"); return ST2_ERR_INVALID; }
    const size_t n = d->prof_recs.size();
    if (n == 0) return 0;
    if (d->prof_events.size() <= n) { st2::set_error("get_profile_launches: event pool inconsistent"); return ST2_ERR_STATE; }
    ST2_CUDA_CHECK(cudaEventSynchronize(d->prof_events[n]));
    const size_t m = n < (size_t)max_n ? n : (size_t)max_n;
    for (size_t i = 0; i < m; ++i) {
        float t = 0.f;
        ST2_CUDA_CHECK(cudaEventElapsedTime(&t, d->prof_events[i], d->prof_events[i + 1]));
        cat[i] = d->prof_recs[i].cat; ms[i] = t; flops[i] = d->prof_recs[i].flops; bytes[i] = d->prof_recs[i].bytes;
    }
    return (int64_t)m;
}

int st2_decoder_get_profile(st2_decoder* d, double* ms, int64_t* launches, double* flops, double* bytes) {
    ST2_REQUIRE(d && ms && launches && flops && bytes, "get_profile: null argument");
    for (int i = 0; i < st2::PC_COUNT; ++i) { ms[i] = 0; launches[i] = 0; flops[i] = 0; bytes[i] = 0; }
    const size_t n = d->prof_recs.size();
    if (n == 0) return ST2_OK;
    ST2_REQUIRE(d->prof_events.size() > n, "get_profile: event pool inconsistent");
    ST2_CUDA_CHECK(cudaEventSynchronize(d->prof_events[n]));
    for (size_t i = 0; i < n; ++i) {
        float t = 0.f;
        ST2_CUDA_CHECK(cudaEventElapsedTime(&t, d->prof_events[i], d->prof_events[i + 1]));
        const auto& r = d->prof_recs[i];
        ms[r.cat] += t;
        launches[r.cat] += 1;
        flops[r.cat] += r.flops;
        bytes[r.cat] += r.bytes;
    }
    return ST2_OK;
}

}  // extern "C"
