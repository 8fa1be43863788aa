// The modules in front of the Decoder: ProsodyPredictor.F0Ntrain (models.py:448-461), the predictor's
// duration half (inference.py:242-245; models.py:468-520, :404-405) and the TextEncoder (models.py:238-285) -- weight packing,
// forward programs on the shared executor (program.cuh) and their C ABI (st2_f0n_*, st2_dur_*, st2_text_* of
// include/st2_b200.h).  The handle type, st2_decoder_set_weight / _finalize / _set_tap / profile calls live in decoder.cu.
#include "program.cuh"

namespace st2 {

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
void pack_predictor(st2_decoder* d, Packer& P) {
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
void pack_text_encoder(st2_decoder* d, Packer& P) {
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
    if (E.live()) E.chk(launch_lstm_bidir(G, w.whh, w.bhh, y, B, T, H, E.st, E.lengths));
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

// inference.py:242-245: d = predictor.text_encoder(t_en, s, lengths, mask) (DurationEncoder.forward, models.py:485-520),
// x = predictor.lstm(d), duration = sigmoid(duration_proj(x)).sum(-1).
// t_en [B, d_hid, L], s [B, style] -> d_out [B, L, d_hid+style] (the reference's layout of `d`), duration [B, L].
// lengths (device, [B], may be null): the padded batch of ProsodyPredictor.forward (models.py:422-442) -- rows behind an
// utterance are zeroed where DurationEncoder.forward masks (models.py:491, :500), every LSTM is the packed one (models.py:503-509,
// :426-435), and duration at a padded token is what duration_proj makes of the zero row there, as in the reference.
static int dur_forward_impl(st2_decoder* d, const float* t_en, const float* s, const int32_t* lengths, float* d_out, float* duration,
                            int B, int L, int prec, void* ws, int64_t ws_bytes, cudaStream_t st, bool dry, int64_t* peak_out) {
    const st2_config& c = d->cfg;
    const int dh = c.dim_in, H = dh / 2, I = dh + c.style_dim;
    Exec E{d, st, dry, prec, B, (char*)ws, ws_bytes};
    E.lengths = lengths;
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
        if (lengths) E.chk(launch_mask_rows(xa, I, I, lengths, B, L, st));         // models.py:491
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
            if (lengths) E.chk(launch_mask_rows(dst, I, I, lengths, B, L, st));    // models.py:500
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

// TextEncoder.forward(x, input_lengths, m) (models.py:258-285): tokens [B, L] int64 -> out [B, channels, L].
// lengths (device, [B], may be null = mask all False): tokens behind an utterance are zeroed after the embedding and after every
// cnn block (models.py:262, :266), the LSTM is the packed one (models.py:270-277), padded columns of `out` are zero (:279-283).
static int text_forward_impl(st2_decoder* d, const int64_t* tokens, const int32_t* lengths, float* out, int B, int L, int prec,
                             void* ws, int64_t ws_bytes, cudaStream_t st, bool dry, int64_t* peak_out) {
    const int C = d->cfg.dim_in, H = C / 2;
    Exec E{d, st, dry, prec, B, (char*)ws, ws_bytes};
    E.lengths = lengths;
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
        if (lengths) E.chk(launch_mask_rows(xa, C, C, lengths, B, L, st));                       // models.py:262
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
        if (E.live()) {
            E.chk(launch_layer_norm_lrelu(xb, d->te_gamma[i], d->te_gamma[i] + C, 0.2f, xa, B, L, C, st));
            if (lengths) E.chk(launch_mask_rows(xa, C, C, lengths, B, L, st));                   // models.py:266
        }
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

extern "C" {

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
    int e = st2::dur_forward_impl(const_cast<st2_decoder*>(d), nullptr, nullptr, nullptr, nullptr, nullptr, B, L, precision,
                                  nullptr, 0, nullptr, true, &peak);
    if (e != ST2_OK) return e;
    return peak + 256;
}

int st2_dur_forward_ragged(st2_decoder* d, const float* t_en, const float* s, const int32_t* lengths, float* d_out, float* duration,
                           int32_t B, int32_t L, int32_t precision, void* workspace, int64_t workspace_bytes, void* stream) {
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
    int e = st2::dur_forward_impl(d, t_en, s, lengths, d_out, duration, B, L, precision, workspace, workspace_bytes,
                                  (cudaStream_t)stream, false, nullptr);
    d->last_launches = st2::g_launch_count;
    return e;
}

int st2_dur_forward(st2_decoder* d, const float* t_en, const float* s, float* d_out, float* duration, int32_t B, int32_t L,
                    int32_t precision, void* workspace, int64_t workspace_bytes, void* stream) {
    return st2_dur_forward_ragged(d, t_en, s, nullptr, d_out, duration, B, L, precision, workspace, workspace_bytes, stream);
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
    int e = st2::text_forward_impl(const_cast<st2_decoder*>(d), nullptr, nullptr, nullptr, B, L, precision, nullptr, 0, nullptr, true,
                                   &peak);
    if (e != ST2_OK) return e;
    return peak + 256;
}

int st2_text_forward_ragged(st2_decoder* d, const int64_t* tokens, const int32_t* lengths, float* out, int32_t B, int32_t L,
                            int32_t precision, void* workspace, int64_t workspace_bytes, void* stream) {
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
    int e = st2::text_forward_impl(d, tokens, lengths, out, B, L, precision, workspace, workspace_bytes, (cudaStream_t)stream, false,
                                   nullptr);
    d->last_launches = st2::g_launch_count;
    return e;
}

int st2_text_forward(st2_decoder* d, const int64_t* tokens, float* out, int32_t B, int32_t L, int32_t precision, void* workspace,
                     int64_t workspace_bytes, void* stream) {
    return st2_text_forward_ragged(d, tokens, nullptr, out, B, L, precision, workspace, workspace_bytes, stream);
}

}  // extern "C"
