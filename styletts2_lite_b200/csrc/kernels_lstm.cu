// Recurrence of the bidirectional LSTM `shared` of ProsodyPredictor.F0Ntrain (reference models.py:407, :449;
// arithmetic: torch nn.LSTM, gates (i, f, g, o), zero initial state, batch_first).
//
// The input half of the gates (x W_ih^T + b_ih, one [B*T, 640] x [640, 2048] contraction for both directions) is
// a 1x1 convolution and runs on the conv kernels.  What is left is T strictly sequential steps of
//     gates = G[t] + h W_hh^T + b_hh ;  c = sig(f) c + sig(i) tanh(g) ;  h = sig(o) tanh(c)
// with a [4H, H] = 1 MB fp32 matrix per direction: too large for one SM's shared memory, latency bound if re-read
// from L2 every step.  One thread-block CLUSTER of 8 CTAs owns one (direction, group of 8 utterances): CTA r keeps
// the 4 x 32 gate columns of hidden units [32r, 32r+32) resident in shared memory (128 KB) for the whole sequence,
// computes its slice of the gates, updates its 32 cells and broadcasts the 32 new h values of each utterance to
// the h buffers of all 8 CTAs through distributed shared memory; one cluster barrier per time step.
#include "common.cuh"

#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace st2 {

constexpr int kLstmCluster = 8;      // CTAs per cluster
constexpr int kLstmBt = 8;           // utterances per cluster
constexpr int kLstmThreads = 256;    // 8 warps: warp = (gate, half of the utterances) in the gate phase, = utterance in the cell phase

__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

// G   [B][T][2][4H]   input half of the gates, direction-major inside a row (fwd 4H | rev 4H), gate order i,f,g,o
// whh [2][H][4H]      W_hh^T per direction
// bhh [2][4H]
// y   [B][T][2H]      forward h | reverse h   (channels-last, what the AdainResBlk1d stacks read)
template <int H>
__global__ void __cluster_dims__(kLstmCluster, 1, 1) __launch_bounds__(kLstmThreads, 1)
lstm_bidir_kernel(const float* __restrict__ G, const float* __restrict__ whh, const float* __restrict__ bhh,
                  float* __restrict__ y, int B, int T) {
    constexpr int UL = H / kLstmCluster;        // hidden units per CTA (32)
    constexpr int COLS = 4 * UL;                // gate columns per CTA (128)
    static_assert(UL == 32, "one warp lane per hidden unit of the CTA");
    extern __shared__ __align__(16) float smem[];
    float* Ws = smem;                                   // [H][COLS]
    float* hbuf = Ws + H * COLS;                        // [2][Bt][H]
    float* gs = hbuf + 2 * kLstmBt * H;                 // [4][Bt][UL]

    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int dir = blockIdx.y;
    const int b0 = blockIdx.z * kLstmBt;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // resident weights: Ws[k][g*32 + ul] = W_hh[g*H + 32*rank + ul][k]
    const float* wsrc = whh + (size_t)dir * H * 4 * H;
    for (int idx = tid; idx < H * COLS; idx += kLstmThreads) {
        const int k = idx / COLS, col = idx % COLS;
        Ws[idx] = wsrc[(size_t)k * 4 * H + (col / UL) * H + rank * UL + (col % UL)];
    }
    for (int idx = tid; idx < kLstmBt * H; idx += kLstmThreads) hbuf[idx] = 0.f;      // h_0 = 0
    // gate phase: this thread's column and utterances
    const int g = warp & 3, bh = (warp >> 2) * 4;
    const int gcol = g * H + rank * UL + lane;          // column inside one direction's 4H
    const float bias = bhh[dir * 4 * H + gcol];
    // cell phase: this thread's (utterance, unit)
    const int cb = warp;
    float c_state = 0.f;
    float* remote[kLstmCluster];
#pragma unroll
    for (int r = 0; r < kLstmCluster; ++r) remote[r] = cluster.map_shared_rank(hbuf, r);
    cluster.sync();                                     // every CTA of the cluster runs and has zeroed its h_0

    for (int step = 0; step < T; ++step) {
        const int t = dir ? T - 1 - step : step;
        const int cur = step & 1;
        // input half of the gates: issued first, consumed after the dot products
        float gin[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int b = b0 + bh + j;
            gin[j] = (b < B) ? __ldg(G + (((size_t)b * T + t) * 2 + dir) * 4 * H + gcol) : 0.f;
        }
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        const float* hc = hbuf + (size_t)cur * kLstmBt * H + bh * H;
        const float* wc = Ws + g * UL + lane;
#pragma unroll 4
        for (int k = 0; k < H; k += 4) {
            const float w0 = wc[(k + 0) * COLS], w1 = wc[(k + 1) * COLS], w2 = wc[(k + 2) * COLS], w3 = wc[(k + 3) * COLS];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 hv = *reinterpret_cast<const float4*>(hc + j * H + k);      // warp-uniform: broadcast
                acc[j] = fmaf(w0, hv.x, acc[j]);
                acc[j] = fmaf(w1, hv.y, acc[j]);
                acc[j] = fmaf(w2, hv.z, acc[j]);
                acc[j] = fmaf(w3, hv.w, acc[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) gs[(g * kLstmBt + bh + j) * UL + lane] = gin[j] + (acc[j] + bias);
        __syncthreads();
        // cell update of (utterance cb, unit rank*32 + lane)
        const float gi = gs[(0 * kLstmBt + cb) * UL + lane], gf = gs[(1 * kLstmBt + cb) * UL + lane];
        const float gg = gs[(2 * kLstmBt + cb) * UL + lane], go = gs[(3 * kLstmBt + cb) * UL + lane];
        c_state = sigmoid_f(gf) * c_state + sigmoid_f(gi) * tanhf(gg);
        const float h = sigmoid_f(go) * tanhf(c_state);
        const int hoff = ((cur ^ 1) * kLstmBt + cb) * H + rank * UL + lane;
#pragma unroll
        for (int r = 0; r < kLstmCluster; ++r) remote[r][hoff] = h;        // 128 B per warp and destination CTA
        if (b0 + cb < B) y[((size_t)(b0 + cb) * T + t) * 2 * H + dir * H + rank * UL + lane] = h;
        cluster.sync();      // h_{t} visible everywhere; everyone is done reading h_{t-1} and the gate exchange buffer
    }
}

int launch_lstm_bidir(const float* G, const float* whh, const float* bhh, float* y, int B, int T, int H, cudaStream_t st) {
    ST2_REQUIRE(H == 256, "lstm: hidden size %d is not supported (d_hid must be 512)", H);
    ST2_REQUIRE(B > 0 && T > 0, "lstm: bad shape B=%d T=%d", B, T);
    constexpr int HH = 256;
    const size_t smem = ((size_t)HH * 4 * (HH / kLstmCluster) + 2 * kLstmBt * HH + 4 * kLstmBt * (HH / kLstmCluster)) * sizeof(float);
    ST2_CUDA_CHECK(cudaFuncSetAttribute(lstm_bidir_kernel<HH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(kLstmCluster, 2, cdiv(B, kLstmBt));
    lstm_bidir_kernel<HH><<<grid, kLstmThreads, smem, st>>>(G, whh, bhh, y, B, T);
    ST2_LAUNCH_CHECK();
    return ST2_OK;
}

}  // namespace st2

// ------------------------------------------------------------------------------------------------------------------
// Duration half of the predictor (reference models.py:372-392, :485-520, inference.py:242-245)
// ------------------------------------------------------------------------------------------------------------------
namespace st2 {

// x[b][l][C .. C+S) = s[b][0 .. S): the style columns DurationEncoder concatenates to every token (models.py:489-490, :499)
__global__ void concat_style_kernel(float* __restrict__ x, int ld, int C, const float* __restrict__ s, int S, int64_t rows, int L) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * S) return;
    const int64_t row = i / S;
    const int c = (int)(i - row * S);
    x[row * ld + C + c] = s[(row / L) * S + c];
}

// AdaLayerNorm (models.py:372-392) on channels-last x [B][L][C]: LayerNorm over C (biased variance, eps 1e-5), then
// (1 + gamma) * xhat + beta with gamma | beta = h[b][h_off .. h_off + 2C); writes y[row][0..C) with pitch ld_y.
// One warp per token; C = 512 -> 16 values per lane kept in registers, mean and variance by warp shuffles (two pass).
template <int C>
__global__ void ada_layer_norm_kernel(const float* __restrict__ x, const float* __restrict__ h, int ld_h, int h_off,
                                      float* __restrict__ y, int ld_y, int64_t rows, int L) {
    constexpr int PER = C / 32;
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* xr = x + row * C;
    float v[PER];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < PER / 4; ++i) {
        const float4 t = *reinterpret_cast<const float4*>(xr + (i * 32 + lane) * 4);
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
        sum += (t.x + t.y) + (t.z + t.w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * (1.f / C);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) { const float dlt = v[i] - mean; sq = fmaf(dlt, dlt, sq); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq * (1.f / C) + 1e-5f);
    const float* hb = h + (row / L) * ld_h + h_off;
    float* yr = y + row * ld_y;
#pragma unroll
    for (int i = 0; i < PER / 4; ++i) {
        const int c = (i * 32 + lane) * 4;
        const float4 g = *reinterpret_cast<const float4*>(hb + c), b = *reinterpret_cast<const float4*>(hb + C + c);
        float4 o;
        o.x = fmaf(1.f + g.x, (v[4 * i] - mean) * rstd, b.x);
        o.y = fmaf(1.f + g.y, (v[4 * i + 1] - mean) * rstd, b.y);
        o.z = fmaf(1.f + g.z, (v[4 * i + 2] - mean) * rstd, b.z);
        o.w = fmaf(1.f + g.w, (v[4 * i + 3] - mean) * rstd, b.w);
        *reinterpret_cast<float4*>(yr + c) = o;
    }
}

// duration[row] = sum_j sigmoid(x[row] . W[j] + bias[j])   (duration_proj + torch.sigmoid(...).sum(-1), inference.py:244-245)
// One warp per token: x row in registers, one warp-reduced dot product per output bin.
template <int C>
__global__ void duration_head_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ bias,
                                     float* __restrict__ duration, int64_t rows, int nbins) {
    constexpr int PER = C / 32;
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    float v[PER];
#pragma unroll
    for (int i = 0; i < PER / 4; ++i) {
        const float4 t = *reinterpret_cast<const float4*>(x + row * C + (i * 32 + lane) * 4);
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
    float total = 0.f;
    for (int j = 0; j < nbins; ++j) {
        const float* wj = W + (size_t)j * C;
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < PER / 4; ++i) {
            const float4 w = __ldg(reinterpret_cast<const float4*>(wj + (i * 32 + lane) * 4));
            acc = fmaf(v[4 * i], w.x, acc); acc = fmaf(v[4 * i + 1], w.y, acc);
            acc = fmaf(v[4 * i + 2], w.z, acc); acc = fmaf(v[4 * i + 3], w.w, acc);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        total += sigmoid_f(acc + __ldg(bias + j));
    }
    if (lane == 0) duration[row] = total;
}

int launch_concat_style(float* x, int ld, int C, const float* s, int S, int B, int L, cudaStream_t st) {
    const int64_t n = (int64_t)B * L * S;
    concat_style_kernel<<<cdiv(n, 256), 256, 0, st>>>(x, ld, C, s, S, (int64_t)B * L, L);
    ST2_LAUNCH_CHECK();
    return ST2_OK;
}

int launch_ada_layer_norm(const float* x, const float* h, int ld_h, int h_off, float* y, int ld_y, int B, int L, int C,
                          cudaStream_t st) {
    ST2_REQUIRE(C == 512 && ld_y % 4 == 0 && ld_h % 4 == 0 && h_off % 4 == 0, "ada_layer_norm: needs 512 channels (got %d)", C);
    const int64_t rows = (int64_t)B * L;
    ada_layer_norm_kernel<512><<<cdiv(rows, 8), 256, 0, st>>>(x, h, ld_h, h_off, y, ld_y, rows, L);
    ST2_LAUNCH_CHECK();
    return ST2_OK;
}

int launch_duration_head(const float* x, const float* W, const float* bias, float* duration, int B, int L, int C, int nbins,
                         cudaStream_t st) {
    ST2_REQUIRE(C == 512 && nbins > 0, "duration_head: needs 512 channels (got %d)", C);
    const int64_t rows = (int64_t)B * L;
    duration_head_kernel<512><<<cdiv(rows, 8), 256, 0, st>>>(x, W, bias, duration, rows, nbins);
    ST2_LAUNCH_CHECK();
    return ST2_OK;
}

}  // namespace st2
