// Length regulator: the duration -> frame expansion of inference.py:257-268.
//
// The reference rounds the predicted durations (torch.round = half-to-even, clamp(min=1)),
// fills a one-hot alignment matrix A[L,F] in a Python loop and multiplies: en = d^T @ A,
// asr = t_en @ A.  Each output column has exactly one 1 in A, so the product is a column
// gather out[:, f] = src[:, tok(f)] (x*1 + sum of zeros == x): integer / copy work, bit-exact,
// HBM-bound (4*C*F bytes read + 4*C*F written per utterance).
#include "common.cuh"

namespace st2 {

__global__ void round_durations_kernel(const float* __restrict__ duration, const int32_t* __restrict__ n_tokens,
                                       int32_t* __restrict__ dur, int32_t* __restrict__ total, int L) {
    const int b = blockIdx.x;
    const int nt = n_tokens ? min(n_tokens[b], L) : L;
    int local = 0;
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
        int d = 0;
        if (l < nt) {
            float r = rintf(duration[(size_t)b * L + l]);      // round half to even
            r = fmaxf(r, 1.0f);                                 // clamp(min=1)
            r = fminf(r, 1.0e6f);
            d = (int)r;
        }
        dur[(size_t)b * L + l] = d;
        local += d;
    }
    __shared__ int red[256];
    red[threadIdx.x] = local;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) total[b] = red[0];
}

int launch_round_durations(const float* duration, const int32_t* n_tokens, int32_t* dur, int32_t* total, int B,
                           int L, cudaStream_t st) {
    if (B <= 0) return ST2_OK;
    round_durations_kernel<<<B, 256, 0, st>>>(duration, n_tokens, dur, total, L);
    ST2_LAUNCH_CHECK();
    return ST2_OK;
}

// ---- duration smoothing (inference.py:248-255) ------------------------------------------------------------------------
// Between the sigmoid-sum of the duration head and the rounding the reference (a) mixes every duration with a draw from
// N(mean, std) of the sentence's own durations -- mean replaced by the previous split's mean duration when there is one -- with
// weight t (inference.py:248-252), (b) replaces |z| > 3 outliers of duration[1:-2] by mean +- 3 * 0.95 * std of that slice
// (inference.py:253, :134-148), (c) divides by the speed (inference.py:255) and returns the mean duration for the next split
// (inference.py:272).  All statistics are per sentence, torch.std is the unbiased one.  One CTA per utterance; sums in fp64
// with a fixed-order tree (deterministic), the element-wise arithmetic in fp32 with the reference's operation order (no FMA
// contraction).  The normal draw comes from a caller-provided N(0, 1) tape z: dur_stats = z * std + mean (what
// torch.Tensor.normal_(mean, std) computes from its own standard-normal draw).
static constexpr int kSmThreads = 128;

__device__ double smooth_block_sum(double v, double* red) {
    __syncthreads();                                       // red may still be read from the previous reduction
    red[threadIdx.x] = v;
    __syncthreads();
    for (int o = kSmThreads / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    return red[0];
}

// unbiased mean / std of v[lo, hi) (hi - lo >= 1; std is NaN for a single element, like torch)
__device__ void smooth_mean_std(const float* v, int lo, int hi, double* red, float* mean_out, float* std_out) {
    double s = 0;
    for (int i = lo + threadIdx.x; i < hi; i += kSmThreads) s += (double)v[i];
    const double mean = smooth_block_sum(s, red) / (double)(hi - lo);
    double q = 0;
    for (int i = lo + threadIdx.x; i < hi; i += kSmThreads) { const double dlt = (double)v[i] - mean; q += dlt * dlt; }
    const double var = smooth_block_sum(q, red) / (double)(hi - lo - 1);      // n = 1: 0 / 0 = NaN
    *mean_out = (float)mean;
    *std_out = (float)sqrt(var);
}

__global__ void __launch_bounds__(kSmThreads)
smooth_durations_kernel(const float* __restrict__ duration, const int32_t* __restrict__ n_tokens, const float* __restrict__ z,
                        const float* __restrict__ prev_mean, float t, float speed, float* __restrict__ out,
                        float* __restrict__ mean_out, int B, int L, int chain) {
    // chain: one CTA walks the utterances in order and hands the mean duration of sentence b - 1 to sentence b as its previous mean
    // (the loop of StyleTTS2.generate, inference.py:312-313); prev_mean[0] seeds the first one.  Otherwise one CTA per utterance.
    __shared__ double red[kSmThreads];
    float carried = (chain && prev_mean) ? prev_mean[0] : 0.f;
    for (int b = chain ? 0 : (int)blockIdx.x; b < (chain ? B : (int)blockIdx.x + 1); ++b) {
    const int n = n_tokens ? min(max(n_tokens[b], 0), L) : L;
    const float* x = duration + (size_t)b * L;
    float* y = out + (size_t)b * L;
    for (int i = n + threadIdx.x; i < L; i += kSmThreads) y[i] = 0.f;          // padded tokens
    if (n == 0) {
        if (threadIdx.x == 0 && mean_out) mean_out[b] = 0.f;
        continue;
    }
    float mean, sd;
    smooth_mean_std(x, 0, n, red, &mean, &sd);
    const float prev = chain ? carried : (prev_mean ? prev_mean[b] : 0.f);
    const float mu = prev != 0.f ? prev : mean;                                  // inference.py:248-251
    const float c1 = (float)(1.0 - (double)t);                                   // Python computes 1 - t in double
    for (int i = threadIdx.x; i < n; i += kSmThreads) {
        const float stats = z ? __fadd_rn(__fmul_rn(z[(size_t)b * L + i], sd), mu) : mu;
        y[i] = __fadd_rn(__fmul_rn(x[i], c1), __fmul_rn(stats, t));              // inference.py:252
    }
    __syncthreads();
    if (n - 3 >= 1) {                                                            // duration[:, 1:-2], inference.py:253
        float m2, sd2;
        smooth_mean_std(y, 1, n - 2, red, &m2, &sd2);
        const float repl = __fmul_rn(__fmul_rn(3.0f, sd2), 0.95f);               // threshold * std * factor, inference.py:143
        for (int i = 1 + threadIdx.x; i < n - 2; i += kSmThreads) {
            const float dlt = __fsub_rn(y[i], m2);
            const float zz = __fdiv_rn(dlt, sd2);
            if (fabsf(zz) > 3.0f) {                                              // NaN (single element): never
                const float sg = dlt > 0.f ? 1.f : (dlt < 0.f ? -1.f : 0.f);
                y[i] = __fadd_rn(m2, __fmul_rn(sg, repl));
            }
        }
        __syncthreads();
    }
    double s = 0;
    for (int i = threadIdx.x; i < n; i += kSmThreads) {
        const float v = __fdiv_rn(y[i], speed);                                  // inference.py:255
        y[i] = v;
        s += (double)v;
    }
    const double tot = smooth_block_sum(s, red);
    carried = (float)(tot / (double)n);                                          // inference.py:272 -> prev_d_mean of the next sentence
    if (threadIdx.x == 0 && mean_out) mean_out[b] = carried;
    __syncthreads();                                                             // y / red are reused by the next utterance
    }
}

int launch_smooth_durations(const float* duration, const int32_t* n_tokens, const float* z, const float* prev_mean, float t,
                            float speed, float* out, float* mean_out, int B, int L, cudaStream_t st, int chain) {
    if (B <= 0 || L <= 0) return ST2_OK;
    smooth_durations_kernel<<<chain ? 1 : B, kSmThreads, 0, st>>>(duration, n_tokens, z, prev_mean, t, speed, out, mean_out, B, L, chain);
    ST2_LAUNCH_CHECK();
    return ST2_OK;
}

static constexpr int kLrThreads = 256;
static constexpr int kLrTile = 128;      // frames per CTA

// One CTA: utterance b, frames [f0, f0+kLrTile).  Inclusive cumsum of dur[b,:] in shared
// memory, one binary search per frame, then the copy over all C channels.
template <bool CL>
__global__ void __launch_bounds__(kLrThreads)
length_regulate_kernel(const float* __restrict__ src, const int32_t* __restrict__ dur, float* __restrict__ out, int C,
                       int L, int F) {
    extern __shared__ int sm[];
    int* cum = sm;                 // [L] inclusive prefix sums
    int* part = sm + L;            // [kLrThreads]
    int* tok = part + kLrThreads;  // [kLrTile]
    const int b = blockIdx.y;
    const int f0 = blockIdx.x * kLrTile;
    const int chunk = (L + kLrThreads - 1) / kLrThreads;
    const int l0 = threadIdx.x * chunk, l1 = min(L, l0 + chunk);
    int s = 0;
    for (int l = l0; l < l1; ++l) s += max(dur[(size_t)b * L + l], 0);
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int i = 0; i < kLrThreads; ++i) {
            int v = part[i];
            part[i] = run;
            run += v;
        }
    }
    __syncthreads();
    s = part[threadIdx.x];
    for (int l = l0; l < l1; ++l) {
        s += max(dur[(size_t)b * L + l], 0);
        cum[l] = s;
    }
    __syncthreads();
    const int total = L > 0 ? cum[L - 1] : 0;
    if (threadIdx.x < kLrTile) {
        const int f = f0 + threadIdx.x;
        int t = -1;
        if (f < total) {           // first l with cum[l] > f
            int lo = 0, hi = L - 1;
            while (lo < hi) {
                int mid = (lo + hi) >> 1;
                if (cum[mid] > f) hi = mid; else lo = mid + 1;
            }
            t = lo;
        }
        tok[threadIdx.x] = t;
    }
    __syncthreads();
    const int nf = min(kLrTile, F - f0);
    const float* sb = src + (size_t)b * C * L;
    if (CL) {      // out [B,F,C]
        for (int i = threadIdx.x; i < nf * C; i += kLrThreads) {
            int fl = i / C, c = i - fl * C;
            int t = tok[fl];
            out[((size_t)b * F + f0 + fl) * C + c] = t >= 0 ? __ldg(sb + (size_t)c * L + t) : 0.f;
        }
    } else {       // out [B,C,F]
        for (int i = threadIdx.x; i < nf * C; i += kLrThreads) {
            int c = i / nf, fl = i - c * nf;
            int t = tok[fl];
            out[((size_t)b * C + c) * F + f0 + fl] = t >= 0 ? __ldg(sb + (size_t)c * L + t) : 0.f;
        }
    }
}

int launch_length_regulate(const float* src, const int32_t* dur, float* out, int B, int C, int L, int F,
                           int channels_last, cudaStream_t st) {
    if (B <= 0 || F <= 0 || C <= 0) return ST2_OK;
    ST2_REQUIRE(L >= 0 && L <= 8192, "length_regulate: L=%d out of range (max 8192 tokens)", L);
    size_t smem = ((size_t)L + kLrThreads + kLrTile) * sizeof(int);
    dim3 grid(cdiv(F, kLrTile), B);
    if (channels_last)
        length_regulate_kernel<true><<<grid, kLrThreads, smem, st>>>(src, dur, out, C, L, F);
    else
        length_regulate_kernel<false><<<grid, kLrThreads, smem, st>>>(src, dur, out, C, L, F);
    ST2_LAUNCH_CHECK();
    return ST2_OK;
}

}  // namespace st2
