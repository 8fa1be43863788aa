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
