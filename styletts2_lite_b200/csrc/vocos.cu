// Vocos decoder variant (reference Modules/vocos.py): the kernels of its Generator that the shared
// executor does not already have.  The ConvNeXt blocks (vocos.py:27-69) are
//     depthwise Conv1d(k = 7, pad 3)                       -> dwconv7_kernel (channels-last, HBM-bound)
//     AdaIN1d (InstanceNorm over time + style affine)      -> in_stats / adain_coef / affine_act of kernels_norm.cu
//     Linear(dim, 3 dim) -> GELU -> Linear(3 dim, dim)     -> 1x1 convolutions on conv_tc (tcgen05) / conv_simt, GELU as the
//                                                             operand transform of the second one (ACT_GELU)
//     gamma * x + residual                                 -> folded into the second Linear's weights, residual in its epilogue
// and the ISTFTHead (vocos.py:235-296, :165-232) is
//     Linear(dim, n_fft + 2)                               -> 1x1 convolution, columns padded to a multiple of 64
//     mag = min(exp(m), 100), (re, im) = mag (cos p, sin p) -> vocos_spec_kernel, writes the GEMM operand (fp32 or fp16)
//     irfft(n_fft) * window                                -> ONE GEMM with a real basis [2 bins, n_fft] that has the window and
//                                                             the 1/n_fft folded in (vocos_basis_kernel builds it at pack time);
//                                                             C2R semantics: the imaginary parts of DC and Nyquist are ignored
//     overlap-add, trim (n_fft - hop)/2, / window envelope -> vocos_ola_kernel (each output sample: n_fft / hop = 4 frames)
#include "common.cuh"

namespace st2 {

// y[b][t][c] = bias[c] + sum_k w7[k][c] * x[b][t + k - 3][c]   (zero padding); one thread per (t, channel quad)
__global__ void __launch_bounds__(256)
dwconv7_kernel(const float* __restrict__ x, const float* __restrict__ w7, const float* __restrict__ bias, float* __restrict__ y,
               int T, int C) {
    const int cq = C >> 2;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (i >= (int64_t)T * cq) return;
    const int t = (int)(i / cq), c = (int)(i - (int64_t)t * cq) * 4;
    const float* xb = x + (size_t)b * T * C + c;
    float4 acc = __ldg(reinterpret_cast<const float4*>(bias + c));
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        const int tt = t + k - 3;
        if (tt < 0 || tt >= T) continue;
        const float4 v = __ldg(reinterpret_cast<const float4*>(xb + (size_t)tt * C));
        const float4 w = __ldg(reinterpret_cast<const float4*>(w7 + (size_t)k * C + c));
        acc.x = fmaf(w.x, v.x, acc.x); acc.y = fmaf(w.y, v.y, acc.y);
        acc.z = fmaf(w.z, v.z, acc.z); acc.w = fmaf(w.w, v.w, acc.w);
    }
    *reinterpret_cast<float4*>(y + ((size_t)b * T + t) * C + c) = acc;
}

int launch_dwconv7(const float* x, const float* w7, const float* bias, float* y, int B, int T, int C, cudaStream_t st) {
    ST2_REQUIRE(C % 4 == 0 && B > 0 && T > 0, "dwconv7: bad shape");
    const int64_t n = (int64_t)T * (C / 4);
    dwconv7_kernel<<<dim3((unsigned)cdiv(n, 256), B), 256, 0, st>>>(x, w7, bias, y, T, C);
    ST2_LAUNCH_CHECK();
    return ST2_OK;
}

// w[r][c] *= scale[c] (r < rows, c < cols, pitch ld), bias[c] *= scale[c]: a per-output-channel layer scale folded into a Linear
__global__ void scale_cols_kernel(float* __restrict__ w, float* __restrict__ bias, int rows, int ld, int cols,
                                  const float* __restrict__ scale) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)(rows + 1) * cols) return;
    const int r = (int)(i / cols), c = (int)(i - (int64_t)r * cols);
    if (r < rows) w[(size_t)r * ld + c] *= scale[c];
    else if (bias != nullptr) bias[c] *= scale[c];
}

int launch_scale_cols(float* w, float* bias, int rows, int ld, int cols, const float* scale, cudaStream_t st) {
    const int64_t n = (int64_t)(rows + 1) * cols;
    scale_cols_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(w, bias, rows, ld, cols, scale);
    ST2_LAUNCH_CHECK();
    return ST2_OK;
}

// basis[j][n], j < k_pad operand columns (re_0 .. re_{N/2}, im_0 .. im_{N/2}, zero padding), n < N:
//   frame[n] * w[n] = sum_j A[j] basis[j][n]   with   irfft(S)[n] = (1/N) (Re S_0 + (-1)^n Re S_{N/2} + 2 sum_{0<k<N/2} (Re S_k cos - Im S_k sin)(2 pi k n / N))
__global__ void vocos_basis_kernel(const float* __restrict__ window, float* __restrict__ basis, int N, int k_pad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)k_pad * N) return;
    const int j = (int)(i / N), n = (int)(i - (int64_t)j * N);
    const int bins = N / 2 + 1;
    double v = 0.0;
    if (j < 2 * bins) {
        const bool im = j >= bins;
        const int k = im ? j - bins : j;
        const bool edge = (k == 0 || k == N / 2);
        const int64_t kn = ((int64_t)k * n) % N;                    // exact argument reduction
        double sn, cs;
        sincospi(2.0 * (double)kn / (double)N, &sn, &cs);
        if (!im) v = (edge ? 1.0 : 2.0) * cs;
        else v = edge ? 0.0 : -2.0 * sn;
        v *= (double)window[n] / (double)N;
    }
    basis[i] = (float)v;
}

int launch_vocos_basis(const float* window, float* basis, int n_fft, int k_pad, cudaStream_t st) {
    const int64_t n = (int64_t)k_pad * n_fft;
    vocos_basis_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(window, basis, n_fft, k_pad);
    ST2_LAUNCH_CHECK();
    return ST2_OK;
}

// o [rows][ld_o]: log-magnitudes in columns [0, bins), phases in [bins, 2 bins)  (x.chunk(2, dim=1), vocos.py:281)
// a [rows][ld_a]: re in [0, bins), im in [bins, 2 bins), zeros up to ld_a
template <int DT>
__global__ void __launch_bounds__(256)
vocos_spec_kernel(const float* __restrict__ o, int ld_o, void* __restrict__ a, int ld_a, int bins) {
    const int64_t row = blockIdx.x;
    const float* orow = o + row * ld_o;
    for (int c = threadIdx.x; c < ld_a; c += blockDim.x) {
        float v = 0.f;
        if (c < 2 * bins) {
            const int k = c < bins ? c : c - bins;
            const float mag = fminf(expf(orow[k]), 100.f);          // torch.exp, torch.clip(max=1e2)  (vocos.py:282-283)
            const float p = orow[bins + k];
            v = mag * (c < bins ? cosf(p) : sinf(p));               // vocos.py:285-292
        }
        if (DT == DT_F32) reinterpret_cast<float*>(a)[row * ld_a + c] = v;
        else if (DT == DT_F16) reinterpret_cast<__half*>(a)[row * ld_a + c] = __float2half_rn(v);
        else reinterpret_cast<__nv_bfloat16*>(a)[row * ld_a + c] = __float2bfloat16_rn(v);
    }
}

int launch_vocos_spec(const float* o, int ld_o, void* a, int ld_a, int out_dtype, int64_t rows, int bins, cudaStream_t st) {
    ST2_REQUIRE(rows > 0 && 2 * bins <= ld_o && 2 * bins <= ld_a, "vocos_spec: bad shape");
    if (out_dtype == DT_F32) vocos_spec_kernel<DT_F32><<<(unsigned)rows, 256, 0, st>>>(o, ld_o, a, ld_a, bins);
    else if (out_dtype == DT_F16) vocos_spec_kernel<DT_F16><<<(unsigned)rows, 256, 0, st>>>(o, ld_o, a, ld_a, bins);
    else vocos_spec_kernel<DT_BF16><<<(unsigned)rows, 256, 0, st>>>(o, ld_o, a, ld_a, bins);
    ST2_LAUNCH_CHECK();
    return ST2_OK;
}

// out[b][s] = (sum_t frames[b][t][s + pad - t hop]) / (sum_t w^2[s + pad - t hop]),  pad = (N - hop) / 2, over the frames that
// cover the sample (vocos.py:218-230: fold, [pad:-pad], divide by the folded squared window)
__global__ void __launch_bounds__(256)
vocos_ola_kernel(const float* __restrict__ frames, const float* __restrict__ window, float* __restrict__ out, int T, int N,
                 int hop) {
    const int64_t S = (int64_t)T * hop;
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (s >= S) return;
    const int64_t pos = s + (N - hop) / 2;
    int t_hi = (int)(pos / hop);
    if (t_hi > T - 1) t_hi = T - 1;
    int t_lo = (int)((pos - N + hop) / hop);                        // ceil((pos - N + 1) / hop) for pos - N + 1 >= 0
    if (pos - N + 1 <= 0) t_lo = 0;
    float y = 0.f, env = 0.f;
    for (int t = t_hi; t >= t_lo; --t) {                            // col2im adds the smallest in-frame offset first
        const int n = (int)(pos - (int64_t)t * hop);
        const float w = __ldg(window + n);
        y += frames[((size_t)b * T + t) * N + n];
        env = fmaf(w, w, env);
    }
    out[(size_t)b * S + s] = y / env;
}

int launch_vocos_ola(const float* frames, const float* window, float* out, int B, int T, int n_fft, int hop, cudaStream_t st) {
    ST2_REQUIRE(B > 0 && T > 0 && hop > 0 && n_fft >= hop && (n_fft - hop) % 2 == 0, "vocos_ola: bad shape");
    const int64_t S = (int64_t)T * hop;
    vocos_ola_kernel<<<dim3((unsigned)cdiv(S, 256), B), 256, 0, st>>>(frames, window, out, T, n_fft, hop);
    ST2_LAUNCH_CHECK();
    return ST2_OK;
}

}  // namespace st2
