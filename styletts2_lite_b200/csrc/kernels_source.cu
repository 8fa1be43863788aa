// NSF harmonic source (SineGen / SourceModuleHnNSF, Modules/hifigan.py:82-268) and the
// CustomSTFT front / iSTFT head of the istftnet variant (Modules/istftnet.py:111-301).
//
// The SineGen phase must match the CPU reference BIT FOR BIT (a 1-ulp phase change moves the
// waveform by ~0.02), so every operation of that chain is an explicitly rounded
// intrinsic (__fmul_rn / __fdiv_rn / __fmaf_rn: never contracted, never fast-math) in exactly
// the reference's order, and the frame cumsum is a sequential double accumulation like ATen's
// CPU cumsum.  This file is compiled without --use_fast_math.
#include "common.cuh"

namespace st2 {

static constexpr int kH = 9;               // harmonic_num + 1 (hifigan.py:106,282)
static constexpr float kSr = 24000.f;      // hifigan.py:280
static constexpr float kPi32 = 3.14159274101257324f;   // float(np.pi)

// torch `%` on floats == python modulo: fmod, then shift into the divisor's sign
__device__ __forceinline__ float remainder1(float q) {
    float r = fmodf(q, 1.0f);
    if (r != 0.f && r < 0.f) r = __fadd_rn(r, 1.0f);
    return r;
}

// frames[b][j][h] = fp32( fp32( fp32(cs*2) * pi32 ) * scale ),  cs = fp32( sum_{i<=j} (double) rad_i )
// rad = ((f0*(h+1)) / 24000) % 1     (hifigan.py:199,123,145-147,154-155)
// One thread per (b,h): a strictly sequential double accumulation keeps the reference's
// summation order; 9*B independent chains of <= 4800 adds cost ~20 us.
__global__ void sinegen_frames_kernel(const float* __restrict__ f0, float* __restrict__ frames, int B, int L2,
                                      float scale) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * kH) return;
    const int b = i / kH, h = i - b * kH;
    const float mult = (float)(h + 1);
    const float* fb = f0 + (size_t)b * L2;
    float* out = frames + (size_t)b * L2 * kH + h;
    double acc = 0.0;
    for (int j = 0; j < L2; ++j) {
        float fn = __fmul_rn(__ldg(fb + j), mult);
        float rad = remainder1(__fdiv_rn(fn, kSr));
        acc = __dadd_rn(acc, (double)rad);
        float cs = __double2float_rn(acc);
        float pf = __fmul_rn(__fmul_rn(__fmul_rn(cs, 2.0f), kPi32), scale);
        out[(size_t)j * kH] = pf;
    }
}

int launch_sinegen_frames(const float* f0, float* frames, int B, int L2, int scale, cudaStream_t st) {
    sinegen_frames_kernel<<<cdiv(B * kH, 64), 64, 0, st>>>(f0, frames, B, L2, (float)scale);
    ST2_LAUNCH_CHECK();
    return ST2_OK;
}

// F.interpolate(phase*scale, scale_factor=scale, mode='linear') (hifigan.py:155-156), ATen CPU
// upsample_linear1d rounding: src = fma(1/scale, n+0.5, -0.5) clamped at 0; lambda1 = src-i0;
// lambda0 = 1-lambda1; result = fma(lambda0, p[i0], fp32(lambda1*p[i1])).
__device__ __forceinline__ void interp_setup(int n, float inv_scale, int L2, int& i0, int& i1, float& l0,
                                             float& l1) {
    float src = __fmaf_rn(inv_scale, __fadd_rn((float)n, 0.5f), -0.5f);
    src = fmaxf(src, 0.f);
    i0 = (int)src;
    i1 = min(i0 + 1, L2 - 1);
    l1 = __fsub_rn(src, (float)i0);
    l0 = __fsub_rn(1.0f, l1);
}

__global__ void sinegen_phase_kernel(const float* __restrict__ frames, float* __restrict__ phase, int L2, int S,
                                     float inv_scale) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (n >= S) return;
    int i0, i1;
    float l0, l1;
    interp_setup(n, inv_scale, L2, i0, i1, l0, l1);
    const float* f0p = frames + ((size_t)b * L2 + i0) * kH;
    const float* f1p = frames + ((size_t)b * L2 + i1) * kH;
    float* o = phase + ((size_t)b * S + n) * kH;
#pragma unroll
    for (int h = 0; h < kH; ++h) o[h] = __fmaf_rn(l0, f0p[h], __fmul_rn(l1, f1p[h]));
}

int launch_sinegen_phase(const float* frames, float* phase, int B, int L2, int scale, cudaStream_t st) {
    const int S = L2 * scale;
    dim3 grid(cdiv(S, 256), B);
    sinegen_phase_kernel<<<grid, 256, 0, st>>>(frames, phase, L2, S, (float)(1.0 / (double)scale));
    ST2_LAUNCH_CHECK();
    return ST2_OK;
}

// ---- Philox4x32-10 (counter-based; used when the caller passes no noise tape) ---------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}

__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& z0, float& z1) {
    float u1 = ((float)a + 0.5f) * 2.3283064365386963e-10f;   // (0,1]
    float u2 = ((float)b + 0.5f) * 2.3283064365386963e-10f;
    u1 = fminf(fmaxf(u1, 1e-12f), 1.0f);
    // hardware log2 / sincos: the draw is this library's own (the reference's randn cannot be reproduced anyway), its argument
    // range is (0, 1] / [0, 2 pi), and the accurate library versions made this the heaviest part of the kernel
    float r = sqrtf(-2.0f * __logf(u1));
    float s, c;
    __sincosf(6.2831853071795864f * u2, &s, &c);
    z0 = r * c;
    z1 = r * s;
}

// har_source[b][n] = tanh( sum_h lin_w[h] * ( sin(phase)*0.1*uv + noise_amp*noise ) + lin_b )
// (SineGen.forward hifigan.py:189-218 + SourceModuleHnNSF.forward :254-264)
static constexpr int kSrcTile = 256;
__global__ void __launch_bounds__(kSrcTile)
har_source_kernel(const float* __restrict__ f0, const float* __restrict__ frames, const float* __restrict__ noise,
                  uint64_t seed, const uint64_t* __restrict__ seed_dev, const float* __restrict__ lin_w,
                  const float* __restrict__ lin_b, float* __restrict__ har, int L2, int S, int scale, float inv_scale) {
    __shared__ float snz[kSrcTile * kH];
    const int b = blockIdx.y;
    const int n0 = blockIdx.x * kSrcTile;
    const int n = n0 + threadIdx.x;
    if (noise != nullptr) {
        const int cnt = min(kSrcTile, S - n0) * kH;
        const float* src = noise + ((size_t)b * S + n0) * kH;
        for (int i = threadIdx.x; i < cnt; i += kSrcTile) snz[i] = __ldg(src + i);
        __syncthreads();
    }
    if (n >= S) return;
    float nz[12];
    if (noise != nullptr) {
#pragma unroll
        for (int h = 0; h < kH; ++h) nz[h] = snz[threadIdx.x * kH + h];
    } else {
        const uint64_t idx = (uint64_t)b * (uint64_t)S + (uint64_t)n;
        if (seed_dev != nullptr) seed = __ldg(seed_dev);         // seed in device memory: a captured CUDA graph replays with new noise
        const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            uint4 x = philox4x32_10(make_uint4((uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)r, 0x5354u), key);
            box_muller(x.x, x.y, nz[4 * r + 0], nz[4 * r + 1]);
            box_muller(x.z, x.w, nz[4 * r + 2], nz[4 * r + 3]);
        }
    }
    const float f0v = __ldg(f0 + (size_t)b * L2 + n / scale);       // nearest x scale (hifigan.py:284,323)
    const bool voiced = f0v > 10.0f;                                  // voiced_threshold (hifigan.py:282)
    const float uv = voiced ? 1.f : 0.f;
    // noise_amp = uv*noise_std + (1-uv)*sine_amp/3   (hifigan.py:212)
    const float namp = __fadd_rn(__fmul_rn(uv, 0.003f), __fdiv_rn(__fmul_rn(__fsub_rn(1.f, uv), 0.1f), 3.0f));
    int i0, i1;
    float l0, l1;
    interp_setup(n, inv_scale, L2, i0, i1, l0, l1);
    const float* f0p = frames + ((size_t)b * L2 + i0) * kH;
    const float* f1p = frames + ((size_t)b * L2 + i1) * kH;
    float acc = __ldg(lin_b);
#pragma unroll
    for (int h = 0; h < kH; ++h) {
        float ph = __fmaf_rn(l0, __ldg(f0p + h), __fmul_rn(l1, __ldg(f1p + h)));
        float sw = __fmul_rn(sinf(ph), 0.1f);
        float v = __fadd_rn(__fmul_rn(sw, uv), __fmul_rn(namp, nz[h]));
        acc = fmaf(__ldg(lin_w + h), v, acc);
    }
    har[(size_t)b * S + n] = tanhf(acc);
}

int launch_har_source(const float* f0, const float* frames, const float* noise, uint64_t seed, const uint64_t* seed_dev,
                      const float* lin_w, const float* lin_b, float* har, int B, int L2, int scale, cudaStream_t st) {
    const int S = L2 * scale;
    dim3 grid(cdiv(S, kSrcTile), B);
    har_source_kernel<<<grid, kSrcTile, 0, st>>>(f0, frames, noise, seed, seed_dev, lin_w, lin_b, har, L2, S, scale,
                                                 (float)(1.0 / (double)scale));
    ST2_LAUNCH_CHECK();
    return ST2_OK;
}

// ---- CustomSTFT.transform (istftnet.py:207-243): replicate pad n_fft/2, windowed DFT as two
// strided convs, magnitude (+1e-14) and atan2 phase with the (im==0 & re<0) -> pi fix-up.
// out[b][f][0:bins] = magnitude, out[b][f][bins:2*bins] = phase  (torch.cat at istftnet.py:550)
static constexpr int kMaxFft = 32;
static constexpr int kMaxBins = kMaxFft / 2 + 1;
__global__ void stft_transform_kernel(const float* __restrict__ har, const float* __restrict__ wr,
                                      const float* __restrict__ wi, float* __restrict__ out, int ld_out, int S,
                                      int frames, int n_fft, int hop) {
    __shared__ float swr[kMaxBins * kMaxFft], swi[kMaxBins * kMaxFft];
    const int bins = n_fft / 2 + 1;
    for (int i = threadIdx.x; i < bins * n_fft; i += blockDim.x) {
        swr[i] = wr[i];
        swi[i] = wi[i];
    }
    __syncthreads();
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (f >= frames) return;
    const float* hb = har + (size_t)b * S;
    float x[kMaxFft];
    const int pad = n_fft / 2;
    for (int n = 0; n < n_fft; ++n) {
        int i = f * hop + n - pad;
        i = max(0, min(S - 1, i));
        x[n] = __ldg(hb + i);
    }
    float* o = out + ((size_t)b * frames + f) * ld_out;
    for (int k = 0; k < bins; ++k) {
        float re = 0.f, im = 0.f;
        for (int n = 0; n < n_fft; ++n) {
            re = fmaf(x[n], swr[k * n_fft + n], re);
            im = fmaf(x[n], swi[k * n_fft + n], im);
        }
        float mag = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im)), 1e-14f));
        float ph = atan2f(im, re);
        if (im == 0.f && re < 0.f) ph = kPi32;
        o[k] = mag;
        o[bins + k] = ph;
    }
}

int launch_stft_transform(const float* har, const float* wr, const float* wi, float* out, int ld_out, int B, int S,
                          int n_fft, int hop, cudaStream_t st) {
    ST2_REQUIRE(n_fft <= kMaxFft && n_fft % 2 == 0, "stft: n_fft=%d unsupported", n_fft);
    const int frames = S / hop + 1;
    dim3 grid(cdiv(frames, 128), B);
    stft_transform_kernel<<<grid, 128, 0, st>>>(har, wr, wi, out, ld_out, S, frames, n_fft, hop);
    ST2_LAUNCH_CHECK();
    return ST2_OK;
}

// ---- iSTFT head (istftnet.py:571-573 + CustomSTFT.inverse :246-293): x = conv_post output
// [B][frames][ld_x] (2*bins channels); spec = exp(x[:bins]), phase = sin(x[bins:]);
// re = spec*cos(phase), im = spec*sin(phase); overlap-add of the two transposed convs with the
// windowed inverse basis, real - imag, trimmed by n_fft/2 on both sides.  Shared-memory staged:
// one CTA produces kIstftTile consecutive samples from the <= (kIstftTile+n_fft)/hop + 2 frames that touch them.
static constexpr int kIstftTile = 256;
__global__ void __launch_bounds__(kIstftTile)
istft_head_kernel(const float* __restrict__ x, int ld_x, const float* __restrict__ wr, const float* __restrict__ wi,
                  float* __restrict__ out, int frames, int S, int n_fft, int hop) {
    extern __shared__ float sm[];
    const int bins = n_fft / 2 + 1;
    const int pad = n_fft / 2;
    float* swr = sm;                         // [bins][n_fft]
    float* swi = swr + bins * n_fft;
    float* sre = swi + bins * n_fft;         // [nf][bins]
    const int b = blockIdx.y;
    const int n0 = blockIdx.x * kIstftTile;
    const int m0 = n0 + pad;                                  // first full-length index of the tile
    const int f_lo = max(0, (m0 - (n_fft - 1) + hop - 1) / hop);
    const int f_hi = min(frames - 1, (m0 + kIstftTile - 1) / hop);
    const int nf = f_hi - f_lo + 1;
    float* sim = sre + ((kIstftTile + n_fft) / hop + 2) * bins;
    for (int i = threadIdx.x; i < bins * n_fft; i += kIstftTile) {
        swr[i] = wr[i];
        swi[i] = wi[i];
    }
    for (int i = threadIdx.x; i < nf * bins; i += kIstftTile) {
        int fl = i / bins, k = i - fl * bins;
        const float* row = x + ((size_t)b * frames + f_lo + fl) * ld_x;
        float spec = expf(row[k]);
        float ph = sinf(row[bins + k]);
        float s, c;
        sincosf(ph, &s, &c);
        sre[i] = spec * c;
        sim[i] = spec * s;
    }
    __syncthreads();
    const int n = n0 + threadIdx.x;
    if (n >= S) return;
    const int m = n + pad;
    float accr = 0.f, acci = 0.f;
    const int fa = max(f_lo, (m - (n_fft - 1) + hop - 1) / hop);
    const int fb = min(f_hi, m / hop);
    for (int f = fa; f <= fb; ++f) {
        const int j = m - f * hop;                       // 0 <= j < n_fft
        const float* r = sre + (f - f_lo) * bins;
        const float* i_ = sim + (f - f_lo) * bins;
        for (int k = 0; k < bins; ++k) {
            accr = fmaf(r[k], swr[k * n_fft + j], accr);
            acci = fmaf(i_[k], swi[k * n_fft + j], acci);
        }
    }
    out[(size_t)b * S + n] = accr - acci;
}

int launch_istft_head(const float* x, int ld_x, const float* wr, const float* wi, float* out, int B, int frames,
                      int S, int n_fft, int hop, cudaStream_t st) {
    ST2_REQUIRE(n_fft <= kMaxFft && n_fft % 2 == 0 && hop >= 1, "istft: n_fft=%d hop=%d unsupported", n_fft, hop);
    const int bins = n_fft / 2 + 1;
    size_t smem = ((size_t)2 * bins * n_fft + (size_t)2 * ((kIstftTile + n_fft) / hop + 2) * bins) * sizeof(float);
    ST2_REQUIRE(smem <= 48 * 1024, "istft: smem %zu too large", smem);
    dim3 grid(cdiv(S, kIstftTile), B);
    istft_head_kernel<<<grid, kIstftTile, smem, st>>>(x, ld_x, wr, wi, out, frames, S, n_fft, hop);
    ST2_LAUNCH_CHECK();
    return ST2_OK;
}

}  // namespace st2
