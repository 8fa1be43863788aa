// AdaIN1d (InstanceNorm + style affine) fused with LeakyReLU / Snake -- the HBM-bound half of
// the decoder.  Replaces Modules/hifigan.py:14-24 (AdaIN1d), :68/:71 (Snake1D) and the
// nn.LeakyReLU(0.2) of AdainResBlk1d (:360,:392,:396) on channels-last activations.
//
//   1. in_stats_kernel    x[B][T][ld] -> per (b, slab, c) partial (sum, sumsq) in double
//   2. adain_coef_kernel  partials + style rows h -> y = a*x + b coefficients per (b,c)
//   3. affine_act_kernel  y = act(a*x + b), 128-bit loads, fp32 / bf16 / fp16 stores
//
// Algorithmic bytes of (3): numel*(4 + sizeof(out)); (1) re-reads x (L2-resident when the
// producer just wrote it).  The reference's CPU statistics are correct to <= 1 fp32 ulp for
// 1.44 M-element rows, hence the double accumulators.
#include "common.cuh"

namespace st2 {

static constexpr int kThreads = 256;
static constexpr int kMaxSlabs = 256;

struct StatsGeom {
    int cq;          // channel quads (C rounded up to 4) / 4
    int qpc;         // quads per CTA (power of two <= 64)
    int rpp;         // rows per pass = 256 / qpc
    int cblocks;     // CTAs along channels
    int nslab;       // CTAs along time
    int slab_rows;
};

static StatsGeom stats_geom(int T, int C) {
    StatsGeom g;
    g.cq = (C + 3) / 4;
    int q = 1;
    while (q < g.cq && q < 64) q <<= 1;
    g.qpc = q;
    g.rpp = kThreads / q;
    g.cblocks = (g.cq + q - 1) / q;
    int target = g.rpp * 32;
    int nslab = (T + target - 1) / target;
    if (nslab > kMaxSlabs) nslab = kMaxSlabs;
    if (nslab < 1) nslab = 1;
    g.nslab = nslab;
    g.slab_rows = (T + nslab - 1) / nslab;
    return g;
}

int64_t adain_scratch_bytes(int B, int T, int C) {
    StatsGeom g = stats_geom(T, C);
    int64_t partial = (int64_t)B * g.nslab * g.cq * 4 * 2 * sizeof(double);
    return (partial + 255) / 256 * 256;
}

template <bool VEC>
__global__ void __launch_bounds__(kThreads)
in_stats_kernel(const float* __restrict__ x, int ld, int T, int C, int qpc_log2, int slab_rows,
                int cq_total, double2* __restrict__ partial) {
    const int qpc = 1 << qpc_log2;
    const int rpp = kThreads >> qpc_log2;
    const int ql = threadIdx.x & (qpc - 1);
    const int rl = threadIdx.x >> qpc_log2;
    const int q = blockIdx.y * qpc + ql;
    const int b = blockIdx.z;
    const int slab = blockIdx.x;
    const int t0 = slab * slab_rows;
    const int t1 = min(T, t0 + slab_rows);
    double s[4] = {0, 0, 0, 0}, ss[4] = {0, 0, 0, 0};
    if (q < cq_total) {
        const float* xb = x + (size_t)b * T * ld + q * 4;
        for (int t = t0 + rl; t < t1; t += rpp) {
            float v[4];
            if (VEC) {
                float4 f = __ldg(reinterpret_cast<const float4*>(xb + (size_t)t * ld));
                v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = (q * 4 + j < C) ? __ldg(xb + (size_t)t * ld + j) : 0.f;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                double d = (double)v[j];
                s[j] += d;
                ss[j] = fma(d, d, ss[j]);
            }
        }
    }
    __shared__ double red[kThreads * 8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        red[(j * 2 + 0) * kThreads + threadIdx.x] = s[j];
        red[(j * 2 + 1) * kThreads + threadIdx.x] = ss[j];
    }
    __syncthreads();
    if (rl == 0 && q < cq_total) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double a = 0, c2 = 0;
            for (int r = 0; r < rpp; ++r) {
                a += red[(j * 2 + 0) * kThreads + r * qpc + ql];
                c2 += red[(j * 2 + 1) * kThreads + r * qpc + ql];
            }
            partial[((size_t)b * gridDim.x + slab) * (cq_total * 4) + q * 4 + j] = make_double2(a, c2);
        }
    }
}

int launch_in_stats(const float* x, int ld, int B, int T, int C, void* scratch, cudaStream_t st) {
    StatsGeom g = stats_geom(T, C);
    int lg = 0;
    while ((1 << lg) < g.qpc) ++lg;
    dim3 grid(g.nslab, g.cblocks, B);
    bool vec = (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && (g.cq * 4 <= ld);
    if (vec)
        in_stats_kernel<true><<<grid, kThreads, 0, st>>>(x, ld, T, C, lg, g.slab_rows, g.cq, (double2*)scratch);
    else
        in_stats_kernel<false><<<grid, kThreads, 0, st>>>(x, ld, T, C, lg, g.slab_rows, g.cq, (double2*)scratch);
    ST2_LAUNCH_CHECK();
    return ST2_OK;
}

// a = (1+gamma)*rstd, b = beta - mean*a   (y = a*x + b  ==  (1+gamma)*IN(x) + beta)
__global__ void adain_coef_kernel(const double2* __restrict__ partial, int nslab, int cq4,
                                  const float* __restrict__ h, int ld_h, int h_off, float* __restrict__ coef,
                                  int T, int C, int Cpad) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (c >= Cpad) return;
    float a = 0.f, bb = 0.f;
    if (c < C) {
        if (h != nullptr) {
            double s = 0, ss = 0;
            for (int i = 0; i < nslab; ++i) {
                double2 p = partial[((size_t)b * nslab + i) * cq4 + c];
                s += p.x;
                ss += p.y;
            }
            double mean = s / (double)T;
            double var = ss / (double)T - mean * mean;   // biased, InstanceNorm1d
            if (var < 0) var = 0;
            double rstd = 1.0 / sqrt(var + 1e-5);
            double gamma = (double)h[(size_t)b * ld_h + h_off + c];
            double beta = (double)h[(size_t)b * ld_h + h_off + C + c];
            double ad = (1.0 + gamma) * rstd;
            a = (float)ad;
            bb = (float)(beta - mean * ad);
        } else {
            a = 1.f;
            bb = 0.f;
        }
    }
    coef[((size_t)b * 2 + 0) * Cpad + c] = a;
    coef[((size_t)b * 2 + 1) * Cpad + c] = bb;
}

int launch_adain_coef(const void* scratch, const float* h, int ld_h, int h_off, float* coef, int B, int T,
                      int C, int Cpad, cudaStream_t st) {
    StatsGeom g = stats_geom(T, C);
    dim3 grid(cdiv(Cpad, 128), B);
    adain_coef_kernel<<<grid, 128, 0, st>>>((const double2*)scratch, g.nslab, g.cq * 4, h, ld_h, h_off, coef, T,
                                           C, Cpad);
    ST2_LAUNCH_CHECK();
    return ST2_OK;
}

template <int DT> struct Store4;
template <> struct Store4<DT_F32> {
    static __device__ __forceinline__ void st(void* y, size_t idx, float4 v) {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + idx) = v;
    }
};
template <> struct Store4<DT_BF16> {
    static __device__ __forceinline__ void st(void* y, size_t idx, float4 v) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y);
        __nv_bfloat162 hi = __floats2bfloat162_rn(v.z, v.w);
        uint2 u;
        u.x = *reinterpret_cast<uint32_t*>(&lo);
        u.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(y) + idx) = u;
    }
};
template <> struct Store4<DT_F16> {
    static __device__ __forceinline__ void st(void* y, size_t idx, float4 v) {
        __half2 lo = __floats2half2_rn(v.x, v.y);
        __half2 hi = __floats2half2_rn(v.z, v.w);
        uint2 u;
        u.x = *reinterpret_cast<uint32_t*>(&lo);
        u.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(y) + idx) = u;
    }
};

template <int ACT, bool FAST>
__device__ __forceinline__ float act_fn(float v, float p0, float p1) {
    if (ACT == ACT_LRELU) return v >= 0.f ? v : v * p0;           // p0 = slope
    if (ACT == ACT_SNAKE) {                                        // p0 = alpha, p1 = 1/alpha
        float sn = FAST ? __sinf(p0 * v) : sinf(p0 * v);
        return fmaf(p1 * sn, sn, v);
    }
    if (ACT == ACT_GELU) return 0.5f * v * (1.f + erff(v * 0.70710678118654752f));   // nn.GELU(), exact erf form (vocos.py:48)
    return v;
}

// One CTA: `rows` consecutive time steps of one utterance, all Cpad channels.
template <int ACT, int DT>
__global__ void __launch_bounds__(kThreads)
affine_act_kernel(const float* __restrict__ x, int ld_x, const float* __restrict__ coef,
                  const float* __restrict__ alpha, float slope, void* __restrict__ y, int ld_y, int T, int Cpad,
                  int rows) {
    extern __shared__ float sm[];
    float* sa = sm;
    float* sb = sm + Cpad;
    float* sal = sm + 2 * Cpad;      // alpha
    float* sial = sm + 3 * Cpad;     // 1/alpha
    const int b = blockIdx.y;
    for (int c = threadIdx.x; c < Cpad; c += kThreads) {
        sa[c] = coef[((size_t)b * 2 + 0) * Cpad + c];
        sb[c] = coef[((size_t)b * 2 + 1) * Cpad + c];
        if (ACT == ACT_SNAKE) {
            float al = alpha[c];
            sal[c] = al;
            sial[c] = 1.f / al;
        }
    }
    __syncthreads();
    const int cq = Cpad >> 2;
    const int t0 = blockIdx.x * rows;
    const int nrow = min(rows, T - t0);
    const int total = nrow * cq;
    const float* xb = x + ((size_t)b * T + t0) * ld_x;
    const size_t ybase = ((size_t)b * T + t0) * ld_y;
    constexpr bool FAST = (DT != DT_F32);
    constexpr int U = 4;
    for (int i0 = threadIdx.x; i0 < total; i0 += kThreads * U) {
        float4 v[U];
        int r[U], q[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            int i = i0 + u * kThreads;
            r[u] = i / cq;
            q[u] = i - r[u] * cq;
            if (i < total) v[u] = __ldg(reinterpret_cast<const float4*>(xb + (size_t)r[u] * ld_x + q[u] * 4));
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            int i = i0 + u * kThreads;
            if (i >= total) continue;
            const int c = q[u] * 4;
            float4 a4 = *reinterpret_cast<const float4*>(sa + c);
            float4 b4 = *reinterpret_cast<const float4*>(sb + c);
            float4 o;
            o.x = fmaf(a4.x, v[u].x, b4.x);
            o.y = fmaf(a4.y, v[u].y, b4.y);
            o.z = fmaf(a4.z, v[u].z, b4.z);
            o.w = fmaf(a4.w, v[u].w, b4.w);
            if (ACT == ACT_SNAKE) {
                float4 al = *reinterpret_cast<const float4*>(sal + c);
                float4 ia = *reinterpret_cast<const float4*>(sial + c);
                o.x = act_fn<ACT, FAST>(o.x, al.x, ia.x);
                o.y = act_fn<ACT, FAST>(o.y, al.y, ia.y);
                o.z = act_fn<ACT, FAST>(o.z, al.z, ia.z);
                o.w = act_fn<ACT, FAST>(o.w, al.w, ia.w);
            } else if (ACT == ACT_LRELU || ACT == ACT_GELU) {
                o.x = act_fn<ACT, FAST>(o.x, slope, 0.f);
                o.y = act_fn<ACT, FAST>(o.y, slope, 0.f);
                o.z = act_fn<ACT, FAST>(o.z, slope, 0.f);
                o.w = act_fn<ACT, FAST>(o.w, slope, 0.f);
            }
            Store4<DT>::st(y, ybase + (size_t)r[u] * ld_y + c, o);
        }
    }
}

template <int ACT>
static int launch_affine_act_dt(const float* x, int ld_x, const float* coef, const float* alpha, float slope,
                                void* y, int ld_y, int out_dtype, int B, int T, int Cpad, cudaStream_t st) {
    int cq = Cpad / 4;
    int rows = 4096 / cq;
    if (rows < 1) rows = 1;
    dim3 grid(cdiv(T, rows), B);
    size_t smem = (size_t)4 * Cpad * sizeof(float);
    switch (out_dtype) {
        case DT_F32:
            affine_act_kernel<ACT, DT_F32><<<grid, kThreads, smem, st>>>(x, ld_x, coef, alpha, slope, y, ld_y, T, Cpad, rows);
            break;
        case DT_BF16:
            affine_act_kernel<ACT, DT_BF16><<<grid, kThreads, smem, st>>>(x, ld_x, coef, alpha, slope, y, ld_y, T, Cpad, rows);
            break;
        case DT_F16:
            affine_act_kernel<ACT, DT_F16><<<grid, kThreads, smem, st>>>(x, ld_x, coef, alpha, slope, y, ld_y, T, Cpad, rows);
            break;
        default:
            set_error("affine_act: bad out_dtype %d", out_dtype);
            return ST2_ERR_INVALID;
    }
    ST2_LAUNCH_CHECK();
    return ST2_OK;
}

int launch_affine_act(const float* x, int ld_x, const float* coef, const float* alpha, int act, float slope,
                      void* y, int ld_y, int out_dtype, int B, int T, int Cpad, cudaStream_t st) {
    ST2_REQUIRE(Cpad % 4 == 0 && ld_x % 4 == 0 && ld_y % 4 == 0 && Cpad <= ld_x && Cpad <= ld_y,
                "affine_act: Cpad=%d ld_x=%d ld_y=%d must be multiples of 4 with Cpad <= ld", Cpad, ld_x, ld_y);
    ST2_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 7) == 0,
                "affine_act: unaligned pointers");
    ST2_REQUIRE(Cpad <= 2048, "affine_act: Cpad=%d too large", Cpad);
    switch (act) {
        case ACT_NONE: return launch_affine_act_dt<ACT_NONE>(x, ld_x, coef, alpha, slope, y, ld_y, out_dtype, B, T, Cpad, st);
        case ACT_LRELU: return launch_affine_act_dt<ACT_LRELU>(x, ld_x, coef, alpha, slope, y, ld_y, out_dtype, B, T, Cpad, st);
        case ACT_GELU: return launch_affine_act_dt<ACT_GELU>(x, ld_x, coef, alpha, slope, y, ld_y, out_dtype, B, T, Cpad, st);
        case ACT_SNAKE:
            ST2_REQUIRE(alpha != nullptr, "affine_act: snake needs alpha");
            return launch_affine_act_dt<ACT_SNAKE>(x, ld_x, coef, alpha, slope, y, ld_y, out_dtype, B, T, Cpad, st);
    }
    set_error("affine_act: bad act %d", act);
    return ST2_ERR_INVALID;
}

}  // namespace st2
