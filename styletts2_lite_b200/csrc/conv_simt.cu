// fp32 SIMT implicit-GEMM Conv1d / ConvTranspose1d on channels-last activations.
//
// This is the fp32 parity backbone (ST2_PREC_FP32): exact fp32 FFMA accumulation, so the whole
// decoder stays inside the <= 1e-4 envelope against the CPU reference.  It replaces every
// nn.Conv1d / nn.ConvTranspose1d call of Modules/hifigan.py (AdainResBlk1d :377-382,
// AdaINResBlock1 :29-46, ups :292-294, noise_convs :298-302, asr_res :438-440) -- the
// tensor-core path (conv_tc.cu) implements the same ConvArgs contract with tcgen05.
//
// GEMM view: M = output time (tile 128), N = Cout (tile 64 or 32), K = taps x Cin (chunks of 16).
// Zero padding is materialised by predicated loads; ConvTranspose1d runs as `stride` polyphase
// sub-convolutions (blockIdx.z), see ConvArgs in common.cuh.
#include "common.cuh"

namespace st2 {

static constexpr int BK = 16;
static constexpr int kConvThreads = 256;

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(kConvThreads)
conv_simt_kernel(const ConvArgs a, const bool vecA, const bool vecB, const bool vecO) {
    constexpr int APITCH = BM + 4;
    constexpr int A_CHUNKS = BM * BK / 4 / kConvThreads;                       // float4 chunks per thread
    constexpr int B_CHUNKS = (BK * BN / 4 + kConvThreads - 1) / kConvThreads;  // 1
    constexpr int NTX = BN / TN;
    static_assert((BM / TM) * NTX == kConvThreads, "tile/thread mismatch");
    static_assert(TM % 4 == 0 && TN == 4, "vector widths");
    __shared__ __align__(16) float As[2][BK][APITCH];
    __shared__ __align__(16) float Bs[2][BK][BN];

    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int b = blockIdx.z / a.phases;
    const int p = blockIdx.z - b * a.phases;
    const int tx = tid % NTX, ty = tid / NTX;

    const float* xb = a.x + (size_t)b * a.Tin * a.ld_x;
    const int kchunks = (a.Cin + BK - 1) / BK;
    const int nit = a.ntaps * kchunks;

    float4 ra[A_CHUNKS];
    float4 rb[B_CHUNKS];

    auto load_tiles = [&](int it) {
        const int j = it / kchunks;
        const int ci0 = (it - j * kchunks) * BK;
#pragma unroll
        for (int i = 0; i < A_CHUNKS; ++i) {
            const int c = tid + i * kConvThreads;
            const int row = c >> 2, kq = c & 3;
            const int m = m0 + row;
            const int t = m * a.in_stride + j * a.tap_step + a.in_off;
            const int ci = ci0 + kq * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < a.M && t >= 0 && t < a.Tin && ci < a.Cin) {
                const float* ptr = xb + (size_t)t * a.ld_x + ci;
                if (vecA) {
                    v = __ldg(reinterpret_cast<const float4*>(ptr));
                } else {
                    v.x = __ldg(ptr);
                    if (ci + 1 < a.Cin) v.y = __ldg(ptr + 1);
                    if (ci + 2 < a.Cin) v.z = __ldg(ptr + 2);
                    if (ci + 3 < a.Cin) v.w = __ldg(ptr + 3);
                }
            }
            ra[i] = v;
        }
        const int widx = p + j * a.w_step;
        const float* wt = a.w + (size_t)widx * a.Cin * a.Cout;
#pragma unroll
        for (int i = 0; i < B_CHUNKS; ++i) {
            const int c = tid + i * kConvThreads;
            const int kr = c / (BN / 4), cq = c - kr * (BN / 4);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (kr < BK) {
                const int ci = ci0 + kr, co = n0 + cq * 4;
                if (ci < a.Cin && co < a.Cout) {
                    const float* ptr = wt + (size_t)ci * a.Cout + co;
                    if (vecB) {
                        v = __ldg(reinterpret_cast<const float4*>(ptr));
                    } else {
                        v.x = __ldg(ptr);
                        if (co + 1 < a.Cout) v.y = __ldg(ptr + 1);
                        if (co + 2 < a.Cout) v.z = __ldg(ptr + 2);
                        if (co + 3 < a.Cout) v.w = __ldg(ptr + 3);
                    }
                }
            }
            rb[i] = v;
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int i = 0; i < A_CHUNKS; ++i) {
            const int c = tid + i * kConvThreads;
            const int row = c >> 2, kq = c & 3;
            As[buf][kq * 4 + 0][row] = ra[i].x;
            As[buf][kq * 4 + 1][row] = ra[i].y;
            As[buf][kq * 4 + 2][row] = ra[i].z;
            As[buf][kq * 4 + 3][row] = ra[i].w;
        }
#pragma unroll
        for (int i = 0; i < B_CHUNKS; ++i) {
            const int c = tid + i * kConvThreads;
            const int kr = c / (BN / 4), cq = c - kr * (BN / 4);
            if (kr < BK) *reinterpret_cast<float4*>(&Bs[buf][kr][cq * 4]) = rb[i];
        }
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    for (int it = 0; it < nit; ++it) {
        const int cur = it & 1;
        if (it + 1 < nit) load_tiles(it + 1);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float av[TM], bv[TN];
#pragma unroll
            for (int i = 0; i < TM; i += 4) {
                float4 t4 = *reinterpret_cast<const float4*>(&As[cur][kk][ty * TM + i]);
                av[i] = t4.x; av[i + 1] = t4.y; av[i + 2] = t4.z; av[i + 3] = t4.w;
            }
            {
                float4 t4 = *reinterpret_cast<const float4*>(&Bs[cur][kk][tx * TN]);
                bv[0] = t4.x; bv[1] = t4.y; bv[2] = t4.z; bv[3] = t4.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (it + 1 < nit) store_tiles(cur ^ 1);
        __syncthreads();
    }

    // epilogue: y = (acc + bias + res) * scale (+ y_old)
    const int co = n0 + tx * TN;
    if (co >= a.Cout) return;
    float bias[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) bias[j] = (a.bias != nullptr && co + j < a.Cout) ? __ldg(a.bias + co + j) : 0.f;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + ty * TM + i;
        if (m >= a.M) continue;
        const int t = m * a.out_stride + p - a.out_pad;
        // mirror (ReflectionPad1d((1,0)) folded into the store): outputs are shifted by one row, so t == 0 would be output
        // index -1 of the convolution; row 0 belongs to the mirror write of row 2 alone (two writers raced here)
        if (t < (a.mirror ? 1 : 0) || t >= a.Tout) continue;
        float v[TN];
#pragma unroll
        for (int j = 0; j < TN; ++j) v[j] = acc[i][j] + bias[j];
        // ReflectionPad1d((1,0)) fused: the value of (shifted) row 2 also lands in row 0
        const int nrep = (a.mirror && t == 2) ? 2 : 1;
        for (int rep = 0; rep < nrep; ++rep) {
            const int tt = rep == 0 ? t : 0;
            float o[TN];
#pragma unroll
            for (int j = 0; j < TN; ++j) o[j] = v[j];
            if (a.res != nullptr) {
                const float* rp = a.res + ((size_t)b * (a.Tout >> a.res_shift) + (tt >> a.res_shift)) * a.ld_res + co;
#pragma unroll
                for (int j = 0; j < TN; ++j)
                    if (co + j < a.Cout) o[j] += rp[j];
            }
            float* yp = a.y + ((size_t)b * a.Tout + tt) * a.ld_y + co;
            if (vecO) {
                float4 r = make_float4(o[0], o[1], o[2], o[3]);
                if (a.accumulate) {
                    float4 old = *reinterpret_cast<const float4*>(yp);
                    r.x += old.x; r.y += old.y; r.z += old.z; r.w += old.w;
                }
                r.x *= a.scale; r.y *= a.scale; r.z *= a.scale; r.w *= a.scale;
                *reinterpret_cast<float4*>(yp) = r;
            } else {
#pragma unroll
                for (int j = 0; j < TN; ++j)
                    if (co + j < a.Cout) yp[j] = ((a.accumulate ? yp[j] : 0.f) + o[j]) * a.scale;
            }
        }
    }
}

int launch_conv_simt(const ConvArgs& a, cudaStream_t st) {
    ST2_REQUIRE(a.B > 0 && a.M > 0 && a.Cin > 0 && a.Cout > 0 && a.ntaps > 0 && a.phases > 0,
                "conv_simt: bad geometry B=%d M=%d Cin=%d Cout=%d taps=%d phases=%d", a.B, a.M, a.Cin, a.Cout,
                a.ntaps, a.phases);
    const bool vecA = (a.Cin % 4 == 0) && (a.ld_x % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.x) & 15) == 0);
    const bool vecB = (a.Cout % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.w) & 15) == 0);
    const bool vecO = (a.Cout % 4 == 0) && (a.ld_y % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.y) & 15) == 0);
    if (a.Cout > 32) {
        dim3 grid(cdiv(a.M, 128), cdiv(a.Cout, 64), a.B * a.phases);
        conv_simt_kernel<128, 64, 8, 4><<<grid, kConvThreads, 0, st>>>(a, vecA, vecB, vecO);
    } else {
        dim3 grid(cdiv(a.M, 128), cdiv(a.Cout, 32), a.B * a.phases);
        conv_simt_kernel<128, 32, 4, 4><<<grid, kConvThreads, 0, st>>>(a, vecA, vecB, vecO);
    }
    ST2_LAUNCH_CHECK();
    return ST2_OK;
}

}  // namespace st2
