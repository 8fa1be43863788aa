// Fused  [AdaIN affine + Snake / LeakyReLU]  ->  Conv1d / ConvTranspose1d (tcgen05)  ->
// [bias + residual + scale + accumulate + InstanceNorm partial statistics]   for sm_100a.
//
// Replaces, in one kernel, the chain of Modules/hifigan.py:67-73 (AdaINResBlock1 iteration:
// n1/n2 -> Snake -> c1/c2 -> + x) and :329-334 (Snake -> ups -> + x_source): the raw fp32
// activation tile is read from HBM exactly once (with its (k-1)*dilation halo rows), the
// per-(b,c) AdaIN affine y = a*x + b and the activation are applied in registers, the result is
// written as the 128-byte-swizzled K-major bf16/fp16 A operand in shared memory, and every tap of
// the convolution is one UMMA whose A descriptor start address is shifted by whole 128-byte rows
// (dilation = row shift; zero padding = rows written as zeros).  Weights stream through a TMA
// ring.  The epilogue drains TMEM through a swizzled per-warp staging tile so that residual
// reads and output writes are 64-byte row segments, and emits per-tile (sum, sum of squares) per
// channel so the next AdaIN needs no extra pass over the tensor.
//
// Algorithmic HBM bytes per element: 4 (x) + 4 (y) [+ 4 residual] -- the kernel is HBM-bound for
// C <= 128; two CTAs are co-resident per SM so that one CTA's load/transform and
// epilogue phases overlap the other's MMA phase.
//
// Warp roles (320 threads): warp 0 = TMA weight producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2-9 = workers (transform, then epilogue).
#include <cuda.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace st2 {

static constexpr int FW = 8;                      // worker warps
static constexpr int F_THREADS = (2 + FW) * 32;   // 320
static constexpr int F_KC = 64;

struct FusedParams {
    // input + transform
    const float* x; int ld_x; int Tin;
    const float* coef;        // [B][2][coef_ld]: a, b
    int coef_ld;
    const float* alpha;       // [Cin] snake alpha (ACT_SNAKE)
    float slope;              // ACT_LRELU
    int Cin;                  // real input channels
    int kchunks;              // CinPad / 64
    // geometry (ConvArgs contract, in_stride == 1)
    int B, M, Tout, Cout;
    int ntaps, tap_step, in_off;
    int phases, w_step, out_stride, out_pad;
    int halo_min;             // smallest input-row offset of any tap relative to the tile's first row
    int rows;                 // MT + span  (rows of the A tile)
    int R;                    // accumulators per tile (MT = 128*R)
    int bn;                   // N tile (CoutPad or 256)
    int ntile_n;
    int mtiles;               // ceil(M / MT)
    int num_tiles;            // B * phases * mtiles * ntile_n
    int stages;               // weight ring depth
    int tmem_cols;
    int is_bf16;
    // epilogue
    const float* bias;
    const float* res; int ld_res; int res_shift;
    float* y; int ld_y;
    float scale; int accumulate; int mirror;
    float2* stats;            // [B][phases*mtiles][Cout] partial (sum, sumsq) or nullptr
};

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void worker_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(FW * 32) : "memory"); }

template <int ACT>
__device__ __forceinline__ float fused_act(float v, float p0, float p1) {
    if (ACT == ACT_LRELU) return v >= 0.f ? v : v * p0;
    if (ACT == ACT_SNAKE) {
        float sn = __sinf(p0 * v);
        return fmaf(p1 * sn, sn, v);
    }
    return v;
}

__device__ __forceinline__ uint32_t pack16(float lo, float hi, int is_bf16) {
    if (is_bf16) {
        __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
        return *reinterpret_cast<uint32_t*>(&t);
    }
    __half2 t = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&t);
}

template <int ACT>
__global__ void __launch_bounds__(F_THREADS, 2)
conv_fused_kernel(const __grid_constant__ CUtensorMap map_b, const FusedParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t a_bytes = ((uint32_t)p.rows * 128u + 1023u) & ~1023u;
    const uint32_t b_stage_bytes = (uint32_t)p.bn * 128u;
    uint8_t* smem_a = smem;                                       // [rows][64] 16-bit, SWIZZLE_128B
    uint8_t* smem_b = smem_a + a_bytes;                           // ring of [bn][64]
    float* staging = reinterpret_cast<float*>(smem_b + (size_t)p.stages * b_stage_bytes);   // [FW][32][16]
    float2* tstats = reinterpret_cast<float2*>(staging + FW * 32 * 16);                      // [4][R][bn]
    uint64_t* bars = reinterpret_cast<uint64_t*>(tstats + 4 * p.R * p.bn);
    uint64_t* b_full = bars;
    uint64_t* b_empty = bars + p.stages;
    uint64_t* a_full = bars + 2 * p.stages;
    uint64_t* a_empty = a_full + 1;
    uint64_t* tmem_full = a_full + 2;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(a_full + 3);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_b);
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&b_full[s], 1);
            mbar_init(&b_empty[s], 1);
        }
        mbar_init(a_full, FW);
        mbar_init(a_empty, 1);
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                     "r"((uint32_t)p.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // channels >= Cin of the single/last chunk are never written by the transform: zero the A tile once
    for (uint32_t i = threadIdx.x; i < a_bytes / 16; i += F_THREADS)
        reinterpret_cast<uint4*>(smem_a)[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const int MT = 128 * p.R;
    const int per_b = p.phases * p.mtiles * p.ntile_n;

    if (warp == 0) {
        // ===== weight producer (TMA ring) =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const int rem = tile % per_b;
                const int ph = rem / (p.mtiles * p.ntile_n);
                const int nt = rem % p.ntile_n;
                for (int kc = 0; kc < p.kchunks; ++kc)
                    for (int j = 0; j < p.ntaps; ++j) {
                        mbar_wait(&b_empty[stage], phase ^ 1);
                        mbar_expect_tx(&b_full[stage], b_stage_bytes);
                        tma_load_3d(smem_b + (size_t)stage * b_stage_bytes, &map_b, &b_full[stage], kc * F_KC, nt * p.bn,
                                    ph + j * p.w_step);
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc = umma_idesc(128, p.bn, p.is_bf16);
            int stage = 0;
            uint32_t phase = 0, a_phase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                for (int kc = 0; kc < p.kchunks; ++kc) {
                    mbar_wait(a_full, a_phase);
                    a_phase ^= 1;
                    tc_fence_after();
                    for (int j = 0; j < p.ntaps; ++j) {
                        mbar_wait(&b_full[stage], phase);
                        tc_fence_after();
                        const uint64_t bdesc = umma_desc_sw128(smem_u32(smem_b + (size_t)stage * b_stage_bytes));
                        const int rowoff = j * p.tap_step + p.in_off - p.halo_min;
                        for (int r = 0; r < p.R; ++r) {
                            const uint64_t adesc = umma_desc_sw128(smem_u32(smem_a + (size_t)(r * 128 + rowoff) * 128));
#pragma unroll
                            for (int k4 = 0; k4 < F_KC / 16; ++k4)
                                umma_f16(tmem_base + (uint32_t)(r * p.bn), adesc + (uint64_t)(k4 * 2), bdesc + (uint64_t)(k4 * 2),
                                         idesc, (kc > 0 || j > 0 || k4 > 0) ? 1u : 0u);
                        }
                        umma_commit(&b_empty[stage]);
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    }
                    umma_commit(a_empty);            // A chunk consumed
                }
                umma_commit(tmem_full);              // accumulators of this tile complete
            }
        }
    } else {
        // ===== workers: transform (x -> act(a*x+b) -> swizzled 16-bit A), then epilogue =====
        const int wi = warp - 2;                          // 0..7
        const int wt = wi * 32 + lane;                    // worker thread id 0..255
        const int q = warp & 3;                           // TMEM lane quarter of this warp
        const int half = wi >> 2;                         // column half handled in the epilogue
        uint32_t a_empty_phase = 0, tmem_phase = 0;
        int chunk_count = 0;
        const int cin_last = p.Cin - (p.kchunks - 1) * F_KC;            // channels in the last chunk (<= 64)
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const int b = tile / per_b;
            const int rem = tile - b * per_b;
            const int ph = rem / (p.mtiles * p.ntile_n);
            const int rem2 = rem - ph * (p.mtiles * p.ntile_n);
            const int mt = rem2 / p.ntile_n;
            const int nt = rem2 - mt * p.ntile_n;
            const int m0 = mt * MT;
            const int n0 = nt * p.bn;
            const float* xb = p.x + (size_t)b * p.Tin * p.ld_x;
            const float* ca = p.coef + (size_t)b * 2 * p.coef_ld;
            const float* cb = ca + p.coef_ld;
            // ---------------- transform, one 64-channel chunk at a time
            for (int kc = 0; kc < p.kchunks; ++kc) {
                const int cch = (kc == p.kchunks - 1) ? cin_last : F_KC;      // channels in this chunk
                const int lpr = cch >> 2;                                       // lanes (float4) per row: 8 or 16
                const int rows_per_pass = (FW * 32) / lpr;
                const int rl = wt / lpr;                                        // row within a pass
                const int c4 = (wt - rl * lpr) * 4;                             // channel (within chunk) of this thread
                const int cg = kc * F_KC + c4;
                const float4 a4 = *reinterpret_cast<const float4*>(ca + cg);
                const float4 b4 = *reinterpret_cast<const float4*>(cb + cg);
                float4 al = make_float4(1.f, 1.f, 1.f, 1.f), ia = al;
                if (ACT == ACT_SNAKE) {
                    al = *reinterpret_cast<const float4*>(p.alpha + cg);
                    ia = make_float4(1.f / al.x, 1.f / al.y, 1.f / al.z, 1.f / al.w);
                } else if (ACT == ACT_LRELU) {
                    al = make_float4(p.slope, p.slope, p.slope, p.slope);
                }
                if (chunk_count > 0) {                                         // previous chunk's MMAs done reading A
                    mbar_wait(a_empty, a_empty_phase);
                    a_empty_phase ^= 1;
                }
                ++chunk_count;
                const int t_base = m0 + p.halo_min;
                const uint32_t cidx = (uint32_t)(c4 >> 3);                      // 16-byte chunk within the 128-byte row
                const uint32_t sub = (uint32_t)(c4 & 4) * 2;                    // 0 or 8 bytes
                constexpr int U = 4;
                for (int r0 = rl; r0 < p.rows; r0 += rows_per_pass * U) {
                    float4 v[U];
                    bool ok[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int r = r0 + u * rows_per_pass;
                        const int t = t_base + r;
                        ok[u] = (r < p.rows) && (t >= 0) && (t < p.Tin);
                        if (ok[u]) v[u] = __ldg(reinterpret_cast<const float4*>(xb + (size_t)t * p.ld_x + cg));
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int r = r0 + u * rows_per_pass;
                        if (r >= p.rows) continue;
                        uint2 o = make_uint2(0u, 0u);                           // conv zero padding
                        if (ok[u]) {
                            float y0 = fused_act<ACT>(fmaf(a4.x, v[u].x, b4.x), al.x, ia.x);
                            float y1 = fused_act<ACT>(fmaf(a4.y, v[u].y, b4.y), al.y, ia.y);
                            float y2 = fused_act<ACT>(fmaf(a4.z, v[u].z, b4.z), al.z, ia.z);
                            float y3 = fused_act<ACT>(fmaf(a4.w, v[u].w, b4.w), al.w, ia.w);
                            o.x = pack16(y0, y1, p.is_bf16);
                            o.y = pack16(y2, y3, p.is_bf16);
                        }
                        const uint32_t off = (uint32_t)r * 128u + ((cidx ^ ((uint32_t)r & 7u)) << 4) + sub;
                        *reinterpret_cast<uint2*>(smem_a + off) = o;
                    }
                }
                fence_proxy_async();            // generic-proxy smem writes -> visible to the tensor core (async proxy)
                tc_fence_before();              // earlier tcgen05.ld of this thread ordered before the MMA warp's next MMAs
                __syncwarp();
                if (lane == 0) mbar_arrive(a_full);
            }
            // ---------------- epilogue
            mbar_wait(tmem_full, tmem_phase);
            tmem_phase ^= 1;
            tc_fence_after();
            float* stg = staging + wi * (32 * 16);
            const int ncol_half = p.bn >> 1;
            const int rr = lane >> 2;                      // 0..7: row within an 8-row pass (coalesced phase)
            const int c4o = (lane & 3) * 4;                // column offset within the 16-column chunk
            for (int r = 0; r < p.R; ++r) {
                for (int cc = half * ncol_half; cc < (half + 1) * ncol_half; cc += 16) {
                    float v[16];
                    tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(r * p.bn + cc), v);
                    // staging[row = lane][16] with 16-byte slots XOR-swizzled by (row>>1)&3: conflict-free both ways
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int slot = i ^ ((lane >> 1) & 3);
                        *reinterpret_cast<float4*>(stg + lane * 16 + slot * 4) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                    }
                    __syncwarp();
                    const int co = n0 + cc + c4o;
                    float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (p.bias != nullptr && co < p.Cout) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + co));
                    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int it = 0; it < 4; ++it) {
                        const int row = it * 8 + rr;                          // row within this warp's 32 lanes
                        const int m = m0 + r * 128 + q * 32 + row;
                        const int t = m * p.out_stride + ph - p.out_pad;
                        const int slot = (lane & 3) ^ ((row >> 1) & 3);
                        float4 a = *reinterpret_cast<const float4*>(stg + row * 16 + slot * 4);
                        if (m >= p.M || t < 0 || t >= p.Tout || co >= p.Cout) continue;
                        a.x += bias4.x; a.y += bias4.y; a.z += bias4.z; a.w += bias4.w;
                        const int nrep = (p.mirror && t == 2) ? 2 : 1;
                        for (int rep = 0; rep < nrep; ++rep) {
                            const int tt = rep == 0 ? t : 0;
                            float4 o = a;
                            if (p.res != nullptr) {
                                const float4 rv = *reinterpret_cast<const float4*>(
                                    p.res + ((size_t)b * (p.Tout >> p.res_shift) + (tt >> p.res_shift)) * p.ld_res + co);
                                o.x += rv.x; o.y += rv.y; o.z += rv.z; o.w += rv.w;
                            }
                            float* yp = p.y + ((size_t)b * p.Tout + tt) * p.ld_y + co;
                            if (p.accumulate) {
                                const float4 old = *reinterpret_cast<const float4*>(yp);
                                o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                            }
                            o.x *= p.scale; o.y *= p.scale; o.z *= p.scale; o.w *= p.scale;
                            *reinterpret_cast<float4*>(yp) = o;
                            s1[0] += o.x; s1[1] += o.y; s1[2] += o.z; s1[3] += o.w;
                            s2[0] = fmaf(o.x, o.x, s2[0]); s2[1] = fmaf(o.y, o.y, s2[1]);
                            s2[2] = fmaf(o.z, o.z, s2[2]); s2[3] = fmaf(o.w, o.w, s2[3]);
                        }
                    }
                    if (p.stats != nullptr) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
#pragma unroll
                            for (int o = 4; o < 32; o <<= 1) {
                                s1[i] += __shfl_xor_sync(0xffffffffu, s1[i], o);
                                s2[i] += __shfl_xor_sync(0xffffffffu, s2[i], o);
                            }
                        }
                        if (lane < 4) {
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                tstats[((size_t)q * p.R + r) * p.bn + cc + c4o + i] = make_float2(s1[i], s2[i]);
                        }
                    }
                    __syncwarp();
                }
            }
            tc_fence_before();
            if (p.stats != nullptr) {
                worker_bar_sync();
                for (int c = wt; c < p.bn; c += FW * 32) {
                    float a = 0.f, s = 0.f;
                    for (int qq = 0; qq < 4; ++qq)
                        for (int r = 0; r < p.R; ++r) {
                            const float2 v = tstats[((size_t)qq * p.R + r) * p.bn + c];
                            a += v.x;
                            s += v.y;
                        }
                    if (n0 + c < p.Cout)
                        p.stats[((size_t)b * (p.phases * p.mtiles) + (size_t)ph * p.mtiles + mt) * p.Cout + n0 + c] = make_float2(a, s);
                }
                worker_bar_sync();
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                     : "memory");
    }
}

// ---------------------------------------------------------------- coefficients from float2 tile partials
__global__ void adain_coef_f2_kernel(const float2* __restrict__ partial, int nparts, const float* __restrict__ h, int ld_h,
                                     int h_off, float* __restrict__ coef, int T, int C, int Cpad) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (c >= Cpad) return;
    float a = 0.f, bb = 0.f;
    if (c < C) {
        double s = 0, ss = 0;
        for (int i = 0; i < nparts; ++i) {
            const float2 v = partial[((size_t)b * nparts + i) * C + c];
            s += (double)v.x;
            ss += (double)v.y;
        }
        const double mean = s / (double)T;
        double var = ss / (double)T - mean * mean;
        if (var < 0) var = 0;
        const double rstd = 1.0 / sqrt(var + 1e-5);
        const double gamma = (double)h[(size_t)b * ld_h + h_off + c];
        const double beta = (double)h[(size_t)b * ld_h + h_off + C + c];
        const double ad = (1.0 + gamma) * rstd;
        a = (float)ad;
        bb = (float)(beta - mean * ad);
    }
    coef[((size_t)b * 2 + 0) * Cpad + c] = a;
    coef[((size_t)b * 2 + 1) * Cpad + c] = bb;
}

int launch_adain_coef_f2(const void* partial, int nparts, const float* h, int ld_h, int h_off, float* coef, int B, int T,
                         int C, int Cpad, cudaStream_t st) {
    dim3 grid(cdiv(Cpad, 128), B);
    adain_coef_f2_kernel<<<grid, 128, 0, st>>>((const float2*)partial, nparts, h, ld_h, h_off, coef, T, C, Cpad);
    ST2_LAUNCH_CHECK();
    return ST2_OK;
}

// ---------------------------------------------------------------- host side
int make_weight_map(CUtensorMap* map, int is_bf16, const void* w16, int cin_pad, int cout_pad, int ktaps, int bn);

static int fused_R(int cout_pad) { return cout_pad <= 64 ? 2 : 1; }

int fused_stats_parts(const ConvArgs& a) {
    const int MT = 128 * fused_R(a.w16_cout_pad);
    return a.phases * cdiv(a.M, MT);
}

bool conv_fused_supported(const ConvArgs& a) {
    if (a.in_stride != 1 || a.w16 == nullptr || a.w16_cin_pad % F_KC != 0 || a.w16_cout_pad % 32 != 0) return false;
    if (a.Cin % 4 != 0 || a.ld_x % 4 != 0 || a.Cout % 4 != 0 || a.ld_y % 4 != 0) return false;
    if (a.res != nullptr && a.ld_res % 4 != 0) return false;
    const int last = a.Cin - (a.w16_cin_pad / F_KC - 1) * F_KC;
    if (last != 32 && last != 64) return false;          // lanes-per-row must divide 256
    const int span = (a.ntaps - 1) * (a.tap_step < 0 ? -a.tap_step : a.tap_step);
    return span <= 64;
}

int launch_conv_fused(const ConvArgs& a, const float* coef, int coef_ld, int act, float slope, const float* alpha,
                      void* stats_out, cudaStream_t st) {
    ST2_REQUIRE(conv_fused_supported(a), "conv_fused: unsupported geometry");
    FusedParams p;
    memset(&p, 0, sizeof(p));
    p.x = a.x; p.ld_x = a.ld_x; p.Tin = a.Tin;
    p.coef = coef; p.coef_ld = coef_ld; p.alpha = alpha; p.slope = slope;
    p.Cin = a.Cin; p.kchunks = a.w16_cin_pad / F_KC;
    p.B = a.B; p.M = a.M; p.Tout = a.Tout; p.Cout = a.Cout;
    p.ntaps = a.ntaps; p.tap_step = a.tap_step; p.in_off = a.in_off;
    p.phases = a.phases; p.w_step = a.w_step; p.out_stride = a.out_stride; p.out_pad = a.out_pad;
    const int span = (a.ntaps - 1) * (a.tap_step < 0 ? -a.tap_step : a.tap_step);
    p.halo_min = a.in_off + (a.tap_step < 0 ? (a.ntaps - 1) * a.tap_step : 0);
    p.R = fused_R(a.w16_cout_pad);
    const int MT = 128 * p.R;
    p.rows = MT + span;
    int bn = a.w16_cout_pad;
    if (bn > 256) {
        bn = 256;
        while (a.w16_cout_pad % bn != 0) bn -= 32;
    }
    p.bn = bn;
    p.ntile_n = a.w16_cout_pad / bn;
    p.mtiles = cdiv(a.M, MT);
    p.num_tiles = a.B * a.phases * p.mtiles * p.ntile_n;
    int cols = 32;
    while (cols < p.R * bn) cols <<= 1;
    p.tmem_cols = cols;
    p.is_bf16 = a.fmt16 == DT_BF16 ? 1 : 0;
    p.bias = a.bias; p.res = a.res; p.ld_res = a.ld_res; p.res_shift = a.res_shift;
    p.y = a.y; p.ld_y = a.ld_y; p.scale = a.scale; p.accumulate = a.accumulate; p.mirror = a.mirror;
    p.stats = (float2*)stats_out;
    ST2_REQUIRE(p.ntile_n == 1 || stats_out == nullptr || true, "conv_fused: internal");
    const size_t a_bytes = ((size_t)p.rows * 128 + 1023) & ~(size_t)1023;
    const size_t b_stage = (size_t)bn * 128;
    const size_t fixed = a_bytes + FW * 32 * 16 * 4 + (size_t)4 * p.R * bn * 8 + 64 * 8 + 1024;
    int stages = (int)((110 * 1024 - fixed) / b_stage);       // keep two CTAs per SM
    if (stages > 8) stages = 8;
    if (stages < 2) stages = 2;
    p.stages = stages;
    const size_t smem = fixed + (size_t)stages * b_stage;
    CUtensorMap map_b;
    const int ktaps_total = a.phases > 1 ? a.ntaps * a.phases : a.ntaps;
    int e = make_weight_map(&map_b, p.is_bf16, a.w16, a.w16_cin_pad, a.w16_cout_pad, ktaps_total, bn);
    if (e != ST2_OK) return e;
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        ST2_CUDA_CHECK(cudaFuncSetAttribute(conv_fused_kernel<ACT_NONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        ST2_CUDA_CHECK(cudaFuncSetAttribute(conv_fused_kernel<ACT_LRELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        ST2_CUDA_CHECK(cudaFuncSetAttribute(conv_fused_kernel<ACT_SNAKE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
    int grid = 2 * num_sms;
    if (grid > p.num_tiles) grid = p.num_tiles;
    switch (act) {
        case ACT_NONE: conv_fused_kernel<ACT_NONE><<<grid, F_THREADS, smem, st>>>(map_b, p); break;
        case ACT_LRELU: conv_fused_kernel<ACT_LRELU><<<grid, F_THREADS, smem, st>>>(map_b, p); break;
        case ACT_SNAKE:
            ST2_REQUIRE(alpha != nullptr, "conv_fused: snake needs alpha");
            conv_fused_kernel<ACT_SNAKE><<<grid, F_THREADS, smem, st>>>(map_b, p);
            break;
        default: set_error("conv_fused: bad act %d", act); return ST2_ERR_INVALID;
    }
    ST2_LAUNCH_CHECK();
    return ST2_OK;
}

}  // namespace st2
