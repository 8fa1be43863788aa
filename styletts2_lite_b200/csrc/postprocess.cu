// SURVEY.md section 8(f) N4 -- what happens to the decoder's waveforms after the path: StyleTTS2.generate
// (inference.py:314-319) drops 4000 samples at both ends of every sentence ("weird pulse and silent tokens"), concatenates the
// sentences and pads 4000 zeros on both sides (np.zeros -> the result is float64); Demo/infer.py:51-54 divides by the peak
// (float64) and writes 24 kHz audio with soundfile, whose default WAV subtype is PCM_16: libsndfile converts a double x with
// lrint(x * 0x7FFF) (pcm.c d2s_array, normalisation on for float input; soundfile 0.13.1 / libsndfile 1.2.2, uv.lock:1925).
// Here that is two kernels over the B sentence waveforms of one decoder batch, bit-exact: the peak is an exact fp32 max, the
// division and the scale are IEEE fp64 operations, lrint is round-half-even.
#include "common.cuh"

namespace st2 {

static constexpr int PP_THREADS = 256;

// kept length of sentence i: max(len_i - 2*trim, 0)   (numpy slicing wav[trim:-trim] of a shorter array is empty)
__device__ __forceinline__ int pp_kept(const int32_t* lengths, int S, int trim, int i) {
    const int len = lengths ? lengths[i] : S;
    const int k = len - 2 * trim;
    return k > 0 ? k : 0;
}

// peak_bits <- max |wav_i[trim : len_i - trim]| over all sentences (bit pattern of a non-negative float orders like an integer);
// block 0 also writes the exclusive prefix sums of the kept lengths and the total output length
__global__ void __launch_bounds__(PP_THREADS)
post_peak_kernel(const float* __restrict__ wav, const int32_t* __restrict__ lengths, int B, int S, int trim, int pad,
                 unsigned int* __restrict__ peak_bits, int64_t* __restrict__ offsets, int64_t* __restrict__ total) {
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
        int64_t o = pad;
        for (int i = 0; i < B; ++i) {
            offsets[i] = o;
            o += pp_kept(lengths, S, trim, i);
        }
        offsets[B] = o;
        *total = o + pad;
    }
    const int i = blockIdx.y;
    if (i >= B) return;                            // B == 0: the grid still has one row of blocks for the offsets
    const int kept = pp_kept(lengths, S, trim, i);
    const float* src = wav + (size_t)i * S + trim;
    float m = 0.f;
    for (int n = blockIdx.x * PP_THREADS + threadIdx.x; n < kept; n += gridDim.x * PP_THREADS) m = fmaxf(m, fabsf(src[n]));
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    __shared__ float sm[PP_THREADS / 32];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < PP_THREADS / 32; ++w) m = fmaxf(m, sm[w]);
        if (m > 0.f) atomicMax(peak_bits, __float_as_uint(m));
    }
}

// out_f64[n] = wav / peak (float64, the array Demo/infer.py hands to soundfile), pcm[n] = lrint(out * 32767)
__global__ void __launch_bounds__(PP_THREADS)
post_write_kernel(const float* __restrict__ wav, const int32_t* __restrict__ lengths, int B, int S, int trim, int pad,
                  const unsigned int* __restrict__ peak_bits, const int64_t* __restrict__ offsets, double* __restrict__ out_f64,
                  int16_t* __restrict__ pcm) {
    const double peak = (double)__uint_as_float(*peak_bits);
    const int i = blockIdx.y;
    if (i == B) {
        // the two zero pads (0 / peak = 0 for any non-zero peak; an all-zero input stays all zero instead of NaN)
        const int64_t tail0 = offsets[B];
        for (int n = blockIdx.x * PP_THREADS + threadIdx.x; n < 2 * pad; n += gridDim.x * PP_THREADS) {
            const int64_t dst = n < pad ? n : tail0 + (n - pad);
            if (out_f64) out_f64[dst] = 0.0;
            if (pcm) pcm[dst] = 0;
        }
        return;
    }
    const int kept = pp_kept(lengths, S, trim, i);
    const float* src = wav + (size_t)i * S + trim;
    const int64_t o = offsets[i];
    for (int n = blockIdx.x * PP_THREADS + threadIdx.x; n < kept; n += gridDim.x * PP_THREADS) {
        const double r = peak > 0.0 ? __ddiv_rn((double)src[n], peak) : 0.0;
        if (out_f64) out_f64[o + n] = r;
        if (pcm) pcm[o + n] = (int16_t)__double2int_rn(__dmul_rn(r, 32767.0));
    }
}

}  // namespace st2

using namespace st2;

extern "C" {

int64_t st2_postprocess_scratch_bytes(int32_t B) {
    if (B < 0) return ST2_ERR_INVALID;
    return 256 + (int64_t)(B + 2) * 8;
}

int64_t st2_postprocess_max_samples(int32_t B, int32_t S, int32_t trim, int32_t pad) {
    if (B < 0 || S < 0 || trim < 0 || pad < 0) return ST2_ERR_INVALID;
    const int64_t kept = S > 2 * trim ? S - 2 * trim : 0;
    return (int64_t)B * kept + 2 * (int64_t)pad;
}

int st2_postprocess(const float* wav, const int32_t* lengths, int32_t B, int32_t S, int32_t trim, int32_t pad, double* out_f64,
                    int16_t* out_pcm, int64_t* out_total, void* scratch, void* stream) {
    ST2_REQUIRE(B >= 0 && S >= 0 && trim >= 0 && pad >= 0, "postprocess: negative size");
    ST2_REQUIRE(out_total != nullptr && scratch != nullptr && (B == 0 || wav != nullptr), "postprocess: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned int* peak = (unsigned int*)scratch;
    int64_t* offsets = (int64_t*)((char*)scratch + 256);
    ST2_CUDA_CHECK(cudaMemsetAsync(peak, 0, 4, st));
    const int nb = S > 0 ? (cdiv(S, PP_THREADS) < 64 ? cdiv(S, PP_THREADS) : 64) : 1;
    post_peak_kernel<<<dim3(nb, B > 0 ? B : 1), PP_THREADS, 0, st>>>(wav, lengths, B, S, trim, pad, peak, offsets, out_total);
    ST2_LAUNCH_CHECK();
    if (out_f64 != nullptr || out_pcm != nullptr) {
        post_write_kernel<<<dim3(nb, B + 1), PP_THREADS, 0, st>>>(wav, lengths, B, S, trim, pad, peak, offsets, out_f64, out_pcm);
        ST2_LAUNCH_CHECK();
    }
    return ST2_OK;
}

}  // extern "C"
