// Log-mel frontend kernels (reference: log_mel_spectrogram, audio.py:238-278; _stft :211-235;
// _build_mel_filterbank :41-80).
//
// Pass 1 (mel_logmel_kernel): persistent CTAs, one 32-frame tile of one utterance at a time.
//   reflect-padded framing -> symmetric Hann window -> 400-point real FFT (16 x 25, fp32,
//   see mel_fft.cuh) -> power -> banded mel filterbank -> log10(max(., 1e-10)) -> packed
//   (128, T_u) fp32 output + per-utterance running max (ordered-uint atomicMax).
// Pass 2 (mel_normalize_kernel): max(x, utt_max - 8) then (x + 4) / 4, in place.
//
// Layout: audio of B utterances packed back to back (sample_offsets[B+1]); mel of utterance
// u is a row-major (128, T_u) block starting at float offset 128 * frame_offsets[u].
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <vector>

#include "mel_fft.cuh"

namespace qasr {

constexpr int kMelNfft = 400;
constexpr int kMelHop = 160;
constexpr int kMelBins = 128;
constexpr int kMelFreqs = 201;
constexpr int kMelMaxTaps = 12;       // max non-zeros per filter row we store (reference: <= 9)
constexpr int kMelFramesPerCta = 32;
constexpr int kMelThreads = 288;      // 9 warps: the 32 x 9 DFT-25 tasks of a tile take exactly one round

struct MelTables {
  const float* window;     // [400]  symmetric Hann, float32 (np.hanning(400).astype(f32))
  const float2* twiddle;   // [9][25] W400^{n2*k1}
  const int* fb_start;     // [128] first non-zero frequency bin of each filter
  const int* fb_count;     // [128] number of stored taps
  const float* fb_weight;  // [128][kMelMaxTaps]
};

// ---------------------------------------------------------------- host: table construction
// Mirrors _build_mel_filterbank (audio.py:41-80): HTK mel formula, triangular filters on
// linspace(0, sr/2, 201), slopes rounded to float32, then divided by the filter width in Hz.
inline void build_mel_filterbank_host(std::vector<float>& fb /*128*201*/, int n_fft = 400, int n_mels = 128,
                                      double sr = 16000.0, double f_min = 0.0, double f_max = 8000.0) {
  const int n_freqs = n_fft / 2 + 1;
  fb.assign(static_cast<size_t>(n_mels) * n_freqs, 0.0f);
  std::vector<double> freqs(n_freqs), fpts(n_mels + 2);
  const double fstep = (sr / 2.0 - 0.0) / (n_freqs - 1);
  for (int k = 0; k < n_freqs; ++k) freqs[k] = k * fstep + 0.0;
  freqs[n_freqs - 1] = sr / 2.0;
  const double mel_min = 2595.0 * log10(1.0 + f_min / 700.0);
  const double mel_max = 2595.0 * log10(1.0 + f_max / 700.0);
  const double mstep = (mel_max - mel_min) / (n_mels + 1);
  for (int i = 0; i < n_mels + 2; ++i) {
    double m = i * mstep + mel_min;
    if (i == n_mels + 1) m = mel_max;
    fpts[i] = 700.0 * (pow(10.0, m / 2595.0) - 1.0);
  }
  for (int i = 0; i < n_mels; ++i) {
    const double fl = fpts[i], fc = fpts[i + 1], fr = fpts[i + 2];
    const double width = fr - fl;
    for (int k = 0; k < n_freqs; ++k) {
      const double up = (freqs[k] - fl) / (fc - fl);
      const double down = (fr - freqs[k]) / (fr - fc);
      const double v = fmax(0.0, fmin(up, down));
      float f32 = static_cast<float>(v);
      if (width > 0.0) f32 = static_cast<float>(static_cast<double>(f32) / width);
      fb[static_cast<size_t>(i) * n_freqs + k] = f32;
    }
  }
}

inline void build_hann_window_host(std::vector<float>& w, int n = 400) {
  // np.hanning(n): 0.5 - 0.5 cos(2 pi i / (n - 1)), computed in float64 then cast.
  w.resize(n);
  for (int i = 0; i < n; ++i) {
    // numpy evaluates hanning on n = arange(1-M, M, 2): 0.5 + 0.5*cos(pi*n/(M-1))
    const double nn = static_cast<double>(1 - n + 2 * i);
    w[i] = static_cast<float>(0.5 + 0.5 * cos(M_PI * nn / (n - 1)));
  }
}

inline void build_twiddle_host(std::vector<float2>& tw) {
  tw.resize(9 * 25);
  for (int k1 = 0; k1 < 9; ++k1)
    for (int n2 = 0; n2 < 25; ++n2) {
      const double a = -2.0 * M_PI * static_cast<double>(n2 * k1) / 400.0;
      tw[k1 * 25 + n2] = make_float2(static_cast<float>(cos(a)), static_cast<float>(sin(a)));
    }
}

// ---------------------------------------------------------------- device helpers
__device__ __forceinline__ unsigned float_to_ordered(float f) {
  unsigned b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}
// NaN-propagating maximum (numpy semantics); the canonical positive quiet NaN orders above +inf in float_to_ordered,
// so the per-utterance atomicMax keeps it.
#define kMelNaN __int_as_float(0x7FC00000)
__device__ __forceinline__ float nan_max(float a, float b) { return (a != a || b != b) ? kMelNaN : fmaxf(a, b); }
// np.pad(mode="reflect") index map (period 2(N-1)); valid for any integer i when N >= 2.
__device__ __forceinline__ long long reflect_index(long long i, long long n) {
  const long long period = 2 * (n - 1);
  long long m = i % period;
  if (m < 0) m += period;
  return (m < n) ? m : period - m;
}
// largest u in [0, B) with offs[u] <= v   (offs is non-decreasing, offs[0] = 0)
template <typename T>
__device__ __forceinline__ int upper_segment(const T* __restrict__ offs, int B, T v) {
  int lo = 0, hi = B;  // invariant: offs[lo] <= v < offs[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(offs + mid) <= v) lo = mid; else hi = mid;
  }
  return lo;
}

// Sample j of the tile's span lives at raw[j + j / 160]: one pad word per hop, so that the 32 frames of a tile (one per lane,
// 160 samples apart = bank stride 0) read their n-th sample from 32 different banks (stride 161).
constexpr int kMelRawPadded = (kMelFramesPerCta - 1) * (kMelHop + 1) + kMelNfft + kMelNfft / kMelHop + 1;  // 5394
static_assert(kMelFramesPerCta == 32 && kMelThreads == 288, "steps A / C / D map the 32 frames of a tile onto the 32 lanes of a warp");

// One mel bin of one frame: CNT taps of the banded filter against the power spectrum column of this lane's frame.  The tap
// count is warp-uniform (warp = mel bin), so the caller dispatches on it with a uniform switch: only the taps that exist
// are issued (3.1 on average against the 12 predicated slots of the generic loop, which made this step half of the
// kernel's instructions).  All loads first, then the FMA chain in ascending tap order (same summation order as before).
template <int CNT>
__device__ __forceinline__ float mel_band(const float* __restrict__ w, const float* __restrict__ p) {
  float pv[CNT], wv[CNT];
#pragma unroll
  for (int j = 0; j < CNT; ++j) {
    pv[j] = p[j * (kMelFramesPerCta + 1)];
    wv[j] = w[j];
  }
  float acc = 0.0f;
#pragma unroll
  for (int j = 0; j < CNT; ++j) acc = fmaf(wv[j], pv[j], acc);
  return acc;
}

struct MelSmem {
  union {  // the raw samples are dead once step A has produced Y; the power spectrum is written in step C
    float raw[kMelRawPadded];
    float P[kMelFreqs][kMelFramesPerCta + 1];
  };
  float window[kMelNfft];
  float2 twiddle[9 * 25];
  float2 Y[kMelFramesPerCta][9][25];
  float fb_weight[kMelBins * kMelMaxTaps];
  int fb_band[kMelBins];  // start | count << 16
  float red[(kMelThreads + 31) / 32];
  int next_tile;
};

// Where a tile (32 frames of one utterance) lives: resolved from the tile index by a binary search over the per-utterance tile
// offsets (warp-uniform, a handful of L1-resident loads).
struct MelTile {
  int u, T, t0, nf, span;
  long long N, f0, first;
  const float* x;
  __device__ __forceinline__ bool interior() const { return first >= 0 && first + span <= N; }
};
__device__ __forceinline__ MelTile mel_tile(int tile, const float* __restrict__ audio, const long long* __restrict__ sample_offsets,
                                            const long long* __restrict__ frame_offsets, const int* __restrict__ block_offsets, int B) {
  MelTile t;
  t.u = upper_segment<int>(block_offsets, B, tile);
  const long long s0 = __ldg(sample_offsets + t.u);
  t.N = __ldg(sample_offsets + t.u + 1) - s0;
  t.f0 = __ldg(frame_offsets + t.u);
  t.T = static_cast<int>(__ldg(frame_offsets + t.u + 1) - t.f0);
  t.t0 = (tile - __ldg(block_offsets + t.u)) * kMelFramesPerCta;
  t.nf = min(kMelFramesPerCta, t.T - t.t0);
  t.span = (t.nf - 1) * kMelHop + kMelNfft;
  t.first = static_cast<long long>(t.t0) * kMelHop - kMelNfft / 2;
  t.x = audio + s0;
  return t;
}

// A warp copies whole hops (160 samples -> 161 padded words of `raw`): hops warp, warp + 9, ... of the tile's span, five
// coalesced loads per hop.  The loads go to registers first, so that they can be issued one tile ahead (mel_prefetch) and land
// while the current tile is still computing; the tiles at the ends of an utterance reflect their out-of-range indices.
constexpr int kMelHopsPerWarp = ((kMelFramesPerCta - 1) * kMelHop + kMelNfft + kMelHop - 1) / kMelHop / (kMelThreads / 32) + 1;  // 4
struct MelPrefetch {
  float v[kMelHopsPerWarp][kMelHop / 32];
};
template <bool kReflect>
__device__ __forceinline__ void mel_prefetch_impl(const MelTile& t, MelPrefetch& r, int tid) {
#pragma unroll
  for (int h = 0; h < kMelHopsPerWarp; ++h) {
    const int hop = (tid >> 5) + h * (kMelThreads / 32);
    const int left = t.span - hop * kMelHop;
#pragma unroll
    for (int k = 0; k < kMelHop / 32; ++k) {
      const int i = (tid & 31) + 32 * k;
      long long j = t.first + hop * kMelHop + i;
      if (kReflect && (j < 0 || j >= t.N)) j = reflect_index(j, t.N);  // np.pad(mode="reflect"): only the ends of an utterance
      r.v[h][k] = (i < left) ? __ldg(t.x + j) : 0.0f;
    }
  }
}
__device__ __forceinline__ void mel_prefetch(const MelTile& t, MelPrefetch& r, int tid) {
  if (t.interior()) mel_prefetch_impl<false>(t, r, tid);
  else mel_prefetch_impl<true>(t, r, tid);
}
__device__ __forceinline__ void mel_store_prefetched(const MelTile& t, const MelPrefetch& r, float* __restrict__ raw, int tid) {
#pragma unroll
  for (int h = 0; h < kMelHopsPerWarp; ++h) {
    const int hop = (tid >> 5) + h * (kMelThreads / 32);
    const int left = t.span - hop * kMelHop;
#pragma unroll
    for (int k = 0; k < kMelHop / 32; ++k) {
      const int i = (tid & 31) + 32 * k;
      if (i < left) raw[hop * (kMelHop + 1) + i] = r.v[h][k];
    }
  }
}

// Persistent CTAs (two per SM): the tables are staged once per CTA, tiles beyond the first wave are handed out by an atomic
// counter (tiles at utterance ends cost less, reflected ones more), and the raw samples of the NEXT tile are fetched into
// registers between step A and step C of the current one (ncu before: a third of the stall samples sat on the shared-memory
// stores waiting for a tile's own global loads, and a sixth of the instructions re-staged the tables).
__global__ void __launch_bounds__(kMelThreads)
mel_logmel_kernel(const float* __restrict__ audio, const long long* __restrict__ sample_offsets,
                  const long long* __restrict__ frame_offsets, const int* __restrict__ block_offsets, int B,
                  MelTables tab, float* __restrict__ mel_out, unsigned* __restrict__ utt_max /* [B] maxima, [B] tile counter (0) */) {
  extern __shared__ __align__(16) uint8_t mel_smem_raw[];
  MelSmem& s = *reinterpret_cast<MelSmem*>(mel_smem_raw);
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  unsigned* __restrict__ tile_counter = utt_max + B;  // tiles beyond the first gridDim.x are handed out dynamically
  const int total_tiles = __ldg(block_offsets + B);

  for (int i = tid; i < kMelNfft; i += kMelThreads) s.window[i] = __ldg(tab.window + i);
  for (int i = tid; i < 9 * 25; i += kMelThreads) s.twiddle[i] = __ldg(tab.twiddle + i);
  for (int i = tid; i < kMelBins * kMelMaxTaps; i += kMelThreads) s.fb_weight[i] = __ldg(tab.fb_weight + i);
  for (int i = tid; i < kMelBins; i += kMelThreads) s.fb_band[i] = __ldg(tab.fb_start + i) | (__ldg(tab.fb_count + i) << 16);

  MelPrefetch pre;
  int tile = static_cast<int>(blockIdx.x);
  if (tile < total_tiles) mel_prefetch(mel_tile(tile, audio, sample_offsets, frame_offsets, block_offsets, B), pre, tid);
  while (tile < total_tiles) {
  const MelTile tl = mel_tile(tile, audio, sample_offsets, frame_offsets, block_offsets, B);
  const int u = tl.u, T = tl.T, t0 = tl.t0, nf = tl.nf;
  const long long f0 = tl.f0;

  // ---- stage the reflect-padded sample span of these frames
  mel_store_prefetched(tl, pre, s.raw, tid);
  __syncthreads();

  // Every step maps the tile's frames onto lanes (lane = frame) and one unit of work onto a warp, so that table lookups
  // (window, twiddles, filter taps) are warp-uniform broadcasts, branches are uniform and every shared-memory access of a
  // warp hits 32 different banks (ncu on the previous task-major mapping: 30 % of the LSU wavefronts were bank conflicts).
  // ---- step A/B: 25 real 16-point DFTs per frame, twiddled; warp = input residue n2 (3 rounds of 9 warps)
  for (int n2 = warp; n2 < 25; n2 += kMelThreads / 32) {
    if (lane < nf) {
      // sample j = 25 n1 + n2 of the frame sits j / 160 pad words further (see kMelRawPadded); only n1 = 6 and n1 = 12
      // straddle a hop boundary depending on n2, every other offset is a compile-time immediate
      const float* fr = s.raw + lane * (kMelHop + 1) + n2;
      const float* wn = s.window + n2;
      const int p6 = (n2 >= 10) ? 1 : 0, p12 = (n2 >= 20) ? 2 : 1;
      static_assert(kMelHop == 160 && kMelNfft == 400, "pad offsets below are written out for hop 160, n_fft 400");
      float v[16];
#pragma unroll
      for (int n1 = 0; n1 < 16; ++n1) {
        const int pad = (n1 < 6) ? 0 : (n1 == 6) ? p6 : (n1 < 12) ? 1 : (n1 == 12) ? p12 : 2;
        v[n1] = fr[25 * n1 + pad] * wn[25 * n1];
      }
      fft::cf y[9];
      fft::rdft16(v, y);
#pragma unroll
      for (int k1 = 0; k1 < 9; ++k1) {
        const float2 w = s.twiddle[k1 * 25 + n2];
        s.Y[lane][k1][n2] = make_float2(y[k1].re * w.x - y[k1].im * w.y, y[k1].re * w.y + y[k1].im * w.x);
      }
    }
  }
  if (tid == 0) s.next_tile = static_cast<int>(gridDim.x + atomicAdd(tile_counter, 1u));  // tiles differ in cost (ragged ends, reflection)
  __syncthreads();

  // ---- the raw samples are dead from here (P aliases them): fetch the next tile's into registers; they land during step C
  // and the filterbank
  const int next = s.next_tile;
  if (next < total_tiles) mel_prefetch(mel_tile(next, audio, sample_offsets, frame_offsets, block_offsets, B), pre, tid);

  // ---- step C: 9 complex 25-point DFTs per frame -> power spectrum P[bin][frame]; warp = k1 (one round)
  if (lane < nf) {
    const int f = lane, k1 = warp;
    fft::cf z[25];
#pragma unroll
    for (int i = 0; i < 25; ++i) {
      const float2 t = s.Y[f][k1][i];
      z[i] = {t.x, t.y};
    }
    fft::dft25(z);
    const bool self_conj = (k1 == 0) || (k1 == 8);
#pragma unroll
    for (int k2 = 0; k2 < 25; ++k2) {
      const int k = k1 + 16 * k2;
      const float pw = z[k2].re * z[k2].re + z[k2].im * z[k2].im;
      if (k <= 200) s.P[k][f] = pw;
      else if (!self_conj) s.P[400 - k][f] = pw;
    }
  }
  __syncthreads();

  // ---- banded mel filterbank + log10; warp = mel bin (15 rounds of 9 warps), lane = frame (coalesced 128-byte rows).
  // All taps of a bin are loaded before the FMA chain starts (the tap count is warp-uniform), so a round costs one
  // shared-memory latency instead of one per tap; the summation order is unchanged.
  float lmax = -INFINITY;
  // this lane's output column; advanced by one warp-round of mel rows per iteration (no 64-bit multiply per bin)
  float* __restrict__ out = mel_out + f0 * kMelBins + static_cast<long long>(warp) * T + t0 + lane;
  const long long out_step = static_cast<long long>(kMelThreads / 32) * T;
  for (int m = warp; m < kMelBins; m += kMelThreads / 32, out += out_step) {
    if (lane < nf) {
      const int band = s.fb_band[m];
      const int st = band & 0xFFFF, cnt = band >> 16;
      const float* w = s.fb_weight + m * kMelMaxTaps;
      const float* pcol = &s.P[st][lane];
      float acc = 0.0f;
      static_assert(kMelMaxTaps == 12, "the dispatch below lists every tap count");
      switch (cnt) {
        case 1: acc = mel_band<1>(w, pcol); break;
        case 2: acc = mel_band<2>(w, pcol); break;
        case 3: acc = mel_band<3>(w, pcol); break;
        case 4: acc = mel_band<4>(w, pcol); break;
        case 5: acc = mel_band<5>(w, pcol); break;
        case 6: acc = mel_band<6>(w, pcol); break;
        case 7: acc = mel_band<7>(w, pcol); break;
        case 8: acc = mel_band<8>(w, pcol); break;
        case 9: acc = mel_band<9>(w, pcol); break;
        case 10: acc = mel_band<10>(w, pcol); break;
        case 11: acc = mel_band<11>(w, pcol); break;
        case 12: acc = mel_band<12>(w, pcol); break;
        default: break;
      }
      // log10(x) = log2(x) * log10(2); lg2.approx is accurate to ~2^-22 relative, i.e. < 2e-6 absolute here.
      // A non-finite sample makes its frames NaN in the reference (np.maximum / .max() propagate NaN, audio.py:274-275)
      // and, through the utterance-wide max, the whole utterance: NaN is carried, not dropped by fmaxf.
      // the argument is >= 1e-10, a normal number: the .ftz form gives the same bits as __log2f without its subnormal rescue
      float lg;
      asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(fmaxf(acc, 1e-10f)));
      const float v = (acc != acc) ? kMelNaN : lg * 0.30102999566398120f;
      *out = v;
      lmax = nan_max(lmax, v);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lmax = nan_max(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
  if ((tid & 31) == 0) s.red[tid >> 5] = lmax;
  __syncthreads();
  if (tid == 0) {
    float m = s.red[0];
#pragma unroll
    for (int i = 1; i < (kMelThreads + 31) / 32; ++i) m = nan_max(m, s.red[i]);
    atomicMax(utt_max + u, float_to_ordered(m));
  }
  __syncthreads();  // the next tile overwrites raw / P and the reduction scratch
  tile = next;
  }  // tile loop
}

// Pass 2: x = (max(x, utt_max - 8) + 4) / 4 over the packed mel buffer (float4 granularity:
// every utterance block is a multiple of 128 floats).
constexpr int kMelNormThreads = 256;
constexpr int kMelNormVecPerCta = 2048;  // float4 per CTA

__global__ void __launch_bounds__(kMelNormThreads)
mel_normalize_kernel(float* __restrict__ mel, const long long* __restrict__ frame_offsets, int B,
                     const unsigned* __restrict__ utt_max, long long total_vec4) {
  const long long base = static_cast<long long>(blockIdx.x) * kMelNormVecPerCta;
  // utterance of the first element of this CTA (element e belongs to u iff 128*fo[u] <= 4e < 128*fo[u+1])
  int u = upper_segment<long long>(frame_offsets, B, (base * 4) / kMelBins);
  long long end_vec = __ldg(frame_offsets + u + 1) * (kMelBins / 4);
  float thr = ordered_to_float(__ldg(utt_max + u)) - 8.0f;
  float4* __restrict__ p = reinterpret_cast<float4*>(mel);
  for (int i = threadIdx.x; i < kMelNormVecPerCta; i += kMelNormThreads) {
    const long long e = base + i;
    if (e >= total_vec4) break;
    while (e >= end_vec) {
      ++u;
      end_vec = __ldg(frame_offsets + u + 1) * (kMelBins / 4);
      thr = ordered_to_float(__ldg(utt_max + u)) - 8.0f;
    }
    float4 v = p[e];
    if (thr != thr) {  // a NaN anywhere in the utterance: np.maximum(log_spec, nan) is NaN everywhere (audio.py:275)
      v = make_float4(thr, thr, thr, thr);
    } else {
      v.x = (fmaxf(v.x, thr) + 4.0f) * 0.25f;
      v.y = (fmaxf(v.y, thr) + 4.0f) * 0.25f;
      v.z = (fmaxf(v.z, thr) + 4.0f) * 0.25f;
      v.w = (fmaxf(v.w, thr) + 4.0f) * 0.25f;
    }
    p[e] = v;
  }
}

}  // namespace qasr
