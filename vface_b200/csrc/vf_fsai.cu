// Frequency Spectrum Attention Interpolation (FSAI), scripts/face_swap_utils.py:425-464.
//
// Reference: out = Re ifft([fft(dst)[:split], fft(donor)[split:]]) along the channel axis (length d)
// of every token row.  The op is linear and real:
//     out = dst + Re ifft( mask_hi * fft(donor - dst) ),     mask_hi[k] = [k >= split]
// and, because the input is real, taking the real part equals filtering with the real symmetric
// response h[k] = 0.5 * (mask_hi[k] + mask_hi[(d-k) % d]).  A real symmetric response commutes with
// packing two real rows into one complex row, so ONE complex FFT of length d serves two outputs:
//     z = (donor - dst_a) + i (donor - dst_b);   y = ifft(h * fft(z));
//     out_a = dst_a + Re y;  out_b = dst_b + Im y.
// In the fused form (vf_fsai_blend2) the pair is the two branches of the same token (they share the
// donor, ldm/models/pnp_utils.py:195+198), in the single form it is two consecutive rows.
//
// Kernel: persistent CTAs; each batch of token pairs is staged into shared memory with coalesced
// 16-byte loads, transformed by a mixed-radix (5,4,2) Stockham FFT in shared memory (fp32), filtered,
// inverse-transformed (conjugate trick) and written back with 16-byte stores.  Twiddles and the
// filter response live in shared memory, built once per CTA.
//
// HBM roofline: algorithmic bytes per row pair = 6*d*e (single) / 5*d*e (fused).
#include "vf_common.cuh"

#include <cstdlib>
#include "vf_fsai_fast.cuh"

namespace vf {

constexpr int kFsaiThreads = 256;
constexpr int kMaxStages = 8;

struct FsaiParams {
  const void* donor;
  const void* dst_a;
  void* out_a;
  const void* dst_b;   // fused mode only
  void* out_b;
  long long rows;
  long long ld_donor, ld_a, ld_out_a, ld_b, ld_out_b;
  int d, split;
  int pairs_per_batch;
  int n_stages;
  int radix[kMaxStages];
  int fused;           // 1: pair = (branch a, branch b) of one row; 0: pair = rows (2p, 2p+1)
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// multiply by -i
__device__ __forceinline__ float2 cmul_mi(float2 a) { return make_float2(a.y, -a.x); }

template <int R> __device__ __forceinline__ void dft(float2 (&v)[R]);

template <> __device__ __forceinline__ void dft<2>(float2 (&v)[2]) {
  float2 a = v[0], b = v[1];
  v[0] = cadd(a, b);
  v[1] = csub(a, b);
}
template <> __device__ __forceinline__ void dft<4>(float2 (&v)[4]) {
  float2 a = cadd(v[0], v[2]), b = csub(v[0], v[2]);
  float2 c = cadd(v[1], v[3]), d = cmul_mi(csub(v[1], v[3]));
  v[0] = cadd(a, c);
  v[1] = cadd(b, d);
  v[2] = csub(a, c);
  v[3] = csub(b, d);
}
template <> __device__ __forceinline__ void dft<5>(float2 (&v)[5]) {
  const float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f;
  const float s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;
  float2 t1 = cadd(v[1], v[4]), t2 = cadd(v[2], v[3]);
  float2 t3 = csub(v[1], v[4]), t4 = csub(v[2], v[3]);
  float2 m1 = make_float2(v[0].x + c1 * t1.x + c2 * t2.x, v[0].y + c1 * t1.y + c2 * t2.y);
  float2 m2 = make_float2(v[0].x + c2 * t1.x + c1 * t2.x, v[0].y + c2 * t1.y + c1 * t2.y);
  float2 n1 = make_float2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y);
  float2 n2 = make_float2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y);
  float2 in1 = cmul_mi(n1), in2 = cmul_mi(n2);   // -i * n
  v[0] = make_float2(v[0].x + t1.x + t2.x, v[0].y + t1.y + t2.y);
  v[1] = cadd(m1, in1);
  v[4] = csub(m1, in1);
  v[2] = cadd(m2, in2);
  v[3] = csub(m2, in2);
}

// One Stockham stage of radix R over `pairs` independent length-n transforms laid out back to back.
// kFilter: the input is first conjugated and multiplied by the (pre-scaled) response h, which turns
// the following forward stages into the inverse transform (ifft(X) = conj(fft(conj(X))) / n).
template <int R, bool kFilter>
__device__ __forceinline__ void stockham_stage(const float2* __restrict__ in, float2* __restrict__ out,
                                               const float2* __restrict__ tw, const float* __restrict__ hresp,
                                               int n, int p, int pairs) {
  const int T = n / R;
  const int tw_step = n / (p * R);
  const int items = T * pairs;
  for (int it = threadIdx.x; it < items; it += kFsaiThreads) {
    const int pr = it / T;
    const int i = it - pr * T;
    const int k = i % p;
    const float2* src = in + pr * n;
    float2* dstp = out + pr * n;
    float2 v[R];
#pragma unroll
    for (int t = 0; t < R; ++t) {
      float2 x = src[i + t * T];
      if (kFilter) {
        float hv = hresp[i + t * T];
        x = make_float2(x.x * hv, -x.y * hv);
      }
      if (t > 0) x = cmul(x, tw[t * k * tw_step]);
      v[t] = x;
    }
    dft<R>(v);
    const int j = (i - k) * R + k;
#pragma unroll
    for (int t = 0; t < R; ++t) dstp[j + t * p] = v[t];
  }
}

template <bool kFilter>
__device__ __forceinline__ void run_stage(int radix, const float2* in, float2* out, const float2* tw,
                                          const float* hresp, int n, int p, int pairs) {
  if (radix == 4) stockham_stage<4, kFilter>(in, out, tw, hresp, n, p, pairs);
  else if (radix == 5) stockham_stage<5, kFilter>(in, out, tw, hresp, n, p, pairs);
  else stockham_stage<2, kFilter>(in, out, tw, hresp, n, p, pairs);
}

template <typename T> struct Row16;   // a 16-byte chunk of a row as floats
template <> struct Row16<float> {
  static constexpr int E = 4;
  static __device__ __forceinline__ void ld(const float* p, float (&v)[4]) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    v[0] = __uint_as_float(u.x); v[1] = __uint_as_float(u.y); v[2] = __uint_as_float(u.z); v[3] = __uint_as_float(u.w);
  }
  static __device__ __forceinline__ void st(float* p, const float (&v)[4]) {
    *reinterpret_cast<uint4*>(p) = make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
  }
};
template <> struct Row16<__nv_bfloat16> {
  static constexpr int E = 8;
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    v[0] = bf16lo(u.x); v[1] = bf16hi(u.x); v[2] = bf16lo(u.y); v[3] = bf16hi(u.y);
    v[4] = bf16lo(u.z); v[5] = bf16hi(u.z); v[6] = bf16lo(u.w); v[7] = bf16hi(u.w);
  }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, const float (&v)[8]) {
    *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
  }
};

template <typename T>
__global__ void __launch_bounds__(kFsaiThreads)
fsai_kernel(const FsaiParams P) {
  constexpr int E = Row16<T>::E;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int d = P.d;
  const int ppb = P.pairs_per_batch;
  float2* tw = reinterpret_cast<float2*>(smem_raw);            // d
  float* hresp = reinterpret_cast<float*>(tw + d);             // d (scaled by 1/d)
  float2* buf0 = reinterpret_cast<float2*>(hresp + d);         // ppb * d
  float2* buf1 = buf0 + (size_t)ppb * d;                       // ppb * d

  for (int i = threadIdx.x; i < d; i += kFsaiThreads) {
    float s, c;
    sincospif(-2.0f * (float)i / (float)d, &s, &c);
    tw[i] = make_float2(c, s);
    float hv = 0.5f * ((i >= P.split ? 1.0f : 0.0f) + ((((d - i) % d) >= P.split) ? 1.0f : 0.0f));
    hresp[i] = hv / (float)d;
  }
  __syncthreads();

  const T* donor = reinterpret_cast<const T*>(P.donor);
  const T* dst_a = reinterpret_cast<const T*>(P.dst_a);
  const T* dst_b = reinterpret_cast<const T*>(P.dst_b);
  T* out_a = reinterpret_cast<T*>(P.out_a);
  T* out_b = reinterpret_cast<T*>(P.out_b);

  const long long n_pairs = P.fused ? P.rows : (P.rows + 1) / 2;
  const long long n_batches = (n_pairs + ppb - 1) / ppb;
  const int chunks = d / E;

  for (long long batch = blockIdx.x; batch < n_batches; batch += gridDim.x) {
    const long long pair0 = batch * ppb;
    const int pairs = (int)min((long long)ppb, n_pairs - pair0);

    // ---- stage in: z = (donor - dst_re) + i (donor' - dst_im) --------------------------------
    for (int it = threadIdx.x; it < pairs * chunks; it += kFsaiThreads) {
      const int pr = it / chunks, ch = it - pr * chunks;
      const long long pair = pair0 + pr;
      float re[E], im[E];
      if (P.fused) {
        float dn[E], a[E], b[E];
        Row16<T>::ld(donor + pair * P.ld_donor + ch * E, dn);
        Row16<T>::ld(dst_a + pair * P.ld_a + ch * E, a);
        Row16<T>::ld(dst_b + pair * P.ld_b + ch * E, b);
#pragma unroll
        for (int j = 0; j < E; ++j) { re[j] = dn[j] - a[j]; im[j] = dn[j] - b[j]; }
      } else {
        const long long r0 = 2 * pair, r1 = 2 * pair + 1;
        float dn[E], a[E];
        Row16<T>::ld(donor + r0 * P.ld_donor + ch * E, dn);
        Row16<T>::ld(dst_a + r0 * P.ld_a + ch * E, a);
#pragma unroll
        for (int j = 0; j < E; ++j) re[j] = dn[j] - a[j];
        if (r1 < P.rows) {
          Row16<T>::ld(donor + r1 * P.ld_donor + ch * E, dn);
          Row16<T>::ld(dst_a + r1 * P.ld_a + ch * E, a);
#pragma unroll
          for (int j = 0; j < E; ++j) im[j] = dn[j] - a[j];
        } else {
#pragma unroll
          for (int j = 0; j < E; ++j) im[j] = 0.0f;
        }
      }
      float2* zp = buf0 + (size_t)pr * d + ch * E;
#pragma unroll
      for (int j = 0; j < E; ++j) zp[j] = make_float2(re[j], im[j]);
    }
    __syncthreads();

    // ---- forward FFT, filter, inverse FFT ----------------------------------------------------
    float2* cur = buf0;
    float2* nxt = buf1;
    int p = 1;
    for (int s = 0; s < P.n_stages; ++s) {
      run_stage<false>(P.radix[s], cur, nxt, tw, hresp, d, p, pairs);
      p *= P.radix[s];
      __syncthreads();
      float2* t = cur; cur = nxt; nxt = t;
    }
    p = 1;
    for (int s = 0; s < P.n_stages; ++s) {
      if (s == 0) run_stage<true>(P.radix[s], cur, nxt, tw, hresp, d, p, pairs);
      else        run_stage<false>(P.radix[s], cur, nxt, tw, hresp, d, p, pairs);
      p *= P.radix[s];
      __syncthreads();
      float2* t = cur; cur = nxt; nxt = t;
    }
    // cur holds conj(y): Re y = cur.x, Im y = -cur.y

    // ---- stage out: out = dst + y --------------------------------------------------------------
    for (int it = threadIdx.x; it < pairs * chunks; it += kFsaiThreads) {
      const int pr = it / chunks, ch = it - pr * chunks;
      const long long pair = pair0 + pr;
      const float2* yp = cur + (size_t)pr * d + ch * E;
      float a[E], o[E];
      if (P.fused) {
        Row16<T>::ld(dst_a + pair * P.ld_a + ch * E, a);
#pragma unroll
        for (int j = 0; j < E; ++j) o[j] = a[j] + yp[j].x;
        Row16<T>::st(out_a + pair * P.ld_out_a + ch * E, o);
        Row16<T>::ld(dst_b + pair * P.ld_b + ch * E, a);
#pragma unroll
        for (int j = 0; j < E; ++j) o[j] = a[j] - yp[j].y;
        Row16<T>::st(out_b + pair * P.ld_out_b + ch * E, o);
      } else {
        const long long r0 = 2 * pair, r1 = 2 * pair + 1;
        Row16<T>::ld(dst_a + r0 * P.ld_a + ch * E, a);
#pragma unroll
        for (int j = 0; j < E; ++j) o[j] = a[j] + yp[j].x;
        Row16<T>::st(out_a + r0 * P.ld_out_a + ch * E, o);
        if (r1 < P.rows) {
          Row16<T>::ld(dst_a + r1 * P.ld_a + ch * E, a);
#pragma unroll
          for (int j = 0; j < E; ++j) o[j] = a[j] - yp[j].y;
          Row16<T>::st(out_a + r1 * P.ld_out_a + ch * E, o);
        }
      }
    }
    __syncthreads();
  }
}

// ---- fast path (vf_fsai_fast.cuh): the three channel widths of the REFace UNet ---------------------------
template <typename C, int kMinBlocks>
__global__ void __launch_bounds__(C::kThreads, kMinBlocks)
fsai_fast_kernel(const fsaifast::Args A) {
  extern __shared__ __align__(16) float sm_fast[];
  float* sm_exch = sm_fast;
  float* sm_tw1 = sm_exch + C::kExchFloats;
  float* sm_rec = sm_tw1 + C::kTw1Floats;
  {
    float2* pre = reinterpret_cast<float2*>(sm_exch);          // W_D^j, only needed to fill the tables
    static_assert(C::D * 2 <= C::kExchFloats, "scratch for the twiddle seed table");
    for (int j = threadIdx.x; j < C::D; j += C::kThreads) {
      float s, c;
      sincospif(-2.0f * (float)j / (float)C::D, &s, &c);
      pre[j] = make_float2(c, s);
    }
    __syncthreads();
    constexpr int kItems = C::L * C::M > C::kSub ? C::L * C::M : C::kSub;
    for (int i = threadIdx.x; i < kItems; i += C::kThreads) fsaifast::build_tables<C>(i, sm_tw1, sm_rec, pre, A.split);
    __syncthreads();
  }
  const long long stride = (long long)gridDim.x * C::RP;
  for (long long base = (long long)blockIdx.x * C::RP; base < A.n_pairs; base += stride) {
    fsaifast::phase_a<C>(A, base, threadIdx.x, sm_exch, sm_tw1);
    if (A.prefetch && base + stride < A.n_pairs) {
      // pull the next iteration's rows into L2 while this one is being transformed: 16 warps per SM
      // cannot cover a DRAM round trip at the start of phase A, an L2 hit they can
      using T = typename C::T;
      constexpr int kLines = C::D * (int)sizeof(T) / 128;              // 128-byte lines per row
      const long long row0 = (A.fused ? 1 : 2) * (base + stride);
      const int n_rows = (A.fused ? 1 : 2) * C::RP;
      const int n_ops = A.fused ? 3 : 2;
      for (int i = threadIdx.x; i < n_ops * n_rows * kLines; i += C::kThreads) {
        const int op = i / (n_rows * kLines);
        const int rem = i - op * (n_rows * kLines);
        const int r = rem / kLines, ln = rem - r * kLines;
        const long long row = row0 + r;
        if (row < A.rows) {
          const T* p = op == 0 ? reinterpret_cast<const T*>(A.donor) + row * A.ld_donor
                     : op == 1 ? reinterpret_cast<const T*>(A.dst_a) + row * A.ld_a
                               : reinterpret_cast<const T*>(A.dst_b) + row * A.ld_b;
          asm volatile("prefetch.global.L2 [%0];" :: "l"(p + ln * (128 / (int)sizeof(T))));
        }
      }
    }
    __syncthreads();
    fsaifast::phase_b<C>(threadIdx.x, sm_exch, sm_rec);
    __syncthreads();
    fsaifast::phase_c<C>(A, base, threadIdx.x, sm_exch, sm_tw1);
    __syncthreads();
  }
}

template <typename C, int kMinBlocks>
static int launch_fast(const FsaiParams& P, cudaStream_t st) {
  fsaifast::Args A{};
  A.donor = P.donor; A.dst_a = P.dst_a; A.out_a = P.out_a; A.dst_b = P.dst_b; A.out_b = P.out_b;
  A.rows = P.rows; A.fused = P.fused;
  A.n_pairs = P.fused ? P.rows : (P.rows + 1) / 2;
  A.ld_donor = P.ld_donor; A.ld_a = P.ld_a; A.ld_out_a = P.ld_out_a; A.ld_b = P.ld_b; A.ld_out_b = P.ld_out_b;
  A.split = P.split;
  {
    static int pf = -1;
    if (pf < 0) { const char* e = getenv("VF_FSAI_PREFETCH"); pf = e ? atoi(e) : 1; }
    A.prefetch = pf;
  }
  constexpr size_t smem = (size_t)C::kSmemFloats * sizeof(float);
  static int blocks_per_sm = 0;
  if (blocks_per_sm == 0) {
    VF_CUDA_TRY(cudaFuncSetAttribute(fsai_fast_kernel<C, kMinBlocks>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int nb = 0;
    VF_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fsai_fast_kernel<C, kMinBlocks>, C::kThreads, smem));
    blocks_per_sm = nb > 0 ? nb : 1;
  }
  const long long iters = (A.n_pairs + C::RP - 1) / C::RP;
  long long grid = (long long)num_sms() * blocks_per_sm;
  if (grid > iters) grid = iters;
  fsai_fast_kernel<C, kMinBlocks><<<(int)grid, C::kThreads, smem, st>>>(A);
  return check_cuda(cudaGetLastError(), "fsai_fast_kernel launch");
}

static bool small_cta() {      // tuning knob: VF_FSAI_SMALL=1 -> half the rows per CTA, five CTAs per SM
  static int v = -1;
  if (v < 0) { const char* e = getenv("VF_FSAI_SMALL"); v = e ? atoi(e) : 0; }
  return v != 0;
}

template <typename T>
static int dispatch_fast(const FsaiParams& P, cudaStream_t st, bool* handled) {
  *handled = true;
  switch (P.d) {
    case 320:
      if (small_cta()) return launch_fast<fsaifast::Cfg<T, 320, 4, 10, 8, 8>, 5>(P, st);
      return launch_fast<fsaifast::Cfg<T, 320, 4, 10, 8, 16>, 3>(P, st);
    case 640:
      if (small_cta()) return launch_fast<fsaifast::Cfg<T, 640, 2, 20, 16, 4>, 5>(P, st);
      return launch_fast<fsaifast::Cfg<T, 640, 2, 20, 16, 8>, 3>(P, st);
    case 1280: return launch_fast<fsaifast::Cfg<T, 1280, 2, 20, 32, 8>, 1>(P, st);
    default: *handled = false; return 0;
  }
}

static int factor_radices(int d, int* radix) {
  int n = 0, r = d;
  while (r % 5 == 0) { if (n == kMaxStages) return -1; radix[n++] = 5; r /= 5; }
  while (r % 4 == 0) { if (n == kMaxStages) return -1; radix[n++] = 4; r /= 4; }
  while (r % 2 == 0) { if (n == kMaxStages) return -1; radix[n++] = 2; r /= 2; }
  return r == 1 ? n : -1;
}

static int launch_fsai(FsaiParams& P, int dtype, cudaStream_t st, const char* who) {
  if (int rc = check_device()) return rc;
  if (P.rows <= 0) return fail("%s: rows=%lld", who, P.rows);
  if (dtype != VF_F32 && dtype != VF_BF16) return fail("%s: bad dtype %d", who, dtype);
  if (P.d < 32 || P.d > 2048) return fail("%s: d=%d out of range [32, 2048]", who, P.d);
  P.n_stages = factor_radices(P.d, P.radix);
  if (P.n_stages < 0) return fail("%s: d=%d is not of the form 2^a * 5^b", who, P.d);
  if (P.split < 0 || P.split > P.d) return fail("%s: split=%d outside [0, d=%d]", who, P.split, P.d);
  const int epc = dtype == VF_F32 ? 4 : 8;
  if (P.d % epc) return fail("%s: d=%d must be a multiple of %d", who, P.d, epc);
  const long long lds[5] = {P.ld_donor, P.ld_a, P.ld_out_a, P.fused ? P.ld_b : P.ld_a, P.fused ? P.ld_out_b : P.ld_out_a};
  for (long long ld : lds)
    if (ld < P.d || (ld % epc)) return fail("%s: row strides must be >= d and 16-byte multiples", who);
  const void* ptrs[5] = {P.donor, P.dst_a, P.out_a, P.fused ? P.dst_b : P.dst_a, P.fused ? P.out_b : P.out_a};
  for (const void* q : ptrs) {
    if (!q) return fail("%s: null pointer", who);
    if (reinterpret_cast<uintptr_t>(q) & 15) return fail("%s: pointers must be 16-byte aligned", who);
  }
  if (!getenv("VF_FSAI_GENERIC")) {
    bool handled = false;
    const int rc = dtype == VF_F32 ? dispatch_fast<float>(P, st, &handled) : dispatch_fast<__nv_bfloat16>(P, st, &handled);
    if (handled) return rc;
  }
  int ppb = 5120 / P.d;
  if (ppb < 1) ppb = 1;
  if (ppb > 16) ppb = 16;
  P.pairs_per_batch = ppb;
  const size_t smem = (size_t)P.d * (sizeof(float2) + sizeof(float)) + 2 * (size_t)ppb * P.d * sizeof(float2);
  const long long n_pairs = P.fused ? P.rows : (P.rows + 1) / 2;
  const long long n_batches = (n_pairs + ppb - 1) / ppb;
  long long grid = n_batches;
  const long long cap = 2LL * num_sms();
  if (grid > cap) grid = cap;
  if (dtype == VF_F32) {
    static bool attr_f32 = false;
    if (!attr_f32) { VF_CUDA_TRY(cudaFuncSetAttribute(fsai_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); attr_f32 = true; }
    fsai_kernel<float><<<(int)grid, kFsaiThreads, smem, st>>>(P);
  } else {
    static bool attr_bf16 = false;
    if (!attr_bf16) { VF_CUDA_TRY(cudaFuncSetAttribute(fsai_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); attr_bf16 = true; }
    fsai_kernel<__nv_bfloat16><<<(int)grid, kFsaiThreads, smem, st>>>(P);
  }
  return check_cuda(cudaGetLastError(), "fsai_kernel launch");
}

}  // namespace vf

extern "C" int vf_fsai_blend(const void* donor, const void* dst, void* out,
                             long long rows, int d, int split,
                             long long ld_donor, long long ld_dst, long long ld_out,
                             int dtype, void* stream) {
  vf::FsaiParams P{};
  P.donor = donor; P.dst_a = dst; P.out_a = out; P.dst_b = nullptr; P.out_b = nullptr;
  P.rows = rows; P.d = d; P.split = split;
  P.ld_donor = ld_donor; P.ld_a = ld_dst; P.ld_out_a = ld_out; P.ld_b = 0; P.ld_out_b = 0;
  P.fused = 0;
  return vf::launch_fsai(P, dtype, (cudaStream_t)stream, "vf_fsai_blend");
}

extern "C" int vf_fsai_blend2(const void* donor, const void* dst_a, void* out_a,
                              const void* dst_b, void* out_b,
                              long long rows, int d, int split,
                              long long ld_donor, long long ld_a, long long ld_out_a,
                              long long ld_b, long long ld_out_b,
                              int dtype, void* stream) {
  vf::FsaiParams P{};
  P.donor = donor; P.dst_a = dst_a; P.out_a = out_a; P.dst_b = dst_b; P.out_b = out_b;
  P.rows = rows; P.d = d; P.split = split;
  P.ld_donor = ld_donor; P.ld_a = ld_a; P.ld_out_a = ld_out_a; P.ld_b = ld_b; P.ld_out_b = ld_out_b;
  P.fused = 1;
  return vf::launch_fsai(P, dtype, (cudaStream_t)stream, "vf_fsai_blend2");
}
