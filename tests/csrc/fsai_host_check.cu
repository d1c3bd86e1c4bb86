// CPU check of the FSAI fast path's index maps (vface_b200/csrc/vf_fsai_fast.cuh): the phase bodies are
// __host__ __device__, so the CTA is emulated thread by thread here and compared with a naive O(D^2)
// double-precision evaluation of  out = Re ifft([fft(dst)[:sp], fft(donor)[sp:]])
// (scripts/face_swap_utils.py:425-464).  Built and run by tests/test_fsai_host_check.py; no GPU involved.
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../vface_b200/csrc/vf_fsai_fast.cuh"

using namespace vf::fsaifast;

static float frand() { return (float)rand() / RAND_MAX * 2.0f - 1.0f; }
static float to_t(float v, float*) { return v; }
static __nv_bfloat16 to_t(float v, __nv_bfloat16*) { return __float2bfloat16_rn(v); }
static float from_t(float v) { return v; }
static float from_t(__nv_bfloat16 v) { return __bfloat162float(v); }

static std::vector<double> naive(const std::vector<double>& donor, const std::vector<double>& dst, int d, int sp) {
  typedef std::complex<double> cd;
  std::vector<cd> f1(d), f2(d), comb(d);
  const double pi = std::acos(-1.0);
  for (int k = 0; k < d; ++k) {
    cd a = 0, b = 0;
    for (int n = 0; n < d; ++n) {
      const cd w = std::polar(1.0, -2.0 * pi * (double)((long long)k * n % d) / d);
      a += donor[n] * w;
      b += dst[n] * w;
    }
    comb[k] = k < sp ? b : a;
  }
  std::vector<double> out(d);
  for (int n = 0; n < d; ++n) {
    cd s = 0;
    for (int k = 0; k < d; ++k) s += comb[k] * std::polar(1.0, 2.0 * pi * (double)((long long)k * n % d) / d);
    out[n] = s.real() / d;
  }
  return out;
}

template <typename C>
static double run_case(int rows, int fused, int split, long long ld, bool inplace) {
  using T = typename C::T;
  constexpr int D = C::D;
  std::vector<T> donor(rows * ld), a(rows * ld), b(rows * ld), oa(rows * ld), ob(rows * ld);
  for (auto* v : {&donor, &a, &b})
    for (auto& x : *v) x = to_t(frand() * 2.0f, (T*)nullptr);
  std::vector<T> a0 = a, b0 = b;
  Args A{};
  A.donor = donor.data(); A.dst_a = a.data(); A.dst_b = b.data();
  A.out_a = inplace ? a.data() : oa.data(); A.out_b = inplace ? b.data() : ob.data();
  A.rows = rows; A.fused = fused; A.n_pairs = fused ? rows : (rows + 1) / 2;
  A.ld_donor = A.ld_a = A.ld_out_a = A.ld_b = A.ld_out_b = ld;
  A.split = split;
  std::vector<float> sm(C::kSmemFloats);
  float* sm_exch = sm.data();
  float* sm_tw1 = sm_exch + C::kExchFloats;
  float* sm_rec = sm_tw1 + C::kTw1Floats;
  std::vector<float2> pre(D);
  const double pi = std::acos(-1.0);
  for (int j = 0; j < D; ++j) pre[j] = make_float2((float)std::cos(2 * pi * j / D), (float)-std::sin(2 * pi * j / D));
  for (int i = 0; i < C::kThreads * 8; ++i) build_tables<C>(i, sm_tw1, sm_rec, pre.data(), split);
  for (long long base = 0; base < A.n_pairs; base += C::RP) {
    for (int tid = 0; tid < C::kThreads; ++tid) phase_a<C>(A, base, tid, sm_exch, sm_tw1);
    for (int tid = 0; tid < C::kThreads; ++tid) phase_b<C>(tid, sm_exch, sm_rec);
    for (int tid = 0; tid < C::kThreads; ++tid) phase_c<C>(A, base, tid, sm_exch, sm_tw1);
  }
  double worst = 0;
  for (int r = 0; r < rows; ++r) {
    std::vector<double> dn(D), da(D), db(D);
    for (int n = 0; n < D; ++n) { dn[n] = from_t(donor[r * ld + n]); da[n] = from_t(a0[r * ld + n]); db[n] = from_t(b0[r * ld + n]); }
    const std::vector<double> wa = naive(dn, da, D, split);
    const T* ga = inplace ? a.data() : oa.data();
    for (int n = 0; n < D; ++n) worst = std::fmax(worst, std::fabs(wa[n] - from_t(ga[r * ld + n])));
    if (fused) {
      const std::vector<double> wb = naive(dn, db, D, split);
      const T* gb = inplace ? b.data() : ob.data();
      for (int n = 0; n < D; ++n) worst = std::fmax(worst, std::fabs(wb[n] - from_t(gb[r * ld + n])));
    }
  }
  return worst;
}

template <int N> static double dft_err() {
  using vf::fftreg::CVec;
  CVec<2> x[N];
  std::vector<std::complex<double>> in(N);
  for (int i = 0; i < N; ++i) {
    const float re = frand(), im = frand();
    in[i] = {re, im};
    x[i].re.p[0] = make_float2(re, 2 * re);
    x[i].im.p[0] = make_float2(im, -im);
  }
  vf::fftreg::dft_inplace<N, false>(x);
  double worst = 0;
  const double pi = std::acos(-1.0);
  for (int k = 0; k < N; ++k) {
    std::complex<double> s = 0;
    for (int n = 0; n < N; ++n) s += in[n] * std::polar(1.0, -2 * pi * k * n / N);
    worst = std::fmax(worst, std::abs(s - std::complex<double>(x[k].re.p[0].x, x[k].im.p[0].x)));
  }
  return worst;
}

int main() {
  int bad = 0;
  auto rep = [&](const char* name, double err, double tol) {
    std::printf("%-44s max|err| = %.3e  (tol %.1e) %s\n", name, err, tol, err < tol ? "ok" : "FAIL");
    if (!(err < tol)) ++bad;
  };
  rep("dft<2>", dft_err<2>(), 1e-5);   rep("dft<4>", dft_err<4>(), 1e-5);   rep("dft<5>", dft_err<5>(), 1e-5);
  rep("dft<8>", dft_err<8>(), 1e-5);   rep("dft<10>", dft_err<10>(), 1e-5); rep("dft<16>", dft_err<16>(), 2e-5);
  rep("dft<20>", dft_err<20>(), 2e-5); rep("dft<32>", dft_err<32>(), 4e-5); rep("dft<40>", dft_err<40>(), 4e-5);
  typedef Cfg<float, 320, 4, 10, 8, 16> C320;
  typedef Cfg<float, 640, 2, 20, 16, 8> C640;
  typedef Cfg<float, 1280, 2, 20, 32, 4> C1280;
  typedef Cfg<__nv_bfloat16, 320, 4, 10, 8, 16> B320;
  typedef Cfg<__nv_bfloat16, 640, 2, 20, 16, 8> B640;
  rep("fsai f32 d=320 fused, 37 rows, in place", run_case<C320>(37, 1, 256, 960, true), 2e-5);
  rep("fsai f32 d=320 single, 37 rows (odd tail)", run_case<C320>(37, 0, 256, 320, false), 2e-5);
  rep("fsai f32 d=320 single, split=100", run_case<C320>(8, 0, 100, 320, false), 2e-5);
  rep("fsai f32 d=640 fused, 19 rows", run_case<C640>(19, 1, 512, 1920, true), 2e-5);
  rep("fsai f32 d=640 single, 9 rows", run_case<C640>(9, 0, 512, 640, false), 2e-5);
  rep("fsai f32 d=1280 fused, 9 rows", run_case<C1280>(9, 1, 1024, 3840, true), 4e-5);
  rep("fsai f32 d=1280 single, 5 rows, split=0", run_case<C1280>(5, 0, 0, 1280, false), 4e-5);
  rep("fsai bf16 d=320 fused, 20 rows, in place", run_case<B320>(20, 1, 256, 960, true), 2e-2);
  rep("fsai bf16 d=640 single, 7 rows", run_case<B640>(7, 0, 512, 640, false), 2e-2);
  std::printf(bad ? "FAILED (%d)\n" : "all ok\n", bad);
  return bad ? 1 : 0;
}
