// Register-resident complex DFTs of compile-time length over V-lane vectors (FSAI fast path).
//
// A CVec<V> is V independent complex numbers that take the SAME butterflies: in the FSAI kernel the V
// lanes are the V interleaved sub-sequences x[V*m + e] of one row, so one vectorised global load of V
// consecutive channels feeds one CVec and every butterfly is issued once for V lanes.  Lanes are stored
// as float2 pairs and go through the Blackwell packed-fp32 instructions (add/mul/fma .f32x2 -> FADD2 /
// FMUL2 / FFMA2): half the issue slots of scalar code for the same fp32 pipe work.
//
// All lengths are compile time and every loop is unrolled, so the arrays live in registers and every
// twiddle is an immediate.  Sizes: 2, 4, 5, 8, 16, 32 (Cooley-Tukey) and 10, 20, 40 (Good-Thomas prime
// factor split 2x5 / 4x5 / 8x5: no twiddles between the factors).
//
// Everything here is __host__ __device__ so that tests/csrc/fsai_host_check.cu can check the index maps
// against a naive DFT on the CPU (the container that builds this has no GPU).
#pragma once

#include <cuda_runtime.h>
#include <type_traits>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
#define VF_PACKED_F32X2 1
#else
#define VF_PACKED_F32X2 0
#endif

#define VF_HD __host__ __device__ __forceinline__

namespace vf {
namespace fftreg {

// ---- compile-time trigonometry ------------------------------------------------------------------------
constexpr double kPi = 3.14159265358979323846264338327950288;

constexpr double sin_series(double x) {   // |x| <= pi/2
  double term = x, sum = x;
  for (int i = 1; i < 14; ++i) {
    term *= -x * x / ((2.0 * i) * (2.0 * i + 1.0));
    sum += term;
  }
  return sum;
}
constexpr double cos_series(double x) {   // |x| <= pi/2
  double term = 1.0, sum = 1.0;
  for (int i = 1; i < 14; ++i) {
    term *= -x * x / ((2.0 * i - 1.0) * (2.0 * i));
    sum += term;
  }
  return sum;
}
// cos / sin of 2*pi*j/n, exact on the axes, series elsewhere (argument folded into [0, pi/2]).
constexpr double cos2pi(long long j, long long n) {
  j %= n;
  if (j < 0) j += n;
  if (4 * j == 0) return 1.0;
  if (4 * j == n) return 0.0;
  if (4 * j == 2 * n) return -1.0;
  if (4 * j == 3 * n) return 0.0;
  if (2 * j > n) j = n - j;                     // cos(2pi - a) = cos a          -> j/n in [0, 1/2]
  if (4 * j > n) return -cos_series(2.0 * kPi * (double)(n - 2 * j) / (2.0 * (double)n));   // cos(pi - a) = -cos a
  return cos_series(2.0 * kPi * (double)j / (double)n);
}
constexpr double sin2pi(long long j, long long n) {
  j %= n;
  if (j < 0) j += n;
  if (4 * j == 0 || 4 * j == 2 * n) return 0.0;
  if (4 * j == n) return 1.0;
  if (4 * j == 3 * n) return -1.0;
  if (2 * j > n) return -sin2pi(n - j, n);      // sin(2pi - a) = -sin a
  if (4 * j > n) return sin_series(2.0 * kPi * (double)(n - 2 * j) / (2.0 * (double)n));    // sin(pi - a) = sin a
  return sin_series(2.0 * kPi * (double)j / (double)n);
}

template <int I, int N, typename F>
VF_HD void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}

// ---- V-lane real vector ---------------------------------------------------------------------------------
template <int V>
struct RVec {
  static_assert(V % 2 == 0, "lanes come in float2 pairs");
  float2 p[V / 2];
  VF_HD float get(int e) const { return (e & 1) ? p[e >> 1].y : p[e >> 1].x; }
  VF_HD void set(int e, float v) { if (e & 1) p[e >> 1].y = v; else p[e >> 1].x = v; }
};

template <int V> VF_HD RVec<V> operator+(const RVec<V>& a, const RVec<V>& b) {
  RVec<V> r;
#pragma unroll
  for (int i = 0; i < V / 2; ++i) {
#if VF_PACKED_F32X2
    r.p[i] = __fadd2_rn(a.p[i], b.p[i]);
#else
    r.p[i] = make_float2(a.p[i].x + b.p[i].x, a.p[i].y + b.p[i].y);
#endif
  }
  return r;
}
template <int V> VF_HD RVec<V> operator-(const RVec<V>& a, const RVec<V>& b) {
  RVec<V> r;
#pragma unroll
  for (int i = 0; i < V / 2; ++i) {
#if VF_PACKED_F32X2
    r.p[i] = __ffma2_rn(b.p[i], make_float2(-1.0f, -1.0f), a.p[i]);
#else
    r.p[i] = make_float2(a.p[i].x - b.p[i].x, a.p[i].y - b.p[i].y);
#endif
  }
  return r;
}
template <int V> VF_HD RVec<V> operator-(const RVec<V>& a) {
  RVec<V> r;
#pragma unroll
  for (int i = 0; i < V / 2; ++i) r.p[i] = make_float2(-a.p[i].x, -a.p[i].y);
  return r;
}
// c * a
template <int V> VF_HD RVec<V> mulc(const RVec<V>& a, float c) {
  RVec<V> r;
#pragma unroll
  for (int i = 0; i < V / 2; ++i) {
#if VF_PACKED_F32X2
    r.p[i] = __fmul2_rn(a.p[i], make_float2(c, c));
#else
    r.p[i] = make_float2(a.p[i].x * c, a.p[i].y * c);
#endif
  }
  return r;
}
// c * a + b
template <int V> VF_HD RVec<V> fmac(const RVec<V>& a, float c, const RVec<V>& b) {
  RVec<V> r;
#pragma unroll
  for (int i = 0; i < V / 2; ++i) {
#if VF_PACKED_F32X2
    r.p[i] = __ffma2_rn(a.p[i], make_float2(c, c), b.p[i]);
#else
    r.p[i] = make_float2(fmaf(a.p[i].x, c, b.p[i].x), fmaf(a.p[i].y, c, b.p[i].y));
#endif
  }
  return r;
}

// ---- V-lane complex vector ---------------------------------------------------------------------------------
template <int V>
struct CVec {
  RVec<V> re, im;
};
template <int V> VF_HD CVec<V> operator+(const CVec<V>& a, const CVec<V>& b) { return {a.re + b.re, a.im + b.im}; }
template <int V> VF_HD CVec<V> operator-(const CVec<V>& a, const CVec<V>& b) { return {a.re - b.re, a.im - b.im}; }
template <int V> VF_HD CVec<V> mul_mi(const CVec<V>& a) { return {a.im, -a.re}; }     // * (-i)
template <int V> VF_HD CVec<V> mul_pi(const CVec<V>& a) { return {-a.im, a.re}; }     // * (+i)
template <int V> VF_HD CVec<V> cswap(const CVec<V>& a) { return {a.im, a.re}; }
// * (c + i s), run-time scalars shared by all lanes
template <int V> VF_HD CVec<V> cmul(const CVec<V>& a, float c, float s) {
  CVec<V> r;
  r.re = fmac(a.im, -s, mulc(a.re, c));
  r.im = fmac(a.im, c, mulc(a.re, s));
  return r;
}
// * W_N^J = exp(-2 pi i J / N), compile time, trivial cases folded
template <int J0, int N, int V>
VF_HD CVec<V> twiddle(const CVec<V>& a) {
  constexpr int J = ((J0 % N) + N) % N;
  if constexpr (J == 0) {
    return a;
  } else if constexpr (4 * J == N) {
    return mul_mi(a);
  } else if constexpr (2 * J == N) {
    return {-a.re, -a.im};
  } else if constexpr (4 * J == 3 * N) {
    return mul_pi(a);
  } else if constexpr (8 * J == N) {            // (1 - i)/sqrt2
    constexpr float h = 0.70710678118654752440f;
    return {mulc(a.re + a.im, h), mulc(a.im - a.re, h)};
  } else if constexpr (8 * J == 3 * N) {        // (-1 - i)/sqrt2
    constexpr float h = 0.70710678118654752440f;
    return {mulc(a.im - a.re, h), mulc(-(a.re + a.im), h)};
  } else if constexpr (8 * J == 5 * N) {        // (-1 + i)/sqrt2
    constexpr float h = 0.70710678118654752440f;
    return {mulc(-(a.re + a.im), h), mulc(a.re - a.im, h)};
  } else if constexpr (8 * J == 7 * N) {        // (1 + i)/sqrt2
    constexpr float h = 0.70710678118654752440f;
    return {mulc(a.re - a.im, h), mulc(a.re + a.im, h)};
  } else {
    constexpr float c = (float)cos2pi(J, N);
    constexpr float s = (float)(-sin2pi(J, N));
    return cmul(a, c, s);
  }
}

// ---- forward DFTs, in place, natural order in and out -------------------------------------------------------
template <int N> struct Dft;

template <> struct Dft<2> {
  template <int V> static VF_HD void run(CVec<V>* x) {
    const CVec<V> a = x[0], b = x[1];
    x[0] = a + b;
    x[1] = a - b;
  }
};
template <> struct Dft<4> {
  template <int V> static VF_HD void run(CVec<V>* x) {
    const CVec<V> a = x[0] + x[2], b = x[0] - x[2];
    const CVec<V> c = x[1] + x[3], d = mul_mi(x[1] - x[3]);
    x[0] = a + c;
    x[1] = b + d;
    x[2] = a - c;
    x[3] = b - d;
  }
};
template <> struct Dft<5> {
  template <int V> static VF_HD void run(CVec<V>* x) {
    constexpr float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f;
    constexpr float s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;
    const CVec<V> t1 = x[1] + x[4], t2 = x[2] + x[3];
    const CVec<V> t3 = x[1] - x[4], t4 = x[2] - x[3];
    CVec<V> m1, m2, n1, n2;
    m1.re = fmac(t2.re, c2, fmac(t1.re, c1, x[0].re));
    m1.im = fmac(t2.im, c2, fmac(t1.im, c1, x[0].im));
    m2.re = fmac(t2.re, c1, fmac(t1.re, c2, x[0].re));
    m2.im = fmac(t2.im, c1, fmac(t1.im, c2, x[0].im));
    n1.re = fmac(t4.re, s2, mulc(t3.re, s1));
    n1.im = fmac(t4.im, s2, mulc(t3.im, s1));
    n2.re = fmac(t4.re, -s1, mulc(t3.re, s2));
    n2.im = fmac(t4.im, -s1, mulc(t3.im, s2));
    const CVec<V> in1 = mul_mi(n1), in2 = mul_mi(n2);
    x[0] = x[0] + t1 + t2;
    x[1] = m1 + in1;
    x[4] = m1 - in1;
    x[2] = m2 + in2;
    x[3] = m2 - in2;
  }
};

// Cooley-Tukey N = N1*N2: n = N2*n1 + n2, k = k1 + N1*k2.
template <int N, int N1, int N2>
struct DftCT {
  template <int V> static VF_HD void run(CVec<V>* x) {
    CVec<V> g[N2][N1];
    static_for<0, N2>([&](auto n2c) {
      constexpr int n2 = decltype(n2c)::value;
      static_for<0, N1>([&](auto n1c) {
        constexpr int n1 = decltype(n1c)::value;
        g[n2][n1] = x[N2 * n1 + n2];
      });
      Dft<N1>::run(g[n2]);
      static_for<1, N1>([&](auto k1c) {
        constexpr int k1 = decltype(k1c)::value;
        g[n2][k1] = twiddle<n2 * k1, N>(g[n2][k1]);
      });
    });
    static_for<0, N1>([&](auto k1c) {
      constexpr int k1 = decltype(k1c)::value;
      CVec<V> h[N2];
      static_for<0, N2>([&](auto n2c) {
        constexpr int n2 = decltype(n2c)::value;
        h[n2] = g[n2][k1];
      });
      Dft<N2>::run(h);
      static_for<0, N2>([&](auto k2c) {
        constexpr int k2 = decltype(k2c)::value;
        x[k1 + N1 * k2] = h[k2];
      });
    });
  }
};
template <> struct Dft<8> : DftCT<8, 4, 2> {};
template <> struct Dft<16> : DftCT<16, 4, 4> {};
template <> struct Dft<32> : DftCT<32, 8, 4> {};

constexpr int mod_inverse(int a, int m) {
  for (int x = 1; x < m; ++x)
    if ((a * x) % m == 1) return x;
  return 0;
}
// Good-Thomas N = N1*N2, gcd(N1, N2) = 1: n = (N2*n1 + N1*n2) mod N,
// k = (N2*(N2^-1 mod N1)*k1 + N1*(N1^-1 mod N2)*k2) mod N; no twiddles.
template <int N, int N1, int N2>
struct DftPFA {
  template <int V> static VF_HD void run(CVec<V>* x) {
    constexpr int A = N2 * mod_inverse(N2 % N1, N1);
    constexpr int B = N1 * mod_inverse(N1 % N2, N2);
    CVec<V> g[N2][N1];
    static_for<0, N2>([&](auto n2c) {
      constexpr int n2 = decltype(n2c)::value;
      static_for<0, N1>([&](auto n1c) {
        constexpr int n1 = decltype(n1c)::value;
        g[n2][n1] = x[(N2 * n1 + N1 * n2) % N];
      });
      Dft<N1>::run(g[n2]);
    });
    static_for<0, N1>([&](auto k1c) {
      constexpr int k1 = decltype(k1c)::value;
      CVec<V> h[N2];
      static_for<0, N2>([&](auto n2c) {
        constexpr int n2 = decltype(n2c)::value;
        h[n2] = g[n2][k1];
      });
      Dft<N2>::run(h);
      static_for<0, N2>([&](auto k2c) {
        constexpr int k2 = decltype(k2c)::value;
        x[(A * k1 + B * k2) % N] = h[k2];
      });
    });
  }
};
template <> struct Dft<10> : DftPFA<10, 2, 5> {};
template <> struct Dft<20> : DftPFA<20, 4, 5> {};
template <> struct Dft<40> : DftPFA<40, 8, 5> {};

// forward / inverse (unnormalised) in place; the inverse is the forward transform with re and im swapped.
template <int N, bool kInverse, int V>
VF_HD void dft_inplace(CVec<V>* x) {
  if constexpr (kInverse) {
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = cswap(x[i]);
    Dft<N>::run(x);
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = cswap(x[i]);
  } else {
    Dft<N>::run(x);
  }
}

}  // namespace fftreg
}  // namespace vf
