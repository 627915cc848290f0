// Scalar helpers shared by all kernels: FP64 real (double) and complex (z128) arithmetic.
#pragma once
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

#define LSA_HD __host__ __device__ __forceinline__

namespace lsa {

struct __align__(16) z128 {
  double x, y;
};

LSA_HD z128 mk(double x, double y) {
  z128 r;
  r.x = x;
  r.y = y;
  return r;
}
LSA_HD z128 operator+(z128 a, z128 b) { return mk(a.x + b.x, a.y + b.y); }
LSA_HD z128 operator-(z128 a, z128 b) { return mk(a.x - b.x, a.y - b.y); }
LSA_HD z128 operator-(z128 a) { return mk(-a.x, -a.y); }
LSA_HD z128 operator*(z128 a, z128 b) { return mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
LSA_HD z128 operator*(double a, z128 b) { return mk(a * b.x, a * b.y); }
LSA_HD z128 operator*(z128 a, double b) { return mk(a.x * b, a.y * b); }
LSA_HD z128& operator+=(z128& a, z128 b) {
  a.x += b.x;
  a.y += b.y;
  return a;
}
LSA_HD z128& operator-=(z128& a, z128 b) {
  a.x -= b.x;
  a.y -= b.y;
  return a;
}
LSA_HD z128 conj_(z128 a) { return mk(a.x, -a.y); }
LSA_HD double conj_(double a) { return a; }
LSA_HD double abs1(z128 a) { return fabs(a.x) + fabs(a.y); }  // LAPACK cabs1, used for pivot search
LSA_HD double abs1(double a) { return fabs(a); }
LSA_HD double abs2(z128 a) { return a.x * a.x + a.y * a.y; }
LSA_HD double abs2(double a) { return a * a; }
LSA_HD double absz(z128 a) { return hypot(a.x, a.y); }
LSA_HD double absz(double a) { return fabs(a); }
// robust complex reciprocal / division (Smith)
LSA_HD z128 recip(z128 a) {
  if (fabs(a.x) >= fabs(a.y)) {
    double t = a.y / a.x, d = a.x + a.y * t;
    return mk(1.0 / d, -t / d);
  }
  double t = a.x / a.y, d = a.x * t + a.y;
  return mk(t / d, -1.0 / d);
}
LSA_HD double recip(double a) { return 1.0 / a; }
LSA_HD z128 operator/(z128 a, z128 b) { return a * recip(b); }

template <class T>
struct scalar_traits;
template <>
struct scalar_traits<double> {
  static constexpr bool is_complex = false;
  LSA_HD static double zero() { return 0.0; }
  LSA_HD static double one() { return 1.0; }
  LSA_HD static double from(z128 a) { return a.x; }
};
template <>
struct scalar_traits<z128> {
  static constexpr bool is_complex = true;
  LSA_HD static z128 zero() { return mk(0, 0); }
  LSA_HD static z128 one() { return mk(1, 0); }
  LSA_HD static z128 from(z128 a) { return a; }
};

// optional conjugation selected at compile time
template <bool CONJ, class T>
LSA_HD T cj(T a) {
  if (CONJ) return conj_(a);
  return a;
}

struct CudaError : std::runtime_error {
  using std::runtime_error::runtime_error;
};
struct NonFiniteError : std::runtime_error {   // NaN / Inf met in the numeric path -> LSA_ERR_NONFINITE
  using std::runtime_error::runtime_error;
};
struct ArgError : std::runtime_error {         // caller error detected below the C ABI -> LSA_ERR_ARG
  using std::runtime_error::runtime_error;
};

#define LSA_CUDA(call)                                                                              \
  do {                                                                                              \
    cudaError_t e_ = (call);                                                                        \
    if (e_ != cudaSuccess)                                                                          \
      throw lsa::CudaError(std::string(#call) + ": " + cudaGetErrorString(e_) + " @" + __FILE__ + \
                           ":" + std::to_string(__LINE__));                                         \
  } while (0)

#define LSA_LAUNCH_CHECK() LSA_CUDA(cudaGetLastError())

// Function attributes (opt-in shared memory, non-portable cluster size) belong to the device a kernel is loaded
// on: `first()` is true once per device and call site, so a process that drives several GPUs sets them on each.
struct PerDeviceOnce {
  bool done[64] = {};
  bool first() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return true;
    if (done[d]) return false;
    done[d] = true;
    return true;
  }
};

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace lsa
