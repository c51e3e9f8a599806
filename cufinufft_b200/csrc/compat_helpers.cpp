// compat_helpers.cpp -- the four array-norm helpers the reference's shared library exports next to
// the C API (contrib/utils_fp.h:53-57, compiled into its libcufinufft.so) and which its own example
// programs and API tests call after linking `-lcufinufft` (examples/example2d1many.cpp:101,
// test/cufinufft2d2api_test.cu:113).  C++ linkage, BIGINT = int (contrib/dataTypes.h), both
// precisions -- so that those programs link against this library without a source change.
// Host-only; not on the transform path.
#include <cmath>
#include <complex>

namespace {
template <typename T> T sq(const std::complex<T> &z) { return z.real() * z.real() + z.imag() * z.imag(); }

template <typename T> T two_norm(int n, const std::complex<T> *a)
{
    T s = 0;
    for (int m = 0; m < n; ++m) s += sq(a[m]);
    return std::sqrt(s);
}
template <typename T> T err_two_norm(int n, const std::complex<T> *a, const std::complex<T> *b)
{
    T s = 0;
    for (int m = 0; m < n; ++m) s += sq(a[m] - b[m]);
    return std::sqrt(s);
}
template <typename T> T inf_norm(int n, const std::complex<T> *a)
{
    T big = 0;
    for (int m = 0; m < n; ++m) big = std::fmax(big, sq(a[m]));
    return std::sqrt(big);
}
}  // namespace

double twonorm(int n, std::complex<double> *a) { return two_norm(n, a); }
float twonorm(int n, std::complex<float> *a) { return two_norm(n, a); }
double errtwonorm(int n, std::complex<double> *a, std::complex<double> *b) { return err_two_norm(n, a, b); }
float errtwonorm(int n, std::complex<float> *a, std::complex<float> *b) { return err_two_norm(n, a, b); }
double relerrtwonorm(int n, std::complex<double> *a, std::complex<double> *b) { return err_two_norm(n, a, b) / two_norm(n, a); }
float relerrtwonorm(int n, std::complex<float> *a, std::complex<float> *b) { return err_two_norm(n, a, b) / two_norm(n, a); }
double infnorm(int n, std::complex<double> *a) { return inf_norm(n, a); }
float infnorm(int n, std::complex<float> *a) { return inf_norm(n, a); }
