// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may build, load or run anything under oracle/.
//
// Scalar types of the CPU restatement.  The reference computes geometry in FP64
// (`Vector = vec3d`, /root/reference/source/rt/imported_types.d:10-11) and colour in FP32
// (`Color{float r,g,b}`, /root/reference/source/rt/color.d:27-35).  Built normally these are
// plain double / float.  Built with -DORC_COUNT_FLOPS they are wrappers that count every
// add/sub/mul/div/sqrt/floor and every sin/cos/tan/atan2/asin/pow call as ONE flop each
// (SURVEY.md §8(d) "Algorithmic FLOP convention"); compare / abs / neg / select / convert
// count zero.  The count is the numerator of bench.py's roofline figure.
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>

namespace orc {

#ifdef ORC_COUNT_FLOPS

struct FlopCounter {
    static uint64_t& tl() { static thread_local uint64_t n = 0; return n; }
};
inline void orc_flops(unsigned n) { FlopCounter::tl() += n; }

template <class T>
struct Counted {
    T v;
    Counted() : v(std::numeric_limits<T>::quiet_NaN()) {}  // D default-inits floating point to NaN
    Counted(double x) : v(T(x)) {}
    template <class U> explicit Counted(const Counted<U>& o) : v(T(o.v)) {}  // narrowing: 0 flops
    explicit operator double() const { return double(v); }
    Counted operator-() const { Counted r; r.v = -v; return r; }
    Counted& operator+=(Counted o) { orc_flops(1); v += o.v; return *this; }
    Counted& operator-=(Counted o) { orc_flops(1); v -= o.v; return *this; }
    Counted& operator*=(Counted o) { orc_flops(1); v *= o.v; return *this; }
    Counted& operator/=(Counted o) { orc_flops(1); v /= o.v; return *this; }
};
#define ORC_BINOP(op)                                                                       \
    template <class T> inline Counted<T> operator op(Counted<T> a, Counted<T> b) {          \
        orc_flops(1); Counted<T> r; r.v = a.v op b.v; return r; }                           \
    template <class T> inline Counted<T> operator op(Counted<T> a, double b) {              \
        orc_flops(1); Counted<T> r; r.v = a.v op T(b); return r; }                          \
    template <class T> inline Counted<T> operator op(double a, Counted<T> b) {              \
        orc_flops(1); Counted<T> r; r.v = T(a) op b.v; return r; }
ORC_BINOP(+) ORC_BINOP(-) ORC_BINOP(*) ORC_BINOP(/)
#undef ORC_BINOP
#define ORC_CMP(op)                                                                         \
    template <class T> inline bool operator op(Counted<T> a, Counted<T> b) { return a.v op b.v; } \
    template <class T> inline bool operator op(Counted<T> a, double b) { return a.v op T(b); }    \
    template <class T> inline bool operator op(double a, Counted<T> b) { return T(a) op b.v; }
ORC_CMP(<) ORC_CMP(>) ORC_CMP(<=) ORC_CMP(>=) ORC_CMP(==) ORC_CMP(!=)
#undef ORC_CMP

using real = Counted<double>;
using colf = Counted<float>;

inline double raw(real x) { return x.v; }
inline float raw(colf x) { return x.v; }
inline real mk_real(double x) { real r; r.v = x; return r; }
inline colf mk_colf(float x) { colf r; r.v = x; return r; }

inline real r_sqrt(real x) { orc_flops(1); return mk_real(std::sqrt(x.v)); }
inline real r_floor(real x) { orc_flops(1); return mk_real(std::floor(x.v)); }
inline real r_sin(real x) { orc_flops(1); return mk_real(std::sin(x.v)); }
inline real r_cos(real x) { orc_flops(1); return mk_real(std::cos(x.v)); }
inline real r_atan2(real y, real x) { orc_flops(1); return mk_real(std::atan2(y.v, x.v)); }
inline real r_asin(real x) { orc_flops(1); return mk_real(std::asin(x.v)); }
inline real r_pow(real x, real y) { orc_flops(1); return mk_real(std::pow(x.v, y.v)); }
inline real r_fabs(real x) { return mk_real(std::fabs(x.v)); }
inline colf f_floor(colf x) { orc_flops(1); return mk_colf(std::floor(x.v)); }

#else  // plain build

inline void orc_flops(unsigned) {}
using real = double;
using colf = float;
inline double raw(double x) { return x; }
inline float raw(float x) { return x; }
inline real mk_real(double x) { return x; }
inline colf mk_colf(float x) { return x; }
inline real r_sqrt(real x) { return std::sqrt(x); }
inline real r_floor(real x) { return std::floor(x); }
inline real r_sin(real x) { return std::sin(x); }
inline real r_cos(real x) { return std::cos(x); }
inline real r_atan2(real y, real x) { return std::atan2(y, x); }
inline real r_asin(real x) { return std::asin(x); }
inline real r_pow(real x, real y) { return std::pow(x, y); }
inline real r_fabs(real x) { return std::fabs(x); }
inline colf f_floor(colf x) { return std::floor(x); }

#endif

// double -> float narrowing exactly where D narrows implicitly (Color op scalar takes a
// float parameter: /root/reference/source/rt/color.d:128-138).  Zero flops.
inline colf narrow(real x) { return mk_colf(float(raw(x))); }

}  // namespace orc
