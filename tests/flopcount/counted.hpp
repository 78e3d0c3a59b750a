// tests/flopcount/counted.hpp -- TEST INFRASTRUCTURE: an op-counting scalar that stands in for `double` when the per-ray
// device code (geoac_b200/csrc/*.cuh, GEOAC_HD) is compiled on the host, to derive the ALGORITHMIC flop count per RK4 step
// of each variant (add/sub/mul/div/sqrt = 1, fma = 2, each exp/sin/cos/asin/atan2/pow/cbrt = 1 and also
// tallied as a transcendental, compares/abs/min/max/negation = 0).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace cnt {
struct Tally { unsigned long long add = 0, mul = 0, fma = 0, div = 0, sqrt = 0, trans = 0; };
inline Tally& tally() { static Tally t; return t; }
inline unsigned long long flops() { const Tally& t = tally(); return t.add + t.mul + 2 * t.fma + t.div + t.sqrt + t.trans; }

struct Cnt {
    double v;
    constexpr Cnt() : v(0.0) {}
    constexpr Cnt(double x) : v(x) {}
    constexpr Cnt(int x) : v((double)x) {}
    constexpr Cnt(long x) : v((double)x) {}
    constexpr Cnt(unsigned x) : v((double)x) {}
    explicit operator double() const { return v; }
    explicit operator int() const { return (int)v; }
    explicit operator bool() const { return v != 0.0; }
    Cnt& operator+=(Cnt o) { tally().add++; v += o.v; return *this; }
    Cnt& operator-=(Cnt o) { tally().add++; v -= o.v; return *this; }
    Cnt& operator*=(Cnt o) { tally().mul++; v *= o.v; return *this; }
    Cnt& operator/=(Cnt o) { tally().div++; v /= o.v; return *this; }
};
inline Cnt operator+(Cnt a, Cnt b) { tally().add++; return Cnt(a.v + b.v); }
inline Cnt operator-(Cnt a, Cnt b) { tally().add++; return Cnt(a.v - b.v); }
inline Cnt operator*(Cnt a, Cnt b) { tally().mul++; return Cnt(a.v * b.v); }
inline Cnt operator/(Cnt a, Cnt b) { tally().div++; return Cnt(a.v / b.v); }
inline Cnt operator-(Cnt a) { return Cnt(-a.v); }
inline Cnt operator+(Cnt a) { return a; }
#define CNT_MIX(op) \
    inline Cnt operator op(Cnt a, double b) { return a op Cnt(b); } inline Cnt operator op(double a, Cnt b) { return Cnt(a) op b; } \
    inline Cnt operator op(Cnt a, int b) { return a op Cnt(b); }    inline Cnt operator op(int a, Cnt b) { return Cnt(a) op b; }
CNT_MIX(+) CNT_MIX(-) CNT_MIX(*) CNT_MIX(/)
#undef CNT_MIX
#define CNT_CMP(op) \
    inline bool operator op(Cnt a, Cnt b) { return a.v op b.v; } inline bool operator op(Cnt a, double b) { return a.v op b; } \
    inline bool operator op(double a, Cnt b) { return a op b.v; } inline bool operator op(Cnt a, int b) { return a.v op b; }
CNT_CMP(<) CNT_CMP(>) CNT_CMP(<=) CNT_CMP(>=) CNT_CMP(==) CNT_CMP(!=)
#undef CNT_CMP
inline Cnt fma(Cnt a, Cnt b, Cnt c) { tally().fma++; return Cnt(std::fma(a.v, b.v, c.v)); }
inline Cnt fma(double a, Cnt b, Cnt c) { return fma(Cnt(a), b, c); }
inline Cnt fma(Cnt a, double b, Cnt c) { return fma(a, Cnt(b), c); }
inline Cnt fma(Cnt a, Cnt b, double c) { return fma(a, b, Cnt(c)); }
inline Cnt fma(double a, Cnt b, double c) { return fma(Cnt(a), b, Cnt(c)); }
inline Cnt fma(Cnt a, double b, double c) { return fma(a, Cnt(b), Cnt(c)); }
inline Cnt fma(double a, double b, Cnt c) { return fma(Cnt(a), Cnt(b), c); }
inline Cnt sqrt(Cnt a) { tally().sqrt++; return Cnt(std::sqrt(a.v)); }
#define CNT_T1(fn) inline Cnt fn(Cnt a) { tally().trans++; return Cnt(std::fn(a.v)); }
CNT_T1(exp) CNT_T1(sin) CNT_T1(cos) CNT_T1(tan) CNT_T1(asin) CNT_T1(cbrt) CNT_T1(log10)
#undef CNT_T1
inline Cnt atan2(Cnt a, Cnt b) { tally().trans++; return Cnt(std::atan2(a.v, b.v)); }
inline Cnt pow(Cnt a, Cnt b) { tally().trans++; return Cnt(std::pow(a.v, b.v)); }
inline Cnt pow(double a, Cnt b) { return pow(Cnt(a), b); }
inline Cnt pow(Cnt a, int b) { tally().mul += (b > 1 ? b - 1 : 0); return Cnt(std::pow(a.v, b)); }
inline void sincos(Cnt a, Cnt* s, Cnt* c) { tally().trans += 2; *s = Cnt(std::sin(a.v)); *c = Cnt(std::cos(a.v)); }
inline Cnt fabs(Cnt a) { return Cnt(std::fabs(a.v)); }
inline Cnt floor(Cnt a) { return Cnt(std::floor(a.v)); }
inline Cnt fmax(Cnt a, Cnt b) { return Cnt(std::fmax(a.v, b.v)); }
inline Cnt fmin(Cnt a, Cnt b) { return Cnt(std::fmin(a.v, b.v)); }
inline Cnt fmax(Cnt a, double b) { return Cnt(std::fmax(a.v, b)); }
inline Cnt fmin(Cnt a, double b) { return Cnt(std::fmin(a.v, b)); }
inline Cnt fmax(double a, Cnt b) { return Cnt(std::fmax(a, b.v)); }
inline Cnt fmin(double a, Cnt b) { return Cnt(std::fmin(a, b.v)); }
inline Cnt ldexp(Cnt a, int n) { return Cnt(std::ldexp(a.v, n)); }
}  // namespace cnt
