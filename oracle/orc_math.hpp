// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_scalar.hpp).
//
// Vector / matrix arithmetic of the reference.  The reference takes these from the dub
// package gfm 7.0.8 (sub-package gfm:math; /root/reference/dub.sdl:10,
// /root/reference/dub.selections.json:6), whose source is NOT under /root/reference.
// What follows restates gfm:math's published vec3d / mat3d algorithms as recalled:
//   * Vector!(T,3): squaredLength = 0 + x*x + y*y + z*z (left to right); length = sqrt of it;
//     normalize() multiplies every component by invLength = 1/length(); dot = 0 + Σ a[i]*b[i];
//     cross = (a.y*b.z - a.z*b.y, a.z*b.x - a.x*b.z, a.x*b.y - a.y*b.x).
//   * Matrix!(T,3,3): row-major storage c[row][col]; `*` is the ordinary product with
//     sum = 0; sum += c[i][k]*x.c[k][j]; rotateAxis!(i,j)(a): identity with c[i][i]=cos a,
//     c[i][j]=-sin a, c[j][i]=sin a, c[j][j]=cos a; rotateX=(1,2) rotateY=(2,0) rotateZ=(0,1);
//     inverse() for 3x3 is the adjugate times invDet = 1/det; transposed() swaps indices.
//   * radians(x) = x * (PI/180) with PI an 80-bit `real`.
// Call sites this is anchored on: /root/reference/source/rt/imported_types.d:13-20 (mul, which
// fixes the row-vector x row-major convention), camera.d:90-112, transform.d:24-55.
// Behavioural pin: tests/test_oracle_kat.py checks that lecture4.sdl's camera (pitch -30)
// looks down: front = (0,-0.5,0.866).  PARITY UNPINNED by reference tests.
#pragma once
#include "orc_scalar.hpp"

namespace orc {

struct Vec3 {
    real x, y, z;
    Vec3() : x(mk_real(std::numeric_limits<double>::quiet_NaN())), y(x), z(x) {}  // D: double.init is NaN
    Vec3(real x_, real y_, real z_) : x(x_), y(y_), z(z_) {}
    real& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    const real& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};

inline Vec3 operator+(const Vec3& a, const Vec3& b) { return Vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline Vec3 operator-(const Vec3& a, const Vec3& b) { return Vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline Vec3 operator-(const Vec3& a) { return Vec3(-a.x, -a.y, -a.z); }
inline Vec3 operator*(const Vec3& a, real s) { return Vec3(a.x * s, a.y * s, a.z * s); }
inline Vec3 operator*(real s, const Vec3& a) { return Vec3(s * a.x, s * a.y, s * a.z); }

inline real sqlen(const Vec3& a) {
    real s = mk_real(0.0);
    s += a.x * a.x; s += a.y * a.y; s += a.z * a.z;
    return s;
}
inline real length(const Vec3& a) { return r_sqrt(sqlen(a)); }
inline real dot(const Vec3& a, const Vec3& b) {
    real s = mk_real(0.0);
    s += a.x * b.x; s += a.y * b.y; s += a.z * b.z;
    return s;
}
inline Vec3 cross(const Vec3& a, const Vec3& b) {
    return Vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline void normalize(Vec3& a) {
    real inv = mk_real(1.0) / length(a);
    a.x *= inv; a.y *= inv; a.z *= inv;
}
inline Vec3 normalized(Vec3 a) { normalize(a); return a; }

struct Mat3 {
    real c[3][3];
    static Mat3 identity() {
        Mat3 m;
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m.c[i][j] = mk_real(i == j ? 1.0 : 0.0);
        return m;
    }
};

// Row vector times row-major matrix: /root/reference/source/rt/imported_types.d:13-20.
inline Vec3 mul(const Vec3& v, const Mat3& m) {
    return Vec3(v.x * m.c[0][0] + v.y * m.c[1][0] + v.z * m.c[2][0],
                v.x * m.c[0][1] + v.y * m.c[1][1] + v.z * m.c[2][1],
                v.x * m.c[0][2] + v.y * m.c[1][2] + v.z * m.c[2][2]);
}

// ---- host-side (load / begin-frame) matrix helpers: plain double, never counted ----------
struct Mat3d {
    double c[3][3];
};
inline Mat3d m3_identity() {
    Mat3d m{};
    m.c[0][0] = m.c[1][1] = m.c[2][2] = 1.0;
    return m;
}
inline Mat3d m3_mul(const Mat3d& a, const Mat3d& b) {
    Mat3d r;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += a.c[i][k] * b.c[k][j];
            r.c[i][j] = s;
        }
    return r;
}
inline Mat3d m3_inverse(const Mat3d& m) {
    const auto& c = m.c;
    double det = c[0][0] * (c[1][1] * c[2][2] - c[2][1] * c[1][2]) -
                 c[0][1] * (c[1][0] * c[2][2] - c[1][2] * c[2][0]) +
                 c[0][2] * (c[1][0] * c[2][1] - c[1][1] * c[2][0]);
    double inv = 1 / det;
    Mat3d r;
    r.c[0][0] = (c[1][1] * c[2][2] - c[2][1] * c[1][2]) * inv;
    r.c[0][1] = -(c[0][1] * c[2][2] - c[0][2] * c[2][1]) * inv;
    r.c[0][2] = (c[0][1] * c[1][2] - c[0][2] * c[1][1]) * inv;
    r.c[1][0] = -(c[1][0] * c[2][2] - c[1][2] * c[2][0]) * inv;
    r.c[1][1] = (c[0][0] * c[2][2] - c[0][2] * c[2][0]) * inv;
    r.c[1][2] = -(c[0][0] * c[1][2] - c[1][0] * c[0][2]) * inv;
    r.c[2][0] = (c[1][0] * c[2][1] - c[2][0] * c[1][1]) * inv;
    r.c[2][1] = -(c[0][0] * c[2][1] - c[2][0] * c[0][1]) * inv;
    r.c[2][2] = (c[0][0] * c[1][1] - c[1][0] * c[0][1]) * inv;
    return r;
}
inline Mat3d m3_transposed(const Mat3d& m) {
    Mat3d r;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.c[i][j] = m.c[j][i];
    return r;
}
// D's PI is an 80-bit real; x * (PI/180) is evaluated in real and rounded once to double.
inline double d_radians(double deg) {
    return (double)((long double)deg * (3.14159265358979323846264338327950288L / 180.0L));
}
inline Mat3d m3_rotate_axis(int i, int j, double angle) {
    Mat3d m = m3_identity();
    double ca = (double)cosl((long double)angle), sa = (double)sinl((long double)angle);
    m.c[i][i] = ca; m.c[i][j] = -sa; m.c[j][i] = sa; m.c[j][j] = ca;
    return m;
}
inline Mat3d m3_rotate_x(double a) { return m3_rotate_axis(1, 2, a); }
inline Mat3d m3_rotate_y(double a) { return m3_rotate_axis(2, 0, a); }
inline Mat3d m3_rotate_z(double a) { return m3_rotate_axis(0, 1, a); }
inline void d_mul(const double v[3], const Mat3d& m, double out[3]) {
    double r0 = v[0] * m.c[0][0] + v[1] * m.c[1][0] + v[2] * m.c[2][0];
    double r1 = v[0] * m.c[0][1] + v[1] * m.c[1][1] + v[2] * m.c[2][1];
    double r2 = v[0] * m.c[0][2] + v[1] * m.c[1][2] + v[2] * m.c[2][2];
    out[0] = r0; out[1] = r1; out[2] = r2;
}
inline Mat3 to_counted(const Mat3d& m) {
    Mat3 r;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.c[i][j] = mk_real(m.c[i][j]);
    return r;
}

}  // namespace orc
