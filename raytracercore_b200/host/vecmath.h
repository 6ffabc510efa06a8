// vecmath.h — host-side 4-vector / 4x4-matrix arithmetic with the evaluation order of the reference's
// Vectors/{Vec4D,Mat4x4D,SIMDHelpers,MatrixTransforms}.cs, so that scenes flattened by this host layer carry the
// same f64 values the reference's SceneLoader would produce. Translation units including this header must be
// compiled with -ffp-contract=off.
#pragma once
#include <cmath>

namespace rtcore {

struct Vec4D {  // Vectors/Vec4D.cs:80-91
  double X = 0, Y = 0, Z = 0, W = 0;
  Vec4D() = default;
  Vec4D(double x, double y, double z, double w) : X(x), Y(y), Z(z), W(w) {}
};

inline Vec4D operator+(const Vec4D& a, const Vec4D& b) { return {a.X + b.X, a.Y + b.Y, a.Z + b.Z, a.W + b.W}; }
inline Vec4D operator-(const Vec4D& a, const Vec4D& b) { return {a.X - b.X, a.Y - b.Y, a.Z - b.Z, a.W - b.W}; }
inline Vec4D operator-(const Vec4D& a) { return {-a.X, -a.Y, -a.Z, -a.W}; }
inline Vec4D operator*(const Vec4D& a, double s) { return {a.X * s, a.Y * s, a.Z * s, a.W * s}; }
inline Vec4D operator/(const Vec4D& a, double s) { return {a.X / s, a.Y / s, a.Z / s, a.W / s}; }
inline bool operator==(const Vec4D& a, const Vec4D& b) { return a.X == b.X && a.Y == b.Y && a.Z == b.Z; }  // Vec4D.cs:480
inline bool operator!=(const Vec4D& a, const Vec4D& b) { return !(a == b); }

inline double Dot(const Vec4D& a, const Vec4D& b) { return a.X * b.X + a.Y * b.Y + a.Z * b.Z + a.W * b.W; }  // :341-347
inline Vec4D Cross(const Vec4D& a, const Vec4D& b) {  // :355-364 (SIMDCross = false)
  return {a.Y * b.Z - a.Z * b.Y, a.Z * b.X - a.X * b.Z, a.X * b.Y - a.Y * b.X, 0};
}
inline double SquaredLength(const Vec4D& a) { return a.X * a.X + a.Y * a.Y + a.Z * a.Z + a.W * a.W; }
inline double Length(const Vec4D& a) { return std::sqrt(SquaredLength(a)); }
inline Vec4D Normalize(const Vec4D& a) {  // SIMDNormalize = true -> SIMDHelpers.Normalize, SIMDHelpers.cs:332-335
  double s = (a.X * a.X + a.Y * a.Y) + (a.Z * a.Z + a.W * a.W);
  double l = std::sqrt(s);
  return {a.X / l, a.Y / l, a.Z / l, a.W / l};
}

struct Mat4x4D {  // Vectors/Mat4x4D.cs:21-39, row-major
  double D[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  static Mat4x4D Identity() { return Mat4x4D(); }
  bool operator==(const Mat4x4D& o) const {
    for (int i = 0; i < 16; i++)
      if (D[i] != o.D[i]) return false;
    return true;
  }
  bool operator!=(const Mat4x4D& o) const { return !(*this == o); }
  Mat4x4D Transpose3x3() const {  // Mat4x4D.cs:78-85
    Mat4x4D r;
    const double* d = D;
    double t[16] = {d[0], d[4], d[8], 0, d[1], d[5], d[9], 0, d[2], d[6], d[10], 0, 0, 0, 0, 1};
    for (int i = 0; i < 16; i++) r.D[i] = t[i];
    return r;
  }
};

// Mat4x4D operator* (Mat4x4D.cs:97-124, SIMD branch): result[y][x] = Vector.Dot(row y of left, column x of right);
// the 4-lane dot is taken pairwise, (p0+p1)+(p2+p3).
inline Mat4x4D operator*(const Mat4x4D& l, const Mat4x4D& r) {
  Mat4x4D o;
  for (int y = 0; y < 4; y++)
    for (int x = 0; x < 4; x++) {
      double p0 = l.D[y * 4 + 0] * r.D[0 * 4 + x], p1 = l.D[y * 4 + 1] * r.D[1 * 4 + x];
      double p2 = l.D[y * 4 + 2] * r.D[2 * 4 + x], p3 = l.D[y * 4 + 3] * r.D[3 * 4 + x];
      o.D[y * 4 + x] = (p0 + p1) + (p2 + p3);
    }
  return o;
}

// Mat4x4D * Vec4D (Mat4x4D.cs:171-180 -> SIMDHelpers.MultiplyMatrixVector/Sum4, SIMDHelpers.cs:111-127,222-237)
inline Vec4D operator*(const Mat4x4D& m, const Vec4D& v) {
  const double* d = m.D;
  return {(d[0] * v.X + d[1] * v.Y) + (d[2] * v.Z + d[3] * v.W), (d[4] * v.X + d[5] * v.Y) + (d[6] * v.Z + d[7] * v.W),
          (d[8] * v.X + d[9] * v.Y) + (d[10] * v.Z + d[11] * v.W),
          (d[12] * v.X + d[13] * v.Y) + (d[14] * v.Z + d[15] * v.W)};
}

namespace MatrixTransforms {  // Vectors/MatrixTransforms.cs
inline Mat4x4D Translate(double x, double y, double z) {
  Mat4x4D m;
  m.D[3] = x;
  m.D[7] = y;
  m.D[11] = z;
  return m;
}
inline Mat4x4D Scale(double x, double y, double z) {
  Mat4x4D m;
  m.D[0] = x;
  m.D[5] = y;
  m.D[10] = z;
  return m;
}
inline Mat4x4D Rotate(double angle, const Vec4D& a) {  // :25-38
  double c = std::cos(angle), s = std::sin(angle), co = 1 - c;
  Mat4x4D m;
  double t[16] = {c + a.X * a.X * co,       a.X * a.Y * co - a.Z * s, a.X * a.Z * co + a.Y * s, 0,
                  a.Y * a.X * co + a.Z * s, c + a.Y * a.Y * co,       a.Y * a.Z * co - a.X * s, 0,
                  a.Z * a.X * co - a.Y * s, a.Z * a.Y * co + a.X * s, c + a.Z * a.Z * co,       0,
                  0,                        0,                        0,                        1};
  for (int i = 0; i < 16; i++) m.D[i] = t[i];
  return m;
}
}  // namespace MatrixTransforms

inline double toRadians(double deg) { return deg * (3.14159265358979323846 / 180); }  // Consts.cs:10-15

}  // namespace rtcore
