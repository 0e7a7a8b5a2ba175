// rt/math.hpp -- host-side FP64 math of the scene surface.
//
// Mirrors the reference's value types so scene code written against the
// reference compiles against this header unchanged:
//   Vector3 / Point3 / Color   reference Vec3.h:10-141
//   Interval                   reference Interval.h:8-75
//   Aabb                       reference AABB.h:18-135
// Only construction-time arithmetic lives here (bounding boxes, cached quad
// constants, camera frame).  Nothing in this header intersects rays: the
// render path is the CUDA library behind include/rt_abi.h and there is no CPU
// fallback.
#pragma once

#include <cfloat>
#include <cmath>

namespace rt {

struct Vector3 {
    double E[3];

    Vector3() : E{0.0, 0.0, 0.0} {}
    Vector3(double x, double y, double z) : E{x, y, z} {}

    double X() const { return E[0]; }
    double Y() const { return E[1]; }
    double Z() const { return E[2]; }
    double operator[](int i) const { return E[i]; }
    double& operator[](int i) { return E[i]; }

    Vector3 operator-() const { return Vector3(-E[0], -E[1], -E[2]); }
    Vector3& operator+=(const Vector3& o)
    {
        for (int k = 0; k < 3; ++k) E[k] += o.E[k];
        return *this;
    }
    Vector3& operator*=(double s)
    {
        for (int k = 0; k < 3; ++k) E[k] *= s;
        return *this;
    }
    // The reference divides by multiplying with the reciprocal (Vec3.h:44-47).
    Vector3& operator/=(double s) { return *this *= 1 / s; }

    double LengthSquared() const { return E[0] * E[0] + E[1] * E[1] + E[2] * E[2]; }
    double Length() const { return std::sqrt(LengthSquared()); }
};

using Point3 = Vector3;
using Color = Vector3;

inline Vector3 operator+(const Vector3& a, const Vector3& b)
{
    return Vector3(a.E[0] + b.E[0], a.E[1] + b.E[1], a.E[2] + b.E[2]);
}
inline Vector3 operator-(const Vector3& a, const Vector3& b)
{
    return Vector3(a.E[0] - b.E[0], a.E[1] - b.E[1], a.E[2] - b.E[2]);
}
inline Vector3 operator*(const Vector3& a, const Vector3& b)
{
    return Vector3(a.E[0] * b.E[0], a.E[1] * b.E[1], a.E[2] * b.E[2]);
}
inline Vector3 operator*(double s, const Vector3& v) { return Vector3(s * v.E[0], s * v.E[1], s * v.E[2]); }
inline Vector3 operator*(const Vector3& v, double s) { return s * v; }
inline Vector3 operator/(const Vector3& v, double s) { return (1 / s) * v; } // Vec3.h:96-99
inline double Dot(const Vector3& a, const Vector3& b)
{
    return a.E[0] * b.E[0] + a.E[1] * b.E[1] + a.E[2] * b.E[2];
}
inline Vector3 Cross(const Vector3& a, const Vector3& b)
{
    return Vector3(a.E[1] * b.E[2] - a.E[2] * b.E[1], a.E[2] * b.E[0] - a.E[0] * b.E[2],
                   a.E[0] * b.E[1] - a.E[1] * b.E[0]);
}
inline Vector3 UnitVector(const Vector3& v) { return v / v.Length(); }

// Interval.h:8-75.  Default = empty (+max, -max).
struct Interval {
    double Min, Max;
    Interval() : Min(+DBL_MAX), Max(-DBL_MAX) {}
    Interval(double lo, double hi) : Min(lo), Max(hi) {}
    Interval(const Interval& a, const Interval& b)
        : Min(a.Min <= b.Min ? a.Min : b.Min), Max(a.Max >= b.Max ? a.Max : b.Max)
    {
    }
    double Size() const { return Max - Min; }
    Interval Expand(double delta) const
    {
        const double pad = delta / 2.0;
        return Interval(Min - pad, Max + pad);
    }
};
inline Interval operator+(const Interval& iv, double d) { return Interval(iv.Min + d, iv.Max + d); }

// AABB.h:18-135.  The point-pair and interval constructors pad thin axes to
// 1e-4 (AABB.h:114-120); the union constructor does not (AABB.h:43-48).
struct Aabb {
    Interval X, Y, Z;

    Aabb() {}
    Aabb(const Interval& x, const Interval& y, const Interval& z) : X(x), Y(y), Z(z) { PadToMinimums(); }
    Aabb(const Point3& a, const Point3& b)
    {
        X = (a[0] <= b[0]) ? Interval(a[0], b[0]) : Interval(b[0], a[0]);
        Y = (a[1] <= b[1]) ? Interval(a[1], b[1]) : Interval(b[1], a[1]);
        Z = (a[2] <= b[2]) ? Interval(a[2], b[2]) : Interval(b[2], a[2]);
        PadToMinimums();
    }
    Aabb(const Aabb& p, const Aabb& q) : X(p.X, q.X), Y(p.Y, q.Y), Z(p.Z, q.Z) {}

    const Interval& AxisInterval(int n) const { return n == 1 ? Y : (n == 2 ? Z : X); }

    // AABB.h:101-107: X only when strictly longest; ties fall towards Z.
    int LongestAxis() const
    {
        if (X.Size() > Y.Size()) return X.Size() > Z.Size() ? 0 : 2;
        return Y.Size() > Z.Size() ? 1 : 2;
    }

private:
    void PadToMinimums()
    {
        const double delta = 0.0001;
        if (X.Size() < delta) X = X.Expand(delta);
        if (Y.Size() < delta) Y = Y.Expand(delta);
        if (Z.Size() < delta) Z = Z.Expand(delta);
    }
};
inline Aabb operator+(const Aabb& b, const Vector3& off)
{
    return Aabb(b.X + off.X(), b.Y + off.Y(), b.Z + off.Z());
}

} // namespace rt
