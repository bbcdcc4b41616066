// geometry.h — host-side vector/matrix types of the drop-in facade.
//
// Same type names and the same arithmetic semantics as the reference's src/geometry.h (Vector<N,T>,
// Matrix<R,C,T>, Dot/Cross/Normalize, Make*Matrix), written from scratch as plain structs.  The operation ORDER
// is part of the contract: the uniforms the host hands to the kernels must equal the
// reference's bit for bit, so
//   * Dot accumulates left to right from 0.f                       (reference geometry.h:881-889)
//   * vector / scalar is "inv = 1/f, then multiply"                (geometry.h:335-341, 519-524)
//   * Normalize(v) = v * 1 / |v|  -> reciprocal then 3 multiplies   (geometry.h:906-910)
//   * matrix x vector = per-row Dot, matrix x matrix = Dot(row, col) (geometry.h:774-793)
//   * inverse = adjugate / det by cofactor expansion along row 0   (geometry.h:604-742)
// This file must be compiled without -ffast-math / -mfma (x86-64 SSE2 scalar, like the oracle).
#pragma once

#include <cmath>
#include <cstddef>
#include <cstdint>

typedef float Float;

static const Float Pi = 3.14159265358979323846;
static const Float InvPi = 0.31830988618379067154;
static const Float Gamma = 2.2;
static const Float InvGamma = 1.f / 2.2f;

inline Float Radians(Float deg) { return deg * Pi / 180.f; }

template <typename T>
inline T Clamp(T v, T lo, T hi)
{
    return std::min(hi, std::max(v, lo));
}

struct Vector2f
{
    union { struct { Float x, y; }; struct { Float u, v; }; struct { Float s, t; }; };
    Vector2f() : x(0), y(0) {}
    Vector2f(Float xx, Float yy) : x(xx), y(yy) {}
    explicit Vector2f(Float a) : x(a), y(a) {}
    Float  operator[](size_t i) const { return i == 0 ? x : y; }
    Float& operator[](size_t i) { return i == 0 ? x : y; }
    Vector2f operator+(const Vector2f& o) const { return Vector2f(x + o.x, y + o.y); }
    Vector2f operator-(const Vector2f& o) const { return Vector2f(x - o.x, y - o.y); }
    Vector2f operator*(Float f) const { return Vector2f(x * f, y * f); }
    Vector2f& operator*=(Float f) { x *= f; y *= f; return *this; }
};

struct Vector2i
{
    int x, y;
    Vector2i() : x(0), y(0) {}
    Vector2i(int xx, int yy) : x(xx), y(yy) {}
};

struct Vector3f
{
    union { struct { Float x, y, z; }; struct { Float r, g, b; }; };
    Vector3f() : x(0), y(0), z(0) {}
    Vector3f(Float xx, Float yy, Float zz) : x(xx), y(yy), z(zz) {}
    explicit Vector3f(Float a) : x(a), y(a), z(a) {}
    Float  operator[](size_t i) const { return i == 0 ? x : (i == 1 ? y : z); }
    Float& operator[](size_t i) { return i == 0 ? x : (i == 1 ? y : z); }
    Vector3f operator-() const { return Vector3f(-x, -y, -z); }
    bool operator==(const Vector3f& o) const { return x == o.x && y == o.y && z == o.z; }
    bool operator!=(const Vector3f& o) const { return !(*this == o); }
    Vector3f operator+(const Vector3f& o) const { return Vector3f(x + o.x, y + o.y, z + o.z); }
    Vector3f operator-(const Vector3f& o) const { return Vector3f(x - o.x, y - o.y, z - o.z); }
    Vector3f& operator+=(const Vector3f& o) { x += o.x; y += o.y; z += o.z; return *this; }
    Vector3f operator*(const Vector3f& o) const { return Vector3f(x * o.x, y * o.y, z * o.z); }
    Vector3f operator*(Float f) const { return Vector3f(x * f, y * f, z * f); }
    Vector3f operator/(Float f) const { Float inv = (Float)1 / f; return Vector3f(x * inv, y * inv, z * inv); }
    Vector3f& operator*=(Float f) { x *= f; y *= f; z *= f; return *this; }
    Float LengthSquared() const { return x * x + y * y + z * z; }
    Float Length() const { return std::sqrt(LengthSquared()); }
};

struct Vector4f
{
    Float x, y, z, w;
    Vector4f() : x(0), y(0), z(0), w(0) {}
    Vector4f(Float xx, Float yy, Float zz, Float ww) : x(xx), y(yy), z(zz), w(ww) {}
    Vector4f(const Vector3f& v, Float ww) : x(v.x), y(v.y), z(v.z), w(ww) {}
    Float  operator[](size_t i) const { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
    Float& operator[](size_t i) { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
    Vector3f xyz() const { return Vector3f(x, y, z); }
    Vector4f operator*(Float f) const { return Vector4f(x * f, y * f, z * f, w * f); }
    Vector4f operator/(Float f) const
    {
        Float inv = (Float)1 / f;
        return Vector4f(x * inv, y * inv, z * inv, w * inv);
    }
};

typedef Vector2f Point2f;
typedef Vector2i Point2i;
typedef Vector3f Point3f;
typedef Vector4f Point4f;
typedef Vector3f Color3;

inline Float Dot(const Vector3f& a, const Vector3f& b)
{
    Float r = 0.f;
    r += a.x * b.x;
    r += a.y * b.y;
    r += a.z * b.z;
    return r;
}
inline Float Dot(const Vector4f& a, const Vector4f& b)
{
    Float r = 0.f;
    r += a.x * b.x;
    r += a.y * b.y;
    r += a.z * b.z;
    r += a.w * b.w;
    return r;
}
inline Vector3f Cross(const Vector3f& a, const Vector3f& b)
{
    return Vector3f(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline Vector3f Normalize(const Vector3f& v)
{
    return (v * (Float)1) / v.Length();
}
inline Vector3f operator*(Float s, const Vector3f& v) { return v * s; }

struct Matrix3x3f
{
    Vector3f rows[3];
    Vector3f&       operator[](size_t i) { return rows[i]; }
    const Vector3f& operator[](size_t i) const { return rows[i]; }
};

struct Matrix4x4f
{
    Vector4f rows[4];
    Matrix4x4f() {}
    explicit Matrix4x4f(Float diag)
    {
        for (int i = 0; i < 4; ++i) rows[i][i] = diag;
    }
    static Matrix4x4f Identity() { return Matrix4x4f(1.f); }
    Vector4f&       operator[](size_t i) { return rows[i]; }
    const Vector4f& operator[](size_t i) const { return rows[i]; }
    Vector4f Col(size_t c) const { return Vector4f(rows[0][c], rows[1][c], rows[2][c], rows[3][c]); }
    void SetCol(size_t c, const Vector4f& v)
    {
        for (int i = 0; i < 4; ++i) rows[i][c] = v[i];
    }
    void SetRow(size_t r, const Vector4f& v) { rows[r] = v; }
};

inline Vector4f operator*(const Matrix4x4f& m, const Vector4f& v)
{
    return Vector4f(Dot(m[0], v), Dot(m[1], v), Dot(m[2], v), Dot(m[3], v));
}
inline Vector3f operator*(const Matrix3x3f& m, const Vector3f& v)
{
    return Vector3f(Dot(m[0], v), Dot(m[1], v), Dot(m[2], v));
}
inline Matrix4x4f operator*(const Matrix4x4f& a, const Matrix4x4f& b)
{
    Matrix4x4f r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) r[i][j] = Dot(a[i], b.Col(j));
    return r;
}

Matrix3x3f MakeNormalMatrix(const Matrix4x4f& m);
Matrix4x4f MakeModelMatrix(const Vector3f& translation, Float yRotate = 0.f, Float scale = 1.f);
Matrix4x4f MakeLookAtMatrix(const Vector3f& eyePos, const Vector3f& center,
                            const Vector3f& worldUp = Vector3f(0.f, 1.f, 0.f));
Matrix4x4f MakePerspectiveMatrix(Float fov, Float aspectRatio, Float n, Float f);
Matrix4x4f MakeOrthographicMatrix(Float l, Float r, Float b, Float t, Float n, Float f);
