// geometry.cpp — host-side matrix builders with the reference's exact expression order.
// The four builders (model, look-at, perspective, orthographic; reference src/geometry.cpp:92-179) are TRANSCRIBED: the uniforms
// they produce have to equal the reference's bit for bit, and for these five-line textbook formulas
// that means the same expressions in the same order (checked word for word against the reference, tests/test_host_math.py).
// The normal matrix (src/geometry.cpp:60-68) and the cofactor-expansion determinant / inverse (src/geometry.h:604-742) are
// written afresh around the same evaluation tree.  See geometry.h for the rules.
#include "geometry.h"

namespace
{
// Determinant by cofactor expansion along row 0, recursing down to 1x1 — the same evaluation tree as the
// reference's dt<DIM,T>::det (geometry.h:604-624): ret = 0; ret += m[0][i] * (det(minor(0,i)) * sign).
Float DetN(const Float* m, int n)
{
    if (n == 1) return m[0];
    Float ret = 0;
    Float minor[9];
    for (int i = 0; i < n; ++i)
    {
        int k = 0;
        for (int r = 1; r < n; ++r)
            for (int c = 0; c < n; ++c)
                if (c != i) minor[k++] = m[r * n + c];
        Float cof = DetN(minor, n - 1) * ((0 + i) % 2 ? -1 : 1);
        ret += m[i] * cof;
    }
    return ret;
}

Float Cofactor4(const Matrix4x4f& m, int row, int col)
{
    Float minor[9];
    int   k = 0;
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c)
            if (r != row && c != col) minor[k++] = m[r][c];
    return DetN(minor, 3) * ((row + col) % 2 ? -1 : 1);
}
}  // namespace

// (M^-1)^T upper-left 3x3: inverse = adjugate / det, adjugate = cofactor matrix transposed, so the
// inverse-transpose is the cofactor matrix scaled by 1/det (reference geometry.cpp:60-68, geometry.h:726-742).
Matrix3x3f MakeNormalMatrix(const Matrix4x4f& m)
{
    Float flat[16];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) flat[r * 4 + c] = m[r][c];
    Float det = DetN(flat, 4);
    Float inv = (Float)1 / det;
    Matrix3x3f out;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) out[r][c] = Cofactor4(m, r, c) * inv;
    return out;
}

// T * R * S with a rotation about +Y (reference geometry.cpp:92-108).  cos/sin are the unqualified C
// functions on a float argument (double evaluation, rounded on assignment), as in the reference.
Matrix4x4f MakeModelMatrix(const Vector3f& translation, Float yRotate, Float scale)
{
    Matrix4x4f S(1.f);
    S[0][0] = S[1][1] = S[2][2] = scale;

    Matrix4x4f R(1.f);
    Float      radVal = Radians(yRotate);
    R[0][0] = cos(radVal);
    R[0][2] = sin(radVal);
    R[2][0] = -sin(radVal);
    R[2][2] = cos(radVal);

    Matrix4x4f T(1.f);
    T.SetCol(3, Vector4f(translation, 1.f));
    return T * R * S;
}

// reference geometry.cpp:110-128
Matrix4x4f MakeLookAtMatrix(const Vector3f& eyePos, const Vector3f& center, const Vector3f& worldUp)
{
    Vector3f front = Normalize(center - eyePos);
    Vector3f right = Normalize(Cross(front, worldUp));
    Vector3f up = Normalize(Cross(right, front));

    Matrix4x4f R(1.f);
    R.SetRow(0, Vector4f(right, 0.f));
    R.SetRow(1, Vector4f(up, 0.f));
    R.SetRow(2, Vector4f(-front, 0.f));

    Matrix4x4f T(1.f);
    T.SetCol(3, Vector4f(-eyePos, 1.f));
    return R * T;
}

// reference geometry.cpp:130-145.  Note m[3][3] keeps the identity's 1 (the reference never clears it).
Matrix4x4f MakePerspectiveMatrix(Float fov, Float aspectRatio, Float n, Float f)
{
    Float      tanFovOver2 = std::tan(Radians(fov / 2.f));
    Matrix4x4f m(1.f);
    m[0][0] = 1.f / (aspectRatio * tanFovOver2);
    m[1][1] = 1.f / tanFovOver2;
    m[2][2] = -(f + n) / (f - n);
    m[2][3] = -2 * f * n / (f - n);
    m[3][2] = -1;
    return m;
}

// reference geometry.cpp:166-179
Matrix4x4f MakeOrthographicMatrix(Float l, Float r, Float b, Float t, Float n, Float f)
{
    Matrix4x4f m(1.f);
    m[0][0] = 2.f / (r - l);
    m[1][1] = 2.f / (t - b);
    m[2][2] = -2.f / (f - n);
    m[0][3] = -(r + l) / (r - l);
    m[1][3] = -(t + b) / (t - b);
    m[2][3] = -(f + n) / (f - n);
    m[3][3] = 1.f;
    return m;
}
