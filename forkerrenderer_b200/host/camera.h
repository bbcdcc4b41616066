// camera.h — Camera (reference src/camera.h:13-50, camera.cpp:34-57): eye / look-at holder delegating to the
// matrix builders in geometry.cpp.
#pragma once
#include "geometry.h"

class Camera
{
public:
    enum ProjectionType { Orthographic, Perspective };

    explicit Camera(const Point3f& lookFrom, const Point3f& lookAt = Point3f(0.f, 0.f, 0.f))
        : m_EyePos(lookFrom), m_LookAtPos(lookAt), m_WorldUp(0.f, 1.f, 0.f)
    {
    }

    void    SetPosition(Float x, Float y, Float z) { m_EyePos = Point3f(x, y, z); }
    void    SetLookAtPos(Float x, Float y, Float z) { m_LookAtPos = Point3f(x, y, z); }
    Point3f GetPosition() const { return m_EyePos; }
    Point3f GetLookAt() const { return m_LookAtPos; }

    Matrix4x4f GetViewMatrix() const { return MakeLookAtMatrix(m_EyePos, m_LookAtPos, m_WorldUp); }
    Matrix4x4f GetPerspectiveMatrix(Float fov, Float aspectRatio, Float n, Float f) const
    {
        return MakePerspectiveMatrix(fov, aspectRatio, n, f);
    }
    Matrix4x4f GetOrthographicMatrix(Float l, Float r, Float b, Float t, Float n, Float f) const
    {
        return MakeOrthographicMatrix(l, r, b, t, n, f);
    }

private:
    Point3f  m_EyePos, m_LookAtPos;
    Vector3f m_WorldUp;
};
