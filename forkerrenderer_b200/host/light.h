// light.h — PointLight / DirLight (reference src/light.h:11-61).  DirLight is parsed but used by no pass.
#pragma once
#include "geometry.h"

class Light
{
public:
    Color3 color;
    explicit Light(const Color3& c) : color(c) {}
    virtual ~Light() = default;
};

class DirLight : public Light
{
public:
    Point3f  position;
    Vector3f direction;
    DirLight() : Light(Color3(1.f)), position(0.f), direction(0, 0, -1) {}
    DirLight(const Vector3f& dir, const Point3f& pos, const Color3& c = Color3(1.f))
        : Light(c), position(pos), direction(Normalize(dir))
    {
    }
};

class PointLight : public Light
{
public:
    Vector3f position;
    PointLight() : Light(Vector3f(1.f)), position(0.f) {}
    PointLight(Float x, Float y, Float z, const Vector3f& c = Vector3f(1.f)) : Light(c), position(x, y, z) {}
    explicit PointLight(const Vector3f& pos, const Vector3f& c = Vector3f(1.f)) : Light(c), position(pos) {}
};
