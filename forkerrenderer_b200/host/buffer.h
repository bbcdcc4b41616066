// buffer.h — Buffer1f / Buffer3f of the drop-in facade (public surface of reference src/buffer.h:32-93).
//
// In the reference these are std::vector-backed host arrays.  Here the pixels live in device planes owned by
// the fgl context (SoA fp32); a Buffer object is a handle (plane id + size) with a lazily
// synchronised host mirror so that the per-pixel accessors GetValue/SetValue keep working for tools and tests.
// They are never used on the fast path.
#pragma once

#include <vector>

#include "geometry.h"
#include "tgaimage.h"

class Buffer
{
public:
    enum InitType { Zero, One, MaxPositive, MinNegative };

    int GetWidth() const { return m_Width; }
    int GetHeight() const { return m_Height; }
    int GetPlane() const { return m_Plane; }

protected:
    Buffer(int w, int h, int plane, int channels) : m_Width(w), m_Height(h), m_Plane(plane), m_Channels(channels) {}
    virtual ~Buffer() = default;

    // host mirror management (device -> host on first read after device work, host -> device before device work)
    void pull() const;
    void push();
    friend struct ForkerGL;

    int                        m_Width, m_Height;
    int                        m_Plane;     // FGL_PLANE_* or a user plane id; -1 = empty default-constructed buffer
    int                        m_Channels;
    mutable std::vector<Float> m_Host;      // AoS mirror, reference layout
    mutable bool               m_HostValid = false;
    bool                       m_HostDirty = false;
};

class Buffer1f : public Buffer
{
public:
    Buffer1f() : Buffer(0, 0, -1, 1) {}
    explicit Buffer1f(int w, int h, InitType type);                 // standalone buffer (user plane)
    Buffer1f(int w, int h, int plane) : Buffer(w, h, plane, 1) {}   // view of a ForkerGL plane

    Float GetValue(int x, int y) const { pull(); return m_Host[(size_t)x + (size_t)y * m_Width]; }
    void  SetValue(int x, int y, Float v) { pull(); m_Host[(size_t)x + (size_t)y * m_Width] = v; m_HostDirty = true; }

    TGAImage GenerateImage(bool inverseColor = false) const;

    void SimpleBlurDenoised();
    void TwoPassGaussianBlurDenoised();
};

class Buffer3f : public Buffer
{
public:
    Buffer3f() : Buffer(0, 0, -1, 3) {}
    explicit Buffer3f(int w, int h, InitType type);
    Buffer3f(int w, int h, int plane) : Buffer(w, h, plane, 3) {}

    Vector3f GetValue(int x, int y) const
    {
        pull();
        const Float* p = &m_Host[((size_t)x + (size_t)y * m_Width) * 3];
        return Vector3f(p[0], p[1], p[2]);
    }
    void SetValue(int x, int y, const Vector3f& v)
    {
        pull();
        Float* p = &m_Host[((size_t)x + (size_t)y * m_Width) * 3];
        p[0] = v.x, p[1] = v.y, p[2] = v.z;
        m_HostDirty = true;
    }

    TGAImage GenerateImage() const;
    void     PaintColor(const Color3& color);

    void SimpleBlurDenoised();
    void TwoPassGaussianBlurDenoised();
};
